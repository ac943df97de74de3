#!/bin/bash
# One gpurun call: full GPU test suite, default bench (both arms), ncu launch list + full captures of the top kernels.
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v Warn | tail -2
echo "=== bench ours"; timeout 900 python bench.py > gpurun_out/bench_final.log 2>gpurun_out/bench_final.err; echo "exit $?"; tail -c 3000 gpurun_out/bench_final.log
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "exit $?"; tail -c 1500 gpurun_out/bench_ref.log
echo "=== read bandwidth sanity"; python - <<'PY'
import torch
x = torch.empty(2 * 1024**3, dtype=torch.bfloat16, device="cuda").fill_(1)
for _ in range(2): x.sum()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): x.sum()
e1.record(); torch.cuda.synchronize()
print("torch sum() read-only bandwidth: %.2f TB/s" % (5 * x.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e12))
y = torch.empty_like(x)
e0.record()
for _ in range(5): y.copy_(x)
e1.record(); torch.cuda.synchronize()
print("torch copy_ read+write bandwidth: %.2f TB/s" % (5 * 2 * x.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e12))
PY
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; tail -1 gpurun_out/prof_step.log
echo "=== ncu full: gemm"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 2 -c 1 -o gpurun_out/prof_gemm_qkv_final python tools/profile_gemm.py qkv > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 2 -c 1 -o gpurun_out/prof_gemm_decgu_final python tools/profile_gemm.py decgu >> gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
echo "=== ncu full: attention + decode kernels"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd" -s 7 -c 2 -o gpurun_out/prof_attn_final python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemv_kernel|decode_attn" -s 30 -c 5 -o gpurun_out/prof_decode_final python tools/profile_decode.py > gpurun_out/ncu_decode.log 2>&1; tail -1 gpurun_out/ncu_decode.log
ls -la gpurun_out/*.ncu-rep
