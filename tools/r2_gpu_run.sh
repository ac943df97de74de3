#!/bin/bash
# One gpurun call (round 2): full GPU test suite, default bench, ncu launch list of one step, ncu --set full of the persistent decode
# kernel and of the attention kernels, in-kernel trace.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r2_prof_attn.ncu-rep gpurun_out/r2_prof_gemm.ncu-rep
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Warning | tail -8
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v Warn | tail -2
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.log 2>gpurun_out/r2_bench.err; echo "exit $?"; tail -c 4500 gpurun_out/r2_bench.log
if [ "${SR_R2_DECODE_CAPTURES:-0}" = "1" ]; then   # the decode kernel did not change after its round-2 captures
echo "=== trace"; timeout 120 python tools/persistent_trace.py > gpurun_out/r2_trace.log 2>&1; tail -28 gpurun_out/r2_trace.log
echo "=== ncu full: persistent decode kernel"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 3 -c 1 -o gpurun_out/r2_prof_decode_persistent python tools/profile_persistent.py > gpurun_out/r2_ncu_dp.log 2>&1; tail -1 gpurun_out/r2_ncu_dp.log
fi
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py > gpurun_out/r2_prof_step.log 2>&1; tail -1 gpurun_out/r2_prof_step.log
echo "=== ncu full: attention"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd" -s 7 -c 2 -o gpurun_out/r2_prof_attn python tools/profile_step.py > gpurun_out/r2_ncu_attn.log 2>&1; tail -1 gpurun_out/r2_ncu_attn.log
echo "=== ncu full: GEMM (ViT qkv, ViT proj)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 2 -c 2 -o gpurun_out/r2_prof_gemm python tools/profile_gemm.py qkv proj > gpurun_out/r2_ncu_gemm.log 2>&1; tail -1 gpurun_out/r2_ncu_gemm.log
echo "=== stand-alone GEMM / attention timings"; timeout 200 python tools/gemm_bench.py 2>&1 | tail -10; timeout 200 python tools/attn_bench.py 2>&1 | grep tiles= | cut -c1-250
ls -la gpurun_out/*.ncu-rep
