#!/bin/bash
# GPU call: persistent decode bring-up (correctness + timing), then e2e tests and bench with both decode paths.
set -u
mkdir -p gpurun_out
echo "=== bring-up"; timeout 600 python tools/persistent_bringup.py 36 151936 > gpurun_out/bringup.log 2>&1; echo "exit $?"; grep -v Warn gpurun_out/bringup.log | tail -70
if grep -q "correctness: ALL OK" gpurun_out/bringup.log; then
  echo "=== pytest persistent + e2e"; timeout 900 python -m pytest tests/test_gpu_decode_persistent.py tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -8
  echo "=== bench persistent"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_persist.log 2>gpurun_out/bench_persist.err; echo "exit $?"; tail -c 2600 gpurun_out/bench_persist.log; tail -3 gpurun_out/bench_persist.err
fi
echo "=== bench per-layer kernels"; SR_DECODE_PERSISTENT=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_old.log 2>gpurun_out/bench_old.err; echo "exit $?"; tail -c 2600 gpurun_out/bench_old.log; tail -3 gpurun_out/bench_old.err
