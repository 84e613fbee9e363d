#!/bin/bash
# one gpurun call: bring-up probe, parity tests, smoke, bench (each under its own timeout)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for v in 0 1; do
  timeout 300 python tools/tc_debug.py $v > gpurun_out/tc_debug_v$v.log 2>&1
  echo "tc_debug v$v exit $?" >> gpurun_out/status.txt
done
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/status.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --engine simt --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_simt.log 2>&1
echo "bench simt exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --engine tc3x --steps 10 --warmup 3 > gpurun_out/bench_tc3x.log 2>&1
echo "bench tc3x exit $?" >> gpurun_out/status.txt
tail -3 gpurun_out/status.txt; tail -5 gpurun_out/tc_debug_v0.log; tail -15 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_tc3x.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --engine tc3x --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/status.txt
