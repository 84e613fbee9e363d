#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/status.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/status.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench ref exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.log 2>&1; echo "bench exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt; tail -2 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; tail -1 gpurun_out/bench_reference.log | cut -c1-300; tail -1 gpurun_out/bench_ours.log
