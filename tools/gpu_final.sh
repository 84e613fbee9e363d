#!/bin/bash
# Round-end check on one B200: GPU tests, smoke, every bench workload (logs copied to profiles/ by hand afterwards)
mkdir -p gpurun_out; rm -f gpurun_out/status.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/status.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench ref exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.log 2> gpurun_out/bench_ours.err; echo "bench exit $?" >> gpurun_out/status.txt
for w in fwd_bwd small_batch config3 train; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "bench $w exit $?" >> gpurun_out/status.txt
done
cat gpurun_out/status.txt; tail -2 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log
