#!/bin/bash
# Round 2, GPU call 3: chain3 with smem biases / look-ahead / 2+4 buffer split / merged scale+pack, gather A/B, bench.
mkdir -p gpurun_out
timeout 400 python tools/r2_chain3_check.py 2>&1 | grep -v Warn > gpurun_out/chain3_check.log; echo "chain3 exit $?" >> gpurun_out/chain3_check.log
timeout 200 python tools/trace_chain3.py mlp 2>&1 | grep -v Warn > gpurun_out/trace3_mlp.log
timeout 200 python tools/trace_chain3.py front 2>&1 | grep -v Warn > gpurun_out/trace3_front.log
timeout 600 python tools/ab_gather.py 2>&1 | grep -v Warn > gpurun_out/ab_gather.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log
cat gpurun_out/chain3_check.log gpurun_out/trace3_mlp.log gpurun_out/trace3_front.log gpurun_out/ab_gather.log gpurun_out/pytest_gpu.log; tail -c 6000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
