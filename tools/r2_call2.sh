#!/bin/bash
# Round 2, GPU call 2: chain3 after the barrier fix (parity, timing, traces, per-stage profile), TMA scattered-row rates.
mkdir -p gpurun_out
timeout 400 python tools/r2_chain3_check.py 2>&1 | grep -v Warn > gpurun_out/chain3_check.log; echo "chain3 exit $?" >> gpurun_out/chain3_check.log
timeout 200 python tools/trace_chain3.py mlp 2>&1 | grep -v Warn > gpurun_out/trace3_mlp.log
timeout 200 python tools/trace_chain3.py front 2>&1 | grep -v Warn > gpurun_out/trace3_front.log
timeout 100 tools/ubench/ubench2.bin bulk > gpurun_out/ubench2.log 2>&1
for cfg in "256 1" "256 4" "128 1" "128 4"; do timeout 60 tools/ubench/ubench2.bin g4 $cfg >> gpurun_out/ubench2.log 2>&1; done
timeout 600 python tools/ab_gather.py 2>&1 | grep -v Warn > gpurun_out/ab_gather.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/chain3_check.log gpurun_out/trace3_mlp.log gpurun_out/trace3_front.log gpurun_out/ubench2.log gpurun_out/ab_gather.log gpurun_out/pytest_gpu.log
