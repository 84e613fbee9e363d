#!/bin/bash
# Round 2, GPU call 1: microbenchmarks that size the chain-kernel redesign + what round 1 left unmeasured.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 120 tools/ubench/ubench.bin > gpurun_out/ubench.log 2>&1; echo "ubench exit $?" >> gpurun_out/ubench.log
timeout 300 python tools/r2_probe.py 2>&1 | grep -v Warn > gpurun_out/r2_probe.log; echo "probe exit $?" >> gpurun_out/r2_probe.log
timeout 900 bash tools/next_round_checks.sh > /dev/null 2>&1
DIAG_V=40000 DN_TC_TMA=1 DN_TC_TMA_TAIL=1 timeout 240 compute-sanitizer --tool racecheck python tools/diag_tail.py > gpurun_out/racecheck_tail.log 2>&1; echo "racecheck exit $?" >> gpurun_out/racecheck_tail.log
cat gpurun_out/ubench.log; tail -8 gpurun_out/r2_probe.log; cat gpurun_out/next_round_checks.log; tail -15 gpurun_out/racecheck_tail.log
