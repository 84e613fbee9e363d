#!/bin/bash
# Round 2, GPU call 1: microbenchmarks that size the chain-kernel redesign, first run of rows_chain3_kernel,
# and the gather variants round 1 left unmeasured.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 120 tools/ubench/ubench.bin > gpurun_out/ubench.log 2>&1; echo "ubench exit $?" >> gpurun_out/ubench.log
timeout 300 python tools/r2_probe.py 2>&1 | grep -v Warn > gpurun_out/r2_probe.log; echo "probe exit $?" >> gpurun_out/r2_probe.log
timeout 400 python tools/r2_chain3_check.py 2>&1 | grep -v Warn > gpurun_out/chain3_check.log; echo "chain3 exit $?" >> gpurun_out/chain3_check.log
DN_TC_CHAIN3=0 timeout 400 python tools/r2_chain3_check.py 2>&1 | grep -v Warn > gpurun_out/chain3_off.log; echo "chain3-off exit $?" >> gpurun_out/chain3_off.log
{
  echo "== patch gather v2 (validated)";            timeout 200 python tools/ab_patch.py 2>&1 | grep -v Warn | tail -8
  echo "== patch gather v3 (async, double buffer)"; DN_SPMM_PATCH_V=3 timeout 200 python tools/ab_patch.py 2>&1 | grep -v Warn | tail -8
  echo "== patched parity tests with v3";           DN_SPMM_PATCH_V=3 timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider -k patched 2>&1 | tail -2
} > gpurun_out/patch_ab.log 2>&1
cat gpurun_out/ubench.log; tail -8 gpurun_out/r2_probe.log; cat gpurun_out/chain3_check.log gpurun_out/chain3_off.log gpurun_out/patch_ab.log
