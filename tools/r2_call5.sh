#!/bin/bash
mkdir -p gpurun_out
for h in 1 0; do echo "== DN_TC_HYBRID=$h"; DN_TC_HYBRID=$h timeout 300 python tools/r2_bwd_debug2.py 2>&1 | grep -v Warn; done > gpurun_out/bwd_debug2.log 2>&1
for h in 1 0; do echo "== DN_TC_HYBRID=$h"; DN_TC_HYBRID=$h timeout 400 python tools/r2_chain3_check.py 2>&1 | grep -v Warn; done > gpurun_out/chain3_hybrid.log 2>&1
cat gpurun_out/bwd_debug2.log gpurun_out/chain3_hybrid.log
