#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/r2_chain3_check.py 2>&1 | grep -v Warn > gpurun_out/chain3_check.log
head -20 gpurun_out/chain3_check.log
