#!/bin/bash
# First GPU call of the next round: measure what round 1 left compiled-but-unmeasured or opt-in.
#   1. staged gather, copy/gather overlapped (DN_SPMM_PATCH_V=3) vs the validated patch kernel vs the plain gather
#   2. TMA-fed chain kernel with and without the TMA tail: time + parity over several calls (tools/diag_tail.py)
# Usage: gpurun --timeout 900 -- 'bash tools/next_round_checks.sh'
mkdir -p gpurun_out
{
  echo "== patch gather v2 (validated)";           timeout 200 python tools/ab_patch.py 2>&1 | grep -v Warn | tail -8
  echo "== patch gather v3 (async, double buffer)"; DN_SPMM_PATCH_V=3 timeout 200 python tools/ab_patch.py 2>&1 | grep -v Warn | tail -8
  echo "== patched parity tests with v3";           DN_SPMM_PATCH_V=3 timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider -k patched 2>&1 | tail -2
  echo "== TMA chain kernel A/B";                   timeout 200 python tools/ab_tma.py 2>&1 | tail -2
  echo "== TMA chain + tail: parity over 8 calls";  DN_TC_TMA=1 DN_TC_TMA_TAIL=1 timeout 200 python tools/diag_tail.py 2>&1 | grep "^call\|quarter\|tile index" | head -30
} 2>&1 | tee gpurun_out/next_round_checks.log
