#!/bin/bash
# Round 2, GPU call 4: chain3 (bias lanes, no probe), tensor-core backward, aux benches, ncu --set full of the block.
mkdir -p gpurun_out
timeout 400 python tools/r2_chain3_check.py 2>&1 | grep -v Warn > gpurun_out/chain3_check.log; echo "chain3 exit $?" >> gpurun_out/chain3_check.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload fwd_bwd --steps 10 --warmup 3 > gpurun_out/bench_fwd_bwd.log 2> gpurun_out/bench_fwd_bwd.err; echo "exit $?" >> gpurun_out/bench_fwd_bwd.log
DN_B200_ENGINE=simt timeout 300 python bench.py --engine simt --workload fwd_bwd --steps 5 --warmup 3 > gpurun_out/bench_fwd_bwd_simt.log 2>&1
timeout 300 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train1.log 2> gpurun_out/bench_train1.err; echo "exit $?" >> gpurun_out/bench_train1.log
timeout 300 python bench.py --workload small_batch --steps 10 --warmup 3 > gpurun_out/bench_small.log 2> gpurun_out/bench_small.err; echo "exit $?" >> gpurun_out/bench_small.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:rows_chain3|spmm_features|to_basis_kernel|pack_weights' -s 5 -c 5 -f -o gpurun_out/r02_block python tools/profile_block.py 3 > gpurun_out/ncu_block.log 2>&1; echo "ncu exit $?" >> gpurun_out/ncu_block.log
DN_SPMM_PIPE=2 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:spmm_features' -s 1 -c 1 -f -o gpurun_out/r02_gather_pipe2 python tools/profile_block.py 2 > gpurun_out/ncu_gather.log 2>&1
cat gpurun_out/chain3_check.log gpurun_out/pytest_gpu.log; for f in fwd_bwd fwd_bwd_simt train1 small; do echo "== $f"; tail -c 2500 gpurun_out/bench_$f.log; tail -3 gpurun_out/bench_$f.err 2>/dev/null; done; tail -3 gpurun_out/ncu_block.log
