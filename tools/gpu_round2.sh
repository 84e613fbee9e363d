#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/status.txt
timeout 300 python tools/tc_debug.py 0 > gpurun_out/tc_debug.log 2>&1; echo "tc_debug exit $?" >> gpurun_out/status.txt
timeout 600 python tools/accuracy_report.py > gpurun_out/accuracy.log 2>&1; echo "accuracy exit $?" >> gpurun_out/status.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --engine tc3x --steps 20 --warmup 3 > gpurun_out/bench_tc3x.log 2>&1; echo "bench exit $?" >> gpurun_out/status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python tools/profile_block.py 4 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list exit $?" >> gpurun_out/status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:rows_chain|to_basis_kernel|spmm_features' -s 8 -c 4 -f -o gpurun_out/prof_full python tools/profile_block.py 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt; grep -v Warn gpurun_out/accuracy.log | tail -8; tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/bench_tc3x.log | cut -c1-400
