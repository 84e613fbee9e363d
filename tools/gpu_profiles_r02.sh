#!/bin/bash
# Round-2 profiles: ncu launch list of the bench command, ncu --set full of one block forward (fp32-grade C=128 and
# bf16 C=256).  Outputs under gpurun_out/; tools/ncu_traffic.py turns the reports into profiles/r02_*.
mkdir -p gpurun_out
K='regex:to_basis_kernel|pack_weights|rows_chain|spmm_features|spmm_gxy'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 14 -c 7 -f -o gpurun_out/r02_block \
  python tools/profile_block.py 3 > gpurun_out/r02_ncu_block.log 2>&1
DN_B200_ENGINE=bf16 DN_PROFILE_C=256 timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 16 -c 8 -f \
  -o gpurun_out/r02_block_c256_bf16 python tools/profile_block.py 3 > gpurun_out/r02_ncu_block_c256.log 2>&1
ls -la gpurun_out/*.ncu-rep
