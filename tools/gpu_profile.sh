#!/bin/bash
# Profiling visit: whole-step launch list + ncu --set full of the dominant kernels (1 GPU, never multi-rank).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --rows 32 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_wx_tc|gemm_dw_tc|tcn_dw" -c 9 \
   -o gpurun_out/prof_block python tools/profile_block.py 32 8 1 > gpurun_out/prof_block.log 2>&1
ls -la gpurun_out | head -20
