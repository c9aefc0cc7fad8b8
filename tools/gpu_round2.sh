#!/bin/bash
# Round-2 evidence run on one B200: bench (both arms), CUPTI per-kernel tables, ncu launch list of the bench command, ncu --set full
# of the recurrence kernels and of the TCN-block GEMMs.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.log 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.log 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/kernel_times.py 32 gpurun_out/r02_kernel_times_spex_n32.md > /dev/null 2>&1
timeout 300 python tools/kernel_times_bsrnn.py 16 4 gpurun_out/r02_kernel_times_pbsrnn_n16.md > /dev/null 2>&1
# launch list of the bench command (Spex+ part: 3 warm-up steps + 1 timed step; the window skips the warm-up)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 1400 --csv --log-file gpurun_out/r02_launches_spex.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pbsrnn > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec -s 2 -c 2 -o gpurun_out/r02_prof_lstm_rec \
   python tools/run_lstm_rec_once.py 501 512 256 128 > gpurun_out/prof_lstm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_wx_tc2|gemm_dw_tc2|tcn_dw" -c 10 -o gpurun_out/r02_prof_block \
   python tools/profile_block.py 32 8 1 > gpurun_out/prof_block.log 2>&1
ls -la gpurun_out | tail -20
