#!/bin/bash
# One GPU-box visit: smoke, parity tests, bench, ncu launch list (+ optional full captures).
# Everything lands in gpurun_out/.  Usage: tools/gpu_round.sh [tests|bench|all]
mkdir -p gpurun_out
MODE=${1:-all}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
if [ "$MODE" != "bench" ]; then
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
fi
if [ "$MODE" = "tests" ]; then exit 0; fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
tail -c 1500 gpurun_out/bench_ref.log
timeout 300 python tools/kernel_times.py 32 gpurun_out/kernel_times_n32.md > /dev/null 2>&1
# launch list of the SAME bench command: the window skips the 3 warm-up steps (~1220 kernels each) and covers the timed one
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3660 -c 1300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --rows 32 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_wx|gemm_dw|tcn_dw" -c 9 \
   -o gpurun_out/prof_block python tools/profile_block.py 32 8 1 > gpurun_out/prof_block.log 2>&1
ls -la gpurun_out
