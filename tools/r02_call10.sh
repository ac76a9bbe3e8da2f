#!/bin/bash
mkdir -p gpurun_out
echo "== 1. one-sided kernels after the careful-retry change (same box A/B against the symmetric path)"
timeout 600 python tools/ab_paths.py 7,8 256 sym_panel_blocks > gpurun_out/r02i_ab_paths.log 2>&1; cat gpurun_out/r02i_ab_paths.log
timeout 600 python tools/bench_kernels.py 7 8 > gpurun_out/r02i_bench_kernels.log 2>&1; tail -12 gpurun_out/r02i_bench_kernels.log
echo "== 2. GPU tests"
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02i_pytest_gpu.log 2>&1; tail -10 gpurun_out/r02i_pytest_gpu.log
echo "== 3. bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; tail -c 300 gpurun_out/r02i_bench_n1.json; tail -3 gpurun_out/r02i_bench_n1.err
