#!/bin/bash
mkdir -p gpurun_out
echo "== 1. GPU tests"
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -30 gpurun_out/r02c_pytest_gpu.log
echo "== 2. velocity builds under A/B"
timeout 600 python tools/ab_paths.py 7,8 0,1,2,3,4 > gpurun_out/r02c_ab_paths.log 2>&1; cat gpurun_out/r02c_ab_paths.log
echo "== 3. stream sums on moved particles"
timeout 600 python tools/stream_after_steps.py 7 3 > gpurun_out/r02c_stream_after_steps.log 2>&1; cat gpurun_out/r02c_stream_after_steps.log
echo "== 4. bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; tail -c 1500 gpurun_out/r02c_bench_n1.json; tail -5 gpurun_out/r02c_bench_n1.err
