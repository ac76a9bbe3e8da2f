#!/bin/bash
# Round 2, call 3: the promoted pair-symmetric default (integer fixed-point accumulation).
mkdir -p gpurun_out
echo "== 1. GPU tests"
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02b_pytest_gpu.log 2>&1; tail -25 gpurun_out/r02b_pytest_gpu.log
echo "== 2. A/B of the paths and the builds under test"
timeout 600 python tools/ab_paths.py 7,8 0,1,2,3 0,1 > gpurun_out/r02b_ab_paths.log 2>&1; cat gpurun_out/r02b_ab_paths.log
echo "== 3. bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; tail -c 3000 gpurun_out/r02b_bench_n1.json; tail -5 gpurun_out/r02b_bench_n1.err
timeout 300 python bench.py --steps 3 --warmup 3 --one-sided --no-cpu --no-parity > gpurun_out/r02b_bench_one_sided.json 2> gpurun_out/r02b_bench_one_sided.err; tail -c 1500 gpurun_out/r02b_bench_one_sided.json
echo "== 4. ncu launch list of a bench run"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity > gpurun_out/r02b_ncu_bench.log 2>&1; tail -2 gpurun_out/r02b_ncu_bench.log
echo "== 5. ncu --set full of the symmetric velocity kernel (icosTri 7) and the symmetric stream kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sym_kernel -c 1 -f -o gpurun_out/r02b_sym_vel_L7 python tools/profile_kernel.py bve_velocity 7 > gpurun_out/r02b_ncu_sym_vel.log 2>&1; tail -2 gpurun_out/r02b_ncu_sym_vel.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sym_kernel -c 1 -f -o gpurun_out/r02b_sym_stream_L7 python tools/profile_kernel.py bve_stream 7 > gpurun_out/r02b_ncu_sym_stream.log 2>&1; tail -2 gpurun_out/r02b_ncu_sym_stream.log
