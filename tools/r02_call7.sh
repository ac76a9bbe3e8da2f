#!/bin/bash
# Round 2 validation of the final default: GPU tests, bench, ncu launch list, ncu --set full of the two velocity kernels at icosTri 8.
mkdir -p gpurun_out
echo "== 1. GPU tests"
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -14 gpurun_out/r02f_pytest_gpu.log
echo "== 2. bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; tail -c 600 gpurun_out/r02f_bench_n1.json; tail -3 gpurun_out/r02f_bench_n1.err
echo "== 3. ncu launch list of a bench run"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity > gpurun_out/r02f_ncu_bench.log 2>&1; tail -1 gpurun_out/r02f_ncu_bench.log | cut -c1-200
echo "== 4. ncu --set full at icosTri 8: triangle kernel, passive one-sided kernel, stream triangle"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sym_kernel -c 1 -f -o gpurun_out/r02f_sym_vel_L8 python tools/profile_kernel.py bve_velocity 8 > gpurun_out/r02f_ncu_a.log 2>&1; tail -1 gpurun_out/r02f_ncu_a.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ds_kernel -c 1 -f -o gpurun_out/r02f_ds_vel_L8 python tools/profile_kernel.py bve_velocity 8 > gpurun_out/r02f_ncu_b.log 2>&1; tail -1 gpurun_out/r02f_ncu_b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sym_kernel -c 1 -f -o gpurun_out/r02f_sym_stream_L7 python tools/profile_kernel.py bve_stream 7 > gpurun_out/r02f_ncu_c.log 2>&1; tail -1 gpurun_out/r02f_ncu_c.log
