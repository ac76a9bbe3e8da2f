#!/bin/bash
# Round 2, call 16: validation of the final kernels on one GPU -- all GPU tests, smoke(), the bench line, then (as far as
# the round's remaining GPU time allows) ncu --set full of the velocity triangle at icosTri 7 and the launch list of a bench run.
mkdir -p gpurun_out
timeout 290 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02o_pytest_gpu.log 2>&1; tail -9 gpurun_out/r02o_pytest_gpu.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 100 python bench.py > gpurun_out/r02o_bench_n1.json 2> gpurun_out/r02o_bench_n1.err; tail -c 1500 gpurun_out/r02o_bench_n1.json; tail -3 gpurun_out/r02o_bench_n1.err
echo "elapsed $SECONDS"
if [ $SECONDS -lt 340 ]; then
  timeout $((445 - SECONDS)) ncu --set full --clock-control none --import-source on -k regex:sym_kernel -c 1 -f -o gpurun_out/r02o_sym_vel_L7 python tools/profile_kernel.py bve_velocity 7 > gpurun_out/r02o_ncu_a.log 2>&1; tail -1 gpurun_out/r02o_ncu_a.log
fi
echo "elapsed $SECONDS"
if [ $SECONDS -lt 385 ]; then
  timeout $((450 - SECONDS)) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02o_launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity > gpurun_out/r02o_ncu_bench.log 2>&1; tail -1 gpurun_out/r02o_ncu_bench.log | cut -c1-200
fi
echo "elapsed $SECONDS"
