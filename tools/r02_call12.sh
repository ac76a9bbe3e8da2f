#!/bin/bash
# Round 2, final validation on one GPU: all GPU tests, smoke(), the bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02k_pytest_gpu.log 2>&1; tail -10 gpurun_out/r02k_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; tail -c 300 gpurun_out/r02k_bench_n1.json; tail -3 gpurun_out/r02k_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/r02k_bench_reference.json 2> gpurun_out/r02k_bench_reference.err; head -c 400 gpurun_out/r02k_bench_reference.json
