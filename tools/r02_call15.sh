#!/bin/bash
# Round 2, call 15: second A/B of triangle-kernel builds (warp reduction overlapped with the last group's own sums, unrolled
# batch loop) and of 6 / 8 targets per thread in the one-sided log kernels.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_sym_gpu.py -m gpu -q -k under_ab -n 4 > gpurun_out/r02n_pytest_builds.log 2>&1; tail -4 gpurun_out/r02n_pytest_builds.log
timeout 240 python tools/ab_builds.py 8 > gpurun_out/r02n_ab_builds.log 2>&1; cat gpurun_out/r02n_ab_builds.log
