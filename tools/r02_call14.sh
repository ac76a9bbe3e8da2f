#!/bin/bash
# Round 2, call 14: A/B of the symmetric-kernel builds (combined warps beside the log table, lane-rotated source order).
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_sym_gpu.py -m gpu -q -k under_ab -n 4 > gpurun_out/r02m_pytest_builds.log 2>&1; tail -4 gpurun_out/r02m_pytest_builds.log
timeout 330 python tools/ab_builds.py 8 > gpurun_out/r02m_ab_builds.log 2>&1; cat gpurun_out/r02m_ab_builds.log
