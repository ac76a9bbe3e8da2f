#!/bin/bash
mkdir -p gpurun_out
echo "== 1. RK4 tests with the fused step end"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sym_gpu.py tests/test_configs_gpu.py tests/test_c_caller.py -m gpu -q -k "rk4 or step or config1 or c_caller or solver" > gpurun_out/r02j_pytest_step.log 2>&1; tail -5 gpurun_out/r02j_pytest_step.log
echo "== 2. step time, fused vs separate"
timeout 600 python tools/ab_step.py 8 > gpurun_out/r02j_ab_step.log 2>&1; cat gpurun_out/r02j_ab_step.log
timeout 300 python tools/ab_step.py 7 >> gpurun_out/r02j_ab_step.log 2>&1; tail -4 gpurun_out/r02j_ab_step.log
