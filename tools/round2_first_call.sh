#!/bin/bash
# The first gpurun call of the next round, in one script (everything below was prepared without GPU time):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Before calling: python __graft_entry__.py (build), and optionally `bash tools/build_renamed.sh` for step 4.
# Every step writes its log under gpurun_out/; a failing step does not stop the rest.
mkdir -p gpurun_out
echo "== 0. toolchain probe"; which gfortran mpirun mpifort flang nvfortran f951 2>&1 | head; nproc; nvidia-smi -L
echo "== 1. GPU parity tests of the default paths"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
echo "== 2. experimental pair-symmetric paths: parity (tests/test_sym_gpu.py)"
# (the n = 20011 cases spend their time in the long-double oracle on the host: left for a later call)
LPM_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_sym_gpu.py -m gpu -q -k "not 20011" > gpurun_out/r02_pytest_sym.log 2>&1; tail -5 gpurun_out/r02_pytest_sym.log
echo "== 3. default vs fenced one-sided (44: velocity, 103: stream) vs symmetric (200-203), icosTri 7 and 8, one box"
timeout 600 python tools/ab_sym.py 7,8 0,44,45,103,200,201,202,203,204,205,206,207,208,209 > gpurun_out/r02_ab_sym.log 2>&1; cat gpurun_out/r02_ab_sym.log
if [ -f build/renamed/lpm_v2_b200/liblpmgpu.so ]; then
  echo "== 4. register-renamed BVE kernel (tools/build_renamed.sh) against the product build"
  { timeout 300 python tools/ab_bve.py . 7; timeout 300 python tools/ab_bve.py build/renamed 7; timeout 300 python tools/ab_bve.py . 7; } > gpurun_out/r02_ab_renamed.log 2>&1
  cat gpurun_out/r02_ab_renamed.log
fi
echo "== 4b. ncu of the symmetric velocity kernel at icosTri 6 (stall reasons, FP64 pipe, L2 atomics); read with ncu -i ... --page raw --csv"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sym_kernel -c 1 -f -o gpurun_out/r02_sym_vel_L6 \
    python tools/profile_bve.py 6 1 201 > gpurun_out/r02_ncu_sym.log 2>&1; tail -2 gpurun_out/r02_ncu_sym.log
echo "== 5. bench line (default) and with the symmetric velocity sum"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 600 gpurun_out/r02_bench_default.json
timeout 600 python bench.py --steps 5 --warmup 3 --variant 200 --no-cpu > gpurun_out/r02_bench_sym.json 2> gpurun_out/r02_bench_sym.err; tail -c 600 gpurun_out/r02_bench_sym.json
