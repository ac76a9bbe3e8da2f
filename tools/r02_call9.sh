#!/bin/bash
mkdir -p gpurun_out
echo "== 1. PSE / SWE tests with the custom exp"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_pse_ops_gpu.py tests/test_swe_gpu.py -m gpu -q > gpurun_out/r02h_pytest_pse.log 2>&1; tail -4 gpurun_out/r02h_pytest_pse.log
echo "== 2. every kernel at a large size"
timeout 900 python tools/bench_kernels.py 7 8 > gpurun_out/r02h_bench_kernels.log 2>&1; cat gpurun_out/r02h_bench_kernels.log | tail -20
echo "== 3. DRAM bytes of the triangle kernel for larger panels"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
for P in 512 1024; do
  LPM_TUNE=sym_panel_blocks=$P timeout 300 ncu --metrics $M --clock-control none -k regex:sym_kernel -c 1 --csv --log-file gpurun_out/r02h_dram_panel$P.csv python tools/profile_kernel.py bve_velocity 8 > /dev/null 2>&1
  grep -E "dram__|lts__|gpu__time" gpurun_out/r02h_dram_panel$P.csv | awk -F'","' '{print "panel '$P':", $(NF-2), $(NF-1), $NF}'
done
