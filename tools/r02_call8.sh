#!/bin/bash
mkdir -p gpurun_out
echo "== 1. launch order of the triangle kernel: panels of N target blocks (4096 = one panel, the previous order)"
timeout 600 python tools/ab_paths.py 8 256,128,512,4096 sym_panel_blocks > gpurun_out/r02g_ab_panels.log 2>&1; cat gpurun_out/r02g_ab_panels.log
echo "== 2. DRAM bytes of the triangle kernel at icosTri 8 (one metrics pass each)"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
for P in 256 4096; do
  LPM_TUNE=sym_panel_blocks=$P timeout 300 ncu --metrics $M --clock-control none -k regex:sym_kernel -c 1 --csv --log-file gpurun_out/r02g_dram_panel$P.csv python tools/profile_kernel.py bve_velocity 8 > /dev/null 2>&1
  grep -E "dram__|lts__|gpu__time" gpurun_out/r02g_dram_panel$P.csv | awk -F'","' '{print "panel '$P':", $(NF-2), $(NF-1), $NF}'
done
LPM_TUNE=sym_panel_blocks=256 timeout 300 ncu --metrics $M --clock-control none -k regex:sym_kernel -c 1 --csv --log-file gpurun_out/r02g_dram_stream_panel256.csv python tools/profile_kernel.py bve_stream 8 > /dev/null 2>&1
grep -E "dram__|lts__|gpu__time" gpurun_out/r02g_dram_stream_panel256.csv | awk -F'","' '{print "stream panel 256:", $(NF-2), $(NF-1), $NF}'
echo "== 3. new tests: SWE solver, symmetric path"
timeout 600 python -m pytest tests/test_swe_gpu.py tests/test_sym_gpu.py -m gpu -q > gpurun_out/r02g_pytest_new.log 2>&1; tail -4 gpurun_out/r02g_pytest_new.log
echo "== 4. compute-sanitizer on the small symmetric cases (memcheck, racecheck)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_sym_gpu.py -m gpu -q -k "513 or 1025 or 127 or coincident or (fixed_point and not 6371000)" > gpurun_out/r02g_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02g_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_sym_gpu.py -m gpu -q -k "velocity_random_ragged and (513 or 1025)" > gpurun_out/r02g_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02g_racecheck.log
echo "== 5. bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; tail -c 400 gpurun_out/r02g_bench_n1.json; tail -3 gpurun_out/r02g_bench_n1.err
