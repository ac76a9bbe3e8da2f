#!/bin/bash
mkdir -p gpurun_out
echo "== 1. chunk length of the triangle kernel"
timeout 600 python tools/ab_paths.py 8 16,8,32,80 sym_chunk_tiles > gpurun_out/r02d_ab_sym_chunks.log 2>&1; cat gpurun_out/r02d_ab_sym_chunks.log
echo "== 2. small slices of the one-sided engine (a rank's grid of an 8-GPU run, on one GPU)"
timeout 600 python tools/slice_sweep.py 6 8 > gpurun_out/r02d_slice_sweep_L6.log 2>&1; cat gpurun_out/r02d_slice_sweep_L6.log
timeout 600 python tools/slice_sweep.py 7 8 > gpurun_out/r02d_slice_sweep_L7.log 2>&1; cat gpurun_out/r02d_slice_sweep_L7.log
echo "== 3. bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; tail -c 1200 gpurun_out/r02d_bench_n1.json; tail -5 gpurun_out/r02d_bench_n1.err
