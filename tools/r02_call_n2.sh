#!/bin/bash
# 2 GPUs: the multi-GPU tests (rank mode, single-process mode, symmetric path bits across rank counts) and a bench line.
mkdir -p gpurun_out
nvidia-smi -L
echo "== 1. multi-GPU tests"
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/r02_pytest_multigpu_n2.log 2>&1; tail -15 gpurun_out/r02_pytest_multigpu_n2.log
echo "== 2. bench, 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu 2> gpurun_out/r02_bench_n2.err | grep '^{' > gpurun_out/r02_bench_n2.json; tail -c 1800 gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err
