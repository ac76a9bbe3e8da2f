#!/bin/bash
# 8 GPUs, final code of round 2: multi-GPU tests and the headline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/r02c_pytest_multigpu_n8.log 2>&1; tail -3 gpurun_out/r02c_pytest_multigpu_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29808 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r02c_n8_L8.err | grep '^{' > gpurun_out/r02c_n8_L8.json
python - <<PY
import json
d = json.load(open("gpurun_out/r02c_n8_L8.json")); r = d["roofline"]
print("L=8 N=8 value %.4g ms/step %.3f e2e %.4g step_frac %.3f rk4 %s parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["step_frac"], r.get("rk4_step_ms"), d["parity_sample"]["max_rel_err"]))
print("   kernels", {k: (round(v["ms_per_launch"], 3), round(v["fp64_pipe_frac"], 3)) for k, v in r["kernels"].items()}, d["clocks"])
PY
