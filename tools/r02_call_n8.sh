#!/bin/bash
# 8 GPUs: multi-GPU tests, the headline at 8 ranks, config 5's upper end (icosTri 9 and 10) with an oracle sample on
# every rank, and the small-slice case (icosTri 6).  Logs under gpurun_out/, copied to profiles/ afterwards.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() {  # level steps warmup extra...
  L=$1; shift; K=$1; shift; W=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600+L)) \
      bench.py --gpus 8 --level $L --steps $K --warmup $W --no-cpu "$@" 2> gpurun_out/r02_n8_L$L.err | grep '^{' > gpurun_out/r02_n8_L$L.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_n8_L$L.json"))
    print("L=$L N=8 value %.4g ms/step %.3f e2e %.4g step_frac %.3f rk4 %s parity %s clocks %s" % (d["value"], d["ms_per_step"],
          d["e2e"]["value"], d["roofline"]["step_frac"], d["roofline"].get("rk4_step_ms"), d.get("parity_sample") and
          (d["parity_sample"]["max_rel_err"], d["parity_sample"]["targets"]), d["clocks"]))
except Exception as e:
    print("L=$L N=8 failed", e); print(open("gpurun_out/r02_n8_L$L.err").read()[-1500:])
PY
}
echo "== 1. multi-GPU tests on 8 GPUs"
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/r02_pytest_multigpu_n8.log 2>&1; tail -8 gpurun_out/r02_pytest_multigpu_n8.log
echo "== 2. headline, 8 ranks"
run 8 5 3
echo "== 3. icosTri 6 (small slices) and 7"
run 6 20 5 --no-rk4
run 7 10 3
echo "== 4. icosTri 9"
run 9 2 1 --parity-targets 512 --e2e-reps 1
echo "== 5. icosTri 10 (31.5 M targets x 21 M sources)"
run 10 1 1 --parity-targets 64 --e2e-reps 1 --no-rk4
