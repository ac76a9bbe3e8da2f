#!/bin/bash
# Config 5: synthetic BVE direct-sum sweep over icosTri levels on N GPUs (development tool).
N=${1:-8}; shift
mkdir -p gpurun_out
for L in "$@"; do
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --level $L --steps 3 --warmup 3 --no-rk4 --no-cpu 2> gpurun_out/sweep_L${L}_n$N.err | grep '^{' > gpurun_out/sweep_L${L}_n$N.json
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+L)) bench.py --gpus $N --level $L --steps 3 --warmup 3 --no-rk4 --no-cpu 2> gpurun_out/sweep_L${L}_n$N.err | grep '^{' > gpurun_out/sweep_L${L}_n$N.json
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sweep_L${L}_n$N.json"))
    print("L=$L N=$N", "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"])
except Exception as e:
    print("L=$L N=$N failed", e)
PY
done
