#!/bin/bash
# A/B of the chunk bound and kernel variant at N GPUs (development tool): bench.py without RK4 / CPU legs.
N=${1:-8}; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  v=${cfg%%:*}; c=${cfg##*:}
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+c+v)) bench.py --gpus $N --steps 5 --warmup 3 --no-rk4 --no-cpu --variant $v --max-chunks $c 2> gpurun_out/chunks_v${v}_c${c}_n$N.err | grep '^{' > gpurun_out/chunks_v${v}_c${c}_n$N.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/chunks_v${v}_c${c}_n$N.json"))
    print("N=$N variant=$v chunks=$c", "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "frac %.4f" % d["roofline"]["frac"])
except Exception as e:
    print("N=$N variant=$v chunks=$c failed", e)
PY
done
