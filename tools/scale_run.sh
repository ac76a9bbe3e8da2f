#!/bin/bash
# Strong-scaling sweep of bench.py on one box (development tool; the driver runs its own).
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ "$n" = 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu 2> gpurun_out/scale_n$n.err | grep '^{' > gpurun_out/scale_n$n.json
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_n$n.json"))
    print("N=$n", "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "rk4_ms", d.get("rk4_step_ms"), "frac", d["roofline"]["frac"], "clk", d["clocks"])
except Exception as e:
    print("N=$n failed", e)
PY
done
