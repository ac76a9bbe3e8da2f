#!/bin/bash
# 8 GPUs, second pass: the headline and the small-slice case after the finer triangle chunks, the serpentine block
# dealing, the overlapped all-reduce and the new one-sided chunking.
mkdir -p gpurun_out
run() {
  L=$1; shift; K=$1; shift; W=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700+L)) \
      bench.py --gpus 8 --level $L --steps $K --warmup $W --no-cpu "$@" 2> gpurun_out/r02b_n8_L$L.err | grep '^{' > gpurun_out/r02b_n8_L$L.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02b_n8_L$L.json")); r = d["roofline"]
    print("L=$L N=8 value %.4g ms/step %.3f e2e %.4g step_frac %.3f rk4 %s parity %s" % (d["value"], d["ms_per_step"],
          d["e2e"]["value"], r["step_frac"], r.get("rk4_step_ms"), d.get("parity_sample") and
          (d["parity_sample"]["max_rel_err"], d["parity_sample"]["reference_vs_extended_precision"], d["parity_sample"]["gpu_vs_extended_precision"])))
    print("   kernels", {k: (round(v["ms_per_launch"], 3), round(v["fp64_pipe_frac"], 3)) for k, v in r["kernels"].items()})
except Exception as e:
    print("L=$L N=8 failed", e); print(open("gpurun_out/r02b_n8_L$L.err").read()[-1500:])
PY
}
run 8 5 3
run 6 20 5 --no-rk4
run 7 10 3 --no-rk4
