"""BVE kernel variant sweep (development tool, run under gpurun)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems
levels = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [7]
variants = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(1, 16))
api.init(1)
api.set_profiling(True)
tf, _ = api.fp64_peak_probe(20000)
print("probe TF", tf, flush=True)
res = {"probe_tflops": tf}
for L in levels:
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
    z = problems.rossby_haurwitz54(m)
    pairs = m.n * m.n_active - m.n_active
    for var in variants:
        api.set_bve_variant(var)
        best = 1e30
        for rep in range(3 if L < 8 else 2):
            api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)
            best = min(best, api.last_kernel_ms())
        res[f"L{L}_v{var}"] = best
        print("L", L, "variant", var, "ms", round(best, 3), "Gpairs/s", round(pairs / best / 1e6, 1),
              "algTF", round(22 * pairs / best / 1e9, 2), "frac", round(22 * pairs / best / 1e9 / tf, 4), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/sweep_bve.json", "w"), indent=1)
