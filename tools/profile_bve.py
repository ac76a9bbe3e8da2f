"""Runs a few BVE velocity evaluations through the C ABI (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems
L = int(sys.argv[1]) if len(sys.argv) > 1 else 7
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
var = int(sys.argv[3]) if len(sys.argv) > 3 else 0
api.init(1)
api.set_bve_variant(var)
m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
z = problems.rossby_haurwitz54(m)
api.set_profiling(True)
for r in range(reps):
    u, v, w = api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)
    print("rep", r, "kernel ms", api.last_kernel_ms(), flush=True)
