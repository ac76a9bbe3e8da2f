"""A/B on one box: time the BVE velocity kernel (and optionally the stream kernel) of the package under
ROOT (argv[1], default the repo) at a sphere level (argv[2], default 7), kernel variant argv[3].  Used to separate
box-to-box variation from code changes:  python tools/ab_bve.py build/ab_head 7; python tools/ab_bve.py . 7"""
import os, sys
root = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, root)
import numpy as np
from lpm_v2_b200 import api, mesh, problems
L = int(sys.argv[2]) if len(sys.argv) > 2 else 7
api.init(1)
api.set_profiling(True)
m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
z = problems.rossby_haurwitz54(m)
av = problems.abs_vorticity(m, z, 2 * np.pi)
if len(sys.argv) > 3:
    api.set_bve_variant(int(sys.argv[3]))
for name, fn in (("bve_velocity", lambda: api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)),
                 ("bve_stream", lambda: api.bve_stream(m.x, m.y, m.z, z, av, m.area, m.is_active, 1.0))):
    ts = []
    for _ in range(4):
        fn()
        ts.append(api.last_kernel_ms())
    print(f"{root} L{L} variant {sys.argv[3] if len(sys.argv) > 3 else 0} {name}: " + " ".join(f"{t:.3f}" for t in ts) + " ms", flush=True)
