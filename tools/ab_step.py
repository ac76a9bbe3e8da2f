"""A/B on one box: the resident BVE RK4 step with its end fused (velocity + stream functions in one pass, the default)
or as two separate sums (lpm_tune("fuse_step_end", 0)), at icosTri level argv[1]; dt scaled with the mesh spacing as
in bench.py.  Prints the step time and the main kernels' times.
    python tools/ab_step.py [level]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems, solvers

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
api.init(1)
api.set_profiling(True)
m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
z = problems.rossby_haurwitz54(m)
dt = 0.01 * float(np.sqrt(6144.0 / m.n_active))
res = {}
for fused in (1, 0, 1):
    api.tune("fuse_step_end", fused)
    sph = solvers.BVEMesh(m, z, 1.0, 2 * np.pi)
    sph.SetVelocityOnMesh()
    sol = solvers.BVESolver(sph)
    sol.Timestep(sph, dt, with_stream=True, copy_back=False)
    api.profile_breakdown(reset=True)
    t0 = time.perf_counter()
    sol.Timestep(sph, dt, with_stream=True, copy_back=False)
    ms = (time.perf_counter() - t0) * 1e3
    ks = api.profile_breakdown(reset=True)
    sol.CopyToMesh(sph, True)
    res[fused] = sph.velocity + [sph.relStream, sph.absStream]
    sol.Delete()
    print(f"L{L} fuse_step_end={fused}: step {ms:.1f} ms; kernels " + ", ".join(f"{k} {c}x {t:.1f} ms" for k, (c, t) in ks.items()), flush=True)
scale = [max(np.abs(a).max(), 1e-300) for a in res[0]]
print("max difference fused vs separate, relative to each field's scale:", [f"{np.abs(a - b).max() / s:.1e}" for a, b, s in zip(res[1], res[0], scale)])
api.tune("fuse_step_end", 1)
