"""Why are the stream-function sums slower inside the resident solver than on the pristine mesh?  (round 2: 1187 / 744 ms
at icosTri 8 on the mesh, 1474 / 1034 ms after two RK4 steps.)  Times lpm_bve_stream on the mesh positions and on
the positions after k RK4 steps, through both engines, and counts the pairs whose argument leaves the table window.
    python tools/stream_after_steps.py [level] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems, solvers

L = int(sys.argv[1]) if len(sys.argv) > 1 else 7
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
api.init(1)
api.set_profiling(True)
api.tune("sym_min_sources", 0)
m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
z = problems.rossby_haurwitz54(m)
av = problems.abs_vorticity(m, z, 2 * np.pi)
sph = solvers.BVEMesh(m, z, 1.0, 2 * np.pi)
sph.SetVelocityOnMesh()
sol = solvers.BVESolver(sph)


def timed(tag, x, y, zc, q):
    for sym in (True, False):
        api.set_symmetric(sym)
        ts = []
        for _ in range(2):
            api.profile_breakdown(reset=True)
            api.bve_stream(x, y, zc, q, av, m.area, m.is_active, 1.0)
            ts.append(api.last_sum_ms())
        ks = api.profile_breakdown(reset=True)
        kern = ", ".join(f"{k.split('/')[1]} {v[1] / v[0]:.3f}" for k, v in ks.items())
        r2 = x * x + y * y + zc * zc
        print(f"L{L} {tag} {'symmetric' if sym else 'one-sided'}: sum ms {min(ts):.3f} (kernels: {kern});  |x|^2 - 1 in "
              f"[{(r2 - 1).min():.2e}, {(r2 - 1).max():.2e}]", flush=True)
    api.set_symmetric(True)


timed("mesh positions", m.x, m.y, m.z, z)
for k in range(1, steps + 1):
    sol.Timestep(sph, 0.01, with_stream=False)
    timed(f"after {k} RK4 step(s)", sph.x.copy(), sph.y.copy(), sph.z.copy(), sph.relVort.copy())
# the same positions projected back to the sphere: is it the radial drift?
r = np.sqrt(sph.x ** 2 + sph.y ** 2 + sph.z ** 2)
timed(f"after {steps} steps, renormalised", sph.x / r, sph.y / r, sph.z / r, sph.relVort.copy())
# the mesh positions with a random tangential jitter of 1e-4 (breaks the icosahedral symmetry only)
rng = np.random.default_rng(0)
j = np.stack([m.x, m.y, m.z], 1) + 1e-4 * rng.normal(size=(m.n, 3))
j /= np.linalg.norm(j, axis=1)[:, None]
timed("mesh positions + 1e-4 jitter", j[:, 0].copy(), j[:, 1].copy(), j[:, 2].copy(), z)
sol.Delete()
