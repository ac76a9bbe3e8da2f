"""Times every direct-sum kernel of the path at a large size (development / profiles tool)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems
Ls = int(sys.argv[1]) if len(sys.argv) > 1 else 7     # sphere level
Lp = int(sys.argv[2]) if len(sys.argv) > 2 else 8     # plane level
api.init(1)
api.set_profiling(True)
tf, _ = api.fp64_peak_probe(20000)
res = {"fp64_probe_tflops": tf, "sphere_level": Ls, "plane_level": Lp, "kernels": {}}
FLOP = {"bve_velocity": 22, "bve_stream": 11, "plane_velocity": 10, "plane_stream": 9, "betaplane_velocity": 14,
        "betaplane_stream": 13, "pse_laplacian_sphere": 35, "pse_laplacian_plane": 17}

def run(name, fn, pairs, reps=3, key=None):
    best, whole = 1e30, 1e30
    for _ in range(reps):
        fn()
        best = min(best, api.last_kernel_ms())
        whole = min(whole, api.last_sum_ms())
    res["kernels"][key or name] = {"ms": best, "ms_with_pack_and_sort": whole, "interactions": pairs, "interactions_per_s": pairs / (best * 1e-3),
                                   "algorithmic_flop_per_interaction": FLOP[name],
                                   "algorithmic_tflops": FLOP[name] * pairs / (best * 1e-3) / 1e12}
    print(f"{key or name:32s} {best:10.3f} ms (whole sum {whole:9.3f})  {pairs / best / 1e6:9.1f} G/s  {FLOP[name] * pairs / best / 1e9:7.2f} algTF", flush=True)

m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, Ls)
z = problems.rossby_haurwitz54(m)
av = problems.abs_vorticity(m, z, 2 * np.pi)
f = problems.spherical_harmonic54(m)
F = m.n_active
print("sphere", m.n, F, flush=True)
run("bve_velocity", lambda: api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0), m.n * F - F)
run("bve_stream", lambda: api.bve_stream(m.x, m.y, m.z, z, av, m.area, m.is_active, 1.0), m.n * F - F)
for pw in (0.75, 0.25):
    eps = m.max_edge_length ** pw
    run("pse_laplacian_sphere", lambda: api.pse_laplacian_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps, 1.0),
        m.n * F, key=f"pse_laplacian_sphere_eps=h^{pw}")
q = mesh.PolyMesh2d(mesh.QUAD_RECT_SEED, Lp, 7.0)
vort = problems.colliding_dipoles(q)
F = q.n_active
print("plane", q.n, F, flush=True)
run("plane_velocity", lambda: api.plane_velocity(q.x, q.y, vort, q.area, q.is_active), q.n * F - F)
run("plane_stream", lambda: api.plane_stream(q.x, q.y, vort, q.area, q.is_active), q.n * F - F)
g = np.sin(q.x) * np.cos(q.y)
run("pse_laplacian_plane", lambda: api.pse_laplacian_plane(q.x, q.y, g, q.area, q.is_active, q.max_edge_length ** 0.75), q.n * F)
b = mesh.PolyMesh2d(mesh.BETA_PLANE_SEED, Lp)
zb = problems.betaplane_gaussian(b)
ab = zb + 1.0 + 2.0 * b.y
F = b.n_active
run("betaplane_velocity", lambda: api.betaplane_velocity(b.x, b.y, zb, b.area, b.is_active), b.n * F - F)
run("betaplane_stream", lambda: api.betaplane_stream(b.x, b.y, zb, ab, b.area, b.is_active), b.n * F - F)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_kernels.json", "w"), indent=1)
