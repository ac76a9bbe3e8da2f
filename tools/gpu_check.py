"""Development check run under gpurun: parity of every kernel vs the oracle at
small sizes, then BVE kernel variant timings.  Writes gpurun_out/gpu_check.json."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lpm_v2_b200 import api, mesh, problems, solvers
from oracle import binding as O

out = {}
def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

nd = api.init(1)
print("devices", nd, flush=True)
tf, ms = api.fp64_peak_probe(20000)
print("dfma probe TF", tf, "ms", ms, flush=True)
out["fp64_probe_tflops"] = tf

for L in (2, 3, 4):
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
    z = problems.gaussian_vortex(m)
    u, v, w = api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)
    ou, ov, ow = O.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)
    print("bve L", L, rel(u, ou), rel(v, ov), rel(w, ow), flush=True)
    out[f"bve_L{L}"] = max(rel(u, ou), rel(v, ov), rel(w, ow))
    al = api.active_list(m.is_active); ol = O.active_list(m.is_active)
    assert np.array_equal(al, ol), "active list mismatch"
    av = problems.abs_vorticity(m, z, 2*np.pi)
    rs, as_ = api.bve_stream(m.x, m.y, m.z, z, av, m.area, m.is_active, 1.0)
    ors, oas = O.bve_stream(m.x, m.y, m.z, z, av, m.area, m.is_active, 1.0)
    print("bve stream L", L, rel(rs, ors), rel(as_, oas), flush=True)
    eps = m.max_edge_length ** 0.6
    f = problems.spherical_harmonic54(m)
    lap = api.pse_laplacian_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps)
    olap = O.pse_laplacian_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps)
    print("pse sphere L", L, rel(lap, olap), flush=True)
    eps2 = m.max_edge_length ** 1.5
    lap = api.pse_laplacian_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps2)
    olap = O.pse_laplacian_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps2)
    print("pse sphere (cutoff active) L", L, rel(lap, olap), flush=True)

for L in (2, 3, 4):
    m = mesh.PolyMesh2d(mesh.QUAD_RECT_SEED, L, 7.0)
    z = problems.colliding_dipoles(m)
    u, v = api.plane_velocity(m.x, m.y, z, m.area, m.is_active)
    ou, ov = O.plane_velocity(m.x, m.y, z, m.area, m.is_active)
    psi = api.plane_stream(m.x, m.y, z, m.area, m.is_active)
    opsi = O.plane_stream(m.x, m.y, z, m.area, m.is_active)
    print("plane L", L, rel(u, ou), rel(v, ov), "stream", rel(psi, opsi), flush=True)
    f = np.sin(m.x) * np.cos(0.5 * m.y)
    eps = m.max_edge_length ** 0.75
    lap = api.pse_laplacian_plane(m.x, m.y, f, m.area, m.is_active, eps)
    olap = O.pse_laplacian_plane(m.x, m.y, f, m.area, m.is_active, eps)
    print("pse plane L", L, rel(lap, olap), flush=True)
    m = mesh.PolyMesh2d(mesh.BETA_PLANE_SEED, L)
    z = problems.betaplane_gaussian(m)
    u, v = api.betaplane_velocity(m.x, m.y, z, m.area, m.is_active)
    ou, ov = O.betaplane_velocity(m.x, m.y, z, m.area, m.is_active)
    lu, lv = O.betaplane_velocity(m.x, m.y, z, m.area, m.is_active, variant="_ld")
    print("beta L", L, "gpu-vs-f64", rel(u, ou), rel(v, ov), "gpu-vs-ld", rel(u, lu), rel(v, lv), "f64-vs-ld", rel(ou, lu), rel(ov, lv), flush=True)
    av = z + 1.0 + 2.0 * m.y
    rs, as_ = api.betaplane_stream(m.x, m.y, z, av, m.area, m.is_active)
    ors, oas = O.betaplane_stream(m.x, m.y, z, av, m.area, m.is_active, variant="_ld")
    print("beta stream L", L, rel(rs, ors), rel(as_, oas), flush=True)

# RK4 step parity (BVE, L3)
m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, 3)
z = problems.gaussian_vortex(m)
sph = solvers.BVEMesh(m, z, 1.0, 2*np.pi)
sph.SetVelocityOnMesh()
u0, v0, w0 = [a.copy() for a in sph.velocity]
sol = solvers.BVESolver(sph)
sol.Timestep(sph, 0.01)
ref = O.bve_rk4_step(m.x, m.y, m.z, z, u0, v0, w0, m.area, m.is_active, 1.0, 2*np.pi, 0.01)
print("rk4 bve", [rel(a, b) for a, b in zip([sph.x, sph.y, sph.z, sph.relVort] + sph.velocity, ref)], flush=True)
print("diag", sol.Diagnostics(), O.total_ke(*sph.velocity, m.area, m.is_active), O.total_enstrophy(sph.relVort, m.area, m.is_active))
sol.Delete()

# timing sweep of BVE variants
api.set_profiling(True)
res = {}
for L in (6, 7):
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
    z = problems.gaussian_vortex(m)
    pairs = m.n * m.n_active - m.n_active
    for var in (1, 2, 3, 4, 5, 7, 8, 9, 10):
        api.set_bve_variant(var)
        best = 1e30
        for rep in range(3):
            u, v, w = api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)
            best = min(best, api.last_kernel_ms())
        res[f"L{L}_v{var}"] = {"ms": best, "pairs_per_s": pairs / (best * 1e-3), "alg_tflops": 22 * pairs / (best * 1e-3) / 1e12}
        print("L", L, "variant", var, "ms", round(best, 3), "Gpairs/s", round(pairs / best / 1e6, 1), "algTF", round(22 * pairs / best / 1e9, 2), flush=True)
out["variants"] = res
api.set_bve_variant(0)
if len(sys.argv) > 1 and sys.argv[1] == "L8":
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, 8)
    z = problems.gaussian_vortex(m)
    pairs = m.n * m.n_active - m.n_active
    for rep in range(2):
        t = time.time()
        u, v, w = api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)
        wall = time.time() - t
        kms = api.last_kernel_ms()
        print("L8 kernel ms", kms, "wall s", wall, "Gpairs/s", pairs / kms / 1e6, "algTF", 22 * pairs / kms / 1e9, flush=True)
    out["L8_ms"] = kms
    # spot check vs oracle on a few targets
    idx = [0, 1, 12345, m.n // 2, m.n - 1]
    for i in idx:
        ou, ov, ow = O.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0, rng=(i, i + 1))
        print("L8 target", i, u[i], ou[i], abs(u[i] - ou[i]), abs(v[i] - ov[i]), abs(w[i] - ow[i]), flush=True)
    out["L8_umax"] = float(np.abs(u).max())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
print("done")
