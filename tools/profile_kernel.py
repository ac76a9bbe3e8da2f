"""Runs one evaluation of a named kernel through the C ABI (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems
name = sys.argv[1]
L = int(sys.argv[2]) if len(sys.argv) > 2 else 6
api.init(1)
api.set_profiling(True)
for kv in filter(None, os.environ.get("LPM_TUNE", "").split(",")):      # e.g. LPM_TUNE=sym_panel_blocks=4096
    k, v = kv.split("=")
    api.tune(k, int(v))
if name in ("bve_stream", "bve_velocity", "pse_sphere"):
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
    z = problems.rossby_haurwitz54(m)
    av = problems.abs_vorticity(m, z, 2 * np.pi)
    if name == "bve_stream":
        api.bve_stream(m.x, m.y, m.z, z, av, m.area, m.is_active, 1.0)
    elif name == "bve_velocity":
        api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)
    else:
        api.pse_laplacian_sphere(m.x, m.y, m.z, z, m.area, m.is_active, m.max_edge_length ** 0.25, 1.0)
elif name in ("plane_velocity", "plane_stream", "swe"):
    q = mesh.PolyMesh2d(mesh.QUAD_RECT_SEED, L, 7.0)
    vort = problems.colliding_dipoles(q)
    if name == "plane_velocity":
        api.plane_velocity(q.x, q.y, vort, q.area, q.is_active)
    elif name == "plane_stream":
        api.plane_stream(q.x, q.y, vort, q.area, q.is_active)
    else:
        api.swe_plane_rhs_integrals(q.x, q.y, vort, 0.1 * vort, 1 + 0 * vort, q.area, q.is_active, q.max_edge_length ** 0.75)
else:
    b = mesh.PolyMesh2d(mesh.BETA_PLANE_SEED, L)
    zb = problems.betaplane_gaussian(b)
    if name == "betaplane_velocity":
        api.betaplane_velocity(b.x, b.y, zb, b.area, b.is_active)
    else:
        api.betaplane_stream(b.x, b.y, zb, zb + 1, b.area, b.is_active)
print(name, "kernel ms", api.last_kernel_ms())
