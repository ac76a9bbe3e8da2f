"""Workload for tests/cuda_emu/race_check.sh: one small call of every kernel family through the C ABI, on the
ThreadSanitizer build of the emulated library.  Shared-memory protocol errors in a kernel (a missing __syncthreads(), a
tile overwritten while still being read, an unsynchronised reduction scratch) are data races between the OS threads that
stand for the CUDA threads of a CTA, and TSan reports them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lpm_v2_b200 import api, mesh, problems, solvers      # noqa: E402
from test_parity_gpu import _rand_sphere                  # noqa: E402

nd = api.init(0)
x, y, z, zeta, area, mask = _rand_sphere(1500, 5, 0.7)
av = zeta + 0.3 * z
api.tune("sym_min_sources", 0)
for sym_on in (False, True):
    api.set_symmetric(sym_on)
    api.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    api.bve_stream(x, y, z, zeta, av, area, mask, 1.0)
api.tune("sym_min_sources", 200000)
m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, 3)
f = problems.rossby_haurwitz54(m)
eps = m.max_edge_length ** 0.75
api.pse_laplacian_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps, 1.0)          # cell order + tile culling
api.pse_gradient_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps, 1.0)
api.pse_double_dot_sphere(m.x, m.y, m.z, -m.y, m.x, m.z * m.x, m.area, m.is_active, eps, 1.0)
q = mesh.PolyMesh2d(mesh.QUAD_RECT_SEED, 3, 7.0)
vq = problems.colliding_dipoles(q)
api.plane_velocity(q.x, q.y, vq, q.area, q.is_active)
api.plane_stream(q.x, q.y, vq, q.area, q.is_active)
api.pse_laplacian_plane(q.x, q.y, vq, q.area, q.is_active, q.max_edge_length ** 0.75)
api.swe_plane_rhs_integrals(q.x, q.y, vq, 0.1 * vq, 1 + 0 * vq, q.area, q.is_active, q.max_edge_length ** 0.75)
b = mesh.PolyMesh2d(mesh.BETA_PLANE_SEED, 3)
zb = problems.betaplane_gaussian(b)
api.betaplane_velocity(b.x, b.y, zb, b.area, b.is_active)
api.betaplane_stream(b.x, b.y, zb, zb + 1, b.area, b.is_active)
sph = solvers.BVEMesh(m, problems.gaussian_vortex(m), 1.0, 2 * np.pi)
sph.velocity = list(api.bve_velocity(m.x, m.y, m.z, sph.relVort, m.area, m.is_active, 1.0))
sol = solvers.BVESolver(sph)
sol.Timestep(sph, 0.01, with_stream=True)
sol.Delete()
print("race_check workload done on", nd, "emulated device(s)")
