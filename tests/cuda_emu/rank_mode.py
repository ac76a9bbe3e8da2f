"""One rank of an emulated rank-mode run (tests/test_emu_abi.py::test_rank_mode_on_emulator): one process per
"GPU", liblpmgpu_emu.so + tests/cuda_emu/fake_nccl/libnccl.so.2.  Through the host API every rank evaluates its
LoadBalance slice and the slices are exchanged -- by peer stores into the shared output slabs between two all-reduce
barriers (one-sided engine) or, on the pair-symmetric path, by an integer all-reduce of the fixed-point accumulators
and the grouped broadcast of the passive slices -- so each rank must end up with all n results.
usage: rank_mode.py <world> <rank> <file for the unique id>"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpm_v2_b200 import api, mesh as M, problems, solvers      # noqa: E402
from oracle import binding as O                                # noqa: E402

world, rank, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
api.init_rank(0)
if world == 1:
    pass
elif rank == 0:
    uid = api.comm_unique_id()
    with open(idfile + ".tmp", "wb") as f:
        f.write(uid)
    os.replace(idfile + ".tmp", idfile)
else:
    for _ in range(6000):
        if os.path.exists(idfile):
            break
        time.sleep(0.01)
    uid = open(idfile, "rb").read()
if world > 1:
    api.comm_init_rank(world, rank, uid)

rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 3)
zeta = problems.rossby_haurwitz54(m)
av = problems.abs_vorticity(m, zeta, 2 * np.pi)
want = O.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
wants = O.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
q = M.PolyMesh2d(M.QUAD_RECT_SEED, 3, 7.0)
vq = problems.colliding_dipoles(q)
wantq = O.plane_velocity(q.x, q.y, vq, q.area, q.is_active)
bp = M.PolyMesh2d(M.BETA_PLANE_SEED, 3)
zb = problems.betaplane_gaussian(bp)
wantb = O.betaplane_velocity(bp.x, bp.y, zb, bp.area, bp.is_active, variant="_ld")
wantqs = O.plane_stream(q.x, q.y, vq, q.area, q.is_active)
wantbs = O.betaplane_stream(bp.x, bp.y, zb, zb + 1.0 + 2.0 * bp.y, bp.area, bp.is_active, variant="_ld")
# sym: every whole BVE evaluation through the pair-symmetric path (block dealing, integer all-reduce of the
# fixed-point accumulators, grouped broadcast of the passive slices); one-sided: peer stores into the shared slabs
for sym_on in ((True,) if world == 1 else (False, True)):
    api.set_symmetric(sym_on)
    api.tune("sym_min_sources", 0)
    got = api.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    assert max(rel(g, w) for g, w in zip(got, want)) <= 1e-12, ("velocity", sym_on)
    gots = api.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
    assert max(rel(g, w) for g, w in zip(gots, wants)) <= 1e-12, ("stream", sym_on)
    if sym_on:      # fixed-point accumulation: the bits must not depend on the number of ranks (checked by the caller)
        np.save(idfile + f".fx.{world}.{rank}.npy", np.stack(list(got) + list(gots)))
    if world == 1:
        continue
    gotq = api.plane_velocity(q.x, q.y, vq, q.area, q.is_active)
    assert max(rel(g, w) for g, w in zip(gotq, wantq)) <= 1e-12, ("plane", sym_on)
    gotb = api.betaplane_velocity(bp.x, bp.y, zb, bp.area, bp.is_active)
    assert max(rel(g, w) for g, w in zip(gotb, wantb)) <= 1e-12, ("betaplane", sym_on)
    assert rel(api.plane_stream(q.x, q.y, vq, q.area, q.is_active), wantqs) <= 1e-12, ("plane stream", sym_on)
    gotbs = api.betaplane_stream(bp.x, bp.y, zb, zb + 1.0 + 2.0 * bp.y, bp.area, bp.is_active)
    assert max(rel(g, w) for g, w in zip(gotbs, wantbs)) <= 1e-12, ("betaplane stream", sym_on)
    # the resident solver: every array in one shared slab, four velocity sums and the stream functions per step
    zg = problems.gaussian_vortex(m)
    u0 = O.bve_velocity(m.x, m.y, m.z, zg, m.area, m.is_active, 1.0)
    sph = solvers.BVEMesh(m, zg, 1.0, 2 * np.pi)
    sph.velocity = [a.copy() for a in u0]
    sol = solvers.BVESolver(sph)
    sol.Timestep(sph, 0.01, with_stream=True)
    sol.Delete()
    ref = O.bve_rk4_step(m.x, m.y, m.z, zg, *u0, m.area, m.is_active, 1.0, 2 * np.pi, 0.01)
    for a, b in zip([sph.x, sph.y, sph.z, sph.relVort] + sph.velocity, ref):
        assert rel(a, b) <= 1e-12, ("rk4", sym_on)
api.tune("sym_min_sources", 200000)
if world == 1:
    api.finalize()
    print("OK rank", rank, "of", world, flush=True)
    sys.exit(0)
lap = api.pse_laplacian_sphere(m.x, m.y, m.z, zeta, m.area, m.is_active, 0.3, 1.0)      # the cell-ordered path's scatter exchange
assert rel(lap, O.pse_laplacian_sphere(m.x, m.y, m.z, zeta, m.area, m.is_active, 0.3, 1.0)) <= 1e-12
api.finalize()
print("OK rank", rank, "of", world, flush=True)
