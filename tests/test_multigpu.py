"""Multi-GPU paths (skipped on a 1-GPU box): single-process mode (all devices of the
box, peer stores + event barrier) in a subprocess, and rank mode (torchrun, NCCL
slice exchange inside liblpmgpu, and the peer-store exchange through CUDA-IPC shared slabs)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")

SINGLE_PROCESS = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from lpm_v2_b200 import api, mesh as M, problems, solvers
from oracle import binding as O
nd = api.init(0)
assert nd >= 2, nd
m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 4)
zeta = problems.gaussian_vortex(m)
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
got = api.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
want = O.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
assert max(rel(g, w) for g, w in zip(got, want)) <= 1e-12
lap = api.pse_laplacian_sphere(m.x, m.y, m.z, zeta, m.area, m.is_active, 0.2, 1.0)
assert rel(lap, O.pse_laplacian_sphere(m.x, m.y, m.z, zeta, m.area, m.is_active, 0.2, 1.0)) <= 1e-12
# kernels whose shared memory exceeds 48 KB need their attribute set on every device
dd = api.pse_double_dot_sphere(m.x, m.y, m.z, -m.y, m.x, m.z * m.x, m.area, m.is_active, 0.2, 1.0)
assert rel(dd, O.pse_double_dot_sphere(m.x, m.y, m.z, -m.y, m.x, m.z * m.x, m.area, m.is_active, 0.2, 1.0)) <= 1e-12
q = M.PolyMesh2d(M.QUAD_RECT_SEED, 4, 3.0)
vq = np.exp(-2 * (q.x ** 2 + q.y ** 2))
sw = api.swe_plane_rhs_integrals(q.x, q.y, vq, 0.1 * vq * q.x, 1 + 0.1 * vq, q.area, q.is_active, 0.3)
so = O.swe_plane_rhs(q.x, q.y, vq, 0.1 * vq * q.x, 1 + 0.1 * vq, q.area, q.is_active, 0.3)
assert max(rel(a, b) for a, b in zip(sw, so)) <= 1e-12
sph = solvers.BVEMesh(m, zeta, 1.0, 2 * np.pi)
sph.velocity = [g.copy() for g in got]
sol = solvers.BVESolver(sph)
ref = [m.x.copy(), m.y.copy(), m.z.copy(), zeta.copy()] + [w.copy() for w in want]
for step in range(2):
    sol.Timestep(sph, 0.01, with_stream=True)
    ref = O.bve_rk4_step(*ref, m.area, m.is_active, 1.0, 2 * np.pi, 0.01)
    for a, b in zip([sph.x, sph.y, sph.z, sph.relVort] + sph.velocity, ref):
        assert rel(a, b) <= 1e-12
rs, as_ = O.bve_stream(ref[0], ref[1], ref[2], ref[3], sph.absVort, m.area, m.is_active, 1.0)
assert rel(sph.relStream, rs) <= 1e-12 and rel(sph.absStream, as_) <= 1e-12
sol.Delete()
api.finalize()
print("OK", nd)
"""

RANK_MODE = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from lpm_v2_b200 import api, torch_api, mesh as M, problems, solvers, dist as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
api.init_rank(local)
uid = D.broadcast_unique_id(api.comm_unique_id() if rank == 0 else None)
api.comm_init_rank(world, rank, uid)
m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 5)
zeta = problems.rossby_haurwitz54(m)
# host API in rank mode: every rank gets the complete arrays
got = api.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
# device API + explicit slice exchange
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (m.x, m.y, m.z, zeta, m.area)]
mask = torch.from_numpy(m.is_active).to(dev)
out = [torch.zeros(m.n, dtype=torch.float64, device=dev) for _ in range(3)]
b, e = D.slice_of(m.n, world, rank)
torch_api.bve_velocity_dev(*t, mask, 1.0, b, e, *out)
torch_api.allgather_slices_dev(out)
torch.cuda.synchronize()
for o, g in zip(out, got):
    assert np.array_equal(o.cpu().numpy(), g)
# device API with the outputs in a shared slab: the finalize step stores every slice to every
# rank over NVLink (no NCCL exchange); bit-identical to the NCCL path
sh, slab = torch_api.shared_tensors(3, m.n, dev)
assert api.comm_is_shared(sh[0].data_ptr(), m.n * 8) and not api.comm_is_shared(out[0].data_ptr(), m.n * 8)
for rep in range(3):
    for s_ in sh:
        s_.fill_(float(rep))
    torch_api.bve_velocity_dev(*t, mask, 1.0, b, e, *sh)
    torch.cuda.synchronize()
    for s_, g in zip(sh, got):
        assert np.array_equal(s_.cpu().numpy(), g), rep
# a PSE sum (cell-ordered targets: the scatter kernel does the peer stores)
lap_sh, slab2 = torch_api.shared_tensors(1, m.n, dev)
f_t = torch.from_numpy(problems.spherical_harmonic54(m)).to(dev)
torch_api.pse_laplacian_sphere_dev(t[0], t[1], t[2], f_t, t[4], mask, 0.1, 1.0, b, e, lap_sh[0])
torch.cuda.synchronize()
lap_host = api.pse_laplacian_sphere(m.x, m.y, m.z, problems.spherical_harmonic54(m), m.area, m.is_active, 0.1, 1.0)
assert np.array_equal(lap_sh[0].cpu().numpy(), lap_host)
del sh, lap_sh
api.comm_free_shared(slab); api.comm_free_shared(slab2)
# the pair-symmetric path in rank mode (target blocks dealt round-robin, integer all-reduce of the fixed-point
# accumulators, grouped broadcast of the passive slices): parity, and the SAME BITS as one GPU alone computed
api.tune("sym_min_sources", 0)
sym_v = api.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
av = problems.abs_vorticity(m, zeta, 2 * np.pi)
sym_s = api.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
api.tune("sym_min_sources", 200000)
one = np.load(%(symfile)r)
assert np.array_equal(np.stack(list(sym_v) + list(sym_s)), one), "symmetric path: bits depend on the rank count"
for a_, b_ in zip(sym_v, got):
    assert float(np.abs(a_ - b_).max() / np.abs(b_).max()) <= 1e-13
# all ranks hold identical results
chk = torch.stack([o.sum() for o in out])
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi)
# resident solver in rank mode
sph = solvers.BVEMesh(m, zeta, 1.0, 2 * np.pi)
sph.velocity = [g.copy() for g in got]
sol = solvers.BVESolver(sph)
sol.Timestep(sph, 0.01, with_stream=True)
sol.Delete()
if rank == 0:
    from oracle import binding as O
    want = O.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert max(rel(g, w) for g, w in zip(got, want)) <= 1e-12
    ref = O.bve_rk4_step(m.x, m.y, m.z, zeta, *got, m.area, m.is_active, 1.0, 2 * np.pi, 0.01)
    for a, b2 in zip([sph.x, sph.y, sph.z, sph.relVort] + sph.velocity, ref):
        assert rel(a, b2) <= 1e-12
    print("OK", world)
dist.destroy_process_group()
"""


SYM_SINGLE = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from lpm_v2_b200 import api, mesh as M, problems
api.init_rank(0)
m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 5)
zeta = problems.rossby_haurwitz54(m)
av = problems.abs_vorticity(m, zeta, 2 * np.pi)
api.tune("sym_min_sources", 0)
v = api.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
s = api.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
np.save(%(symfile)r, np.stack(list(v) + list(s)))
api.finalize()
print("OK")
"""


@needs2
def test_single_process_all_gpus(tmp_path):
    script = tmp_path / "sp.py"
    script.write_text(SINGLE_PROCESS % {"root": ROOT})
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@needs2
def test_rank_mode_torchrun(tmp_path):
    symfile = str(tmp_path / "sym_one_gpu.npy")
    one = tmp_path / "one.py"
    one.write_text(SYM_SINGLE % {"root": ROOT, "symfile": symfile})
    r = subprocess.run([sys.executable, str(one)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    script = tmp_path / "rk.py"
    script.write_text(RANK_MODE % {"root": ROOT, "symfile": symfile})
    n = min(_ngpu(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
