"""World-size-2 (and 3) gloo tests of the host-side multi-rank logic: LoadBalance
slices, the slice exchange, unique-id broadcast.  Each rank evaluates its slice
with the CPU oracle (standing in for the GPU kernel, which needs a B200) and the
exchanged result must equal the single-rank evaluation bit for bit -- the
reference's own reproducibility property."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lpm_v2_b200 import dist as D, mesh as M, problems, api
        from oracle import binding as O
        m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 2)
        zeta = problems.gaussian_vortex(m)
        b, e = D.slice_of(m.n, world, rank)
        s, en, ln = api.load_balance(m.n, world)
        assert (b, e) == (int(s[rank]) - 1, int(en[rank])) and e - b == int(ln[rank])
        u, v, w = O.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0, rng=(b, e))
        ts = [torch.from_numpy(a) for a in (u, v, w)]
        D.exchange_slices(ts)
        uid = D.broadcast_unique_id(bytes(range(128)) if rank == 0 else None)
        assert uid == bytes(range(128))
        tmax = D.max_over_ranks(float(rank + 1))
        assert tmax == float(world)
        if rank == 0:
            full = O.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
            ok = all(np.array_equal(t.numpy(), f) for t, f in zip(ts, full))
            q.put("ok" if ok else "mismatch")
    except Exception as ex:     # pragma: no cover
        if rank == 0:
            q.put(f"error: {ex!r}")
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slice_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == "ok"
