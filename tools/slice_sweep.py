"""Small-slice tail of the one-sided engine, measured on ONE GPU: evaluates the LoadBalance slice that rank r of
`world` ranks owns (the launch shape depends on the whole particle set, so this is exactly the grid a rank of a
`world`-GPU run launches) for different source chunkings and targets per thread, and prints the main kernel's
time against the FP64-pipe bound of the slice's interactions (9 FP64 instructions each).
    python tools/slice_sweep.py [level] [world]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lpm_v2_b200 import api, mesh, problems, torch_api, dist as D

L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
api.init_rank(0)
api.set_profiling(True)
api.set_symmetric(False)
peak_tf, _ = api.fp64_peak_probe(20000)
m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
z = problems.rossby_haurwitz54(m)
dev = torch.device("cuda:0")
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (m.x, m.y, m.z, z, m.area)]
mask = torch.from_numpy(m.is_active).to(dev)
out = [torch.zeros(m.n, dtype=torch.float64, device=dev) for _ in range(3)]
b, e = D.slice_of(m.n, world, world // 2)
inter = (e - b) * m.n_active
bound_ms = inter * 9 * 2 / (peak_tf * 1e12) * 1e3
print(f"L{L}: slice of rank {world // 2} of {world}: {e - b} targets x {m.n_active} sources, FP64-pipe bound {bound_ms:.4f} ms "
      f"(peak {peak_tf:.2f} TF)", flush=True)
for T in (0, 1, 2, 4, 8):
    for cmin in (8192, 4096, 2048, 1024, 512):
        for cap in (64, 256):
            api.tune("force_T", T); api.tune("chunk_min", cmin); api.tune("max_chunks", cap)
            best, sums = 1e30, 1e30
            for _ in range(5):
                api.profile_breakdown(reset=True)
                torch_api.bve_velocity_dev(*t, mask, 1.0, b, e, *out)
                torch.cuda.synchronize()
                ks = api.profile_breakdown(reset=True)
                k = ks["bve_velocity/one_sided"]
                best = min(best, k[1] / k[0]); sums = min(sums, api.last_sum_ms())
            print(f"  T={T or 'auto'} chunk_min={cmin} max_chunks={cap}: kernel {best:.4f} ms = {bound_ms / best:.3f} of the bound; "
                  f"whole sum {sums:.4f} ms = {bound_ms / sums:.3f}", flush=True)
api.tune("force_T", 0); api.tune("chunk_min", 1024); api.tune("max_chunks", 64)
