"""Statement-order sweep of the pair-symmetric velocity kernel on the GPU (a library built with -DLPM_SYM_ORDER_SWEEP
holds one instantiation per order behind lpm_tune("sym_vel_order")):
    mkdir -p build/order_sweep/lpm_v2_b200 && (cd lpm_v2_b200 && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a \
        -Xcompiler -fPIC -I../include -Icsrc -DLPM_SYM_ORDER_SWEEP -shared -o ../build/order_sweep/lpm_v2_b200/liblpmgpu.so csrc/lpm_gpu.cu -ldl)
    LPM_GPU_LIBRARY=build/order_sweep/lpm_v2_b200/liblpmgpu.so python tools/order_sweep.py [level]
ORDER bits (SymBveVel::batch): 0 denominators coordinate-major, 1 a-phase component-major, 2-3 sources per group
(1, 2, 4), 4 transposed sums target-major, 5 scheduling fence per group."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems

L = int(sys.argv[1]) if len(sys.argv) > 1 else 7
orders = list(range(0, 12)) + list(range(16, 27)) + [27] + [40, 41, 42, 43, 56, 57, 58, 59]      # the default (11) is among them
api.init(1)
api.set_profiling(True)
api.tune("sym_min_sources", 0)
res = {}
for lev in ([L] if L == 7 else [7, L]):
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, lev)
    z = problems.rossby_haurwitz54(m)
    todo = orders if lev == 7 else sorted(res, key=res.get)[:5] + [11]
    ref = None
    for o in todo:
        api.tune("sym_vel_order", o)
        best = 1e30
        for _ in range(3 if lev == 7 else 2):
            api.profile_breakdown(reset=True)
            out = api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)
            ks = api.profile_breakdown(reset=True)
            k = ks["bve_velocity/symmetric"]
            best = min(best, k[1] / k[0])
        if ref is None:
            ref = out
        diff = max(np.abs(a - b).max() for a, b in zip(out, ref)) / max(np.abs(r).max() for r in ref)
        if lev == 7:
            res[o] = best
        print(f"L{lev} order {o:2d}: triangle kernel {best:.3f} ms   (max diff from the first order {diff:.1e})", flush=True)
api.tune("sym_vel_order", 11)
