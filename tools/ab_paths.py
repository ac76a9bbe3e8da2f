"""A/B on one box (run under gpurun): the one-sided engine against the pair-symmetric path of the BVE velocity and
stream-function sums at icosTri levels argv[1] (default "7,8"), and any builds still under A/B
(a key of csrc/lpm_gpu_tuning.h, argv[3]).  Prints, per build, the whole-sum time (pack,
kernels, fixed-point conversion, finalize / gather / scatter), the main kernels' own times, interactions/s, and
the largest difference from the one-sided result relative to the field scale.
    python tools/ab_paths.py 7,8 [velocity builds, e.g. 0,1,2,3] [tuning key]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems

levels = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [7, 8]
vshapes = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
sshapes = [0]
vkey = sys.argv[3] if len(sys.argv) > 3 else None
api.init(1)
api.set_profiling(True)
api.tune("sym_min_sources", 0)
for L in levels:
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
    z = problems.rossby_haurwitz54(m)
    pairs = m.n * m.n_active - m.n_active
    av = problems.abs_vorticity(m, z, 2 * np.pi)
    sums = (("bve_velocity", vkey, vshapes, lambda: api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)),
            ("bve_stream", None, sshapes, lambda: api.bve_stream(m.x, m.y, m.z, z, av, m.area, m.is_active, 1.0)))
    for name, key, shapes, fn in sums:
        ref = None
        for build in [None] + list(shapes):
            api.set_symmetric(build is not None)
            if build is not None and key:
                api.tune(key, build)
            ts, ks = [], None
            for _ in range(3 if L < 8 else 2):
                api.profile_breakdown(reset=True)
                out = fn()
                ts.append(api.last_sum_ms())
                ks = api.profile_breakdown(reset=True)
            if ref is None:
                ref = out
            scale = max(np.abs(r).max() for r in ref)
            diff = max(np.abs(a - b).max() for a, b in zip(out, ref)) / scale
            label = "one-sided" if build is None else f"symmetric shape {build}"
            kern = ", ".join(f"{k.split('/')[1]} {v[1] / v[0]:.3f}" for k, v in ks.items())
            print(f"L{L} {name} {label}: sum ms " + " ".join(f"{t:.3f}" for t in ts) + f"  (kernels: {kern})"
                  f"  -> {pairs / min(ts) / 1e9:.1f} G interactions/s, max diff from one-sided: {diff:.2e}", flush=True)
        if key:
            api.tune(key, shapes[0])        # the first value listed is the default
api.set_symmetric(True)
api.tune("sym_min_sources", 200000)
