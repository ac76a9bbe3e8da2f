"""A/B on one box (run under gpurun): the default BVE velocity / stream-function paths against the EXPERIMENTAL pair-symmetric ones
(lpm_set_bve_variant(200 .. 203), csrc/symmetric.cuh) at icosTri levels argv[1] (default "6,7").
Prints, per variant, the whole-sum time (pack, kernels, finalize / gather / scatter), the interactions/s, and the
largest difference from the default path's result relative to the field scale.
    python tools/ab_sym.py 7,8"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems

levels = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [6, 7]
variants = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 200, 201, 202, 203, 204, 205, 206, 207, 208, 209]
api.init(1)
api.set_profiling(True)
for L in levels:
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
    z = problems.rossby_haurwitz54(m)
    pairs = m.n * m.n_active - m.n_active
    av = problems.abs_vorticity(m, z, 2 * np.pi)
    for name, fn in (("bve_velocity", lambda: api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)),
                     ("bve_stream", lambda: api.bve_stream(m.x, m.y, m.z, z, av, m.area, m.is_active, 1.0))):
        ref = None
        for var in variants:
            api.set_bve_variant(var)
            ts = []
            for _ in range(3 if L < 8 else 2):
                out = fn()
                ts.append(api.last_sum_ms())
            if ref is None:
                ref = out
            scale = max(np.abs(r).max() for r in ref)
            diff = max(np.abs(a - b).max() for a, b in zip(out, ref)) / scale
            print(f"L{L} {name} variant {var}: sum ms " + " ".join(f"{t:.3f}" for t in ts) +
                  f"  -> {pairs / min(ts) / 1e9:.1f} G interactions/s, max diff from variant {variants[0]}: {diff:.2e}", flush=True)
for L in levels:
    q = mesh.PolyMesh2d(mesh.QUAD_RECT_SEED, L + 1, 7.0)
    vort = problems.colliding_dipoles(q)
    pairs = q.n * q.n_active - q.n_active
    ref = None
    for var in variants:
        api.set_bve_variant(var)
        ts = []
        for _ in range(3):
            out = api.plane_velocity(q.x, q.y, vort, q.area, q.is_active)
            ts.append(api.last_sum_ms())
        if ref is None:
            ref = out
        scale = max(np.abs(r).max() for r in ref)
        diff = max(np.abs(a - b).max() for a, b in zip(out, ref)) / scale
        print(f"quadRect L{L + 1} plane_velocity variant {var}: sum ms " + " ".join(f"{t:.3f}" for t in ts) +
              f"  -> {pairs / min(ts) / 1e9:.1f} G interactions/s, max diff from variant {variants[0]}: {diff:.2e}", flush=True)
for L in levels:
    b = mesh.PolyMesh2d(mesh.BETA_PLANE_SEED, L)
    zb = problems.betaplane_gaussian(b)
    pairs = b.n * b.n_active - b.n_active
    for name, fn in (("betaplane_velocity", lambda: api.betaplane_velocity(b.x, b.y, zb, b.area, b.is_active)),
                     ("betaplane_stream", lambda: api.betaplane_stream(b.x, b.y, zb, zb + 1.0, b.area, b.is_active))):
        ref = None
        for var in variants:
            api.set_bve_variant(var)
            ts = []
            for _ in range(3):
                out = fn()
                ts.append(api.last_sum_ms())
            if ref is None:
                ref = out
            scale = max(np.abs(r).max() for r in ref)
            diff = max(np.abs(a - c).max() for a, c in zip(out, ref)) / scale
            print(f"betaPlane L{L} {name} variant {var}: sum ms " + " ".join(f"{t:.3f}" for t in ts) +
                  f"  -> {pairs / min(ts) / 1e9:.1f} G interactions/s, max diff from variant {variants[0]}: {diff:.2e}", flush=True)
api.set_bve_variant(0)
