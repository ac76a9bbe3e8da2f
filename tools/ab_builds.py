"""A/B on one box (run under gpurun) of builds of the three triangle kernels behind tuning keys "sym_vel_build",
"sym_stream_build" and "sym_velstream_build" (a `switch (rt().sym_..._build)` in SymVel / SymStream / SymVelStream::launch,
csrc/symmetric.cuh) at icosTri level argv[1] (default 8): per build, the triangle kernel's own time (CUDA events inside
the library) and the largest difference from build 0 relative to the field scale.
    python tools/ab_builds.py [level] [vel builds] [stream builds] [velstream builds]      (lists like 0,1,2; "-" = skip)
The keys exist only while a round measures builds.  profiles/r02m_ab_builds.log (round 2) was produced with these
(launch_sym<K, targets per thread, threads, source batch, CTAs per SM, statement order, warps combined, tile>; "rot" =
lane-rotated source order, now unconditional in sym_reduce_red):
  velocity   0 <8,128,8,1,43,comb,256> (the default until then)   1-4 the same, rot, orders 43 / 27 / 59 / 11 (4 won)
  stream     0 <4,256,4,2,-,no,256> (the default until then)  1 <4,256,4,2,-,comb,128>  2 <8,128,4,2,-,comb,128>
             3 <8,128,8,2,-,comb,128>  4 = 1 with tiles of 64  5 <4,512,4,1,-,comb,128>  6 = 1 rot  7 = 0 rot
             8 = 3 rot (won)  9 <4,256,8,2,-,comb,128> rot
  fused end  0 <4,128,4,2,0,no,256> (the default until then)  1 <4,128,4,2,0,comb,64>  2 <4,256,4,1,0,comb,128>
             3 <4,256,4,1,0,no,256>  4 = 1 rot (won)  5 = 0 rot  6 = 2 rot
profiles/r02n_ab_builds.log (this version of the script; none of these beat the defaults and all were deleted):
  "sym_vel_build"        0 order 11 (default)  1 order 43 + the last group's transposed sums first (the warp reduction overlaps
                         that group's own sums), fence after the first group only  2 the same in groups of 2 sources
                         3 = 1 with the batch loop unrolled by 2  4 order 43 unrolled by 2  5 order 11, last group's
                         transposed sums first  6 groups of 2 sources, fenced
  "ds_stream_T"          0 default (4 targets per thread, 2 CTAs of 256 threads)  8: 8 targets, one CTA per SM (ds_kernel<BveStream>)
  "sym_velstream_build"  0 default  1 transposed sums before own sums per source
  "ds_velstream_T"       0 default (4)  6 / 8 targets per thread (ds_kernel<BveVelStream>)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lpm_v2_b200 import api, mesh, problems, solvers


def lst(i, default):
    a = sys.argv[i] if len(sys.argv) > i else default
    return [] if a == "-" else [int(v) for v in a.split(",")]


L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
vel, stream, velstream = lst(2, "0,1,2,3,4,5,6"), lst(3, "0,8"), lst(4, "0,1,0")
steps_T = lst(5, "0,6,8")      # "ds_velstream_T" of the RK4 steps, paired with the velstream builds
api.init(1)
api.set_profiling(True)
api.tune("sym_min_sources", 0)
m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, L)
z = problems.rossby_haurwitz54(m)
av = problems.abs_vorticity(m, z, 2 * np.pi)
for name, key, builds, fn in (("bve_velocity", "sym_vel_build", vel, lambda: api.bve_velocity(m.x, m.y, m.z, z, m.area, m.is_active, 1.0)),
                              ("bve_stream", "ds_stream_T", stream, lambda: api.bve_stream(m.x, m.y, m.z, z, av, m.area, m.is_active, 1.0))):
    ref = None
    for b in builds:
        api.tune(key, b)
        ts = []
        for _ in range(2):
            api.profile_breakdown(reset=True)
            out = fn()
            ks = api.profile_breakdown(reset=True)
            ts.append({k.split("/")[1]: v[1] / v[0] for k, v in ks.items()})
        if ref is None:
            ref = out
        scale = max(np.abs(r).max() for r in ref)
        diff = max(np.abs(a - c).max() for a, c in zip(out, ref)) / scale
        print(f"L{L} {name} {key} {b}: " + " | ".join(", ".join(f"{k} {v:.2f}" for k, v in t.items()) for t in ts) +
              f" ms; whole sum {api.last_sum_ms():.2f} ms; max diff from {builds[0]}: {diff:.2e}", flush=True)
    if builds:
        api.tune(key, 0)
if velstream:
    dt = 0.01 * float(np.sqrt(6144.0 / m.n_active))
    ref = None
    combos = list(zip(velstream, steps_T))
    for i, (b, T) in enumerate([combos[0]] + combos):        # the first step also warms the solver's buffers up
        api.tune("sym_velstream_build", b)
        api.tune("ds_velstream_T", T)
        sph = solvers.BVEMesh(m, z, 1.0, 2 * np.pi)
        sph.SetVelocityOnMesh()
        sol = solvers.BVESolver(sph)
        api.profile_breakdown(reset=True)
        t0 = time.perf_counter()
        sol.Timestep(sph, dt, with_stream=True, copy_back=False)
        ms = (time.perf_counter() - t0) * 1e3
        ks = api.profile_breakdown(reset=True)
        sol.CopyToMesh(sph, True)
        out = sph.velocity + [sph.relStream, sph.absStream]
        sol.Delete()
        if i == 0:
            continue
        if ref is None:
            ref = out
        diff = max(np.abs(a - c).max() / max(np.abs(c).max(), 1e-300) for a, c in zip(out, ref))
        print(f"L{L} RK4 step, sym_velstream_build {b}, ds_velstream_T {T}: step {ms:.1f} ms; kernels " +
              ", ".join(f"{k} {c}x {t / c:.1f}" for k, (c, t) in ks.items()) + f" ms; max diff from the first: {diff:.2e}", flush=True)
    api.tune("sym_velstream_build", 0)
    api.tune("ds_velstream_T", 0)
api.tune("sym_min_sources", 200000)
