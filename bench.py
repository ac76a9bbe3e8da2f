#!/usr/bin/env python
"""bench.py -- the hot-path benchmark the driver runs.

Metric (BASELINE.json): FP64 BVE direct-sum interactions/s at icosTri level 8
(N = 1 966 082 targets x F = 1 310 720 active panels, 2.577e12 interactions per
evaluation), 1/2/4/8 B200, strong scaling.  One "step" = one evaluation of
BVESphereVelocity (src/SphereBVESolver.f90:377-430) over all targets + the
slice exchange that replaces its MPI_BCAST loop.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--level L]
  torchrun ... bench.py --gpus N ...        (one rank per GPU)

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_INTERACTION = 22.0     # dot 5, R^2-dot 1, divide 1, cross 9, accumulate 6
METRIC = "bve_direct_sum_interactions_per_s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--level", type=int, default=8, help="icosTri refinement level (8 = headline)")
    ap.add_argument("--no-rk4", action="store_true", help="skip the RK4 step-time measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the per-rank oracle parity sample")
    ap.add_argument("--parity-targets", type=int, default=4096, help="contiguous targets per rank in the parity sample")
    ap.add_argument("--e2e-reps", type=int, default=0, help="timed end-to-end calls (0 = min(steps, 3))")
    ap.add_argument("--one-sided", action="store_true", help="A/B: lpm_set_symmetric(0), every ordered pair evaluated")
    return ap.parse_args()


def workload(level):
    from lpm_v2_b200 import mesh, problems
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, level)
    zeta = problems.rossby_haurwitz54(m)          # config 4: RH54 vorticity (examples/rh54.namelist)
    return m, zeta


def config_dict(level, m, gpus):
    """The workload, identical in both arms (the driver compares the two lines' configs)."""
    return {
        "workload": f"RossbyHaurwitz54 BVE direct sum, icosTri level {level} (faceKind=3, initNest={level}): "
                    f"{m.n} targets x {m.n_active} active panels, RH54 vorticity (examples/rh54.namelist), R=1",
        "interactions_per_evaluation": int(m.n) * int(m.n_active) - int(m.n_active),
        "partition": f"targets split by the reference's LoadBalance rule (src/MPISetup.f90:132-146) over --gpus {gpus} "
                     "ranks, sources replicated on every rank",
        "l2": "256 MiB memset between timed steps (time included); sources (63 MB) are meant to live in L2",
    }


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU legs
def cpu_leg(m, zeta, sample_targets, threads, steps=1, warmup=0):
    """The oracle port (kind "port"; the Fortran reference cannot be built here) on the
    host cores: `sample_targets` contiguous targets x all sources, LoadBalance-split
    over `threads` workers (stand-in for mpirun -np threads).  Returns interactions/s."""
    from oracle import binding as O
    tb = (m.n - sample_targets) // 2
    te = tb + sample_targets
    for _ in range(warmup):
        O.bve_velocity_mt(threads, m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0, tb, min(te, tb + 64 * threads))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.bve_velocity_mt(threads, m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0, tb, te)
    dt = (time.perf_counter() - t0) / steps
    act = m.is_active[tb:te] != 0
    inter = sample_targets * m.n_active - int(act.sum())
    return inter / dt, dt, f"{sample_targets} contiguous targets [{tb},{te}) x all {m.n_active} active sources", inter


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the same path, all host threads.  One step = one
    bounded sample of the evaluation (a contiguous block of targets x all sources): `value` is the sample's
    interactions over the sample's time, `ms_per_step` the sample's time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m, zeta = workload(args.level)
    threads = os.cpu_count() or 1
    per_thread = 512 if args.level >= 7 else max(8, m.n // threads // 4)
    sample = min(m.n, per_thread * threads)
    val, dt, desc, inter = cpu_leg(m, zeta, sample, threads, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "interactions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.level, m, args.gpus),
        "cpu_baseline": {"value": val, "unit": "interactions/s", "cores": threads, "kind": "port",
                         "interactions_per_step": inter,
                         "sample": "one step = " + desc + f" = {inter} interactions ({100.0 * sample / m.n:.2f} % of an "
                                   "evaluation), ms_per_step is that sample's time; C restatement of BVESphereVelocity "
                                   "(oracle/lpm_oracle.c, -O3 -march=native), pthread workers on the LoadBalance split; "
                                   "the Fortran + MPI reference cannot be built in this image (no Fortran compiler)"},
        "e2e": {"value": val, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from lpm_v2_b200 import api, torch_api, solvers, dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    api.init_rank(local)
    if world > 1:
        uid = D.broadcast_unique_id(api.comm_unique_id() if rank == 0 else None)
        api.comm_init_rank(world, rank, uid)
    api.set_symmetric(not args.one_sided)

    m, zeta = workload(args.level)
    n, F = m.n, m.n_active
    inter = n * F - F
    ibeg, iend = D.slice_of(n, world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident inputs
    host = {k: torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for k, a in
            dict(x=m.x, y=m.y, z=m.z, q=zeta, a=m.area).items()}
    hmask = torch.from_numpy(np.ascontiguousarray(m.is_active)).pin_memory()
    d = {k: t.to(dev) for k, t in host.items()}
    dmask = hmask.to(dev)
    # N > 1: the outputs live in a CUDA-IPC shared slab, so the sum's finalize step stores every
    # slice into every rank's copy over NVLink (the MPI_BCAST loop fused into the sum)
    if world > 1:
        out, out_slab = torch_api.shared_tensors(3, n, dev)
    else:
        out = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        torch_api.bve_velocity_dev(d["x"], d["y"], d["z"], d["q"], d["a"], dmask, 1.0, ibeg, iend, *out, stream=stream)
        if world > 1 and not exchange_fused:
            torch_api.allgather_slices_dev(out, stream=stream)

    exchange_fused = world > 1 and api.comm_is_shared(out[0].data_ptr(), n * 8)
    api.set_profiling(True)
    for _ in range(args.warmup):
        step()
    barrier()
    api.profile_breakdown(reset=True)
    api.launch_count(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        step()
        if k + 1 < args.steps:
            flush.zero_()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = api.launch_count()
    kernels = api.profile_breakdown(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = D.max_over_ranks(ms_total / args.steps, dev) if world > 1 else ms_total / args.steps
    value = inter / (ms_step * 1e-3)

    # ---- end to end through the reference-facing C-ABI host call (host buffers, H2D + D2H inside)
    hout = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]

    def e2e_step():
        api.check(api.lib.lpm_bve_velocity(
            n, *[api.C.cast(host[k].data_ptr(), api._d) for k in ("x", "y", "z", "q", "a")],
            api.C.cast(hmask.data_ptr(), api._i32), 1.0, *[api.C.cast(t.data_ptr(), api._d) for t in hout]))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_reps = args.e2e_reps or max(1, min(args.steps, 3))
    for _ in range(e2e_reps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_reps
    if world > 1:
        e2e_s = D.max_over_ranks(e2e_s, dev)
    e2e_ok = bool(np.array_equal(hout[0].numpy(), out[0].cpu().numpy()))

    # ---- parity sample: >= 4096 contiguous targets of THIS rank's slice against the parity build of the oracle
    # (-O2, no contraction, the reference's own loop order), relative to the field scale, max over ranks
    parity = None
    if not args.no_parity:
        from oracle import binding as O
        cnt = min(args.parity_targets, iend - ibeg)
        tb = ibeg + (iend - ibeg - cnt) // 2
        threads = max(1, (os.cpu_count() or 1) // world)
        ref = O.bve_velocity_mt(threads, m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0, tb, tb + cnt, fast=False)
        err = 0.0
        for g, r in zip(out, ref):
            gs = g[tb:tb + cnt].cpu().numpy()
            err = max(err, float(np.abs(gs - r[tb:tb + cnt]).max() / g.abs().max().item()))
        # how far the as-written FP64 reference itself is from the exact sum (long double adjudicator) on the first
        # targets of the sample, and how far the GPU is: R^2 - x_i.x_j cancels to ~h^2 for neighbours, so from
        # icosTri 9 upwards BOTH carry ~1e-12 of rounding in the nearest terms and differ by that much from each other
        nld = min(8, cnt)
        ld = O.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0, rng=(tb, tb + nld), variant="_ld")
        ref_ld = max(float(np.abs(r[tb:tb + nld] - l[tb:tb + nld]).max() / g.abs().max().item()) for g, r, l in zip(out, ref, ld))
        gpu_ld = max(float(np.abs(g[tb:tb + nld].cpu().numpy() - l[tb:tb + nld]).max() / g.abs().max().item()) for g, l in zip(out, ld))
        if world > 1:
            err = D.max_over_ranks(err, dev)
            ref_ld = D.max_over_ranks(ref_ld, dev)
            gpu_ld = D.max_over_ranks(gpu_ld, dev)
        parity = {"max_rel_err": err, "tolerance": 1e-12, "targets": int(cnt * world), "ranks": world,
                  "reference_vs_extended_precision": ref_ld, "gpu_vs_extended_precision": gpu_ld,
                  "extended_precision_targets": int(nld * world),
                  "against": "oracle/lpm_oracle.c parity build (restatement of src/SphereBVESolver.f90:396-420; it reproduces the reference's own source text, executed by oracle/fortran_subset.py, bit for bit: tests/test_refsrc_golden.py), "
                             f"{cnt} contiguous targets from the middle of every rank's slice x all sources; "
                             "error = max |u_gpu - u_ref| over the sample / max |u| over all targets, per component"}

    # ---- RK4 step time through the resident solver (4 velocity sums + stream functions)
    rk4_ms, rk4_kernels = None, None
    if not args.no_rk4:
        sph = solvers.BVEMesh(m, zeta, 1.0, 2.0 * np.pi)
        sph.velocity = [t.cpu().numpy().copy() for t in out]
        sol = solvers.BVESolver(sph)
        # examples/rh54.namelist ships dt = 0.01 for 6144 panels; scaled with the mesh spacing, because the reference's
        # RK4 does not re-project stage positions onto the sphere: at dt |u| >> h the denominators R^2 - x_i.x_j of
        # neighbouring particles change sign (in the reference too) and the step is no longer a meaningful workload
        dt = 0.01 * float(np.sqrt(6144.0 / F))
        sol.Timestep(sph, dt, with_stream=True, copy_back=False)      # warm-up step
        barrier()
        api.profile_breakdown(reset=True)
        t0 = time.perf_counter()
        sol.Timestep(sph, dt, with_stream=True, copy_back=False)
        barrier()
        rk4_ms = (time.perf_counter() - t0) * 1e3
        rk4_kernels = api.profile_breakdown(reset=True)
        if world > 1:
            rk4_ms = D.max_over_ranks(rk4_ms, dev)
        sol.Delete()

    # ---- roofline of the dominant kernel on this rank
    # symmetric path: sym_kernel over this rank's share of the F^2 - F active x active interactions (target blocks
    # dealt round-robin) + ds_kernel over its slice of the (n - F) F passive-target interactions
    probe_tf, _ = api.fp64_peak_probe(20000)
    nv = n - F
    vb, ve = D.slice_of(nv, world, rank)
    work = {"bve_velocity/symmetric": (F * F - F) / world, "bve_velocity/one_sided": (ve - vb) * F}
    if "bve_velocity/symmetric" not in kernels:
        work["bve_velocity/one_sided"] = (iend - ibeg) * F - (F if world == 1 else 0)
    # FP64-pipe instructions per interaction in the hot loops (cuobjdump -sass)
    executed = {"bve_velocity/symmetric": 6.3, "bve_velocity/one_sided": 9.0}
    per_kernel = {}
    for name, (cnt_k, ms_k) in kernels.items():
        w = work.get(name)
        if not w or not cnt_k:
            continue
        t = ms_k / cnt_k * 1e-3
        per_kernel[name] = {"launches": cnt_k, "ms_per_launch": ms_k / cnt_k, "share_of_step": ms_k / cnt_k / (ms_total / args.steps),
                            "interactions_per_launch": int(w),
                            "achieved_tflops": FLOP_PER_INTERACTION * w / t / 1e12,
                            "fp64_pipe_frac": executed[name] * 2.0 * w / t / 1e12 / probe_tf}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_launch"]) if per_kernel else None
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "bve_velocity_ncu_summary.json")
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            traffic = pj.get("dram_bytes_per_launch")
            traffic_src = ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from "
                           "the committed ncu --set full capture " + str(pj.get("capture", "profiles/bve_velocity_ncu_summary.json")))
        except Exception:
            traffic = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    domk = per_kernel.get(dom, {})
    line = {
        "metric": METRIC, "value": value, "unit": "interactions/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.level, m, args.gpus),
        "roofline": {
            "bound": "fp64", "kernel": dom, "achieved": domk.get("achieved_tflops"), "peak": probe_tf, "unit": "TFLOP/s",
            "frac": (domk.get("achieved_tflops") or 0.0) / probe_tf, "traffic": traffic, "traffic_source": traffic_src,
            "fp64_pipe_frac": domk.get("fp64_pipe_frac"),
            "step_frac": FLOP_PER_INTERACTION * inter / world / (ms_total / args.steps * 1e-3) / 1e12 / probe_tf,
            "kernels": per_kernel,
            "rk4_step_ms": rk4_ms,
            "rk4_dt": 0.01 * float(np.sqrt(6144.0 / F)),
            "rk4_kernels": {k: {"launches": c, "ms": t} for k, (c, t) in (rk4_kernels or {}).items()},
            "note": f"achieved = algorithmic {FLOP_PER_INTERACTION:.0f} FLOP/interaction (SURVEY.md 8d) x the interactions one "
                    "launch of `kernel` covers on rank 0 / its mean duration (CUDA events on the launching stream inside the "
                    "timed region); peak = DFMA probe measured in this run (MEASURED_PEAKS.json has no FP64 figure; nominal "
                    "148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2).  frac exceeds 1 because the kernels execute fewer FP64 "
                    "instructions than the algorithmic count: the pair-symmetric kernel shares the denominator and its "
                    "reciprocal between i<-j and j<-i (6.3 FP64 instructions = 12.6 flop per interaction), the one-sided kernel "
                    "factors the cross product out of the pair loop (9 = 18 flop); fp64_pipe_frac = executed FP64 "
                    "instructions / pipe peak is the utilisation figure.  step_frac = the whole step (all kernels, finalize, "
                    "L2 flush) on the same algorithmic count.  rk4_step_ms = one resident BVESolver Timestep (4 velocity "
                    "sums + stream functions, no host copies), max over ranks.",
        },
        "e2e": {"value": inter / e2e_s, "unit": "interactions/s",
                "h2d_bytes_per_step": int(world * (5 * 8 * n + 4 * n)), "d2h_bytes_per_step": int(world * 3 * 8 * n),
                "ms_per_step": e2e_s * 1e3, "matches_resident_result": e2e_ok,
                "call": "lpm_bve_velocity (host pointers, pinned) per rank"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity_sample": parity,
        "path": "one-sided (lpm_set_symmetric(0))" if args.one_sided else "default (pair-symmetric for whole evaluations)",
    }
    if not args.no_cpu and world == 1:
        threads = os.cpu_count() or 1
        sample = min(n, (1024 if args.level >= 7 else 64) * threads)
        val, dt, desc, cinter = cpu_leg(m, zeta, sample, threads)
        line["cpu_baseline"] = {"value": val, "unit": "interactions/s", "cores": threads, "kind": "port",
                                "sample": desc + f" ({dt:.1f} s); C restatement of the reference loop "
                                                 "(oracle/lpm_oracle.c, -O3 -march=native), one pthread worker per core "
                                                 "on the LoadBalance split; the Fortran+MPI reference cannot be built in this image"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
