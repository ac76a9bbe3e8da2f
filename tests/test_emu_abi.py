"""The whole C ABI on the SIMT emulator (CPU tier).

tests/cuda_emu/build_emu.sh compiles lpm_v2_b200/csrc -- kernels AND host code -- with g++ against the emulator
(convert.py rewrites the <<<>>> launches; emu_runtime.h stands in for the CUDA runtime, device memory = host memory)
into tests/cuda_emu/liblpmgpu_emu.so, which exports the same symbols as liblpmgpu.so.  A subprocess then runs a
subset of the GPU parity tests against it (LPM_GPU_LIBRARY): the tests that finish within seconds under emulation
(tests/cuda_emu/emu_subset.txt: 40-odd of tests/test_parity_gpu.py, test_pse_ops_gpu.py, test_swe_gpu.py) and the small
cases of the experimental pair-symmetric paths (tests/test_sym_gpu.py).

What this buys: kernel and host logic (indexing, pipelines, culling, retry paths, chunking, the symmetric schedule
and its gather / scatter plumbing) is exercised on every CPU run, before any GPU time is spent.  What it is not: a
CPU path of the product -- liblpmgpu.so has none (tests/test_abi.py::test_no_cpu_fallback), nothing under
lpm_v2_b200/ refers to the emulator, and timings under it mean nothing.
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU = os.path.join(HERE, "cuda_emu")


@pytest.fixture(scope="module")
def emu_env():
    subprocess.check_call(["bash", os.path.join(EMU, "build_emu.sh")])
    env = dict(os.environ)
    env["LPM_GPU_LIBRARY"] = os.path.join(EMU, "liblpmgpu_emu.so")
    return env


def _run(env, args, timeout):
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"] + args,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    return tail


def test_default_paths_on_emulator(emu_env):
    ids = [l.strip() for l in open(os.path.join(EMU, "emu_subset.txt")) if l.strip()]
    tail = _run(emu_env, ids, 900)
    assert f"{len(ids)} passed" in tail, tail


def test_symmetric_paths_on_emulator(emu_env):
    """The pair-symmetric path (tests/test_sym_gpu.py lowers its size threshold) through the real host code:
    velocity, stream functions (several target blocks, so tiles above the diagonal; a radius whose logarithms leave the
    table's default window), fixed-point accumulation, panels and chunks of the launch order (with the stream kernel's
    shorter source tiles), the coincident-particle case."""
    tail = _run(emu_env, ["tests/test_sym_gpu.py", "-k",
                          "(velocity_random_ragged and (127 or 513 or 1025 or 3000 or 3-0.5)) "
                          "or (stream_random_ragged and (513 or 3-0.5 or 4099 or 6000)) or (launch_geometry and 2-2) "
                          "or (fixed_point and not 6371000 and not 3e-07) or coincident"], 900)
    assert " passed" in tail and "failed" not in tail, tail


def test_single_process_multi_device_on_emulator(emu_env, tmp_path):
    """The single-process mode of tests/test_multigpu.py (every device evaluates its LoadBalance slice and stores it
    to all replicas, event barrier, RK4 steps on every replica) on three emulated devices, one level smaller."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_tm", os.path.join(HERE, "test_multigpu.py"))
    tm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tm)
    text = (tm.SINGLE_PROCESS % {"root": ROOT}).replace("ICOS_TRI_SPHERE_SEED, 4)", "ICOS_TRI_SPHERE_SEED, 3)")
    text = text.replace("QUAD_RECT_SEED, 4, 3.0)", "QUAD_RECT_SEED, 3, 3.0)")
    assert "SEED, 3)" in text and "SEED, 3, 3.0)" in text
    script = tmp_path / "sp.py"
    script.write_text(text)
    env = dict(emu_env, LPM_EMU_DEVICES="3")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "OK 3" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_rank_mode_on_emulator(emu_env, tmp_path):
    """One process per "GPU", three ranks (a ragged LoadBalance split): tests/cuda_emu/rank_mode.py.  The NCCL calls
    go to tests/cuda_emu/fake_nccl (shared memory + a process-shared barrier); "device" allocations are shared-memory
    segments, so the CUDA-IPC slabs and the peer stores into them are real cross-process stores.  Covers the fused
    peer-store exchange with its barriers, the resident solver in a shared slab, the scatter exchange of the
    cell-ordered PSE path, and the integer all-reduce + grouped broadcast of the pair-symmetric path."""
    st = os.statvfs("/dev/shm")
    if st.f_bavail * st.f_frsize < (512 << 20):
        pytest.skip("needs 512 MB of /dev/shm for the emulated device memory of three ranks")
    world = 3
    env = dict(emu_env)
    env["LD_LIBRARY_PATH"] = os.path.join(EMU, "fake_nccl") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    idfile = str(tmp_path / "uid")
    procs = [subprocess.Popen([sys.executable, os.path.join(EMU, "rank_mode.py"), str(world), str(r), idfile],
                              env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK rank {r} of {world}" in out, out[-3000:]
    # fixed-point accumulation of the symmetric sums: one rank alone gets the same BITS as three
    r1 = subprocess.run([sys.executable, os.path.join(EMU, "rank_mode.py"), "1", "0", idfile], env=env,
                        capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0 and "OK rank 0 of 1" in r1.stdout, (r1.stdout + r1.stderr)[-3000:]
    import numpy as np
    one = np.load(idfile + ".fx.1.0.npy")
    for r in range(world):
        assert np.array_equal(np.load(idfile + f".fx.{world}.{r}.npy"), one), f"rank {r} of {world} differs from the single-rank bits"


@pytest.mark.skipif(os.environ.get("LPM_RACE_CHECK") != "1", reason="2 minutes: set LPM_RACE_CHECK=1")
def test_kernels_are_race_free_under_thread_sanitizer(tmp_path):
    """tests/cuda_emu/race_check.sh: every kernel family once through the ThreadSanitizer build of the emulated
    library; a shared-memory protocol error in a kernel is a data race between the threads that stand for a CTA."""
    log = tmp_path / "race.log"
    r = subprocess.run(["bash", os.path.join(EMU, "race_check.sh"), str(log)], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "workloads completed: 2" in r.stdout and "ThreadSanitizer reports: 0" in r.stdout, r.stdout + log.read_text()[-4000:]
