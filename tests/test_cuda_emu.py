"""The kernel SOURCES of the experimental pair-symmetric paths, run on CPU threads.

tests/cuda_emu is a minimal SIMT emulator (one OS thread per CUDA thread, pthread barriers for
__syncthreads() and the warp shuffles, memcpy + phase flip for the TMA bulk copies, a truncated 1/d for
MUFU.RCP64H): lpm_v2_b200/csrc/{directsum,pairs,sym_kernels}.cuh are compiled with g++ as they are and
their kernels executed -- sym_kernel<SymBveVel / SymBveStream> for the active x active triangle,
ds_kernel<BveVel / BveStream> on the gathered passive targets with an all-zero scan -- following the
host steps of csrc/symmetric.cuh, and the results compared with the CPU oracle.

This checks logic that had no GPU time in round 1 (triangle schedule, two-stage tile pipeline,
recursive-halving warp reduction, per-batch retry of the table logarithm, rank dealing); it is test
infrastructure only and says nothing about speed.  The GPU parity tests of the same paths are in
tests/test_sym_gpu.py (LPM_EXPERIMENTAL=1).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_parity_gpu import _rand_sphere

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU = os.path.join(HERE, "cuda_emu")
TOL = 1e-12


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU, "libcuda_emu.so")
    subprocess.check_call(["bash", os.path.join(EMU, "build_emu.sh")])      # also converts csrc for g++ (convert.py)
    lib = C.CDLL(so)
    d = np.ctypeslib.ndpointer(np.float64, flags="C")
    i32 = np.ctypeslib.ndpointer(np.int32, flags="C")
    lib.emu_sym_bve_velocity.argtypes = [C.c_int64, d, d, d, d, d, i32, C.c_double, C.c_int, C.c_int, C.c_int, d, d, d]
    lib.emu_default_bve_velocity.argtypes = [C.c_int64, d, d, d, d, d, i32, C.c_double, d, d, d]
    lib.emu_sym_plane_velocity.argtypes = [C.c_int64, d, d, d, d, i32, C.c_int, C.c_int, C.c_int, d, d]
    lib.emu_sym_bve_stream.argtypes = [C.c_int64, d, d, d, d, d, d, i32, C.c_double, C.c_int, C.c_int, C.c_int, d, d]
    return lib


def _check(got, want, ld):
    """The conditioning-aware bound of tests/test_parity_gpu.py: 1e-12 of the field scale plus the
    as-written FP64 reference's own distance from the extended-precision sum."""
    scale = max(np.abs(w).max() for w in want)
    ref_err = max(np.abs(w - l).max() for w, l in zip(want, ld))
    assert all(np.all(np.isfinite(g)) for g in got)
    assert max(np.abs(g - w).max() for g, w in zip(got, want)) <= TOL * scale + 2.0 * ref_err
    assert max(np.abs(g - l).max() for g, l in zip(got, ld)) <= TOL * scale + 2.0 * ref_err


@pytest.mark.parametrize("n,frac,shape,chunk_tiles,world", [
    (2600, 1.0, 0, 2, 1),       # every particle active: 11 tiles, 6 blocks of 512, chunks of 2 tiles
    (3000, 0.6, 1, 3, 2),       # 8 targets per thread, passive targets through the one-sided engine, two "ranks"
    (300, 0.5, 0, 1, 1),        # one block: diagonal tiles only
    (2100, 0.9, 3, 2, 3),       # an unfenced statement order, three "ranks"
])
def test_emulated_symmetric_velocity(emu, oracle, n, frac, shape, chunk_tiles, world):
    x, y, z, zeta, area, mask = _rand_sphere(n, 5, frac)
    R = 1.7
    x, y, z = R * x, R * y, R * z
    got = [np.full(n, np.nan) for _ in range(3)]
    assert emu.emu_sym_bve_velocity(n, x, y, z, zeta, area, mask, R, shape, chunk_tiles, world, *got) == 0
    want = oracle.bve_velocity(x, y, z, zeta, area, mask, R)
    ld = oracle.bve_velocity(x, y, z, zeta, area, mask, R, variant="_ld")
    _check(got, want, ld)


def test_emulated_default_velocity(emu, oracle):
    """The default one-sided kernel under the same emulator (a check of the emulator as much as of the kernel)."""
    n = 2000
    x, y, z, zeta, area, mask = _rand_sphere(n, 11, 0.7)
    got = [np.full(n, np.nan) for _ in range(3)]
    assert emu.emu_default_bve_velocity(n, x, y, z, zeta, area, mask, 1.0, *got) == 0
    want = oracle.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    ld = oracle.bve_velocity(x, y, z, zeta, area, mask, 1.0, variant="_ld")
    _check(got, want, ld)


@pytest.mark.parametrize("n,frac,R,shape,chunk_tiles,world,close", [
    (2500, 1.0, 6.371e6, 1, 3, 3, False),
    (2600, 1.0, 1.0, 0, 2, 1, True),        # arguments below the table window: the per-source library-log retry
    (2600, 1.0, 1.0, 2, 2, 1, True),        # ... and the per-batch one
])
def test_emulated_symmetric_stream(emu, oracle, n, frac, R, shape, chunk_tiles, world, close):
    x, y, z, zeta, area, mask = _rand_sphere(n, 5, frac)
    if close:       # nearly coincident active points, one pair in different blocks and one inside a block
        act = np.flatnonzero(mask)
        for a, b in ((act[3], act[-5]), (act[700], act[701])):
            p = np.array([x[a], y[a], z[a]]) + 1e-6 * np.array([0.3, -0.2, 0.5])
            p /= np.linalg.norm(p)
            x[b], y[b], z[b] = p
    x, y, z = R * x, R * y, R * z
    av = zeta + 0.3 * z / R
    got = [np.full(n, np.nan) for _ in range(2)]
    assert emu.emu_sym_bve_stream(n, x, y, z, zeta, av, area, mask, R, shape, chunk_tiles, world, *got) == 0
    want = oracle.bve_stream(x, y, z, zeta, av, area, mask, R)
    ld = oracle.bve_stream(x, y, z, zeta, av, area, mask, R, variant="_ld")
    _check(got, want, ld)


def _rand_plane(n, seed, frac):
    """Jittered lattice in [-7, 7]^2 (order shuffled), random mask and vorticity (as tests/test_parity_gpu.py)."""
    rng = np.random.default_rng(seed)
    m = int(np.ceil(np.sqrt(n)))
    gx, gy = np.meshgrid(np.arange(m), np.arange(m))
    p = np.stack([gx.ravel(), gy.ravel()]).astype(np.float64)[:, :n]
    h = 14.0 / m
    p = -7.0 + h * (p + 0.5) + 0.2 * h * rng.normal(size=(2, n))
    p = p[:, rng.permutation(n)]
    mask = (rng.random(n) < frac).astype(np.int32)
    area = np.where(mask != 0, h * h, 0.0)
    vort = rng.uniform(-1.0, 1.0, n)
    return p[0].copy(), p[1].copy(), vort, area, mask


@pytest.mark.parametrize("n,frac,shape,chunk_tiles,world", [
    (2500, 0.5, 0, 2, 1),       # quadRect-like: half of the particles active; ragged last block and tile
    (1500, 0.7, 3, 5, 3),
])
def test_emulated_symmetric_plane_velocity(emu, oracle, n, frac, shape, chunk_tiles, world):
    x, y, vort, area, mask = _rand_plane(n, 3, frac)
    got = [np.full(n, np.nan) for _ in range(2)]
    assert emu.emu_sym_plane_velocity(n, x, y, vort, area, mask, shape, chunk_tiles, world, *got) == 0
    want = oracle.plane_velocity(x, y, vort, area, mask)
    ld = oracle.plane_velocity(x, y, vort, area, mask, variant="_ld")
    _check(got, want, ld)
