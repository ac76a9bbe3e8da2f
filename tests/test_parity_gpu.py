"""GPU parity tests: every kernel of the hot path, through the C ABI, against
the CPU oracle on the same inputs.

Tolerance (BASELINE.json north_star): max_i |gpu - oracle| / max_i |oracle|
<= 1e-12 per output component -- the GPU sums the same FP64 terms in another
order (tiles, chunks) and factors the cross product, so bit equality is not
expected; index lists (active-source compaction, LoadBalance) are bit-exact.
"""
import json
import os

import numpy as np
import pytest

from lpm_v2_b200 import mesh as M, problems, solvers
from conftest import relerr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12
PI = problems.PI


def _rand_sphere(n, seed, active_frac=0.6, clustered=False):
    """Quasi-uniform points (jittered Fibonacci spiral, order shuffled) with a random
    mask and random vorticity.  `clustered=True` gives i.i.d. uniform points instead,
    which contain near-coincident pairs (an ill-conditioned input, see
    test_bve_near_coincident_points)."""
    rng = np.random.default_rng(seed)
    if clustered:
        p = rng.normal(size=(3, n))
    else:
        k = np.arange(n) + 0.5
        zc = 1.0 - 2.0 * k / n
        phi = k * PI * (3.0 - np.sqrt(5.0))
        rr = np.sqrt(1.0 - zc * zc)
        p = np.stack([rr * np.cos(phi), rr * np.sin(phi), zc])
        p = p + 0.15 * np.sqrt(4.0 * PI / n) * rng.normal(size=(3, n))     # jitter << spacing
        p = p[:, rng.permutation(n)]
    p /= np.linalg.norm(p, axis=0)
    mask = (rng.random(n) < active_frac).astype(np.int32)
    area = np.where(mask != 0, 4 * PI / max(mask.sum(), 1), 0.0)
    zeta = rng.uniform(-1.0, 1.0, n)
    return p[0].copy(), p[1].copy(), p[2].copy(), zeta, area, mask


# ---------------------------------------------------------------- BVE velocity
@pytest.mark.parametrize("seed,L", [(M.ICOS_TRI_SPHERE_SEED, 2), (M.ICOS_TRI_SPHERE_SEED, 4),
                                    (M.ICOS_TRI_SPHERE_SEED, 5), (M.CUBED_SPHERE_SEED, 5)])
@pytest.mark.parametrize("ic", ["gauss", "rh54"])
def test_bve_velocity_meshes(gpu, oracle, get_mesh, seed, L, ic):
    """Config 1 (icosTri L5 Gaussian vortex) and config 4 as shipped (cubed sphere L5 RH54)."""
    m = get_mesh(seed, L)
    zeta = problems.gaussian_vortex(m) if ic == "gauss" else problems.rossby_haurwitz54(m)
    got = gpu.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    want = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    for g, w in zip(got, want):
        assert relerr(g, w) <= TOL
    # per-target check too: every component of every target within 1e-12 of the field scale
    scale = max(np.abs(w).max() for w in want)
    assert max(np.abs(g - w).max() for g, w in zip(got, want)) <= TOL * scale


def test_bve_velocity_golden_fixture(gpu, get_mesh):
    g = np.load(os.path.join(HERE, "golden", "oracle_bve_icos2.npz"))
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 2)
    u, v, w = gpu.bve_velocity(m.x, m.y, m.z, g["zeta"], m.area, m.is_active, 1.0)
    assert max(relerr(u, g["u"]), relerr(v, g["v"]), relerr(w, g["w"])) <= TOL
    rs, as_ = gpu.bve_stream(m.x, m.y, m.z, g["zeta"], g["absvort"], m.area, m.is_active, 1.0)
    assert max(relerr(rs, g["relstream"]), relerr(as_, g["absstream"])) <= TOL


@pytest.mark.parametrize("n,frac,seed", [(1, 1.0, 1), (2, 1.0, 2), (3, 0.5, 3), (127, 0.3, 4), (513, 0.9, 5),
                                         (1025, 0.05, 6), (4099, 0.6, 12345), (20011, 0.55, 7)])
def test_bve_velocity_random_ragged(gpu, oracle, n, frac, seed):
    """Ragged sizes, sparse and dense masks, radius != 1; stress set of SURVEY 8(d)."""
    x, y, z, zeta, area, mask = _rand_sphere(n, seed, frac)
    R = 1.7
    x, y, z = R * x, R * y, R * z
    got = gpu.bve_velocity(x, y, z, zeta, area, mask, R)
    want = oracle.bve_velocity(x, y, z, zeta, area, mask, R)
    ld = oracle.bve_velocity(x, y, z, zeta, area, mask, R, variant="_ld")
    scale = max(max(np.abs(w).max() for w in want), 1e-300)
    assert all(np.all(np.isfinite(g)) for g in got)
    # Random-sign vorticity makes the sums cancel (|u| << sum |terms|), so the as-written
    # FP64 reference is itself ~1e-12 away from the exact sum here; allow the GPU the
    # reference's own rounding error on top of the 1e-12 budget.
    ref_err = max(np.abs(w - l).max() for w, l in zip(want, ld))
    assert max(np.abs(g - w).max() for g, w in zip(got, want)) <= TOL * scale + 2.0 * ref_err
    assert max(np.abs(g - l).max() for g, l in zip(got, ld)) <= TOL * scale + 2.0 * ref_err


def test_bve_near_coincident_points(gpu, oracle):
    """i.i.d. random points contain pairs with R^2 - x_i.x_j ~ 1e-8, where the reference
    expression itself loses ~8 digits to cancellation: its FP64 result is only good to
    ~1e-8 there.  The GPU must be at least as close to the extended-precision sum as the
    as-written FP64 reference is (it is closer: its denominator is an FMA chain)."""
    x, y, z, zeta, area, mask = _rand_sphere(3000, 77, 0.8, clustered=True)
    got = gpu.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    f64 = oracle.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    ld = oracle.bve_velocity(x, y, z, zeta, area, mask, 1.0, variant="_ld")
    assert all(np.all(np.isfinite(g)) for g in got)
    for g, f, l in zip(got, f64, ld):
        assert relerr(g, l) <= max(TOL, 2.0 * relerr(f, l))


def test_bve_velocity_no_active_sources(gpu, oracle):
    x, y, z, zeta, area, mask = _rand_sphere(300, 9, 0.0)
    mask[:] = 0
    got = gpu.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    assert all(np.all(g == 0.0) for g in got)
    mask[17] = 1                      # exactly one source: its own velocity is zero
    area[17] = 0.3
    got = gpu.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    want = oracle.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    assert got[0][17] == 0.0 and got[1][17] == 0.0 and got[2][17] == 0.0
    assert max(relerr(g, w) for g, w in zip(got, want)) <= TOL


def test_bve_all_active_self_exclusion(gpu, oracle):
    """Every target is also a source: the j == i exclusion must hit in every tile row."""
    x, y, z, zeta, area, mask = _rand_sphere(3000, 21, 1.0)
    got = gpu.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    want = oracle.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    assert all(np.all(np.isfinite(g)) for g in got)
    assert max(relerr(g, w) for g, w in zip(got, want)) <= TOL


@pytest.mark.parametrize("L", [2, 4, 5, 6])
def test_bve_kernel_shapes_deterministic(gpu, oracle, get_mesh, L):
    """The one-sided engine picks 1, 2, 4 or 8 targets per thread from the problem size (runtime.cuh,
    launch_auto<BveVel>): levels 2..6 cover the shapes; each meets the parity tolerance and is bit-identical
    from run to run."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
    zeta = problems.gaussian_vortex(m)
    want = oracle.bve_velocity_mt(8, m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0, 0, m.n, fast=False)
    got = gpu.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    again = gpu.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    for a, b, c in zip(got, want, again):
        assert relerr(a, b) <= TOL
        assert np.array_equal(a, c)          # run-to-run deterministic


@pytest.mark.parametrize("world", [2, 3, 7, 8])
def test_bve_slices_bitwise_equal_full_evaluation(gpu, get_mesh, world):
    """Rank-mode path: LoadBalance slices through the device-pointer API.  Slice starts
    are not multiples of the target-block size, yet every target gets bit-identical
    results to the 1-GPU evaluation (the reference's bit-reproducibility across ranks)."""
    import torch
    from lpm_v2_b200 import torch_api
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 5)
    zeta = problems.rossby_haurwitz54(m)
    full = gpu.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (m.x, m.y, m.z, zeta, m.area)]
    mask = torch.from_numpy(m.is_active).to(dev)
    out = [torch.full((m.n,), float("nan"), dtype=torch.float64, device=dev) for _ in range(3)]
    s, e, _ = gpu.load_balance(m.n, world)
    for r in range(world):
        torch_api.bve_velocity_dev(*t, mask, 1.0, int(s[r]) - 1, int(e[r]), *out)
    torch.cuda.synchronize()
    for o, f in zip(out, full):
        assert np.array_equal(o.cpu().numpy(), f)


# ---------------------------------------------------------------- index lists
@pytest.mark.parametrize("n,frac", [(1, 1.0), (1023, 0.5), (1024, 0.5), (1025, 0.01), (100003, 0.66), (2500000, 0.667)])
def test_active_list_bit_exact(gpu, oracle, n, frac):
    """Device compaction == Fortran pack([(j,j=1,n)], mask), bit for bit."""
    rng = np.random.default_rng(n)
    mask = (rng.random(n) < frac).astype(np.int32)
    got = gpu.active_list(mask)
    want = oracle.active_list(mask)
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_active_list_mesh(gpu, oracle, get_mesh):
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 5)
    got = gpu.active_list(m.is_active)
    assert np.array_equal(got, oracle.active_list(m.is_active))
    assert got.size == 20 * 4 ** 5


# ---------------------------------------------------------------- stream functions
def test_bve_stream(gpu, oracle, get_mesh):
    for L in (3, 5):
        m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
        zeta = problems.rossby_haurwitz54(m)
        absv = problems.abs_vorticity(m, zeta, 2 * PI)
        got = gpu.bve_stream(m.x, m.y, m.z, zeta, absv, m.area, m.is_active, 1.0)
        want = oracle.bve_stream(m.x, m.y, m.z, zeta, absv, m.area, m.is_active, 1.0)
        assert max(relerr(g, w) for g, w in zip(got, want)) <= TOL


@pytest.mark.parametrize("radius", [1.0, 6371.22e3, 3.0e-7])
def test_bve_stream_any_radius(gpu, oracle, get_mesh, radius):
    """The log table window (pairs.cuh) follows 2 R^2: Earth radius in metres and a tiny
    sphere must be as accurate as the unit sphere."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 4)
    zeta = problems.rossby_haurwitz54(m)
    absv = problems.abs_vorticity(m, zeta, 2 * PI)
    x, y, z, area = m.x * radius, m.y * radius, m.z * radius, m.area * radius * radius
    got = gpu.bve_stream(x, y, z, zeta, absv, area, m.is_active, radius)
    want = oracle.bve_stream(x, y, z, zeta, absv, area, m.is_active, radius)
    assert max(relerr(g, w) for g, w in zip(got, want)) <= TOL


def test_stream_arguments_outside_the_log_window(gpu, oracle):
    """Pairs 1e-9 apart next to pairs 1e7 apart: r^2 spans 106 binades, far more than the
    32-binade shared-memory window, so part of the sum takes the library log(); coincident
    particles give log(0) = -inf as in the reference."""
    rng = np.random.default_rng(5)
    n = 1500
    x = np.concatenate([rng.uniform(-1e7, 1e7, n // 3), rng.uniform(-1, 1, n // 3), 0.25 + rng.uniform(-1e-9, 1e-9, n - 2 * (n // 3))])
    y = np.concatenate([rng.uniform(-1e7, 1e7, n // 3), rng.uniform(-1, 1, n // 3), -0.5 + rng.uniform(-1e-9, 1e-9, n - 2 * (n // 3))])
    vort = rng.uniform(-1, 1, n)
    area = rng.uniform(0.5, 1.5, n)
    mask = (rng.uniform(size=n) < 0.7).astype(np.int32)
    got = gpu.plane_stream(x, y, vort, area, mask)
    want = oracle.plane_stream(x, y, vort, area, mask)
    assert np.all(np.isfinite(got))
    assert relerr(got, want) <= TOL
    # two coincident active particles: the reference adds log(0) * w = -+inf to both
    x[10], y[10] = x[11], y[11]
    mask[10] = mask[11] = 1
    vort[10] = vort[11] = 0.5
    got = gpu.plane_stream(x, y, vort, area, mask)
    want = oracle.plane_stream(x, y, vort, area, mask)
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    assert np.isneginf(got[10]) and np.isneginf(got[11]) and np.isneginf(want[10])
    ok = np.isfinite(want)
    assert relerr(got[ok], want[ok]) <= TOL


# ---------------------------------------------------------------- planar
@pytest.mark.parametrize("L", [2, 4, 5])
def test_plane_velocity_and_stream(gpu, oracle, get_mesh, L):
    """Config 2: quadRect, meshRadius 7, two Lamb dipoles (L5 is the shipped namelist)."""
    m = get_mesh(M.QUAD_RECT_SEED, L, 7.0)
    vort = problems.colliding_dipoles(m)
    got = gpu.plane_velocity(m.x, m.y, vort, m.area, m.is_active)
    want = oracle.plane_velocity(m.x, m.y, vort, m.area, m.is_active)
    assert max(relerr(g, w) for g, w in zip(got, want)) <= TOL
    psi = gpu.plane_stream(m.x, m.y, vort, m.area, m.is_active)
    assert relerr(psi, oracle.plane_stream(m.x, m.y, vort, m.area, m.is_active)) <= TOL


def test_plane_golden_and_trihex(gpu, oracle, get_mesh):
    g = np.load(os.path.join(HERE, "golden", "oracle_plane_quad3.npz"))
    m = get_mesh(M.QUAD_RECT_SEED, 3, 7.0)
    u, v = gpu.plane_velocity(m.x, m.y, g["vort"], m.area, m.is_active)
    assert max(relerr(u, g["u"]), relerr(v, g["v"])) <= TOL
    h = get_mesh(M.TRI_HEX_SEED, 4)
    vort = np.exp(-4 * (h.x ** 2 + h.y ** 2))
    got = gpu.plane_velocity(h.x, h.y, vort, h.area, h.is_active)
    want = oracle.plane_velocity(h.x, h.y, vort, h.area, h.is_active)
    assert max(relerr(a, b) for a, b in zip(got, want)) <= TOL


# ---------------------------------------------------------------- beta plane
@pytest.mark.parametrize("L", [2, 3, 5])
def test_betaplane_velocity(gpu, oracle, get_mesh, L):
    """The reference expression cosh(2 pi dy) - cos(2 pi dx) cancels for near
    pairs, so the as-written FP64 sum is itself only accurate to ~1e-12..1e-11
    (test_oracle_golden.py::test_oracle_vs_extended_precision).  The GPU kernel
    evaluates the same denominator as 2 sinh^2 + 2 sin^2: it must match the
    extended-precision evaluation to 1e-12 and be no further from the as-written
    oracle than that oracle is from the exact sum."""
    m = get_mesh(M.BETA_PLANE_SEED, L)
    zeta = problems.betaplane_gaussian(m)
    got = gpu.betaplane_velocity(m.x, m.y, zeta, m.area, m.is_active)
    ld = oracle.betaplane_velocity(m.x, m.y, zeta, m.area, m.is_active, variant="_ld")
    f64 = oracle.betaplane_velocity(m.x, m.y, zeta, m.area, m.is_active)
    for g, l, f in zip(got, ld, f64):
        assert relerr(g, l) <= TOL
        assert relerr(g, f) <= max(TOL, 2.0 * relerr(f, l) + 1e-14)


def test_betaplane_stream(gpu, oracle, get_mesh):
    m = get_mesh(M.BETA_PLANE_SEED, 4)
    zeta = problems.betaplane_gaussian(m)
    absv = zeta + 1.0 + 2.0 * m.y
    got = gpu.betaplane_stream(m.x, m.y, zeta, absv, m.area, m.is_active)
    ld = oracle.betaplane_stream(m.x, m.y, zeta, absv, m.area, m.is_active, variant="_ld")
    f64 = oracle.betaplane_stream(m.x, m.y, zeta, absv, m.area, m.is_active)
    for g, l, f in zip(got, ld, f64):
        assert relerr(g, l) <= TOL
        assert relerr(g, f) <= 1e-11


# ---------------------------------------------------------------- PSE
def test_pse_sphere_reference_thresholds_on_gpu(gpu, oracle, get_mesh):
    """Config 3: the reference's own regression bounds (SpherePSEConvTest.f90:373-390),
    evaluated on the GPU output."""
    th = json.load(open(os.path.join(HERE, "golden", "reference_thresholds.json")))
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 3)
    pse = solvers.PSE(m, th["mesh"]["pse_power"])
    harm = problems.spherical_harmonic54(m)
    exact = -30.0 * harm
    lap = pse.SphereLaplacianAtParticles(m, harm)
    err = np.abs(lap - exact)
    assert err.max() / np.abs(exact).max() <= th["particlesLinfHarmLap_max"]
    s, e, _ = oracle.load_balance(m.n, 4)
    sl = slice(int(s[0]) - 1, int(e[0]))
    a = m.is_active[sl] != 0
    assert np.sum(err[sl][a] ** 2 * m.area[sl][a]) / np.sum(exact[sl][a] ** 2 * m.area[sl][a]) <= th["particlesL2HarmLap_rank0_max"]
    lap_c = pse.SphereLaplacianAtParticles(m, np.full(m.n, th["const_value"]))
    assert np.abs(lap_c).max() <= th["particlesLinfConstLap_max"]
    want = oracle.pse_laplacian_sphere(m.x, m.y, m.z, harm, m.area, m.is_active, pse.eps, 1.0)
    assert relerr(lap, want) <= TOL


@pytest.mark.parametrize("L,power", [(3, 0.75), (4, 0.6), (5, 1.5), (5, 0.75)])
def test_pse_sphere(gpu, oracle, get_mesh, L, power):
    """power 1.5 makes eps small enough that the far-field cut-off (k > 8) is active."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
    f = problems.spherical_harmonic54(m)
    eps = m.max_edge_length ** power
    got = gpu.pse_laplacian_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps, 1.0)
    want = oracle.pse_laplacian_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps, 1.0)
    assert relerr(got, want) <= TOL


def test_pse_plane(gpu, oracle, get_mesh):
    for L, power in [(3, 0.75), (5, 0.75), (5, 1.2)]:
        m = get_mesh(M.QUAD_RECT_SEED, L, 2.0)
        f = np.sin(1.3 * m.x) * np.cos(0.7 * m.y) + 0.1 * m.x * m.y
        eps = m.max_edge_length ** power
        got = gpu.pse_laplacian_plane(m.x, m.y, f, m.area, m.is_active, eps)
        want = oracle.pse_laplacian_plane(m.x, m.y, f, m.area, m.is_active, eps)
        assert relerr(got, want) <= TOL


# ---------------------------------------------------------------- RK4 steps (resident solvers)
def test_bve_rk4_steps(gpu, oracle, get_mesh):
    """Config 1 time loop (dt = 0.01) for 3 steps at L3: New / Timestep / Delete."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 3)
    omega = 2 * PI
    zeta = problems.gaussian_vortex(m)
    sph = solvers.BVEMesh(m, zeta, 1.0, omega)
    sph.SetVelocityOnMesh()
    ref = [m.x.copy(), m.y.copy(), m.z.copy(), zeta.copy()] + list(
        oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0))
    sol = solvers.BVESolver(sph)
    for step in range(3):
        sol.Timestep(sph, 0.01, with_stream=True)
        ref = oracle.bve_rk4_step(*ref, m.area, m.is_active, 1.0, omega, 0.01)
        got = [sph.x, sph.y, sph.z, sph.relVort] + sph.velocity
        for name, a, b in zip("x y z zeta u v w".split(), got, ref):
            assert relerr(a, b) <= TOL, (step, name)
        rs, as_ = oracle.bve_stream(ref[0], ref[1], ref[2], ref[3], sph.absVort, m.area, m.is_active, 1.0)
        assert relerr(sph.relStream, rs) <= TOL and relerr(sph.absStream, as_) <= TOL
    ke, en = sol.Diagnostics()
    assert abs(ke - oracle.total_ke(*sph.velocity, m.area, m.is_active)) <= 1e-12 * ke
    assert abs(en - oracle.total_enstrophy(sph.relVort, m.area, m.is_active)) <= 1e-12 * en
    sol.Delete()


@pytest.mark.parametrize("symmetric", [False, True])
def test_bve_step_end_fused_equals_separate_sums(gpu, oracle, get_mesh, symmetric):
    """The resident BVE step ends with ONE pass that yields the velocity and the stream functions of the new state
    (src/SphereBVESolver.f90:345-352; BveVelStream / SymBveVelStream).  With the fusion switched off the same step
    ends with the two separate sums: both must meet parity, and they agree with each other to rounding."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 4)
    omega = 2 * PI
    zeta = problems.rossby_haurwitz54(m)
    u0 = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    ref = oracle.bve_rk4_step(m.x, m.y, m.z, zeta, *u0, m.area, m.is_active, 1.0, omega, 0.002)
    res = {}
    gpu.set_symmetric(symmetric)
    gpu.tune("sym_min_sources", 0 if symmetric else 200000)
    try:
        for fused in (1, 0):
            gpu.tune("fuse_step_end", fused)
            sph = solvers.BVEMesh(m, zeta, 1.0, omega)
            sph.velocity = [a.copy() for a in u0]
            sol = solvers.BVESolver(sph)
            sol.Timestep(sph, 0.002, with_stream=True)
            sol.Delete()
            got = [sph.x, sph.y, sph.z, sph.relVort] + sph.velocity
            for name, a, b in zip("x y z zeta u v w".split(), got, ref):
                assert relerr(a, b) <= TOL, (fused, name)
            rs, as_ = oracle.bve_stream(ref[0], ref[1], ref[2], ref[3], sph.absVort, m.area, m.is_active, 1.0)
            assert relerr(sph.relStream, rs) <= TOL and relerr(sph.absStream, as_) <= TOL
            res[fused] = sph.velocity + [sph.relStream, sph.absStream]
    finally:
        gpu.tune("fuse_step_end", 1)
        gpu.set_symmetric(True)
        gpu.tune("sym_min_sources", 200000)
    for a, b in zip(res[1], res[0]):
        assert relerr(a, b) <= 1e-13


def test_plane_rk4_steps(gpu, oracle, get_mesh):
    m = get_mesh(M.QUAD_RECT_SEED, 3, 7.0)
    vort = problems.colliding_dipoles(m)
    pl = solvers.PlaneMeshIncompressible(m, vort)
    pl.SetVelocityOnMesh()
    ref = [m.x.copy(), m.y.copy()] + list(oracle.plane_velocity(m.x, m.y, vort, m.area, m.is_active))
    sol = solvers.PlaneSolver(pl)
    for step in range(2):
        sol.Timestep(pl, 0.01, with_stream=True)
        x, y, u, v = oracle.plane_rk4_step(ref[0], ref[1], vort, ref[2], ref[3], m.area, m.is_active, 0.01)
        ref = [x, y, u, v]
        for name, a, b in zip("x y u v".split(), [pl.x, pl.y] + pl.velocity, ref):
            assert relerr(a, b) <= TOL, (step, name)
        assert relerr(pl.streamFn, oracle.plane_stream(x, y, vort, m.area, m.is_active)) <= TOL
    sol.Delete()


def test_betaplane_rk4_step(gpu, oracle, get_mesh):
    m = get_mesh(M.BETA_PLANE_SEED, 3)
    zeta = problems.betaplane_gaussian(m)
    bp = solvers.BetaPlaneMesh(m, zeta, f0=0.0, beta=4 * PI)
    bp.SetVelocityOnMesh()
    u0, v0 = [a.copy() for a in bp.velocity]
    sol = solvers.BetaPlaneSolver(bp)
    sol.Timestep(bp, 0.05, with_stream=True)
    sol.Delete()
    x, y, z1, u, v = oracle.betaplane_rk4_step(m.x, m.y, zeta, u0, v0, m.area, m.is_active, 4 * PI, 0.05)
    for name, a, b in zip("x y zeta u v".split(), [bp.x, bp.y, bp.relVort] + bp.velocity, [x, y, z1, u, v]):
        assert relerr(a, b) <= 1e-11, name          # beta-plane budget, see test_betaplane_velocity


# ---------------------------------------------------------------- error behaviour
def test_solver_handles_and_finalize(gpu, get_mesh):
    """lpm_gpu_finalize with a solver still alive frees it; afterwards Delete of that handle is a no-op and any
    other use is an error (not a dereference of freed device state), and the library can be initialised again."""
    from lpm_v2_b200 import LpmError
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 2)
    sph = solvers.BVEMesh(m, problems.gaussian_vortex(m), 1.0, 2 * PI)
    sph.SetVelocityOnMesh()
    sol = solvers.BVESolver(sph)
    sol.Timestep(sph, 0.01, with_stream=True)
    try:
        gpu.finalize()
        with pytest.raises(LpmError) as ei:
            sol.Timestep(sph, 0.01, with_stream=True)
        assert ei.value.code == 1 and "stale solver handle" in str(ei.value)
        sol.Delete()                                    # no-op, no crash
        sol.Delete()
    finally:
        gpu.init(1)
    sol2 = solvers.BVESolver(sph)                       # a fresh runtime works as before
    sol2.Timestep(sph, 0.01, with_stream=False)
    sol2.Delete()
    sol2.Delete()                                       # deleting twice is harmless too


def test_invalid_arguments_are_reported_not_fatal(gpu):
    from lpm_v2_b200 import LpmError
    x = np.ones(4)
    m = np.ones(4, np.int32)
    with pytest.raises(LpmError) as ei:
        gpu.pse_laplacian_sphere(x, x, x, x, x, m, 0.0, 1.0)
    assert ei.value.code == 1
    # and the library keeps working afterwards
    u, v, w = gpu.bve_velocity(x, 2 * x, 3 * x, x, x, np.zeros(4, np.int32), 1.0)
    assert np.all(u == 0)
