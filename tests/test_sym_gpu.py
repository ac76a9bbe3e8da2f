"""GPU parity of the pair-symmetric BVE sums (lpm_v2_b200/csrc/symmetric.cuh): the default path of whole
velocity / stream-function evaluations with >= 200 000 active particles.  The tests lower that threshold
(csrc/lpm_gpu_tuning.h, "sym_min_sources") so that sizes the oracle finishes in seconds take the same code,
and check: parity with the oracle (meshes, ragged random sets, extreme radii), agreement with the one-sided
engine, bit-identical results from run to run (fixed-point accumulation), NaN on coincident particles, and
the resident RK4 step.  Reference loops: src/SphereBVESolver.f90:396-420, src/SphereBVE.f90:454-475.
"""
import numpy as np
import pytest

from lpm_v2_b200 import mesh as M, problems, solvers
from conftest import relerr
from test_parity_gpu import _rand_sphere

pytestmark = [pytest.mark.gpu]
TOL = 1e-12


def _check(got, want, ld):
    """As tests/test_parity_gpu.py on random sets: 1e-12 of the field scale plus the as-written FP64
    reference's own distance from the extended-precision sum (random-sign weights cancel)."""
    scale = max(max(np.abs(w).max() for w in want), 1e-300)
    ref_err = max(np.abs(w - l).max() for w, l in zip(want, ld))
    assert all(np.all(np.isfinite(g)) for g in got)
    assert max(np.abs(g - w).max() for g, w in zip(got, want)) <= TOL * scale + 2.0 * ref_err
    assert max(np.abs(g - l).max() for g, l in zip(got, ld)) <= TOL * scale + 2.0 * ref_err


@pytest.fixture
def sym(gpu):
    """Every whole BVE evaluation takes the symmetric path, whatever its size."""
    gpu.set_symmetric(True)
    gpu.tune("sym_min_sources", 0)
    yield gpu
    gpu.tune("sym_min_sources", 200000)


@pytest.mark.parametrize("seed,L", [(M.ICOS_TRI_SPHERE_SEED, 2), (M.ICOS_TRI_SPHERE_SEED, 5), (M.CUBED_SPHERE_SEED, 5)])
def test_sym_bve_velocity_meshes(sym, oracle, get_mesh, seed, L):
    m = get_mesh(seed, L)
    zeta = problems.rossby_haurwitz54(m)
    got = sym.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    want = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    scale = max(np.abs(w).max() for w in want)
    assert max(np.abs(g - w).max() for g, w in zip(got, want)) <= TOL * scale


@pytest.mark.parametrize("n,frac,seed", [(1, 1.0, 1), (2, 1.0, 2), (3, 0.5, 3), (127, 0.3, 4), (513, 0.9, 5),
                                         (1025, 0.05, 6), (4099, 0.6, 12345), (20011, 0.55, 7), (20011, 1.0, 8),
                                         (3000, 0.0, 9)])
def test_sym_bve_velocity_random_ragged(sym, oracle, n, frac, seed):
    """Ragged sizes (active counts that are not multiples of a tile or a block), all-active and
    all-passive masks, radius != 1."""
    x, y, z, zeta, area, mask = _rand_sphere(n, seed, frac)
    R = 1.7
    x, y, z = R * x, R * y, R * z
    got = sym.bve_velocity(x, y, z, zeta, area, mask, R)
    want = oracle.bve_velocity(x, y, z, zeta, area, mask, R)
    ld = oracle.bve_velocity(x, y, z, zeta, area, mask, R, variant="_ld")
    _check(got, want, ld)


@pytest.mark.parametrize("seed,L", [(M.ICOS_TRI_SPHERE_SEED, 2), (M.ICOS_TRI_SPHERE_SEED, 5), (M.CUBED_SPHERE_SEED, 5)])
def test_sym_bve_stream_meshes(sym, oracle, get_mesh, seed, L):
    m = get_mesh(seed, L)
    zeta = problems.rossby_haurwitz54(m)
    av = problems.abs_vorticity(m, zeta, 2.0 * np.pi)
    got = sym.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
    want = oracle.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
    for g, w in zip(got, want):
        assert relerr(g, w) <= TOL


@pytest.mark.parametrize("n,frac,seed,R", [(3, 0.5, 3, 1.0), (513, 0.9, 5, 1.0), (4099, 0.6, 12345, 1.7),
                                           (20011, 0.55, 7, 6.371e6), (6000, 1.0, 8, 3.0e-7), (3000, 0.0, 9, 1.0)])
def test_sym_bve_stream_random_ragged(sym, oracle, n, frac, seed, R):
    """Ragged sizes and radii whose arguments fall in very different binades of the log table (and of the
    fixed-point window)."""
    x, y, z, zeta, area, mask = _rand_sphere(n, seed, frac)
    x, y, z = R * x, R * y, R * z
    av = zeta + 0.3 * z / R
    got = sym.bve_stream(x, y, z, zeta, av, area, mask, R)
    want = oracle.bve_stream(x, y, z, zeta, av, area, mask, R)
    ld = oracle.bve_stream(x, y, z, zeta, av, area, mask, R, variant="_ld")
    _check(got, want, ld)


def test_sym_matches_one_sided_path(sym, get_mesh):
    """Same sum, other order: within a few ulp of the one-sided engine at icosTri 6."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 6)
    zeta = problems.gaussian_vortex(m)
    av = problems.abs_vorticity(m, zeta, 2.0 * np.pi)
    a = sym.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    sa = sym.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
    sym.set_symmetric(False)
    try:
        b = sym.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
        sb = sym.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
    finally:
        sym.set_symmetric(True)
    for g, w in zip(a + sa, b + sb):
        assert relerr(g, w) <= 1e-13


@pytest.mark.parametrize("panel,chunk", [(3, 1), (2, 2), (1, 16), (256, 16)])
def test_sym_launch_geometry_does_not_change_the_bits(sym, oracle, panel, chunk):
    """Panels of target blocks (ragged last panel) and the chunk length of the triangle kernel only reorder exact
    fixed-point additions ... of per-(block, tile) values.  The chunk length also regroups each CTA's own running
    sums, so bits are compared per chunk length; every geometry must meet parity."""
    x, y, z, zeta, area, mask = _rand_sphere(7001, 31, 0.8)          # 5601 active: 6 blocks of 1024, 22 tiles
    av = zeta + 0.3 * z
    want = oracle.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    ld = oracle.bve_velocity(x, y, z, zeta, area, mask, 1.0, variant="_ld")
    sym.tune("sym_chunk_tiles", chunk)
    try:
        sym.tune("sym_panel_blocks", 256)
        base = sym.bve_velocity(x, y, z, zeta, area, mask, 1.0) + sym.bve_stream(x, y, z, zeta, av, area, mask, 1.0)
        sym.tune("sym_panel_blocks", panel)
        got = sym.bve_velocity(x, y, z, zeta, area, mask, 1.0) + sym.bve_stream(x, y, z, zeta, av, area, mask, 1.0)
    finally:
        sym.tune("sym_panel_blocks", 256)
        sym.tune("sym_chunk_tiles", 16)
    _check(got[:3], want, ld)
    assert all(np.array_equal(a, b) for a, b in zip(got, base))


def test_sym_rk4_step(sym, oracle, get_mesh):
    """The resident solver takes the symmetric path for its four velocity sums and the stream functions."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 4)
    zeta = problems.gaussian_vortex(m)
    omega = 2.0 * np.pi
    u, v, w = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    sph = solvers.BVEMesh(m, zeta, 1.0, omega)
    sph.velocity = [u.copy(), v.copy(), w.copy()]
    sol = solvers.BVESolver(sph)
    sol.Timestep(sph, 0.01, with_stream=True)
    sol.Delete()
    ref = oracle.bve_rk4_step(m.x, m.y, m.z, zeta, u, v, w, m.area, m.is_active, 1.0, omega, 0.01)
    got = [sph.x, sph.y, sph.z, sph.relVort] + sph.velocity
    for a, b in zip(got, ref):
        assert relerr(a, b) <= TOL


@pytest.mark.parametrize("R", [1.0, 6.371e6, 3.0e-7])
def test_sym_fixed_point_accumulation(sym, oracle, R):
    """Order-independent (fixed-point) accumulation: parity at radii that put the window in very different
    places, and bit-identical results from run to run whatever order the CTAs finish in."""
    x, y, z, zeta, area, mask = _rand_sphere(1700, 5, 0.7)
    x, y, z = R * x, R * y, R * z
    av = zeta + 0.3 * z / R
    got = sym.bve_velocity(x, y, z, zeta, area, mask, R)
    _check(got, oracle.bve_velocity(x, y, z, zeta, area, mask, R), oracle.bve_velocity(x, y, z, zeta, area, mask, R, variant="_ld"))
    gots = sym.bve_stream(x, y, z, zeta, av, area, mask, R)
    _check(gots, oracle.bve_stream(x, y, z, zeta, av, area, mask, R), oracle.bve_stream(x, y, z, zeta, av, area, mask, R, variant="_ld"))
    for _ in range(2):
        again = sym.bve_velocity(x, y, z, zeta, area, mask, R) + sym.bve_stream(x, y, z, zeta, av, area, mask, R)
        assert all(np.array_equal(a, b) for a, b in zip(again, got + gots))


def test_sym_bitwise_reproducible_large(sym, get_mesh):
    """icosTri 6 (81 920 active particles, 80 target blocks x 10 chunks racing on the accumulators): three runs, same bits."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 6)
    zeta = problems.rossby_haurwitz54(m)
    first = sym.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    for _ in range(2):
        again = sym.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
        assert all(np.array_equal(a, b) for a, b in zip(again, first))


@pytest.mark.parametrize("symmetric", [True, False])
@pytest.mark.parametrize("n,other", [(900, 400), (3000, 2500)])
def test_coincident_particles_fail_loudly(sym, oracle, symmetric, n, other):
    """Two coincident active particles: the reference's strength is -zeta A / (4 pi R * 0) = Inf and the cross
    product 0, so both targets get NaN (src/SphereBVESolver.f90:403-407) and every other target is untouched.
    One-sided engine: the same -- a tile whose sums come out non-finite is re-run with the reciprocals taken one by
    one (rcp_group), so the NaN stays with its two targets.  Pair-symmetric path: the fixed-point accumulators
    cannot hold Inf or NaN, so a value above their window raises the accumulator's overflow counter and it reads as
    NaN; the block's own (diagonal) tiles also take reciprocals one by one (n = 900: one block), but above the
    diagonal the reciprocals of the 8 targets a thread holds share one MUFU, so the NaN may also reach those
    thread-mates (n = 3000, particles 17 and 2500 in different blocks: at most 7 more per coincident particle;
    documented in INTEGRATION.md).  Every other target must be finite and correct: the failure is loud and local."""
    x, y, z, zeta, area, mask = _rand_sphere(n, 21, 1.0)
    x[17], y[17], z[17] = 1.0, 0.0, 0.0
    x[other], y[other], z[other] = 1.0, 0.0, 0.0
    sym.set_symmetric(symmetric)
    try:
        got = sym.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    finally:
        sym.set_symmetric(True)
    with np.errstate(all="ignore"):
        want = oracle.bve_velocity(x, y, z, zeta, area, mask, 1.0)
    bad = np.zeros(n, bool)
    bad[[17, other]] = True
    assert not np.any(np.isfinite(want[1][bad])) and not np.any(np.isfinite(want[2][bad]))
    nonfinite = ~(np.isfinite(got[0]) & np.isfinite(got[1]) & np.isfinite(got[2]))
    assert np.all(nonfinite[bad])
    if not symmetric or n == 900:
        assert nonfinite.sum() == 2
    else:
        assert nonfinite.sum() <= 2 + 14
    ok = ~nonfinite
    for g, w in zip(got, want):
        assert np.abs(g[ok] - w[ok]).max() <= 1e-11 * np.abs(w[ok]).max()


def test_coincident_particles_planar_kernel(gpu, oracle):
    """The planar velocity kernel (one-sided engine): r = 0 gives 0 * Inf = NaN at the two coincident targets in the
    reference (src/PlaneIncompressibleSolver.f90:294-307) and here, and nowhere else."""
    rng = np.random.default_rng(5)
    n = 1500
    x, y = rng.uniform(0.05, 0.95, n), rng.uniform(-0.4, 0.4, n)
    q, area, mask = rng.uniform(-1, 1, n), np.full(n, 1.0 / n), np.ones(n, np.int32)
    x[700], y[700] = x[33], y[33]
    with np.errstate(all="ignore"):
        got, want = gpu.plane_velocity(x, y, q, area, mask), oracle.plane_velocity(x, y, q, area, mask)
    bad = np.zeros(n, bool)
    bad[[33, 700]] = True
    for g, w in zip(got, want):
        assert not np.any(np.isfinite(w[bad])) and not np.any(np.isfinite(g[bad]))
        assert np.all(np.isfinite(g[~bad]))
        assert np.abs(g[~bad] - w[~bad]).max() <= 1e-10 * np.abs(w[~bad]).max()
