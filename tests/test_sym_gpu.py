"""GPU parity of the EXPERIMENTAL paths: the pair-symmetric sums (lpm_set_bve_variant(200 .. 209),
lpm_v2_b200/csrc/symmetric.cuh: BVE / planar / beta-plane velocity and stream functions; 204, 205 with fixed-point
accumulation, 206, 207 with the warps' sums combined in shared memory) and the fenced one-sided kernels (44, 45, 103).

They were written at the end of round 1 without GPU time left, so they have not run on a GPU yet: these tests
are skipped unless LPM_EXPERIMENTAL=1, and the default path does not depend on them.  On the SIMT emulator
(tests/test_emu_abi.py) the small cases pass.  First thing to run on a B200 in the next round:

    LPM_EXPERIMENTAL=1 python -m pytest tests/test_sym_gpu.py -m gpu -x -q
    python tools/ab_sym.py 7            # default vs symmetric timings, one box
"""
import os

import numpy as np
import pytest

from lpm_v2_b200 import mesh as M, problems, solvers
from conftest import relerr
from test_parity_gpu import _rand_sphere

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("LPM_EXPERIMENTAL") != "1",
                                 reason="experimental symmetric path: set LPM_EXPERIMENTAL=1")]
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
def sym(gpu, request):
    gpu.set_bve_variant(request.param)
    yield gpu
    gpu.set_bve_variant(0)


@pytest.mark.parametrize("sym", [200, 201, 202, 203], indirect=True)
@pytest.mark.parametrize("seed,L", [(M.ICOS_TRI_SPHERE_SEED, 2), (M.ICOS_TRI_SPHERE_SEED, 5), (M.CUBED_SPHERE_SEED, 5)])
def test_sym_bve_velocity_meshes(sym, oracle, get_mesh, seed, L):
    m = get_mesh(seed, L)
    zeta = problems.rossby_haurwitz54(m)
    got = sym.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    want = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    scale = max(np.abs(w).max() for w in want)
    assert max(np.abs(g - w).max() for g, w in zip(got, want)) <= TOL * scale


@pytest.mark.parametrize("sym", [200, 201, 202, 203, 206, 207], indirect=True)
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


@pytest.mark.parametrize("sym", [200, 201, 202, 203], indirect=True)
@pytest.mark.parametrize("seed,L", [(M.ICOS_TRI_SPHERE_SEED, 2), (M.ICOS_TRI_SPHERE_SEED, 5), (M.CUBED_SPHERE_SEED, 5)])
def test_sym_bve_stream_meshes(sym, oracle, get_mesh, seed, L):
    m = get_mesh(seed, L)
    zeta = problems.rossby_haurwitz54(m)
    av = problems.abs_vorticity(m, zeta, 2.0 * np.pi)
    got = sym.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
    want = oracle.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
    for g, w in zip(got, want):
        assert relerr(g, w) <= TOL


@pytest.mark.parametrize("sym", [200, 201, 202, 203], indirect=True)
@pytest.mark.parametrize("n,frac,seed,R", [(3, 0.5, 3, 1.0), (513, 0.9, 5, 1.0), (4099, 0.6, 12345, 1.7),
                                           (20011, 0.55, 7, 6.371e6), (6000, 1.0, 8, 3.0e-7), (3000, 0.0, 9, 1.0)])
def test_sym_bve_stream_random_ragged(sym, oracle, n, frac, seed, R):
    """Ragged sizes and radii whose arguments fall in very different binades of the log table."""
    x, y, z, zeta, area, mask = _rand_sphere(n, seed, frac)
    x, y, z = R * x, R * y, R * z
    av = zeta + 0.3 * z / R
    got = sym.bve_stream(x, y, z, zeta, av, area, mask, R)
    want = oracle.bve_stream(x, y, z, zeta, av, area, mask, R)
    ld = oracle.bve_stream(x, y, z, zeta, av, area, mask, R, variant="_ld")
    _check(got, want, ld)


@pytest.mark.parametrize("sym", [200, 201, 202, 203], indirect=True)
@pytest.mark.parametrize("L", [3, 5])
def test_sym_plane_velocity_mesh(sym, oracle, get_mesh, L):
    """Config 2 (colliding dipoles on quadRect): the planar Biot-Savart sum through the symmetric path."""
    q = get_mesh(M.QUAD_RECT_SEED, L, 7.0)
    vort = problems.colliding_dipoles(q)
    got = sym.plane_velocity(q.x, q.y, vort, q.area, q.is_active)
    want = oracle.plane_velocity(q.x, q.y, vort, q.area, q.is_active)
    ld = oracle.plane_velocity(q.x, q.y, vort, q.area, q.is_active, variant="_ld")
    _check(got, want, ld)


@pytest.mark.parametrize("sym", [200, 201], indirect=True)
@pytest.mark.parametrize("n,frac,seed", [(2, 1.0, 2), (513, 0.9, 5), (4099, 0.6, 12345), (20011, 0.5, 7), (3000, 0.0, 9)])
def test_sym_plane_velocity_random_ragged(sym, oracle, n, frac, seed):
    from test_cuda_emu import _rand_plane
    x, y, vort, area, mask = _rand_plane(n, seed, frac)
    got = sym.plane_velocity(x, y, vort, area, mask)
    want = oracle.plane_velocity(x, y, vort, area, mask)
    ld = oracle.plane_velocity(x, y, vort, area, mask, variant="_ld")
    _check(got, want, ld)


@pytest.mark.parametrize("sym", [200, 201, 202, 203], indirect=True)
@pytest.mark.parametrize("L", [2, 3, 4])
def test_sym_betaplane_velocity(sym, oracle, get_mesh, L):
    """As tests/test_parity_gpu.py::test_betaplane_velocity, through the symmetric path (ragged sizes: the beta-plane
    meshes have 3/4 of their particles active and no power-of-two counts)."""
    m = get_mesh(M.BETA_PLANE_SEED, L)
    zeta = problems.betaplane_gaussian(m)
    got = sym.betaplane_velocity(m.x, m.y, zeta, m.area, m.is_active)
    ld = oracle.betaplane_velocity(m.x, m.y, zeta, m.area, m.is_active, variant="_ld")
    f64 = oracle.betaplane_velocity(m.x, m.y, zeta, m.area, m.is_active)
    for g, l, f in zip(got, ld, f64):
        assert relerr(g, l) <= TOL
        assert relerr(g, f) <= max(TOL, 2.0 * relerr(f, l) + 1e-14)


@pytest.mark.parametrize("sym", [200, 201], indirect=True)
def test_sym_plane_and_betaplane_stream(sym, oracle, get_mesh):
    """The planar and beta-plane stream functions through the generic symmetric log kernel (SymLogStream)."""
    q = get_mesh(M.QUAD_RECT_SEED, 4, 7.0)
    vq = problems.colliding_dipoles(q)
    got = sym.plane_stream(q.x, q.y, vq, q.area, q.is_active)
    assert relerr(got, oracle.plane_stream(q.x, q.y, vq, q.area, q.is_active)) <= TOL
    m = get_mesh(M.BETA_PLANE_SEED, 4)
    zeta = problems.betaplane_gaussian(m)
    absv = zeta + 1.0 + 2.0 * m.y
    gotb = sym.betaplane_stream(m.x, m.y, zeta, absv, m.area, m.is_active)
    ld = oracle.betaplane_stream(m.x, m.y, zeta, absv, m.area, m.is_active, variant="_ld")
    f64 = oracle.betaplane_stream(m.x, m.y, zeta, absv, m.area, m.is_active)
    for g, l, f in zip(gotb, ld, f64):
        assert relerr(g, l) <= TOL
        assert relerr(g, f) <= 1e-11


@pytest.mark.parametrize("sym", [200], indirect=True)
def test_sym_plane_rk4_steps(sym, oracle, get_mesh):
    """The resident planar solver: symmetric velocity sums and stream function."""
    import test_parity_gpu as tp
    tp.test_plane_rk4_steps(sym, oracle, get_mesh)


@pytest.mark.parametrize("sym", [200], indirect=True)
def test_sym_betaplane_rk4_step(sym, oracle, get_mesh):
    """The resident beta-plane solver takes the symmetric path for its velocity sums."""
    import test_parity_gpu as tp
    tp.test_betaplane_rk4_step(sym, oracle, get_mesh)


@pytest.mark.parametrize("sym", [200, 201, 204, 205, 206, 207], indirect=True)
def test_sym_matches_default_path(sym, get_mesh):
    """Same sum, other order: within a few ulp of the default kernel at icosTri 6."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 6)
    zeta = problems.gaussian_vortex(m)
    a = sym.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    sym.set_bve_variant(0)
    b = sym.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    for g, w in zip(a, b):
        assert relerr(g, w) <= 1e-13


@pytest.mark.parametrize("sym", [200], indirect=True)
def test_sym_rk4_step(sym, oracle, get_mesh):
    """The resident solver takes the symmetric path for its four velocity sums."""
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


@pytest.mark.parametrize("sym", [44, 45, 103], indirect=True)
def test_fenced_one_sided_variants(sym, oracle, get_mesh):
    """The one-sided kernels with a scheduling fence after every source (variant 44: BVE velocity, 103: the
    stream-function kernels): same arithmetic, another instruction schedule -- also not measured yet."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 4)
    zeta = problems.gaussian_vortex(m)
    av = problems.abs_vorticity(m, zeta, 2.0 * np.pi)
    got = sym.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    want = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    for g, w in zip(got, want):
        assert relerr(g, w) <= TOL
    gots = sym.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
    wants = oracle.bve_stream(m.x, m.y, m.z, zeta, av, m.area, m.is_active, 1.0)
    for g, w in zip(gots, wants):
        assert relerr(g, w) <= TOL
    q = get_mesh(M.QUAD_RECT_SEED, 4, 7.0)
    vq = problems.colliding_dipoles(q)
    assert relerr(sym.plane_stream(q.x, q.y, vq, q.area, q.is_active), oracle.plane_stream(q.x, q.y, vq, q.area, q.is_active)) <= TOL


@pytest.mark.parametrize("sym", [204, 205, 208, 209], indirect=True)
@pytest.mark.parametrize("R", [1.0, 6.371e6, 3.0e-7])
def test_sym_fixed_point_accumulation(sym, oracle, R):
    """Variants 204 / 205: the symmetric sums with order-independent (fixed-point) accumulation -- parity as the
    other variants, and bit-identical results from run to run whatever order the CTAs finish in."""
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
