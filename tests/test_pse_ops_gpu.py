"""GPU parity for the remaining PSE operators: interpolation, gradient,
second partials, double dot product, divergence -- through the C ABI against the oracle."""
import json
import os

import numpy as np
import pytest

from lpm_v2_b200 import mesh as M, problems
from conftest import relerr
from test_oracle_golden import _latlon_grid

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12


@pytest.mark.parametrize("L,power", [(3, 0.6), (4, 0.75), (4, 1.4)])
def test_sphere_operators(gpu, oracle, get_mesh, L, power):
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
    eps = m.max_edge_length ** power          # 1.4: the far-field cut-off is active
    f = problems.spherical_harmonic54(m)
    u, v, w = -m.y + 0.3 * m.z * m.x, m.x * m.x, 0.5 * m.y - m.z
    args = (m.x, m.y, m.z)
    got = gpu.pse_gradient_sphere(*args, f, m.area, m.is_active, eps)
    want = oracle.pse_gradient_sphere(*args, f, m.area, m.is_active, eps)
    scale = max(np.abs(a).max() for a in want)
    assert max(np.abs(a - b).max() for a, b in zip(got, want)) <= TOL * scale
    assert relerr(gpu.pse_divergence_sphere(*args, u, v, w, m.area, m.is_active, eps),
                  oracle.pse_divergence_sphere(*args, u, v, w, m.area, m.is_active, eps)) <= TOL
    assert relerr(gpu.pse_double_dot_sphere(*args, u, v, w, m.area, m.is_active, eps),
                  oracle.pse_double_dot_sphere(*args, u, v, w, m.area, m.is_active, eps)) <= TOL
    rng = np.random.default_rng(L)
    t = rng.normal(size=(3, 777))
    t /= np.linalg.norm(t, axis=0)
    assert relerr(gpu.pse_interpolate_sphere(*args, f, m.area, m.is_active, eps, t[0], t[1], t[2]),
                  oracle.pse_interpolate(*args, f, m.area, m.is_active, eps, t[0], t[1], t[2])) <= TOL


def test_sphere_interpolation_reference_thresholds_on_gpu(gpu, oracle, get_mesh):
    """The reference's interpolation bounds (SpherePSEConvTest.f90:379-385) on GPU output:
    1922 particle targets and the 181 x 360 lat-lon grid (65160 separate targets)."""
    th = json.load(open(os.path.join(HERE, "golden", "reference_thresholds.json")))
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 3)
    eps = m.max_edge_length ** th["mesh"]["pse_power"]
    harm = problems.spherical_harmonic54(m)
    hi = gpu.pse_interpolate_sphere(m.x, m.y, m.z, harm, m.area, m.is_active, eps, m.x, m.y, m.z)
    assert np.abs(hi - harm).max() / np.abs(harm).max() <= th["particlesLinfHarm_max"]
    g = _latlon_grid(th["grid"]["nLat"], th["grid"]["nLon"])
    hd = problems.spherical_harmonic54(g)
    gi = gpu.pse_interpolate_sphere(m.x, m.y, m.z, harm, m.area, m.is_active, eps, g.x, g.y, g.z)
    assert np.abs(gi - hd).max() / np.abs(hd).max() <= th["unifLinfHarm_max"]
    assert relerr(gi, oracle.pse_interpolate(m.x, m.y, m.z, harm, m.area, m.is_active, eps, g.x, g.y, g.z)) <= TOL


@pytest.mark.parametrize("L,power", [(2, 0.5), (4, 0.75), (5, 0.75), (5, 1.3)])
def test_swe_plane_rhs_integrals(gpu, oracle, get_mesh, L, power):
    """SWEPlaneRHSIntegrals (src/SWEPlaneSolver.f90:457-560): fused velocity + double dot +
    surface Laplacian; inputs in the style of examples/PlaneSWEGravityWaves.f90."""
    q = get_mesh(M.QUAD_RECT_SEED, L, 3.0)
    eps = q.max_edge_length ** power
    r2 = q.x ** 2 + q.y ** 2
    vort = np.exp(-2.0 * r2) * (1.0 + 0.3 * q.x)
    div = 0.2 * np.sin(q.x) * np.exp(-r2)
    surf = 1.0 + 0.1 * np.exp(-3.0 * ((q.x - 0.4) ** 2 + q.y ** 2)) + 0.05 * np.cos(q.y)
    got = gpu.swe_plane_rhs_integrals(q.x, q.y, vort, div, surf, q.area, q.is_active, eps)
    want = oracle.swe_plane_rhs(q.x, q.y, vort, div, surf, q.area, q.is_active, eps)
    for name, g, w in zip(("u", "v", "doubleDot", "lapSurf"), got, want):
        assert relerr(g, w) <= TOL, name
    # with zero divergence the velocity is the planar Biot-Savart velocity
    u0, v0 = gpu.swe_plane_rhs_integrals(q.x, q.y, vort, 0 * div, surf, q.area, q.is_active, eps)[:2]
    pu, pv = gpu.plane_velocity(q.x, q.y, vort, q.area, q.is_active)
    assert relerr(u0, pu) <= 1e-13 and relerr(v0, pv) <= 1e-13


@pytest.mark.parametrize("L,power", [(3, 0.5), (5, 0.75), (5, 1.3)])
def test_plane_operators(gpu, oracle, get_mesh, L, power):
    q = get_mesh(M.QUAD_RECT_SEED, L, 2.0)
    eps = q.max_edge_length ** power
    f = np.sin(1.3 * q.x) * np.cos(0.7 * q.y) + 0.1 * q.x * q.y
    got = gpu.pse_gradient_plane(q.x, q.y, f, q.area, q.is_active, eps)
    want = oracle.pse_gradient_plane(q.x, q.y, f, q.area, q.is_active, eps)
    assert max(relerr(a, b) for a, b in zip(got, want)) <= TOL
    got2 = gpu.pse_second_partials_plane(q.x, q.y, want[0], want[1], q.area, q.is_active, eps)
    want2 = oracle.pse_second_partials_plane(q.x, q.y, want[0], want[1], q.area, q.is_active, eps)
    assert max(relerr(a, b) for a, b in zip(got2, want2)) <= TOL
    u, v = q.y ** 2 - q.x, q.x * q.y
    assert relerr(gpu.pse_double_dot_plane(q.x, q.y, u, v, q.area, q.is_active, eps),
                  oracle.pse_double_dot_plane(q.x, q.y, u, v, q.area, q.is_active, eps)) <= TOL
    rng = np.random.default_rng(L)
    tx, ty = rng.uniform(-2, 2, 501), rng.uniform(-2, 2, 501)
    assert relerr(gpu.pse_interpolate_plane(q.x, q.y, f, q.area, q.is_active, eps, tx, ty),
                  oracle.pse_interpolate(q.x, q.y, None, f, q.area, q.is_active, eps, tx, ty)) <= TOL


def _three_orders(gpu, run):
    """run() under the three evaluation modes of the PSE kernels (lpm_set_pse_culling)."""
    res = {}
    try:
        for mode in (0, 2, 1):
            gpu.set_pse_culling(mode)
            res[mode] = run()
    finally:
        gpu.set_pse_culling(1)
    return res


@pytest.mark.parametrize("seed,L,power", [(M.ICOS_TRI_SPHERE_SEED, 6, 0.75), (M.ICOS_TRI_SPHERE_SEED, 5, 1.5),
                                          (M.CUBED_SPHERE_SEED, 5, 0.75)])
def test_cell_order_and_tile_culling_sphere(gpu, oracle, get_mesh, seed, L, power):
    """The PSE kernels evaluate sources and targets in cell (Morton) order and skip source
    tiles out of reach of a target block (sorted.cuh, directsum.cuh).  Culling only drops
    tiles whose every pair the per-pair cut-off rejects: switching it off (mode 2) must not
    change a single bit.  The reorder changes the summation order only: the reference order
    (mode 0) agrees to ~1e-15.  A sample of targets is checked against the oracle, which
    evaluates every pair in the reference's order."""
    m = get_mesh(seed, L)
    eps = m.max_edge_length ** power
    f = problems.spherical_harmonic54(m)
    u, v, w = -m.y + 0.3 * m.z * m.x, m.x * m.x, 0.5 * m.y - m.z
    args = (m.x, m.y, m.z)

    def run():
        return [gpu.pse_laplacian_sphere(*args, f, m.area, m.is_active, eps, 1.0),
                *gpu.pse_gradient_sphere(*args, f, m.area, m.is_active, eps),
                gpu.pse_divergence_sphere(*args, u, v, w, m.area, m.is_active, eps),
                gpu.pse_double_dot_sphere(*args, u, v, w, m.area, m.is_active, eps),
                gpu.pse_interpolate_sphere(*args, f, m.area, m.is_active, eps, m.x[::7], m.y[::7], m.z[::7])]
    res = _three_orders(gpu, run)
    for a, b, c in zip(res[2], res[1], res[0]):
        assert np.array_equal(a, b)
        assert relerr(b, c) <= 1e-13
    rng = np.random.default_rng(L)
    idx = np.concatenate([[0, 1, 12, m.n - 1], rng.integers(0, m.n, 28)])
    lap = res[1][0]
    scale = np.abs(lap).max()
    for i in idx:
        want = oracle.pse_laplacian_sphere(*args, f, m.area, m.is_active, eps, 1.0, rng=(int(i), int(i) + 1))
        assert abs(lap[i] - want[i]) <= TOL * scale


def test_cell_order_and_tile_culling_plane(gpu, get_mesh):
    q = get_mesh(M.QUAD_RECT_SEED, 7, 2.0)
    eps = q.max_edge_length ** 0.75
    f = np.sin(1.3 * q.x) * np.cos(0.7 * q.y) + 0.1 * q.x * q.y
    rng = np.random.default_rng(3)
    tx, ty = rng.uniform(-2, 2, 5001), rng.uniform(-2, 2, 5001)

    def run():
        return [gpu.pse_laplacian_plane(q.x, q.y, f, q.area, q.is_active, eps),
                *gpu.pse_gradient_plane(q.x, q.y, f, q.area, q.is_active, eps),
                *gpu.pse_second_partials_plane(q.x, q.y, q.y ** 2 - q.x, q.x * q.y, q.area, q.is_active, eps),
                gpu.pse_double_dot_plane(q.x, q.y, q.y ** 2 - q.x, q.x * q.y, q.area, q.is_active, eps),
                gpu.pse_interpolate_plane(q.x, q.y, f, q.area, q.is_active, eps, tx, ty)]
    res = _three_orders(gpu, run)
    for a, b, c in zip(res[2], res[1], res[0]):
        assert np.array_equal(a, b)
        assert relerr(b, c) <= 1e-13


@pytest.mark.parametrize("L,power", [(5, 0.75), (4, 0.9), (6, 1.2)])
def test_sphere_distance_series_matches_atan2(gpu, oracle, get_mesh, L, power):
    """Inside the cut-off the sphere kernels get (d/eps)^2 from tan^2(theta/2) and a series for
    atan^2 (pairs.cuh, sphere_k2) instead of the reference's sqrt + atan2: same sums to 1e-13,
    and both within 1e-12 of the oracle (which calls atan2 like SphereGeometry.f90:107-125)."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
    eps = m.max_edge_length ** power
    f = problems.spherical_harmonic54(m)
    u, v, w = -m.y + 0.3 * m.z * m.x, m.x * m.x, 0.5 * m.y - m.z
    args = (m.x, m.y, m.z)

    def run():
        return [gpu.pse_laplacian_sphere(*args, f, m.area, m.is_active, eps, 1.0),
                *gpu.pse_gradient_sphere(*args, f, m.area, m.is_active, eps),
                gpu.pse_divergence_sphere(*args, u, v, w, m.area, m.is_active, eps)]
    try:
        gpu.set_pse_series(False)
        plain = run()
    finally:
        gpu.set_pse_series(True)
    series = run()
    for a, b in zip(series, plain):
        assert relerr(a, b) <= 1e-13
    if L <= 5:
        assert relerr(series[0], oracle.pse_laplacian_sphere(*args, f, m.area, m.is_active, eps, 1.0)) <= TOL


@pytest.mark.parametrize("L,power", [(2, 0.9), (4, 0.75)])
def test_sphere_padding_sources(gpu, oracle, get_mesh, L, power):
    """Active counts that are not a multiple of the 256-source tile: the padding records must
    not reach the distance evaluation (0/0 in its series form).  L2 has 320 sources; at L4 a
    ragged mask leaves 5120 - 173."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
    mask = m.is_active.copy()
    if L == 4:
        act = np.nonzero(mask)[0]
        mask[act[::30][:173]] = 0
    assert int(mask.sum()) % 256 != 0
    eps = m.max_edge_length ** power
    f = problems.spherical_harmonic54(m)
    u, v, w = -m.y + 0.3 * m.z * m.x, m.x * m.x, 0.5 * m.y - m.z
    args = (m.x, m.y, m.z)
    lap = gpu.pse_laplacian_sphere(*args, f, m.area, mask, eps, 1.0)
    assert np.all(np.isfinite(lap))
    assert relerr(lap, oracle.pse_laplacian_sphere(*args, f, m.area, mask, eps, 1.0)) <= TOL
    g = gpu.pse_gradient_sphere(*args, f, m.area, mask, eps)
    want = oracle.pse_gradient_sphere(*args, f, m.area, mask, eps)
    scale = max(np.abs(a).max() for a in want)
    assert max(np.abs(a - b).max() for a, b in zip(g, want)) <= TOL * scale
    assert relerr(gpu.pse_divergence_sphere(*args, u, v, w, m.area, mask, eps),
                  oracle.pse_divergence_sphere(*args, u, v, w, m.area, mask, eps)) <= TOL
    assert relerr(gpu.pse_interpolate_sphere(*args, f, m.area, mask, eps, m.x[::3], m.y[::3], m.z[::3]),
                  oracle.pse_interpolate(*args, f, m.area, mask, eps, m.x[::3], m.y[::3], m.z[::3])) <= TOL
