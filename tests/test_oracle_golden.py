"""Pins the CPU oracle before anything trusts it.

Reference-derived known answers: the PSE Laplacian regression thresholds of
tests/SpherePSEConvTest.f90:373-390 (tests/golden/reference_thresholds.json).
The BVE / planar / beta-plane sums are "parity unpinned" in the reference (no
test asserts a value); for them the oracle is checked against the analytic
solutions the reference's own examples log, against an independent numpy
restatement of the Fortran expressions, against an extended-precision
evaluation, and against committed oracle fixtures (drift guard).
"""
import json
import os

import numpy as np
import pytest

from lpm_v2_b200 import mesh as M, problems
from conftest import relerr

HERE = os.path.dirname(os.path.abspath(__file__))
PI = problems.PI


# ---------------------------------------------------------------- reference known answers
def test_pse_laplacian_reference_thresholds(oracle, get_mesh):
    th = json.load(open(os.path.join(HERE, "golden", "reference_thresholds.json")))
    m = get_mesh(th["mesh"]["seed"], th["mesh"]["init_nest"])
    eps = m.max_edge_length ** th["mesh"]["pse_power"]                  # PSEDirectSum.f90:103-111
    act = m.is_active != 0
    const = np.full(m.n, th["const_value"])
    lap_c = oracle.pse_laplacian_sphere(m.x, m.y, m.z, const, m.area, m.is_active, eps, 1.0)
    assert np.abs(lap_c).max() <= th["particlesLinfConstLap_max"]      # :373-377
    harm = problems.spherical_harmonic54(m)
    exact = -30.0 * harm
    lap = oracle.pse_laplacian_sphere(m.x, m.y, m.z, harm, m.area, m.is_active, eps, 1.0)
    err = np.abs(lap - exact)
    linf = err.max() / np.abs(exact).max()                              # :293
    assert linf <= th["particlesLinfHarmLap_max"]                       # :387
    # the threshold is a tight regression bound: we sit within 1 % of it
    assert linf > 0.99 * th["particlesLinfHarmLap_max"]
    s, e, _ = oracle.load_balance(m.n, th["mesh"]["np"])
    sl = slice(int(s[0]) - 1, int(e[0]))
    a = act[sl]
    l2 = np.sum(err[sl][a] ** 2 * m.area[sl][a]) / np.sum(exact[sl][a] ** 2 * m.area[sl][a])   # :301-319
    assert l2 <= th["particlesL2HarmLap_rank0_max"]
    assert l2 > 0.99 * th["particlesL2HarmLap_rank0_max"]


def _latlon_grid(n_lat, n_lon):
    """tests/SpherePSEConvTest.f90:155-172 (radius 1)."""
    dlam = 360.0 / n_lon
    d2r = PI / 180.0
    lats = -0.5 * PI + np.arange(n_lat) * dlam * d2r
    lons = np.arange(n_lon) * dlam * d2r
    lo, la = np.meshgrid(lons, lats)

    class Grid:
        x = (np.cos(lo) * np.cos(la)).ravel()
        y = (np.sin(lo) * np.cos(la)).ravel()
        z = np.sin(la).ravel()
    return Grid


def test_pse_interpolation_reference_thresholds(oracle, get_mesh):
    """PSESphereInterpolateScalar (PSEDirectSum.f90:151-168) against the interpolation
    bounds of SpherePSEConvTest.f90:379-385: at the particles (:250-251, :292, :310) and on
    the 181 x 360 lat-lon grid (:196-197, :344)."""
    th = json.load(open(os.path.join(HERE, "golden", "reference_thresholds.json")))
    m = get_mesh(th["mesh"]["seed"], th["mesh"]["init_nest"])
    eps = m.max_edge_length ** th["mesh"]["pse_power"]
    harm = problems.spherical_harmonic54(m)
    act = m.is_active != 0
    hi = oracle.pse_interpolate(m.x, m.y, m.z, harm, m.area, m.is_active, eps, m.x, m.y, m.z)
    err = np.abs(hi - harm)
    linf = err.max() / np.abs(harm).max()
    assert 0.99 * th["particlesLinfHarm_max"] < linf <= th["particlesLinfHarm_max"]
    s, e, _ = oracle.load_balance(m.n, th["mesh"]["np"])
    sl = slice(int(s[0]) - 1, int(e[0]))
    a = act[sl]
    l2 = np.sum(err[sl][a] ** 2 * m.area[sl][a]) / np.sum(harm[sl][a] ** 2 * m.area[sl][a])
    assert 0.99 * th["particlesL2Harm_rank0_max"] < l2 <= th["particlesL2Harm_rank0_max"]
    g = _latlon_grid(th["grid"]["nLat"], th["grid"]["nLon"])
    hd = problems.spherical_harmonic54(g)
    gi = oracle.pse_interpolate(m.x, m.y, m.z, harm, m.area, m.is_active, eps, g.x, g.y, g.z)
    unif = np.abs(gi - hd).max() / np.abs(hd).max()
    assert 0.99 * th["unifLinfHarm_max"] < unif <= th["unifLinfHarm_max"]
    ci = oracle.pse_interpolate(m.x, m.y, m.z, np.full(m.n, 2.0), m.area, m.is_active, eps, m.x, m.y, m.z)
    assert np.abs(ci - 2.0).max() / 2.0 < 2e-2           # constants reproduced to ~1 % at L3 (logged only, :290)


def test_pse_grid_laplacian_reference_threshold(oracle, get_mesh):
    """unifLinfHarmLap <= 0.01805 (SpherePSEConvTest.f90:379): the PSE Laplacian of Y_5^4 evaluated at
    the 181 x 360 lat-lon grid points (:196-207, f_target = the exact harmonic there) against -30 Y_5^4
    (:342).  Here: 0.018048 -- the bound is the reference's own value rounded up in the fifth digit."""
    th = json.load(open(os.path.join(HERE, "golden", "reference_thresholds.json")))
    m = get_mesh(th["mesh"]["seed"], th["mesh"]["init_nest"])
    eps = m.max_edge_length ** th["mesh"]["pse_power"]
    harm = problems.spherical_harmonic54(m)
    g = _latlon_grid(th["grid"]["nLat"], th["grid"]["nLon"])
    hd = problems.spherical_harmonic54(g)
    lap = oracle.pse_laplacian_sphere_at_points(m.x, m.y, m.z, harm, m.area, m.is_active, eps, g.x, g.y, g.z, hd)
    exact = -30.0 * hd
    linf = np.abs(lap - exact).max() / np.abs(exact).max()
    assert 0.999 * th["unifLinfHarmLap_max"] < linf <= th["unifLinfHarmLap_max"]
    const = oracle.pse_laplacian_sphere_at_points(m.x, m.y, m.z, np.full(m.n, 2.0), m.area, m.is_active, eps,
                                                  g.x, g.y, g.z, np.full(g.x.size, 2.0))
    assert np.abs(const).max() <= th["ZERO_TOL"]                      # unifLinfConstLap, :373-377


def test_pse_operators_on_smooth_fields(oracle, get_mesh):
    """Gradient / divergence / second partials / double dot restatements against analytic
    derivatives (consistency of sign, projection and eps scaling; O(eps^8 + (h/eps)^p) error)."""
    q = get_mesh(M.QUAD_RECT_SEED, 5, 1.0)
    eps = q.max_edge_length ** 0.5
    inner = (np.abs(q.x) < 0.45) & (np.abs(q.y) < 0.45)       # away from the free boundary
    f = np.sin(2.0 * q.x) * np.cos(q.y)
    gx, gy = oracle.pse_gradient_plane(q.x, q.y, f, q.area, q.is_active, eps)
    ex, ey = 2.0 * np.cos(2.0 * q.x) * np.cos(q.y), -np.sin(2.0 * q.x) * np.sin(q.y)
    assert np.abs(gx - ex)[inner].max() < 2e-2 and np.abs(gy - ey)[inner].max() < 2e-2
    dxx, dxy, dyy = oracle.pse_second_partials_plane(q.x, q.y, ex, ey, q.area, q.is_active, eps)
    assert np.abs(dxx + 4.0 * f)[inner].max() < 5e-2 and np.abs(dyy + f)[inner].max() < 5e-2
    assert np.abs(dxy + 2.0 * np.cos(2.0 * q.x) * np.sin(q.y))[inner].max() < 5e-2
    u, v = q.y ** 2, q.x * q.y                               # u_x = 0, u_y = 2y, v_x = y, v_y = x
    dd = oracle.pse_double_dot_plane(q.x, q.y, u, v, q.area, q.is_active, eps)
    assert np.abs(dd - (4.0 * q.y ** 2 + q.x ** 2))[inner].max() < 5e-2
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 4)
    eps = m.max_edge_length ** 0.4      # at L4 the O(h/eps) quadrature error needs a wide kernel
    g3 = oracle.pse_gradient_sphere(m.x, m.y, m.z, m.z, m.area, m.is_active, eps)      # grad z = e_z - z x
    ez = [-m.z * m.x, -m.z * m.y, 1.0 - m.z * m.z]
    assert max(np.abs(a - b).max() for a, b in zip(g3, ez)) < 5e-2
    div = oracle.pse_divergence_sphere(m.x, m.y, m.z, *ez, m.area, m.is_active, eps)   # div grad z = -2 z
    assert np.abs(div + 2.0 * m.z).max() < 5e-2
    rot = oracle.pse_divergence_sphere(m.x, m.y, m.z, -m.y, m.x, 0 * m.x, m.area, m.is_active, eps)
    assert np.abs(rot).max() < 5e-2                                                   # solid-body rotation is divergence free


def test_pse_kernel_values(oracle):
    """bivariateLaplacianKernel8 (PSEDirectSum.f90:622-627) closed form."""
    k8 = oracle.get().oracle_pse_laplacian_kernel8
    assert k8(0.0) == 40.0 / PI
    for r in (0.3, 1.0, 2.5):
        want = (40 - 40 * r ** 2 + 10 * r ** 4 - 2 * r ** 6 / 3) * np.exp(-r * r) / PI
        assert abs(k8(r) - want) <= 1e-15 * max(1.0, abs(want))


# ---------------------------------------------------------------- analytic solutions
def test_bve_solid_body_rotation_converges(oracle, get_mesh):
    """examples/BVESolidBody.f90:231-243: zeta = 2 Omega z/R => u = Omega (-y, x, 0).
    Checks sign, the -1/(4 pi R) normalisation and the cross-product order; the
    midpoint-rule error of the singular kernel must fall with refinement."""
    errs = []
    for L in (2, 3, 4):
        m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
        zeta, (ue, ve, we) = problems.solid_body(m)
        u, v, w = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
        errs.append(np.sqrt((u - ue) ** 2 + (v - ve) ** 2 + (w - we) ** 2).max() / (2 * PI))
    assert errs[0] < 0.02 and errs[2] < 0.006
    assert errs[0] > errs[1] > errs[2]


def test_bve_stream_function_rh54_eigenfunction(oracle, get_mesh):
    """Y_5^4 is a Laplace eigenfunction (tests/SpherePSEConvTest.f90:508-512):
    with the reference's sign, g = -log(R^2 - x.x')/(4 pi) inverts -Laplacian, so
    the stream function of zeta = 30 Y is psi = +zeta/30 (+ constant)."""
    errs = []
    for L in (3, 4):
        m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
        zeta = problems.rossby_haurwitz54(m)
        rs, _ = oracle.bve_stream(m.x, m.y, m.z, zeta, zeta, m.area, m.is_active, 1.0)
        act = m.is_active != 0
        psi_exact = zeta / 30.0
        d = (rs - psi_exact)[act]
        d = d - np.average(d, weights=m.area[act])
        errs.append(np.abs(d).max() / np.abs(psi_exact).max())
    # midpoint rule with the self panel skipped: O(h^2 log h), ~6 % at level 4
    assert errs[1] < 0.08 and errs[1] < 0.5 * errs[0]


def test_plane_point_vortex(oracle):
    """One active vortex of circulation G at the origin: u = G/(2 pi r) e_theta."""
    x = np.array([0.0, 1.0, 0.0, -2.0]); y = np.array([0.0, 0.0, 0.5, 0.0])
    vort = np.array([3.0, 0, 0, 0]); area = np.array([0.5, 0, 0, 0]); mask = np.array([1, 0, 0, 0])
    u, v = oracle.plane_velocity(x, y, vort, area, mask)
    G = 1.5
    assert u[0] == 0 and v[0] == 0
    np.testing.assert_allclose([u[1], v[1]], [0, G / (2 * PI)], atol=1e-16)
    np.testing.assert_allclose([u[2], v[2]], [-G / (2 * PI * 0.5), 0], atol=1e-16)
    np.testing.assert_allclose([u[3], v[3]], [0, -G / (2 * PI * 2)], atol=1e-16)


def test_betaplane_kernel_periodic_and_small_distance_limit(oracle):
    """The beta-plane kernel has period 1 in x and tends to the planar
    Biot-Savart kernel as r -> 0 (BetaPlaneSolver.f90:245-248)."""
    x = np.array([0.3, 0.3 + 1e-3, 1.3 + 1e-3]); y = np.array([0.1, 0.1 + 2e-3, 0.1 + 2e-3])
    q = np.array([2.0, 0, 0]); area = np.array([0.25, 0, 0]); mask = np.array([1, 0, 0])
    u, v = oracle.betaplane_velocity(x, y, q, area, mask)
    pu, pv = oracle.plane_velocity(x, y, q, area, mask)
    assert abs(u[1] - pu[1]) / abs(pu[1]) < 1e-4 and abs(v[1] - pv[1]) / abs(pv[1]) < 1e-4
    assert abs(u[2] - u[1]) < 1e-9 * abs(u[1]) and abs(v[2] - v[1]) < 1e-9 * abs(v[1])


# ---------------------------------------------------------------- independent restatement
def _numpy_bve(x, y, z, zeta, area, mask, R):
    """Vectorised restatement of SphereBVESolver.f90:403-407, written
    independently of oracle/lpm_oracle.c."""
    n = x.size
    act = np.nonzero(mask)[0]
    u = np.zeros(n); v = np.zeros(n); w = np.zeros(n)
    for i in range(n):
        j = act[act != i]
        s = -zeta[j] * area[j] / (4.0 * PI * R * (R * R - x[i] * x[j] - y[i] * y[j] - z[i] * z[j]))
        u[i] = np.sum((y[i] * z[j] - z[i] * y[j]) * s)
        v[i] = np.sum((z[i] * x[j] - x[i] * z[j]) * s)
        w[i] = np.sum((x[i] * y[j] - y[i] * x[j]) * s)
    return u, v, w


def test_oracle_matches_independent_numpy_restatement(oracle, get_mesh):
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 2)
    zeta = problems.gaussian_vortex(m)
    u, v, w = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    nu, nv, nw = _numpy_bve(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    assert max(relerr(u, nu), relerr(v, nv), relerr(w, nw)) < 1e-13
    mu, mv, mw = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0, variant="_mesh")
    assert max(relerr(u, mu), relerr(v, mv), relerr(w, mw)) < 1e-13     # SphereBVE.f90:489-531 twin


def test_oracle_vs_extended_precision(oracle, get_mesh):
    """Rounding error of the as-written FP64 sums (bounds the parity budget)."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 3)
    zeta = problems.gaussian_vortex(m)
    a = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    b = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0, variant="_ld")
    assert max(relerr(p, q) for p, q in zip(a, b)) < 1e-13
    absv = problems.abs_vorticity(m, zeta, 2 * PI)
    a = oracle.bve_stream(m.x, m.y, m.z, zeta, absv, m.area, m.is_active, 1.0)
    b = oracle.bve_stream(m.x, m.y, m.z, zeta, absv, m.area, m.is_active, 1.0, variant="_ld")
    assert max(relerr(p, q) for p, q in zip(a, b)) < 1e-13
    q = get_mesh(M.QUAD_RECT_SEED, 3, 7.0)
    vort = problems.colliding_dipoles(q)
    a = oracle.plane_velocity(q.x, q.y, vort, q.area, q.is_active)
    b = oracle.plane_velocity(q.x, q.y, vort, q.area, q.is_active, variant="_ld")
    assert max(relerr(p, r) for p, r in zip(a, b)) < 1e-13
    bp = get_mesh(M.BETA_PLANE_SEED, 4)
    zb = problems.betaplane_gaussian(bp)
    a = oracle.betaplane_velocity(bp.x, bp.y, zb, bp.area, bp.is_active)
    b = oracle.betaplane_velocity(bp.x, bp.y, zb, bp.area, bp.is_active, variant="_ld")
    # cosh - cos cancels for near pairs: the reference expression itself is only
    # good to ~1e-12 here -- recorded, and looser than the others
    assert max(relerr(p, r) for p, r in zip(a, b)) < 1e-11


def test_slices_do_not_change_results(oracle, get_mesh):
    """The reference is bit-reproducible across rank counts: a
    target's sum does not depend on who owns it."""
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 2)
    zeta = problems.gaussian_vortex(m)
    full = oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    s, e, _ = oracle.load_balance(m.n, 3)
    parts = [oracle.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0, rng=(s[r] - 1, e[r])) for r in range(3)]
    for c in range(3):
        got = sum(p[c] for p in parts)
        assert np.array_equal(got, full[c])


def test_active_list_is_pack_order(oracle):
    rng = np.random.default_rng(12345)
    mask = (rng.random(1000) < 0.37).astype(np.int32)
    assert np.array_equal(oracle.active_list(mask), np.nonzero(mask)[0].astype(np.int32))


# ---------------------------------------------------------------- drift guard
@pytest.mark.parametrize("name", ["oracle_bve_icos2", "oracle_pse_icos3", "oracle_plane_quad3", "oracle_beta3"])
def test_oracle_fixtures(oracle, get_mesh, name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    tol = 1e-14       # same source, possibly another libm / compiler
    if name == "oracle_bve_icos2":
        m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 2)
        u, v, w = oracle.bve_velocity(m.x, m.y, m.z, g["zeta"], m.area, m.is_active, 1.0)
        assert max(relerr(u, g["u"]), relerr(v, g["v"]), relerr(w, g["w"])) <= tol
        rs, as_ = oracle.bve_stream(m.x, m.y, m.z, g["zeta"], g["absvort"], m.area, m.is_active, 1.0)
        assert max(relerr(rs, g["relstream"]), relerr(as_, g["absstream"])) <= tol
        st = oracle.bve_rk4_step(m.x, m.y, m.z, g["zeta"], g["u"], g["v"], g["w"], m.area, m.is_active, 1.0, 2 * PI, 0.01)
        for k, a in zip("x y z zeta u v w".split(), st):
            assert relerr(a, g["rk4_" + k]) <= 1e-13
    elif name == "oracle_pse_icos3":
        m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 3)
        lap = oracle.pse_laplacian_sphere(m.x, m.y, m.z, g["f"], m.area, m.is_active, float(g["eps"]), 1.0)
        assert relerr(lap, g["lap"]) <= tol
    elif name == "oracle_plane_quad3":
        m = get_mesh(M.QUAD_RECT_SEED, 3, 7.0)
        u, v = oracle.plane_velocity(m.x, m.y, g["vort"], m.area, m.is_active)
        assert max(relerr(u, g["u"]), relerr(v, g["v"])) <= tol
        assert relerr(oracle.plane_stream(m.x, m.y, g["vort"], m.area, m.is_active), g["psi"]) <= tol
    else:
        m = get_mesh(M.BETA_PLANE_SEED, 3)
        u, v = oracle.betaplane_velocity(m.x, m.y, g["zeta"], m.area, m.is_active)
        assert max(relerr(u, g["u"]), relerr(v, g["v"])) <= 1e-12
