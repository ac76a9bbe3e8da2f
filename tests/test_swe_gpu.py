"""Shallow-water direct sums beyond SWEPlaneRHSIntegrals (tests/test_pse_ops_gpu.py): the planar
velocity from vorticity + divergence (src/PlanarSWE.f90:469-494) and the spherical right-hand-side
integrals as far as the reference computes them (src/SphereSWESolver.f90:296-375).
CPU part: the oracle restatements against independent numpy evaluations and against the
operators they specialise to.  GPU part: parity through the C ABI."""
import numpy as np
import pytest

from lpm_v2_b200 import mesh as M, problems
from conftest import relerr

TOL = 1e-12
PI = problems.PI


def _plane_fields(q):
    r2 = q.x ** 2 + q.y ** 2
    vort = np.exp(-2.0 * r2) * (1.0 + 0.3 * q.x)
    div = 0.2 * np.sin(q.x) * np.exp(-r2)
    return vort, div


def _sphere_fields(m, R):
    zeta = problems.rossby_haurwitz54(m)
    div = 0.3 * m.x * m.y + 0.1 * m.z
    surf = 1.0 + 0.05 * m.z ** 2 + 0.02 * m.x * m.y
    return zeta, div, surf


# ------------------------------------------------------------------ CPU: the oracle itself
def test_oracle_swe_plane_velocity(oracle, get_mesh):
    q = get_mesh(M.QUAD_RECT_SEED, 3, 3.0)
    vort, div = _plane_fields(q)
    u, v = oracle.swe_plane_velocity(q.x, q.y, vort, div, q.area, q.is_active)
    # numpy restatement of PlanarSWE.f90:473-488
    dx = q.x[:, None] - q.x[None, :]
    dy = q.y[:, None] - q.y[None, :]
    sq = dx * dx + dy * dy
    np.fill_diagonal(sq, 1.0)
    act = (q.is_active != 0)[None, :] & ~np.eye(q.n, dtype=bool)
    rot = np.where(act, vort[None, :] * q.area[None, :] / (2 * PI * sq), 0.0)
    pot = np.where(act, div[None, :] * q.area[None, :] / (2 * PI * sq), 0.0)
    assert relerr(u, (-dy * rot + dx * pot).sum(1)) <= 1e-13
    assert relerr(v, (dx * rot + dy * pot).sum(1)) <= 1e-13
    # the same sums are the velocity part of SWEPlaneRHSIntegrals, and with zero divergence the planar Biot-Savart velocity
    surf = 1.0 + 0.1 * q.x
    ru, rv = oracle.swe_plane_rhs(q.x, q.y, vort, div, surf, q.area, q.is_active, 0.3)[:2]
    assert np.array_equal(u, ru) and np.array_equal(v, rv)
    u0, v0 = oracle.swe_plane_velocity(q.x, q.y, vort, 0 * div, q.area, q.is_active)
    pu, pv = oracle.plane_velocity(q.x, q.y, vort, q.area, q.is_active)
    assert relerr(u0, pu) <= 1e-13 and relerr(v0, pv) <= 1e-13


@pytest.mark.parametrize("R", [1.0, 2.5])
def test_oracle_swe_sphere_rhs(oracle, get_mesh, R):
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 2)
    x, y, z, area = m.x * R, m.y * R, m.z * R, m.area * R * R
    zeta, div, surf = _sphere_fields(m, R)
    eps = (m.max_edge_length * R) ** 0.6
    u, v, w, dd, lap = oracle.swe_sphere_rhs(x, y, z, zeta, div, surf, area, m.is_active, R, eps)
    assert not dd.any()                                   # the reference never accumulates it
    # numpy restatement of SphereSWESolver.f90:344-371
    X = np.stack([x, y, z], 1)
    dot = X @ X.T
    den = 4 * PI * R * R * (R * R - dot)
    np.fill_diagonal(den, 1.0)
    act = (m.is_active != 0)[None, :] & ~np.eye(m.n, dtype=bool)
    rot = np.where(act, zeta[None, :] * area[None, :] / den, 0.0)
    pot = np.where(act, R * div[None, :] * area[None, :] / den, 0.0)
    cr = np.cross(X[:, None, :], X[None, :, :])
    vel = -(cr * rot[:, :, None]).sum(1) - (X[None, :, :] * pot[:, :, None]).sum(1)
    scale = np.abs(vel).max()
    assert max(np.abs(u - vel[:, 0]).max(), np.abs(v - vel[:, 1]).max(), np.abs(w - vel[:, 2]).max()) <= 1e-13 * scale
    # the Laplacian sum is PSESphereLaplacianAtParticles without its trailing 1/eps^2
    pl = oracle.pse_laplacian_sphere(x, y, z, surf, area, m.is_active, eps, R)
    assert relerr(lap, pl * eps * eps) <= 1e-13
    # zero divergence: the BVE velocity divided by R (1/(4 pi R^2 d) against 1/(4 pi R d))
    u0, v0, w0 = oracle.swe_sphere_rhs(x, y, z, zeta, 0 * div, surf, area, m.is_active, R, eps)[:3]
    bu, bv, bw = oracle.bve_velocity(x, y, z, zeta, area, m.is_active, R)
    assert max(relerr(u0, bu / R), relerr(v0, bv / R), relerr(w0, bw / R)) <= 1e-13


# ------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("L", [2, 4, 5])
def test_swe_plane_velocity(gpu, oracle, get_mesh, L):
    q = get_mesh(M.QUAD_RECT_SEED, L, 3.0)
    vort, div = _plane_fields(q)
    got = gpu.swe_plane_velocity(q.x, q.y, vort, div, q.area, q.is_active)
    want = oracle.swe_plane_velocity(q.x, q.y, vort, div, q.area, q.is_active)
    assert max(relerr(g, w) for g, w in zip(got, want)) <= TOL
    surf = 1.0 + 0.1 * q.x
    ru, rv = gpu.swe_plane_rhs_integrals(q.x, q.y, vort, div, surf, q.area, q.is_active, 0.3)[:2]
    assert relerr(got[0], ru) <= 1e-13 and relerr(got[1], rv) <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("L,power,R", [(3, 0.6, 1.0), (4, 0.75, 1.0), (4, 1.3, 6371.22), (5, 0.75, 1.0)])
def test_swe_sphere_rhs_integrals(gpu, oracle, get_mesh, L, power, R):
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
    x, y, z, area = m.x * R, m.y * R, m.z * R, m.area * R * R
    zeta, div, surf = _sphere_fields(m, R)
    eps = (m.max_edge_length * R) ** power if R == 1.0 else R * m.max_edge_length ** power
    got = gpu.swe_sphere_rhs_integrals(x, y, z, zeta, div, surf, area, m.is_active, R, eps)
    want = oracle.swe_sphere_rhs(x, y, z, zeta, div, surf, area, m.is_active, R, eps)
    scale = max(np.abs(a).max() for a in want[:3])
    assert max(np.abs(g - w).max() for g, w in zip(got[:3], want[:3])) <= TOL * scale
    assert not got[3].any()
    assert relerr(got[4], want[4]) <= TOL
    # pieces: BVE velocity / R at zero divergence, PSE Laplacian without its 1/eps^2
    u0, v0, w0 = gpu.swe_sphere_rhs_integrals(x, y, z, zeta, 0 * div, surf, area, m.is_active, R, eps)[:3]
    bu, bv, bw = gpu.bve_velocity(x, y, z, zeta, area, m.is_active, R)
    assert max(relerr(u0, bu / R), relerr(v0, bv / R), relerr(w0, bw / R)) <= 1e-13
    pl = gpu.pse_laplacian_sphere(x, y, z, surf, area, m.is_active, eps, R)
    assert relerr(got[4], pl * eps * eps) <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("with_topo", [False, True])
def test_swe_plane_rk4_steps(gpu, oracle, get_mesh, with_topo):
    """type SWESolver New / Timestep (src/SWEPlaneSolver.f90:137-205, 298-429): three RK4 steps of the planar
    shallow-water equations on quadRect 3 against the as-written restatement (including the whole-array assignment of
    the stage-1 vorticity / divergence tendencies, :312-315), with a flat bottom and with a user topography function
    called on the host at every stage."""
    from lpm_v2_b200 import solvers
    q = get_mesh(M.QUAD_RECT_SEED, 3, 3.0)
    x, y = q.x, q.y
    rv = np.exp(-2.0 * (x ** 2 + y ** 2))
    dv = 0.1 * rv * x
    h = 1.0 + 0.1 * rv
    f0, beta, g, dt = 1.0, 0.5, 9.80616, 0.01
    eps = q.max_edge_length ** 0.75
    topo = (lambda a, b: 0.05 * float(np.exp(-(a * a + b * b)))) if with_topo else None
    plane = solvers.SWEMeshPlane(q, rv, dv, h, f0=f0, beta=beta, g=g, pseEps=eps)
    sol = solvers.SWEPlaneSolver(plane, topo)
    surf = h + (np.array([topo(a, b) for a, b in zip(x, y)]) if with_topo else 0.0)
    ref = [x.copy(), y.copy(), rv.copy(), dv.copy(), h.copy(), q.area.copy()] + list(
        oracle.swe_plane_rhs(x, y, rv, dv, surf, q.area, q.is_active, eps))
    names = "x y relVort div h area u v doubleDot lapSurf".split()
    got = [plane.x, plane.y, plane.relVort, plane.divergence, plane.h, plane.area] + plane.velocity + [plane.doubleDot, plane.lapSurf]
    for nm, a, b in zip(names, got, ref):                     # New(): the right-hand side at the initial state
        assert relerr(a, b) <= TOL, ("new", nm)
    for step in range(3):
        sol.Timestep(plane, dt)
        ref = oracle.swe_plane_rk4_step(*ref, q.is_active, f0, beta, g, eps, dt, topo)
        got = [plane.x, plane.y, plane.relVort, plane.divergence, plane.h, plane.area] + plane.velocity + [plane.doubleDot, plane.lapSurf]
        for nm, a, b in zip(names, got, ref):
            assert relerr(a, b) <= TOL, (step, nm, relerr(a, b))
    sol.Delete()
