"""make_refsrc_fixtures.py -- golden vectors computed by executing the REFERENCE'S OWN SOURCE TEXT.

TEST INFRASTRUCTURE ONLY.  Runs in the development container, where /root/reference exists:

    python oracle/make_refsrc_fixtures.py            # writes tests/golden/refsrc_*.npz

Each fixture holds the inputs, the outputs, and the reference file:line of the procedure that produced them.  The
procedures are read from the lpm-v2 tree and executed by oracle/fortran_subset.py (see its header for what that means
and what it does not); nothing here restates a formula -- this script only builds the derived-type values the
procedures take (a mesh's particles, fields, an MPISetup for one rank) and calls them, the solver objects included
(their own newPrivate runs under the interpreter).  tests/test_refsrc_golden.py then requires the C oracle
(oracle/lpm_oracle.c) to reproduce every output bit for bit, and the CUDA path to meet its parity bound against them.

Inputs are small meshes from the repo's mesh generator (reference refinement order) and ragged random sets; the
interpreter costs tens of microseconds per pair, so the sizes are a few hundred particles.
"""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fortran_subset as F      # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def A(a, kind=float):
    return F.FArr.of(np.asarray(a).tolist(), kind=kind)


def N(a):
    return np.array(a.tolist())


def mpi_setup(prog, n, nprocs=1):
    """An MPISetup for `nprocs` ranks filled by the reference's LoadBalance (src/MPISetup.f90:132-146)."""
    ms = F.Obj(indexStart=F.FArr([0] * nprocs, lb=0), indexEnd=F.FArr([0] * nprocs, lb=0),
               messageLength=F.FArr([0] * nprocs, lb=0), n=0)
    prog.call("LoadBalance", ms, int(n), int(nprocs))
    return ms


def particles(x, y, z, area, mask):
    return F.Obj(x=A(x), y=A(y), z=None if z is None else A(z), area=A(area), isActive=A(mask, bool), N=len(x))


def field(n, values=None, ndim=1):
    z = lambda: F.FArr.zeros(n)
    if ndim == 1:
        return F.Obj(scalar=A(values) if values is not None else z(), N=n, nDim=1)
    return F.Obj(xComp=z(), yComp=z(), zComp=z(), N=n, nDim=ndim)


def rand_sphere(n, seed, frac, R):
    rng = np.random.default_rng(seed)
    p = rng.normal(size=(n, 3))
    p *= R / np.linalg.norm(p, axis=1)[:, None]
    return p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy(), rng.normal(size=n), rng.random(n) * 4 * np.pi * R * R / n, rng.random(n) < frac


def rand_plane(n, seed, frac, half):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-half, half, n), rng.uniform(-half, half, n), rng.normal(size=n),
            rng.random(n) * (2 * half) ** 2 / n, rng.random(n) < frac)


def save(name, where, **arrays):
    path = os.path.join(GOLDEN, f"refsrc_{name}.npz")
    np.savez_compressed(path, reference=np.array(where), **arrays)
    print(f"  wrote {os.path.relpath(path, ROOT)}  ({where})", flush=True)


def sphere_cases():
    from lpm_v2_b200 import mesh as M, problems
    m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 2)
    yield "icos2", m.x, m.y, m.z, problems.rossby_haurwitz54(m), m.area, m.is_active.astype(bool), 1.0
    x, y, z, zeta, area, mask = rand_sphere(157, 11, 0.6, 1.7)
    yield "rand157", x, y, z, zeta, area, mask, 1.7


def plane_cases():
    from lpm_v2_b200 import mesh as M, problems
    q = M.PolyMesh2d(M.QUAD_RECT_SEED, 3, 3.0)
    yield "quad3", q.x, q.y, problems.colliding_dipoles(q), q.area, q.is_active.astype(bool)
    x, y, vort, area, mask = rand_plane(149, 12, 0.7, 2.0)
    yield "rand149", x, y, vort, area, mask


def beta_cases():
    from lpm_v2_b200 import mesh as M, problems
    b = M.PolyMesh2d(M.BETA_PLANE_SEED, 2)
    yield "beta2", b.x, b.y, problems.betaplane_gaussian(b), b.area, b.is_active.astype(bool)
    x, y, vort, area, mask = rand_plane(131, 13, 0.65, 0.5)
    yield "rand131", x, y, vort, area, mask


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    t0 = time.time()

    # ---- LoadBalance -------------------------------------------------------------------------------------------------
    prog = F.Program(["src/MPISetup.f90"])
    rows = []
    for n, p in [(1966082, 1), (1966082, 8), (10, 3), (7, 7), (3, 8), (491522, 5), (122882, 2)]:
        ms = mpi_setup(prog, n, p)
        for r in range(p):
            rows.append((n, p, r, ms.indexStart.get(r), ms.indexEnd.get(r), ms.messageLength.get(r)))
    save("load_balance", prog.where("LoadBalance"), rows=np.array(rows, dtype=np.int64))

    # ---- sphere: BVESphereVelocity, the mesh-side twin, stream functions, one RK4 step ------------------------------------
    prog = F.Program(["src/SphereBVESolver.f90", "src/SphereBVE.f90", "src/Particles.f90", "src/MPISetup.f90"])
    for tag, x, y, z, zeta, area, mask, R in sphere_cases():
        n = len(x)
        ms = mpi_setup(prog, n)
        u, v, w = (F.FArr.zeros(n) for _ in range(3))
        prog.call("BVESphereVelocity", u, v, w, A(x), A(y), A(z), A(zeta), A(area), float(R), 2 * np.pi, A(mask, bool), ms)
        save(f"bve_velocity_{tag}", prog.where("BVESphereVelocity"), x=x, y=y, z=z, relvort=zeta, area=area, mask=mask, R=R,
             u=N(u), v=N(v), w=N(w))
        omega = 2 * np.pi
        absv = zeta + 2 * omega * np.asarray(z) / R
        bve = F.Obj(mesh=F.Obj(particles=particles(x, y, z, area, mask)), relVort=field(n, zeta), absVort=field(n, absv),
                    relStream=field(n), absStream=field(n), velocity=field(n, ndim=3), radius=float(R), rotationRate=omega,
                    mpiParticles=ms)
        prog.call("setVelocityFromVorticity", bve, file="src/SphereBVE.f90")
        prog.call("SetStreamFunctionsOnMesh", bve, file="src/SphereBVE.f90")
        save(f"bve_mesh_{tag}", prog.where("setVelocityFromVorticity", "src/SphereBVE.f90") + " + " +
             prog.where("SetStreamFunctionsOnMesh", "src/SphereBVE.f90"),
             x=x, y=y, z=z, relvort=zeta, absvort=absv, area=area, mask=mask, R=R,
             u=N(bve.velocity.xComp), v=N(bve.velocity.yComp), w=N(bve.velocity.zComp),
             relstream=N(bve.relStream.scalar), absstream=N(bve.absStream.scalar))
    # three ranks: each evaluates its LoadBalance slice (the MPI_BCAST that follows is the exchange the library replaces)
    x, y, z, zeta, area, mask = rand_sphere(157, 11, 0.6, 1.7)
    n, P = len(x), 3
    u, v, w = (F.FArr.zeros(n) for _ in range(3))
    bounds = []
    for r in range(P):
        pr = F.Program(["src/SphereBVESolver.f90", "src/MPISetup.f90"], num_procs=P, proc_rank=r)
        ms3 = mpi_setup(pr, n, P)
        pr.call("BVESphereVelocity", u, v, w, A(x), A(y), A(z), A(zeta), A(area), 1.7, 2 * np.pi, A(mask, bool), ms3)
        bounds.append((ms3.indexStart.get(r), ms3.indexEnd.get(r)))
    save("bve_velocity_rand157_3ranks", prog.where("BVESphereVelocity"), x=x, y=y, z=z, relvort=zeta, area=area, mask=mask,
         R=1.7, bounds=np.array(bounds, dtype=np.int64), u=N(u), v=N(v), w=N(w))
    # RK4: icosTri 1 (122 particles), two steps of dt = 0.01 through the reference's BVESolver New + timestepPrivate
    from lpm_v2_b200 import mesh as M, problems
    m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 1)
    zeta = problems.gaussian_vortex(m)
    n, R, omega, dt = m.n, 1.0, 2 * np.pi, 0.01
    ms = mpi_setup(prog, n)
    absv = zeta + 2 * omega * m.z / R
    bve = F.Obj(mesh=F.Obj(particles=particles(m.x, m.y, m.z, m.area, m.is_active.astype(bool))), relVort=field(n, zeta),
                absVort=field(n, absv), relStream=field(n), absStream=field(n), velocity=field(n, ndim=3), radius=R,
                rotationRate=omega, mpiParticles=ms)
    prog.call("setVelocityFromVorticity", bve, file="src/SphereBVE.f90")
    u0, v0, w0 = N(bve.velocity.xComp), N(bve.velocity.yComp), N(bve.velocity.zComp)
    solver = F.Obj()
    prog.call("newPrivate", solver, bve, file="src/SphereBVESolver.f90")
    steps = []
    for _ in range(2):
        prog.call("timestepPrivate", solver, bve, dt, file="src/SphereBVESolver.f90")
        p = bve.mesh.particles
        steps.append([N(p.x), N(p.y), N(p.z), N(bve.relVort.scalar), N(bve.velocity.xComp), N(bve.velocity.yComp),
                      N(bve.velocity.zComp), N(bve.relStream.scalar), N(bve.absStream.scalar)])
    save("bve_rk4_icos1", prog.where("timestepPrivate", "src/SphereBVESolver.f90"), x=m.x, y=m.y, z=m.z, relvort=zeta,
         absvort=absv, area=m.area, mask=m.is_active.astype(bool), R=R, omega=omega, dt=dt, u0=u0, v0=v0, w0=w0,
         steps=np.array(steps))

    # ---- plane ---------------------------------------------------------------------------------------------------------
    prog = F.Program(["src/PlaneIncompressibleSolver.f90", "src/PlanarIncompressible.f90", "src/Particles.f90", "src/MPISetup.f90"])
    for tag, x, y, vort, area, mask in plane_cases():
        n = len(x)
        ms = mpi_setup(prog, n)
        u, v = F.FArr.zeros(n), F.FArr.zeros(n)
        prog.call("planarIncompressibleVelocity", u, v, A(x), A(y), A(vort), A(area), A(mask, bool), ms)
        plane = F.Obj(mesh=F.Obj(particles=particles(x, y, None, area, mask)), vorticity=field(n, vort), streamFn=field(n),
                      velocity=field(n, ndim=2), mpiParticles=ms)
        prog.call("setVelocityFromVorticity", plane, file="src/PlanarIncompressible.f90")
        prog.call("SetStreamFunctionOnMesh", plane, file="src/PlanarIncompressible.f90")
        save(f"plane_{tag}", prog.where("planarIncompressibleVelocity") + " + " +
             prog.where("SetStreamFunctionOnMesh", "src/PlanarIncompressible.f90"),
             x=x, y=y, vort=vort, area=area, mask=mask, u=N(u), v=N(v), u_mesh=N(plane.velocity.xComp),
             v_mesh=N(plane.velocity.yComp), stream=N(plane.streamFn.scalar))
    q = M.PolyMesh2d(M.QUAD_RECT_SEED, 2, 3.0)
    vort = problems.colliding_dipoles(q)
    n, dt = q.n, 0.01
    ms = mpi_setup(prog, n)
    plane = F.Obj(mesh=F.Obj(particles=particles(q.x, q.y, None, q.area, q.is_active.astype(bool))), vorticity=field(n, vort),
                  streamFn=field(n), velocity=field(n, ndim=2), mpiParticles=ms)
    prog.call("setVelocityFromVorticity", plane, file="src/PlanarIncompressible.f90")
    u0, v0 = N(plane.velocity.xComp), N(plane.velocity.yComp)
    solver = F.Obj()
    prog.call("newPrivate", solver, plane, file="src/PlaneIncompressibleSolver.f90")
    steps = []
    for _ in range(2):
        prog.call("timestepPrivate", solver, plane, dt, file="src/PlaneIncompressibleSolver.f90")
        p = plane.mesh.particles
        steps.append([N(p.x), N(p.y), N(plane.velocity.xComp), N(plane.velocity.yComp), N(plane.streamFn.scalar)])
    save("plane_rk4_quad2", prog.where("timestepPrivate", "src/PlaneIncompressibleSolver.f90"), x=q.x, y=q.y, vort=vort,
         area=q.area, mask=q.is_active.astype(bool), dt=dt, u0=u0, v0=v0, steps=np.array(steps))

    # ---- beta plane ------------------------------------------------------------------------------------------------------
    prog = F.Program(["src/BetaPlaneSolver.f90", "src/BetaPlane.f90", "src/Particles.f90", "src/MPISetup.f90"])
    for tag, x, y, zeta, area, mask in beta_cases():
        n = len(x)
        ms = mpi_setup(prog, n)
        u, v = F.FArr.zeros(n), F.FArr.zeros(n)
        prog.call("BetaPlaneVelocity", u, v, A(x), A(y), A(zeta), A(area), A(mask, bool), ms)
        absv = zeta + 1.0
        bp = F.Obj(mesh=F.Obj(particles=particles(x, y, None, area, mask)), relVort=field(n, zeta), absVort=field(n, absv),
                   relStream=field(n), absStream=field(n), velocity=field(n, ndim=2), mpiParticles=ms, beta=1.0)
        prog.call("setVelocityFromVorticity", bp, file="src/BetaPlane.f90")
        prog.call("SetStreamFunctionsOnMesh", bp, file="src/BetaPlane.f90")
        save(f"beta_{tag}", prog.where("BetaPlaneVelocity") + " + " + prog.where("SetStreamFunctionsOnMesh", "src/BetaPlane.f90"),
             x=x, y=y, relvort=zeta, absvort=absv, area=area, mask=mask, u=N(u), v=N(v), u_mesh=N(bp.velocity.xComp),
             v_mesh=N(bp.velocity.yComp), relstream=N(bp.relStream.scalar), absstream=N(bp.absStream.scalar))
    b = M.PolyMesh2d(M.BETA_PLANE_SEED, 1)
    zeta = problems.betaplane_gaussian(b)
    n, dt, beta = b.n, 0.01, 1.5
    ms = mpi_setup(prog, n)
    bp = F.Obj(mesh=F.Obj(particles=particles(b.x, b.y, None, b.area, b.is_active.astype(bool))), relVort=field(n, zeta),
               absVort=field(n, zeta + beta * b.y), relStream=field(n), absStream=field(n), velocity=field(n, ndim=2),
               mpiParticles=ms, beta=beta)
    prog.call("setVelocityFromVorticity", bp, file="src/BetaPlane.f90")
    u0, v0 = N(bp.velocity.xComp), N(bp.velocity.yComp)
    solver = F.Obj()
    prog.call("newPrivate", solver, bp, file="src/BetaPlaneSolver.f90")
    steps = []
    for _ in range(2):
        prog.call("timestepPrivate", solver, bp, dt, file="src/BetaPlaneSolver.f90")
        p = bp.mesh.particles
        steps.append([N(p.x), N(p.y), N(bp.relVort.scalar), N(bp.velocity.xComp), N(bp.velocity.yComp),
                      N(bp.relStream.scalar), N(bp.absStream.scalar)])
    save("beta_rk4_beta1", prog.where("timestepPrivate", "src/BetaPlaneSolver.f90"), x=b.x, y=b.y, relvort=zeta,
         absvort=zeta + beta * b.y, area=b.area, mask=b.is_active.astype(bool), beta=beta, dt=dt, u0=u0, v0=v0,
         steps=np.array(steps))

    # ---- PSE Laplacians ------------------------------------------------------------------------------------------------
    prog = F.Program(["src/PSEDirectSum.f90", "src/SphereGeometry.f90", "src/Particles.f90", "src/Field.f90", "src/MPISetup.f90"])
    m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 1)
    for tag, x, y, z, f, area, mask, R, eps in (
            ("icos1", m.x, m.y, m.z, problems.rossby_haurwitz54(m), m.area, m.is_active.astype(bool), 1.0, m.max_edge_length ** 0.6),
            ("rand97",) + rand_sphere(97, 14, 0.7, 2.5) + (2.5, 0.9)):
        n = len(x)
        ms = mpi_setup(prog, n)
        prog.globals["sphereradius"] = float(R)         # SetSphereRadius (module variable of TypeDefs.f90:74)
        mesh = F.Obj(particles=particles(x, y, z, area, mask))
        lap = field(n)
        prog.call("PSESphereLaplacianAtParticles", F.Obj(eps=float(eps)), mesh, field(n, f), lap, ms)
        save(f"pse_sphere_{tag}", prog.where("PSESphereLaplacianAtParticles"), x=x, y=y, z=z, f=f, area=area, mask=mask, R=R,
             eps=eps, lap=N(lap.scalar))
    prog.globals["sphereradius"] = 1.0
    q = M.PolyMesh2d(M.QUAD_RECT_SEED, 2, 3.0)
    f = problems.colliding_dipoles(q)
    n, eps = q.n, q.max_edge_length ** 0.75
    ms = mpi_setup(prog, n)
    lap = field(n)
    prog.call("PSEPlaneLaplacianAtParticles", F.Obj(eps=float(eps)), F.Obj(particles=particles(q.x, q.y, None, q.area, q.is_active.astype(bool))),
              field(n, f), lap, ms)
    save("pse_plane_quad2", prog.where("PSEPlaneLaplacianAtParticles"), x=q.x, y=q.y, f=f, area=q.area,
         mask=q.is_active.astype(bool), eps=eps, lap=N(lap.scalar))
    # ---- the other PSE operators (src/PSEDirectSum.f90:128-456, 537-579) ---------------------------------------------------
    prog = F.Program(["src/PSEDirectSum.f90", "src/SphereGeometry.f90", "src/Particles.f90", "src/Field.f90", "src/MPISetup.f90"])
    q = M.PolyMesh2d(M.QUAD_RECT_SEED, 2, 3.0)
    n, eps = q.n, float(q.max_edge_length ** 0.75)
    ms = mpi_setup(prog, n)
    pse = F.Obj(eps=eps)
    qmesh = F.Obj(particles=particles(q.x, q.y, None, q.area, q.is_active.astype(bool)))
    f = problems.colliding_dipoles(q)
    rng = np.random.default_rng(21)
    tx, ty = rng.uniform(-2, 2, 24), rng.uniform(-2, 2, 24)
    interp = [prog.call("PSEPlaneInterpolateScalar", pse, qmesh, field(n, f), F.FArr([float(a), float(b), 0.0])) for a, b in zip(tx, ty)]
    grad = field(n, ndim=2)
    prog.call("PSEPlaneGradientAtParticles", pse, qmesh, field(n, f), grad, ms)
    second = field(n, ndim=3)
    prog.call("PSEPlaneSecondPartialsAtParticles", pse, qmesh, grad, second, ms)
    uu, vv = np.sin(q.x) * np.cos(0.5 * q.y), np.cos(0.7 * q.x) + q.y * q.y
    vec = field(n, ndim=2)
    vec.xComp.assign(A(uu))
    vec.yComp.assign(A(vv))
    dd = field(n)
    prog.call("PSEPlaneDoubleDotProductAtParticles", pse, qmesh, vec, dd, ms)
    save("pse_ops_plane_quad2", prog.where("PSEPlaneInterpolateScalar") + ", " + prog.where("PSEPlaneGradientAtParticles") + ", " +
         prog.where("PSEPlaneSecondPartialsAtParticles") + ", " + prog.where("PSEPlaneDoubleDotProductAtParticles"),
         x=q.x, y=q.y, f=f, area=q.area, mask=q.is_active.astype(bool), eps=eps, tx=tx, ty=ty, interp=np.array(interp),
         gx=N(grad.xComp), gy=N(grad.yComp), dxx=N(second.xComp), dxy=N(second.yComp), dyy=N(second.zComp), u=uu, v=vv,
         double_dot=N(dd.scalar))
    m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 1)
    n, R, eps = m.n, 1.0, float(m.max_edge_length ** 0.6)
    prog.globals["sphereradius"] = R
    ms = mpi_setup(prog, n)
    pse = F.Obj(eps=eps)
    smesh = F.Obj(particles=particles(m.x, m.y, m.z, m.area, m.is_active.astype(bool)))
    f = problems.rossby_haurwitz54(m)
    pts = rng.normal(size=(16, 3))
    pts /= np.linalg.norm(pts, axis=1)[:, None]
    interp = [prog.call("PSESphereInterpolateScalar", pse, smesh, field(n, f), F.FArr([float(c) for c in pnt])) for pnt in pts]
    grad = field(n, ndim=3)
    prog.call("PSESphereGradientAtParticles", pse, smesh, field(n, f), grad, ms)
    uu, vv, ww = -m.y + 0.3 * m.z * m.x, m.x + 0.2 * m.y * m.z, 0.1 * m.x * m.y
    vec = field(n, ndim=3)
    for comp, val in zip((vec.xComp, vec.yComp, vec.zComp), (uu, vv, ww)):
        comp.assign(A(val))
    dd, dv = field(n), field(n)
    prog.call("PSESphereDoubleDotProductAtParticles", pse, smesh, vec, dd, ms)
    prog.call("PSESphereDivergenceAtParticles", pse, smesh, vec, dv, ms)
    save("pse_ops_sphere_icos1", prog.where("PSESphereInterpolateScalar") + ", " + prog.where("PSESphereGradientAtParticles") + ", " +
         prog.where("PSESphereDoubleDotProductAtParticles") + ", " + prog.where("PSESphereDivergenceAtParticles"),
         x=m.x, y=m.y, z=m.z, f=f, area=m.area, mask=m.is_active.astype(bool), eps=eps, R=R, tx=pts[:, 0], ty=pts[:, 1],
         tz=pts[:, 2], interp=np.array(interp), gx=N(grad.xComp), gy=N(grad.yComp), gz=N(grad.zComp), u=uu, v=vv, w=ww,
         double_dot=N(dd.scalar), divergence=N(dv.scalar))

    # ---- shallow water: the fused right-hand sides and the planar RK4 step ------------------------------------------------
    prog = F.Program(["src/SWEPlaneSolver.f90", "src/SphereSWESolver.f90", "src/PSEDirectSum.f90", "src/SphereGeometry.f90",
                      "src/MPISetup.f90"], ignore_calls=("setbottomheightonmesh",))
    q = M.PolyMesh2d(M.QUAD_RECT_SEED, 2, 3.0)
    n, eps = q.n, float(q.max_edge_length ** 0.75)
    ms = mpi_setup(prog, n)
    vort = problems.colliding_dipoles(q)
    div = 0.1 * np.cos(q.x) * np.sin(q.y)
    h = 1.0 + 0.05 * np.exp(-(q.x ** 2 + q.y ** 2))
    flat = lambda x, y: 0.0
    u, v, ddot, lap = (F.FArr.zeros(n) for _ in range(4))
    prog.call("SWEPlaneRHSIntegrals", u, v, ddot, lap, A(q.x), A(q.y), A(vort), A(div), A(h), flat, A(q.area),
              A(q.is_active, bool), eps, ms)
    save("swe_plane_rhs_quad2", prog.where("SWEPlaneRHSIntegrals"), x=q.x, y=q.y, vort=vort, div=div, h=h, area=q.area,
         mask=q.is_active.astype(bool), eps=eps, u=N(u), v=N(v), double_dot=N(ddot), lap_surf=N(lap))
    hill = lambda x, y: 0.1 * math.exp(-2.0 * (x * x + y * y))          # the bottom topography handed to the solver
    f0, beta, g, dt = 0.5, 0.2, 9.80616, 0.005
    vel = field(n, ndim=2)
    plane = F.Obj(mesh=F.Obj(particles=particles(q.x, q.y, None, q.area, q.is_active.astype(bool))), relVort=field(n, vort),
                  potVort=field(n, (vort + f0) / h), divergence=field(n, div), h=field(n, h), velocity=vel, pseEps=eps,
                  mpiParticles=ms, f0=f0, beta=beta, g=g)
    solver = F.Obj()
    prog.call("newPrivate", solver, plane, hill, file="src/SWEPlaneSolver.f90")
    start = [N(solver.u), N(solver.v), N(solver.doubleDot), N(solver.lapSurf)]
    steps = []
    for _ in range(2):
        prog.call("timestepPrivate", solver, plane, dt, hill, file="src/SWEPlaneSolver.f90")
        p = plane.mesh.particles
        steps.append([N(p.x), N(p.y), N(plane.relVort.scalar), N(plane.divergence.scalar), N(plane.h.scalar), N(p.area),
                      N(plane.velocity.xComp), N(plane.velocity.yComp), N(solver.doubleDot), N(solver.lapSurf)])
    save("swe_plane_rk4_quad2", prog.where("timestepPrivate", "src/SWEPlaneSolver.f90"), x=q.x, y=q.y, vort=vort, div=div, h=h,
         area=q.area, mask=q.is_active.astype(bool), eps=eps, f0=f0, beta=beta, g=g, dt=dt, start=np.array(start),
         steps=np.array(steps), topo=np.array("0.1 * exp(-2 (x^2 + y^2))"))
    m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 1)
    n, R, eps = m.n, 1.0, float(m.max_edge_length ** 0.6)
    prog.globals["sphereradius"] = R
    ms = mpi_setup(prog, n)
    vort = problems.rossby_haurwitz54(m)
    div = 0.05 * m.x * m.z
    h = 1.0 + 0.02 * m.y
    flat3 = lambda x, y, z: 0.0
    u, v, w, ddot, lap = (F.FArr.zeros(n) for _ in range(5))
    prog.call("SWESphereRHSIntegrals", u, v, w, ddot, lap, A(m.x), A(m.y), A(m.z), A(vort), A(div), A(h), A(m.area), flat3,
              A(m.is_active, bool), R, eps, ms)
    save("swe_sphere_rhs_icos1", prog.where("SWESphereRHSIntegrals"), x=m.x, y=m.y, z=m.z, vort=vort, div=div, h=h, area=m.area,
         mask=m.is_active.astype(bool), R=R, eps=eps, u=N(u), v=N(v), w=N(w), double_dot=N(ddot), lap_surf=N(lap))

    # ---- diagnostics (src/SphereBVE.f90:410-441) -------------------------------------------------------------------------------
    prog = F.Program(["src/SphereBVE.f90"])
    d = np.load(os.path.join(GOLDEN, "refsrc_bve_mesh_icos2.npz"))
    n = d["x"].size
    vel = field(n, ndim=3)
    for comp, val in zip((vel.xComp, vel.yComp, vel.zComp), (d["u"], d["v"], d["w"])):
        comp.assign(A(val))
    bve = F.Obj(mesh=F.Obj(particles=particles(d["x"], d["y"], d["z"], d["area"], d["mask"])), velocity=vel,
                relVort=field(n, d["relvort"]))
    save("bve_diagnostics_icos2", prog.where("TotalKE") + ", " + prog.where("TotalEnstrophy"), u=d["u"], v=d["v"], w=d["w"],
         relvort=d["relvort"], area=d["area"], mask=d["mask"], ke=prog.call("TotalKE", bve), enstrophy=prog.call("TotalEnstrophy", bve))
    # ---- the workloads' vorticity fields (examples/RossbyHaurwitz54.f90:413-428 with rh54.namelist; config 1's Gaussian vortex,
    # examples/BVESingleGaussianVortex.f90:120-124, 335-357 with bveSingleGaussVort.namelist) -------------------------------------
    prog = F.Program(["examples/RossbyHaurwitz54.f90", "examples/BVESingleGaussianVortex.f90", "src/SphereGeometry.f90"])
    m = M.PolyMesh2d(M.ICOS_TRI_SPHERE_SEED, 2)
    prog.globals.update(radius=1.0, zonalwind=0.0, rhwaveamplitude=1.0)             # rh54.namelist
    rh = [prog.call("RossbyHaurwitz54Vorticity", float(a), float(b), float(c)) for a, b, c in zip(m.x, m.y, m.z)]
    init_lat, init_lon, R = 0.157079632679490, 0.0, 1.0                             # bveSingleGaussVort.namelist
    center = [R * math.cos(init_lon) * math.cos(init_lat), R * math.sin(init_lon) * math.cos(init_lat), R * math.sin(init_lat)]
    prog.globals.update(radius=R, shapeparam=4.0, vortstrength=12.566370614359172, vortcenter=F.FArr(center), gauss_const=0.0)
    g0 = [prog.call("GaussianVortexVorticity", float(a), float(b), float(c)) for a, b, c in zip(m.x, m.y, m.z)]
    bve = F.Obj(mesh=F.Obj(particles=particles(m.x, m.y, m.z, m.area, m.is_active.astype(bool))), relVort=field(m.n, g0), radius=R)
    const = prog.call("SetGaussConst", bve)
    prog.globals["gauss_const"] = const
    g1 = [prog.call("GaussianVortexVorticity", float(a), float(b), float(c)) for a, b, c in zip(m.x, m.y, m.z)]
    save("workload_vorticity_icos2", prog.where("RossbyHaurwitz54Vorticity") + ", " + prog.where("GaussianVortexVorticity") + ", " +
         prog.where("SetGaussConst"), x=m.x, y=m.y, z=m.z, area=m.area, mask=m.is_active.astype(bool), rh54=np.array(rh),
         gauss_const=const, gaussian=np.array(g1))
    # ---- the mesh itself: the reference's PolyMesh2d New (seed file, uniform refinement) -------------------------------------
    # src/PolyMesh2d.f90:135-195 (newPrivate), :795-939 (initializeMeshFromSeed), :956-1043 (readSeedFile, reading the
    # reference's own *Seed.dat), src/Faces.f90:529-858 (DivideQuadFace, DivideTriFace), :917-975 (face areas),
    # src/Edges.f90:211-233, 567-621 (InsertEdge, divideLinearEdge), src/Particles.f90 (New, InsertParticle, ...),
    # src/SphereGeometry.f90 / PlaneGeometry.f90 (midpoints, centroids, triangle areas), src/Edges.f90:260-275 (MaxEdgeLength).
    # Left out: the per-particle incident-edge lists (RecordIncidentEdgeAtParticles, SortIncidentEdgesAtParticle: they do
    # not enter positions, areas, order or face connectivity).
    files = ["src/PolyMesh2d.f90", "src/Particles.f90", "src/Edges.f90", "src/Faces.f90", "src/SphereGeometry.f90", "src/PlaneGeometry.f90"]
    off_path = ("recordincidentedgeatparticles", "replaceincidentedgewithchild", "sortincidentedgesatparticle", "startsection", "endsection")
    for seed, amp, levels in ((M.ICOS_TRI_SPHERE_SEED, 1.0, (0, 1, 2, 3)), (M.CUBED_SPHERE_SEED, 1.0, (0, 1, 2, 3)),
                              (M.QUAD_RECT_SEED, 3.0, (0, 1, 2, 3)), (M.TRI_HEX_SEED, 1.0, (0, 1, 2)), (M.BETA_PLANE_SEED, 1.0, (0, 1, 2))):
        for L in levels:
            prog = F.Program(files, ignore_calls=off_path)
            mesh = F.Obj(_type="polymesh2d", particles=F.Obj(_type="particles"), edges=F.Obj(_type="edges"), faces=F.Obj(_type="faces"))
            prog.call("newPrivate", mesh, int(seed), L, L, 0, float(amp), file="src/PolyMesh2d.f90")
            p, fc = mesh.particles, mesh.faces
            n = p.n
            col = lambda name: np.array(p.f[name].tolist()[:n]) if p.f.get(name) is not None else np.zeros(n)
            leaf = [i for i in range(1, fc.n + 1) if not fc.hasChildren.get(i)]
            verts = np.array([[fc.vertices.get(r, i) for r in range(1, fc.vertices.n1 + 1)] for i in leaf], dtype=np.int32)
            save(f"mesh_seed{int(seed)}_L{L}", prog.where("newPrivate", "src/PolyMesh2d.f90"), seed=int(seed), level=L, amp=amp, n=n,
                 n_faces=fc.n, n_edges=mesh.edges.n, x=col("x"), y=col("y"), z=col("z"), area=col("area"),
                 mask=np.array(p.isActive.tolist()[:n], dtype=bool), max_edge_length=prog.call("MaxEdgeLength", mesh.edges, p),
                 leaf_face_vertices=verts, leaf_face_center=np.array([fc.centerParticle.get(i) for i in leaf], dtype=np.int32))
    print(f"done in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
