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
    print(f"done in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
