"""Host-side mirror of the reference's solver interface for the hot path.

The reference is Fortran: its drivers call ``New(solver, mesh)``,
``Timestep(solver, mesh, dt)``, ``Delete(solver)`` (src/SphereBVESolver.f90:35-36,
80-90; src/PlaneIncompressibleSolver.f90:34-35; src/BetaPlaneSolver.f90:34-35),
``SetVelocityOnMesh`` / ``SetStreamFunctionsOnMesh`` (src/SphereBVE.f90:110-113)
and ``PSE{Plane,Sphere}LaplacianAtParticles`` (src/PSEDirectSum.f90:467,502).
These classes expose the same operations with the same argument meaning, over
the C ABI; the Fortran shim in lpm_v2_b200/fortran/ binds the same entry points.
Errors follow the reference's convention (log, do not abort) one level up: a
failed call raises LpmError carrying lpm_gpu_last_error().
"""
import ctypes as C

import numpy as np

from . import api
from ._lib import lib, check

_d = C.POINTER(C.c_double)
_i32 = C.POINTER(C.c_int32)


def _pd(a):
    return a.ctypes.data_as(_d) if a is not None else None


class MPISetup:
    """type MPISetup, src/MPISetup.f90:36-43: indexStart/indexEnd/messageLength
    (1-based inclusive), filled by LoadBalance (:132-146)."""

    def __init__(self, n_items, n_procs):
        self.n = int(n_items)
        self.indexStart, self.indexEnd, self.messageLength = api.load_balance(n_items, n_procs)

    def slice0(self, rank):
        """0-based half-open [ibeg, iend) of `rank`."""
        return int(self.indexStart[rank] - 1), int(self.indexEnd[rank])


class BVEMesh:
    """The fields of type BVEMesh the hot path touches (src/SphereBVE.f90:71-86)."""

    def __init__(self, mesh, relvort, radius=1.0, rotation_rate=0.0, absvort=None):
        self.mesh = mesh
        self.x, self.y, self.z = mesh.x.copy(), mesh.y.copy(), mesh.z.copy()
        self.area = mesh.area
        self.is_active = mesh.is_active
        self.radius = float(radius)
        self.rotationRate = float(rotation_rate)
        self.relVort = np.array(relvort, dtype=np.float64)
        self.absVort = (np.array(absvort, dtype=np.float64) if absvort is not None
                        else self.relVort + 2.0 * self.rotationRate * mesh.z / self.radius)
        n = mesh.n
        self.velocity = [np.zeros(n), np.zeros(n), np.zeros(n)]
        self.relStream = np.zeros(n)
        self.absStream = np.zeros(n)

    def SetVelocityOnMesh(self):
        """src/SphereBVE.f90:489-531."""
        self.velocity = list(api.bve_velocity(self.x, self.y, self.z, self.relVort, self.area, self.is_active, self.radius))

    def SetStreamFunctionsOnMesh(self):
        """src/SphereBVE.f90:445-485."""
        self.relStream, self.absStream = api.bve_stream(self.x, self.y, self.z, self.relVort, self.absVort,
                                                        self.area, self.is_active, self.radius)


class BVESolver:
    """type BVESolver: New / Timestep / Delete (src/SphereBVESolver.f90)."""

    def __init__(self, sphere: BVEMesh):            # New(solver, sphereBVE), :112-168
        self._h = C.c_void_p()
        self.n = sphere.mesh.n
        m = np.ascontiguousarray((np.asarray(sphere.is_active) != 0).astype(np.int32))
        u, v, w = (np.ascontiguousarray(a, dtype=np.float64) for a in sphere.velocity)
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (sphere.x, sphere.y, sphere.z, sphere.relVort, sphere.absVort)]
        area = np.ascontiguousarray(sphere.area, dtype=np.float64)
        check(lib.lpm_bve_solver_new(self.n, *[_pd(a) for a in arrs], _pd(u), _pd(v), _pd(w), _pd(area),
                                     m.ctypes.data_as(_i32), sphere.radius, sphere.rotationRate, C.byref(self._h)))

    def Timestep(self, sphere: BVEMesh, dt, with_stream=True, copy_back=True):   # :219-353
        check(lib.lpm_bve_solver_timestep(self._h, float(dt), 1 if with_stream else 0))
        if copy_back:
            self.CopyToMesh(sphere, with_stream)

    def CopyToMesh(self, sphere: BVEMesh, with_stream=True):
        """:333-336, :348-350: positions, vorticity and velocity back to the mesh."""
        u, v, w = sphere.velocity
        rs = sphere.relStream if with_stream else None
        as_ = sphere.absStream if with_stream else None
        check(lib.lpm_bve_solver_get_state(self._h, _pd(sphere.x), _pd(sphere.y), _pd(sphere.z), _pd(sphere.relVort),
                                           _pd(u), _pd(v), _pd(w), _pd(rs), _pd(as_)))

    def Diagnostics(self):
        """TotalKE, TotalEnstrophy (src/SphereBVE.f90:410-441)."""
        ke, en = C.c_double(0), C.c_double(0)
        check(lib.lpm_bve_solver_diagnostics(self._h, C.byref(ke), C.byref(en)))
        return ke.value, en.value

    def Delete(self):                               # :172-213
        if self._h:
            check(lib.lpm_bve_solver_delete(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Delete()
        except Exception:
            pass


class PlaneMeshIncompressible:
    """src/PlanarIncompressible.f90 fields on the hot path."""

    def __init__(self, mesh, vorticity):
        self.mesh = mesh
        self.x, self.y = mesh.x.copy(), mesh.y.copy()
        self.area, self.is_active = mesh.area, mesh.is_active
        self.vorticity = np.array(vorticity, dtype=np.float64)
        self.velocity = [np.zeros(mesh.n), np.zeros(mesh.n)]
        self.streamFn = np.zeros(mesh.n)

    def SetVelocityOnMesh(self):                    # :426-466
        self.velocity = list(api.plane_velocity(self.x, self.y, self.vorticity, self.area, self.is_active))

    def SetStreamFunctionOnMesh(self):              # :470-505
        self.streamFn = api.plane_stream(self.x, self.y, self.vorticity, self.area, self.is_active)


class PlaneSolver:
    """type PlaneSolver (src/PlaneIncompressibleSolver.f90:37-59)."""

    def __init__(self, plane: PlaneMeshIncompressible):
        self._h = C.c_void_p()
        self.n = plane.mesh.n
        m = np.ascontiguousarray((np.asarray(plane.is_active) != 0).astype(np.int32))
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (plane.x, plane.y, plane.vorticity, plane.velocity[0], plane.velocity[1], plane.area)]
        check(lib.lpm_plane_solver_new(self.n, *[_pd(a) for a in arrs], m.ctypes.data_as(_i32), C.byref(self._h)))

    def Timestep(self, plane, dt, with_stream=True):     # :171-259
        check(lib.lpm_plane_solver_timestep(self._h, float(dt), 1 if with_stream else 0))
        check(lib.lpm_plane_solver_get_state(self._h, _pd(plane.x), _pd(plane.y), _pd(plane.velocity[0]),
                                             _pd(plane.velocity[1]), _pd(plane.streamFn) if with_stream else None))

    def Delete(self):
        if self._h:
            check(lib.lpm_plane_solver_delete(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Delete()
        except Exception:
            pass


class BetaPlaneMesh:
    """src/BetaPlane.f90:42-57 fields on the hot path."""

    def __init__(self, mesh, relvort, f0=0.0, beta=0.0):
        self.mesh = mesh
        self.x, self.y = mesh.x.copy(), mesh.y.copy()
        self.area, self.is_active = mesh.area, mesh.is_active
        self.f0, self.beta = float(f0), float(beta)
        self.relVort = np.array(relvort, dtype=np.float64)
        self.absVort = self.relVort + self.f0 + self.beta * mesh.y     # :259
        self.velocity = [np.zeros(mesh.n), np.zeros(mesh.n)]
        self.relStream, self.absStream = np.zeros(mesh.n), np.zeros(mesh.n)

    def SetVelocityOnMesh(self):                    # :359-397
        self.velocity = list(api.betaplane_velocity(self.x, self.y, self.relVort, self.area, self.is_active))

    def SetStreamFunctionsOnMesh(self):             # :399-442
        self.relStream, self.absStream = api.betaplane_stream(self.x, self.y, self.relVort, self.absVort,
                                                              self.area, self.is_active)


class BetaPlaneSolver:
    """type BetaPlaneSolver (src/BetaPlaneSolver.f90:36-56)."""

    def __init__(self, bp: BetaPlaneMesh):
        self._h = C.c_void_p()
        self.n = bp.mesh.n
        m = np.ascontiguousarray((np.asarray(bp.is_active) != 0).astype(np.int32))
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (bp.x, bp.y, bp.relVort, bp.absVort, bp.velocity[0], bp.velocity[1], bp.area)]
        check(lib.lpm_betaplane_solver_new(self.n, *[_pd(a) for a in arrs], m.ctypes.data_as(_i32), bp.beta,
                                           C.byref(self._h)))

    def Timestep(self, bp, dt, with_stream=True):   # :142-219
        check(lib.lpm_betaplane_solver_timestep(self._h, float(dt), 1 if with_stream else 0))
        check(lib.lpm_betaplane_solver_get_state(self._h, _pd(bp.x), _pd(bp.y), _pd(bp.relVort), _pd(bp.velocity[0]),
                                                 _pd(bp.velocity[1]), _pd(bp.relStream) if with_stream else None,
                                                 _pd(bp.absStream) if with_stream else None))

    def Delete(self):
        if self._h:
            check(lib.lpm_betaplane_solver_delete(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Delete()
        except Exception:
            pass


class PSE:
    """type PSE (src/PSEDirectSum.f90:56-60): eps = MaxEdgeLength ** pow, pow
    defaulting to 0.75 (New, :93-113)."""

    def __init__(self, mesh, radius_multiplier=None):
        pw = 0.75 if radius_multiplier is None else float(radius_multiplier)
        self.eps = mesh.max_edge_length ** pw

    def SphereLaplacianAtParticles(self, mesh, scalar_field, sphere_radius=1.0):    # :502-535
        return api.pse_laplacian_sphere(mesh.x, mesh.y, mesh.z, scalar_field, mesh.area, mesh.is_active,
                                        self.eps, sphere_radius)

    def PlaneLaplacianAtParticles(self, mesh, scalar_field):                        # :467-500
        return api.pse_laplacian_plane(mesh.x, mesh.y, scalar_field, mesh.area, mesh.is_active, self.eps)

    # the remaining operators of the module
    def SphereInterpolateScalar(self, mesh, scalar_field, tx, ty, tz, sphere_radius=1.0):   # :151-168, batched
        return api.pse_interpolate_sphere(mesh.x, mesh.y, mesh.z, scalar_field, mesh.area, mesh.is_active, self.eps,
                                          tx, ty, tz, sphere_radius)

    def PlaneInterpolateScalar(self, mesh, scalar_field, tx, ty):                           # :128-149
        return api.pse_interpolate_plane(mesh.x, mesh.y, scalar_field, mesh.area, mesh.is_active, self.eps, tx, ty)

    def SphereGradientAtParticles(self, mesh, scalar_field, sphere_radius=1.0):             # :221-267
        return api.pse_gradient_sphere(mesh.x, mesh.y, mesh.z, scalar_field, mesh.area, mesh.is_active, self.eps,
                                       sphere_radius)

    def PlaneGradientAtParticles(self, mesh, scalar_field):                                 # :180-218
        return api.pse_gradient_plane(mesh.x, mesh.y, scalar_field, mesh.area, mesh.is_active, self.eps)

    def PlaneSecondPartialsAtParticles(self, mesh, grad_x, grad_y):                         # :269-320
        return api.pse_second_partials_plane(mesh.x, mesh.y, grad_x, grad_y, mesh.area, mesh.is_active, self.eps)

    def PlaneDoubleDotProductAtParticles(self, mesh, u, v):                                 # :322-365
        return api.pse_double_dot_plane(mesh.x, mesh.y, u, v, mesh.area, mesh.is_active, self.eps)

    def SphereDoubleDotProductAtParticles(self, mesh, u, v, w, sphere_radius=1.0):          # :367-420
        return api.pse_double_dot_sphere(mesh.x, mesh.y, mesh.z, u, v, w, mesh.area, mesh.is_active, self.eps,
                                         sphere_radius)

    def SphereDivergenceAtParticles(self, mesh, u, v, w, sphere_radius=1.0):                # :537-579
        return api.pse_divergence_sphere(mesh.x, mesh.y, mesh.z, u, v, w, mesh.area, mesh.is_active, self.eps,
                                         sphere_radius)


def SWEPlaneRHSIntegrals(mesh, vort, div, h, topo_fn, pse_eps):
    """src/SWEPlaneSolver.f90:457-560: returns u, v, doubleDot, lapSurf.  `topo_fn(x, y)` is the
    bottom topography function the reference passes as a procedure argument."""
    surf = np.asarray(h, dtype=np.float64) + topo_fn(mesh.x, mesh.y)
    return api.swe_plane_rhs_integrals(mesh.x, mesh.y, vort, div, surf, mesh.area, mesh.is_active, pse_eps)


class SWEMeshPlane:
    """The fields of type SWEMesh (src/PlanarSWE.f90) that the planar shallow-water solver reads and writes."""

    def __init__(self, mesh, relVort, divergence, h, f0=0.0, beta=0.0, g=1.0, pseEps=None):
        self.mesh = mesh
        self.x, self.y = mesh.x.copy(), mesh.y.copy()
        self.area = np.array(mesh.area, dtype=np.float64)          # the SWE solver advances the areas too
        self.is_active = mesh.is_active
        self.relVort = np.array(relVort, dtype=np.float64)
        self.divergence = np.array(divergence, dtype=np.float64)
        self.h = np.array(h, dtype=np.float64)
        self.f0, self.beta, self.g = float(f0), float(beta), float(g)
        self.pseEps = float(pseEps if pseEps is not None else mesh.max_edge_length ** 0.75)
        self.velocity = [np.zeros(mesh.n), np.zeros(mesh.n)]
        self.doubleDot = np.zeros(mesh.n)
        self.lapSurf = np.zeros(mesh.n)


class SWEPlaneSolver:
    """type SWESolver: New / Timestep / Delete (src/SWEPlaneSolver.f90:137-205, 298-429).  `topoFn(x, y) -> float`
    is the bottom topography, as in the reference's interfaces (None = flat bottom)."""

    def __init__(self, plane: SWEMeshPlane, topoFn=None):        # New(solver, plane, topoFn)
        from ._lib import TOPOGRAPHY_FN
        self._h = C.c_void_p()
        self.n = plane.mesh.n
        self._cb = TOPOGRAPHY_FN(lambda x, y, user: float(topoFn(x, y))) if topoFn is not None else None
        m = np.ascontiguousarray((np.asarray(plane.is_active) != 0).astype(np.int32))
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (plane.x, plane.y, plane.relVort, plane.divergence, plane.h, plane.area)]
        cb = C.cast(self._cb, C.c_void_p) if self._cb is not None else None
        check(lib.lpm_swe_plane_solver_new(self.n, *[_pd(a) for a in arrs], m.ctypes.data_as(_i32), plane.f0, plane.beta,
                                           plane.g, plane.pseEps, cb, None, C.byref(self._h)))
        self.CopyToMesh(plane)

    def Timestep(self, plane: SWEMeshPlane, dt):                  # Timestep(solver, plane, dt, topoFn), :298-429
        check(lib.lpm_swe_plane_solver_timestep(self._h, float(dt)))
        self.CopyToMesh(plane)

    def CopyToMesh(self, plane: SWEMeshPlane):                    # :419-426
        check(lib.lpm_swe_plane_solver_get_state(self._h, _pd(plane.x), _pd(plane.y), _pd(plane.relVort), _pd(plane.divergence),
                                                 _pd(plane.h), _pd(plane.area), _pd(plane.velocity[0]), _pd(plane.velocity[1]),
                                                 _pd(plane.doubleDot), _pd(plane.lapSurf)))

    def Delete(self):
        if self._h:
            check(lib.lpm_swe_plane_solver_delete(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Delete()
        except Exception:
            pass
