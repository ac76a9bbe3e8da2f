"""Synthetic inputs: the initial conditions of the reference's drivers, evaluated
with numpy on a PolyMesh2d (input generation only -- no direct sum happens here).

Each function cites the driver routine that defines the field.  Meshes are at
t = 0, so Lagrangian coordinates equal physical coordinates.
"""
import numpy as np

PI = 3.1415926535897932384626433832795027975     # src/TypeDefs.f90:31


def atan4(y, x):
    """Longitude in [0, 2 pi), src/SphereGeometry.f90:513-548."""
    t = np.arctan2(np.abs(y), np.abs(x))
    out = np.where((x > 0) & (y > 0), t, 0.0)
    out = np.where((x < 0) & (y > 0), PI - t, out)
    out = np.where((x < 0) & (y < 0), PI + t, out)
    out = np.where((x > 0) & (y < 0), 2.0 * PI - t, out)
    out = np.where((x == 0) & (y > 0), PI / 2.0, out)
    out = np.where((x == 0) & (y < 0), 3.0 * PI / 2.0, out)
    out = np.where((y == 0) & (x < 0), PI, out)
    return out


def latitude(x, y, z):
    """src/SphereGeometry.f90:498-502."""
    return np.arctan2(z, np.sqrt(x * x + y * y))


def gaussian_vortex(mesh, radius=1.0, shape_param=4.0, vort_strength=12.566370614359172,
                    init_lon=0.0, init_lat=0.157079632679490):
    """Config 1: examples/BVESingleGaussianVortex.f90:120-124, 335-357 with
    examples/bveSingleGaussVort.namelist.  Two passes: the Gaussian with C = 0,
    then C = sum_active(zeta A) / (4 pi R) subtracted."""
    c = radius * np.array([np.cos(init_lon) * np.cos(init_lat), np.sin(init_lon) * np.cos(init_lat), np.sin(init_lat)])
    g = vort_strength * np.exp(-shape_param * shape_param * (radius * radius - mesh.x * c[0] - mesh.y * c[1] - mesh.z * c[2]))
    act = mesh.is_active != 0
    const = np.sum(g[act] * mesh.area[act]) / (4.0 * PI * radius)
    return g - const


def rossby_haurwitz54(mesh, radius=1.0, amplitude=1.0, zonal_wind=0.0):
    """Config 4: examples/RossbyHaurwitz54.f90:413-428 with examples/rh54.namelist."""
    lon = atan4(mesh.y, mesh.x)
    zz = mesh.z / radius
    leg = zz * (zz * zz - 1.0) * (zz * zz - 1.0)
    return 2.0 * zonal_wind * mesh.z / radius + 30.0 * amplitude * np.cos(4.0 * lon) * leg / radius


def solid_body(mesh, radius=1.0, omega=2.0 * PI):
    """examples/BVESolidBody.f90:231-243: zeta = 2 Omega z / R, u = Omega (-y, x, 0)."""
    zeta = 2.0 * omega * mesh.z / radius
    return zeta, (-omega * mesh.y, omega * mesh.x, np.zeros_like(mesh.x))


def abs_vorticity(mesh, relvort, rotation_rate, radius=1.0):
    """src/SphereBVE.f90:346."""
    return relvort + 2.0 * rotation_rate * mesh.z / radius


def spherical_harmonic54(mesh):
    """tests/SpherePSEConvTest.f90:495-506 (and its exact Laplacian, -30 Y)."""
    lat = latitude(mesh.x, mesh.y, mesh.z)
    lon = atan4(mesh.y, mesh.x)
    return 3.0 * np.sqrt(35.0) * np.cos(4.0 * lon) * np.sin(lat) * (-1.0 + np.sin(lat) * np.sin(lat)) ** 2


def lamb_dipole(x, y, lamb_r, u0, xc, yc):
    """examples/CollidingDipoles.f90:391-413 (Bessel functions from scipy
    instead of the driver's BESSJ0/BESSJ1 rational approximations)."""
    from scipy.special import j0, j1
    k0 = 3.8317
    r = np.sqrt((x - xc) * (x - xc) + (y - yc) * (y - yc))
    inside = (r <= lamb_r) & (r >= 1.0e-10)
    rs = np.where(inside, r, 1.0)
    val = -2.0 * u0 * (k0 / lamb_r) * j1(k0 / lamb_r * rs) * (y / rs) / j0(k0)
    return np.where(inside, val, 0.0)


def colliding_dipoles(mesh):
    """Config 2: examples/CollidingDipoles.f90:376-382 with examples/collidingDipoles.namelist."""
    return lamb_dipole(mesh.x, mesh.y, 1.0, 1.0, -2.0, 0.0) + lamb_dipole(mesh.x, mesh.y, 1.0, -1.0, 2.0, 0.0)


def betaplane_gaussian(mesh, shape_param=16.0, vort_max=0.62831853071795864769, x0=0.5, y0=0.025):
    """examples/BetaPlaneGaussianVortex.f90:241-245 with examples/betaPlaneGaussVort.namelist."""
    return vort_max * np.exp(-shape_param * shape_param * ((mesh.x - x0) ** 2 + (mesh.y - y0) ** 2))
