"""ctypes binding of the CPU oracle (oracle/lpm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() as the
checker and by bench.py's cpu_baseline / --impl reference legs.  Nothing in
lpm_v2_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_d = C.POINTER(C.c_double)
_i32 = C.POINTER(C.c_int32)
_n = C.c_int64
_dbl = C.c_double


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_libs = {}


def get(fast=False):
    key = "fast" if fast else "parity"
    if key not in _libs:
        lib = _load("liblpm_oracle_fast.so" if fast else "liblpm_oracle.so")
        _declare(lib)
        _libs[key] = lib
    return _libs[key]


def _declare(lib):
    vel3 = [_n, _d, _d, _d, _d, _d, _i32, _dbl, _n, _n, _d, _d, _d]
    for nm in ("oracle_bve_velocity", "oracle_bve_velocity_mesh", "oracle_bve_velocity_ld"):
        getattr(lib, nm).argtypes = vel3
        getattr(lib, nm).restype = None
    for nm in ("oracle_bve_stream", "oracle_bve_stream_ld"):
        getattr(lib, nm).argtypes = [_n, _d, _d, _d, _d, _d, _d, _i32, _dbl, _n, _n, _d, _d]
        getattr(lib, nm).restype = None
    for nm in ("oracle_plane_velocity", "oracle_plane_velocity_ld", "oracle_betaplane_velocity",
               "oracle_betaplane_velocity_ld"):
        getattr(lib, nm).argtypes = [_n, _d, _d, _d, _d, _i32, _n, _n, _d, _d]
        getattr(lib, nm).restype = None
    for nm in ("oracle_plane_stream", "oracle_plane_stream_ld"):
        getattr(lib, nm).argtypes = [_n, _d, _d, _d, _d, _i32, _n, _n, _d]
        getattr(lib, nm).restype = None
    for nm in ("oracle_betaplane_stream", "oracle_betaplane_stream_ld"):
        getattr(lib, nm).argtypes = [_n, _d, _d, _d, _d, _d, _i32, _n, _n, _d, _d]
        getattr(lib, nm).restype = None
    for nm in ("oracle_pse_laplacian_sphere", "oracle_pse_laplacian_sphere_ld"):
        getattr(lib, nm).argtypes = [_n, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _n, _n, _d]
        getattr(lib, nm).restype = None
    lib.oracle_pse_laplacian_plane.argtypes = [_n, _d, _d, _d, _d, _i32, _dbl, _n, _n, _d]
    lib.oracle_pse_laplacian_plane.restype = None
    lib.oracle_pse_laplacian_kernel8.argtypes = [_dbl]
    lib.oracle_pse_laplacian_kernel8.restype = _dbl
    lib.oracle_bve_rk4_step.argtypes = [_n, _d, _d, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _dbl]
    lib.oracle_bve_rk4_step.restype = None
    lib.oracle_plane_rk4_step.argtypes = [_n, _d, _d, _d, _d, _d, _d, _i32, _dbl]
    lib.oracle_plane_rk4_step.restype = None
    lib.oracle_betaplane_rk4_step.argtypes = [_n, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl]
    lib.oracle_betaplane_rk4_step.restype = None
    lib.oracle_total_ke.argtypes = [_n, _d, _d, _d, _d, _i32]
    lib.oracle_total_ke.restype = _dbl
    lib.oracle_total_enstrophy.argtypes = [_n, _d, _d, _i32]
    lib.oracle_total_enstrophy.restype = _dbl
    lib.oracle_load_balance.argtypes = [C.c_int32, C.c_int32, _i32, _i32, _i32]
    lib.oracle_load_balance.restype = None
    lib.oracle_active_list.argtypes = [_n, _i32, _i32]
    lib.oracle_active_list.restype = C.c_int64
    lib.oracle_bve_velocity_mt.argtypes = [C.c_int, _n, _d, _d, _d, _d, _d, _i32, _dbl, _n, _n, _d, _d, _d]
    lib.oracle_bve_velocity_mt.restype = None
    lib.oracle_set_threads.argtypes = [C.c_int]
    lib.oracle_set_threads.restype = None
    lib.oracle_get_threads.restype = C.c_int


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_d)


def _m(mask):
    m = np.asarray(mask)
    if m.dtype != np.int32:
        m = (m != 0).astype(np.int32)
    return np.ascontiguousarray(m)


def _rng(n, rng):
    if rng is None:
        return 0, n
    return int(rng[0]), int(rng[1])


def load_balance(n, nprocs):
    s, e, m = (np.zeros(nprocs, np.int32) for _ in range(3))
    get().oracle_load_balance(n, nprocs, s.ctypes.data_as(_i32), e.ctypes.data_as(_i32), m.ctypes.data_as(_i32))
    return s, e, m


def active_list(mask):
    m = _m(mask)
    out = np.zeros(m.size, np.int32)
    c = get().oracle_active_list(m.size, m.ctypes.data_as(_i32), out.ctypes.data_as(_i32))
    return out[:c].copy()


def bve_velocity(x, y, z, relvort, area, mask, radius=1.0, rng=None, variant=""):
    x, y, z, relvort, area = map(_f, (x, y, z, relvort, area))
    m = _m(mask)
    n = x.size
    b, e = _rng(n, rng)
    u, v, w = (np.zeros(n) for _ in range(3))
    fn = getattr(get(), "oracle_bve_velocity" + variant)
    fn(n, _p(x), _p(y), _p(z), _p(relvort), _p(area), m.ctypes.data_as(_i32), radius, b, e, _p(u), _p(v), _p(w))
    return u, v, w


def bve_stream(x, y, z, relvort, absvort, area, mask, radius=1.0, rng=None, variant=""):
    x, y, z, relvort, absvort, area = map(_f, (x, y, z, relvort, absvort, area))
    m = _m(mask)
    n = x.size
    b, e = _rng(n, rng)
    rs, as_ = np.zeros(n), np.zeros(n)
    fn = getattr(get(), "oracle_bve_stream" + variant)
    fn(n, _p(x), _p(y), _p(z), _p(relvort), _p(absvort), _p(area), m.ctypes.data_as(_i32), radius, b, e, _p(rs), _p(as_))
    return rs, as_


def _vel2(name, x, y, q, area, mask, rng, variant):
    x, y, q, area = map(_f, (x, y, q, area))
    m = _m(mask)
    n = x.size
    b, e = _rng(n, rng)
    u, v = np.zeros(n), np.zeros(n)
    getattr(get(), name + variant)(n, _p(x), _p(y), _p(q), _p(area), m.ctypes.data_as(_i32), b, e, _p(u), _p(v))
    return u, v


def plane_velocity(x, y, vort, area, mask, rng=None, variant=""):
    return _vel2("oracle_plane_velocity", x, y, vort, area, mask, rng, variant)


def betaplane_velocity(x, y, relvort, area, mask, rng=None, variant=""):
    return _vel2("oracle_betaplane_velocity", x, y, relvort, area, mask, rng, variant)


def plane_stream(x, y, vort, area, mask, rng=None, variant=""):
    x, y, vort, area = map(_f, (x, y, vort, area))
    m = _m(mask)
    n = x.size
    b, e = _rng(n, rng)
    psi = np.zeros(n)
    getattr(get(), "oracle_plane_stream" + variant)(n, _p(x), _p(y), _p(vort), _p(area), m.ctypes.data_as(_i32), b, e, _p(psi))
    return psi


def betaplane_stream(x, y, relvort, absvort, area, mask, rng=None, variant=""):
    x, y, relvort, absvort, area = map(_f, (x, y, relvort, absvort, area))
    m = _m(mask)
    n = x.size
    b, e = _rng(n, rng)
    rs, as_ = np.zeros(n), np.zeros(n)
    getattr(get(), "oracle_betaplane_stream" + variant)(n, _p(x), _p(y), _p(relvort), _p(absvort), _p(area),
                                                         m.ctypes.data_as(_i32), b, e, _p(rs), _p(as_))
    return rs, as_


def pse_laplacian_sphere(x, y, z, f, area, mask, eps, sphere_radius=1.0, rng=None, variant=""):
    x, y, z, f, area = map(_f, (x, y, z, f, area))
    m = _m(mask)
    n = x.size
    b, e = _rng(n, rng)
    lap = np.zeros(n)
    getattr(get(), "oracle_pse_laplacian_sphere" + variant)(n, _p(x), _p(y), _p(z), _p(f), _p(area),
                                                             m.ctypes.data_as(_i32), eps, sphere_radius, b, e, _p(lap))
    return lap


def pse_laplacian_plane(x, y, f, area, mask, eps, rng=None):
    x, y, f, area = map(_f, (x, y, f, area))
    m = _m(mask)
    n = x.size
    b, e = _rng(n, rng)
    lap = np.zeros(n)
    get().oracle_pse_laplacian_plane(n, _p(x), _p(y), _p(f), _p(area), m.ctypes.data_as(_i32), eps, b, e, _p(lap))
    return lap


def bve_rk4_step(x, y, z, relvort, u, v, w, area, mask, radius, omega, dt):
    """In-place on copies; returns the new (x, y, z, relvort, u, v, w)."""
    st = [np.array(a, dtype=np.float64, copy=True) for a in (x, y, z, relvort, u, v, w)]
    area = _f(area)
    m = _m(mask)
    get().oracle_bve_rk4_step(st[0].size, *[_p(a) for a in st], _p(area), m.ctypes.data_as(_i32), radius, omega, dt)
    return st


def plane_rk4_step(x, y, vort, u, v, area, mask, dt):
    xx, yy, uu, vv = (np.array(a, dtype=np.float64, copy=True) for a in (x, y, u, v))
    vort, area = _f(vort), _f(area)
    m = _m(mask)
    get().oracle_plane_rk4_step(xx.size, _p(xx), _p(yy), _p(vort), _p(uu), _p(vv), _p(area), m.ctypes.data_as(_i32), dt)
    return xx, yy, uu, vv


def betaplane_rk4_step(x, y, relvort, u, v, area, mask, beta, dt):
    xx, yy, zz, uu, vv = (np.array(a, dtype=np.float64, copy=True) for a in (x, y, relvort, u, v))
    area = _f(area)
    m = _m(mask)
    get().oracle_betaplane_rk4_step(xx.size, _p(xx), _p(yy), _p(zz), _p(uu), _p(vv), _p(area),
                                    m.ctypes.data_as(_i32), beta, dt)
    return xx, yy, zz, uu, vv


def total_ke(u, v, w, area, mask):
    u, v, w, area = map(_f, (u, v, w, area))
    m = _m(mask)
    return get().oracle_total_ke(u.size, _p(u), _p(v), _p(w), _p(area), m.ctypes.data_as(_i32))


def total_enstrophy(relvort, area, mask):
    relvort, area = _f(relvort), _f(area)
    m = _m(mask)
    return get().oracle_total_enstrophy(relvort.size, _p(relvort), _p(area), m.ctypes.data_as(_i32))


def bve_velocity_mt(nthreads, x, y, z, relvort, area, mask, radius, tbeg, tend, fast=True):
    """Targets [tbeg, tend) over `nthreads` workers on the LoadBalance split.  fast=True: the timing build
    (bench.py's CPU legs); fast=False: the parity build (-O2, no contraction) -- bit-identical to the
    single-threaded oracle_bve_velocity, because each target is summed by one worker in j order."""
    x, y, z, relvort, area = map(_f, (x, y, z, relvort, area))
    m = _m(mask)
    n = x.size
    u, v, w = (np.zeros(n) for _ in range(3))
    get(fast=fast).oracle_bve_velocity_mt(nthreads, n, _p(x), _p(y), _p(z), _p(relvort), _p(area),
                                          m.ctypes.data_as(_i32), radius, tbeg, tend, _p(u), _p(v), _p(w))
    return u, v, w


def set_threads(nthreads):
    """Host threads for the target loops of the PARITY build (OpenMP over targets; results do not depend on it)."""
    get().oracle_set_threads(int(nthreads))


class threads:
    """with oracle.threads(8): ...  -- parity-build oracle calls inside run their target loops on 8 threads."""

    def __init__(self, n):
        self.n = n

    def __enter__(self):
        self.old = get().oracle_get_threads()
        set_threads(self.n)

    def __exit__(self, *a):
        set_threads(self.old)


# ---- remaining PSE operators ---------------------------------------------------------
def _declare_pse_ops(lib):
    lib.oracle_pse_interpolate.argtypes = [C.c_int, _n, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _n, _d, _d, _d, _d]
    lib.oracle_pse_gradient_plane.argtypes = [_n, _d, _d, _d, _d, _i32, _dbl, _n, _n, _d, _d]
    lib.oracle_pse_gradient_sphere.argtypes = [_n, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _n, _n, _d, _d, _d]
    lib.oracle_pse_second_partials_plane.argtypes = [_n, _d, _d, _d, _d, _d, _i32, _dbl, _n, _n, _d, _d, _d]
    lib.oracle_pse_double_dot_plane.argtypes = [_n, _d, _d, _d, _d, _d, _i32, _dbl, _n, _n, _d]
    lib.oracle_pse_double_dot_sphere.argtypes = [_n, _d, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _n, _n, _d]
    lib.oracle_pse_divergence_sphere.argtypes = [_n, _d, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _n, _n, _d]
    for nm in ("oracle_pse_interpolate", "oracle_pse_gradient_plane", "oracle_pse_gradient_sphere",
               "oracle_pse_second_partials_plane", "oracle_pse_double_dot_plane", "oracle_pse_double_dot_sphere",
               "oracle_pse_divergence_sphere"):
        getattr(lib, nm).restype = None
    return lib


def _L():
    lib = get()
    if not getattr(lib, "_pse_ops", False):
        _declare_pse_ops(lib)
        lib._pse_ops = True
    return lib


def pse_interpolate(x, y, z, f, area, mask, eps, tx, ty, tz=None, sphere_radius=1.0):
    sphere = tz is not None
    x, y, f, area, tx, ty = map(_f, (x, y, f, area, tx, ty))
    z = _f(z) if sphere else np.zeros_like(x)
    tz = _f(tz) if sphere else np.zeros_like(tx)
    m = _m(mask)
    out = np.zeros(tx.size)
    _L().oracle_pse_interpolate(1 if sphere else 0, x.size, _p(x), _p(y), _p(z), _p(f), _p(area), m.ctypes.data_as(_i32),
                                eps, sphere_radius, tx.size, _p(tx), _p(ty), _p(tz), _p(out))
    return out


def pse_gradient_plane(x, y, f, area, mask, eps):
    x, y, f, area = map(_f, (x, y, f, area))
    m = _m(mask)
    g = [np.zeros(x.size) for _ in range(2)]
    _L().oracle_pse_gradient_plane(x.size, _p(x), _p(y), _p(f), _p(area), m.ctypes.data_as(_i32), eps, 0, x.size, *[_p(a) for a in g])
    return g


def pse_gradient_sphere(x, y, z, f, area, mask, eps, sphere_radius=1.0):
    x, y, z, f, area = map(_f, (x, y, z, f, area))
    m = _m(mask)
    g = [np.zeros(x.size) for _ in range(3)]
    _L().oracle_pse_gradient_sphere(x.size, _p(x), _p(y), _p(z), _p(f), _p(area), m.ctypes.data_as(_i32), eps,
                                    sphere_radius, 0, x.size, *[_p(a) for a in g])
    return g


def pse_second_partials_plane(x, y, gx, gy, area, mask, eps):
    x, y, gx, gy, area = map(_f, (x, y, gx, gy, area))
    m = _m(mask)
    o = [np.zeros(x.size) for _ in range(3)]
    _L().oracle_pse_second_partials_plane(x.size, _p(x), _p(y), _p(gx), _p(gy), _p(area), m.ctypes.data_as(_i32), eps,
                                          0, x.size, *[_p(a) for a in o])
    return o


def pse_double_dot_plane(x, y, u, v, area, mask, eps):
    x, y, u, v, area = map(_f, (x, y, u, v, area))
    m = _m(mask)
    dd = np.zeros(x.size)
    _L().oracle_pse_double_dot_plane(x.size, _p(x), _p(y), _p(u), _p(v), _p(area), m.ctypes.data_as(_i32), eps, 0, x.size, _p(dd))
    return dd


def pse_double_dot_sphere(x, y, z, u, v, w, area, mask, eps, sphere_radius=1.0):
    x, y, z, u, v, w, area = map(_f, (x, y, z, u, v, w, area))
    m = _m(mask)
    dd = np.zeros(x.size)
    _L().oracle_pse_double_dot_sphere(x.size, _p(x), _p(y), _p(z), _p(u), _p(v), _p(w), _p(area), m.ctypes.data_as(_i32),
                                      eps, sphere_radius, 0, x.size, _p(dd))
    return dd


def pse_divergence_sphere(x, y, z, u, v, w, area, mask, eps, sphere_radius=1.0):
    x, y, z, u, v, w, area = map(_f, (x, y, z, u, v, w, area))
    m = _m(mask)
    div = np.zeros(x.size)
    _L().oracle_pse_divergence_sphere(x.size, _p(x), _p(y), _p(z), _p(u), _p(v), _p(w), _p(area), m.ctypes.data_as(_i32),
                                      eps, sphere_radius, 0, x.size, _p(div))
    return div


def swe_plane_rhs(x, y, vort, div, surf, area, mask, pse_eps):
    x, y, vort, div, surf, area = map(_f, (x, y, vort, div, surf, area))
    m = _m(mask)
    lib = get()
    lib.oracle_swe_plane_rhs.argtypes = [_n, _d, _d, _d, _d, _d, _d, _i32, _dbl, _n, _n, _d, _d, _d, _d]
    lib.oracle_swe_plane_rhs.restype = None
    o = [np.zeros(x.size) for _ in range(4)]
    lib.oracle_swe_plane_rhs(x.size, _p(x), _p(y), _p(vort), _p(div), _p(surf), _p(area), m.ctypes.data_as(_i32),
                             pse_eps, 0, x.size, *[_p(a) for a in o])
    return o


ORACLE_TOPO_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)


def swe_plane_rk4_step(x, y, relvort, div, h, area, u, v, double_dot, lap_surf, mask, f0, beta, g, pse_eps, dt, topo=None):
    """src/SWEPlaneSolver.f90:298-429 as written.  Returns the new (x, y, relvort, div, h, area, u, v, double_dot,
    lap_surf); `topo` is a Python callable topo(x, y) -> float or None (flat bottom)."""
    arrs = [np.array(a, dtype=np.float64, copy=True) for a in (x, y, relvort, div, h, area, u, v, double_dot, lap_surf)]
    m = _m(mask)
    lib = get()
    lib.oracle_swe_plane_rk4_step.argtypes = [_n] + [_d] * 10 + [_i32] + [_dbl] * 5 + [ORACLE_TOPO_FN]
    lib.oracle_swe_plane_rk4_step.restype = None
    cb = ORACLE_TOPO_FN(topo) if topo is not None else C.cast(None, ORACLE_TOPO_FN)
    lib.oracle_swe_plane_rk4_step(arrs[0].size, *[_p(a) for a in arrs], m.ctypes.data_as(_i32), f0, beta, g, pse_eps, dt, cb)
    return arrs


def swe_plane_velocity(x, y, vort, div, area, mask):
    x, y, vort, div, area = map(_f, (x, y, vort, div, area))
    m = _m(mask)
    lib = get()
    lib.oracle_swe_plane_velocity.argtypes = [_n, _d, _d, _d, _d, _d, _i32, _n, _n, _d, _d]
    lib.oracle_swe_plane_velocity.restype = None
    o = [np.zeros(x.size) for _ in range(2)]
    lib.oracle_swe_plane_velocity(x.size, _p(x), _p(y), _p(vort), _p(div), _p(area), m.ctypes.data_as(_i32), 0, x.size,
                                  *[_p(a) for a in o])
    return o


def swe_sphere_rhs(x, y, z, vort, div, surf, area, mask, radius, pse_eps):
    x, y, z, vort, div, surf, area = map(_f, (x, y, z, vort, div, surf, area))
    m = _m(mask)
    lib = get()
    lib.oracle_swe_sphere_rhs.argtypes = [_n, _d, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _n, _n, _d, _d, _d, _d, _d]
    lib.oracle_swe_sphere_rhs.restype = None
    o = [np.zeros(x.size) for _ in range(5)]
    lib.oracle_swe_sphere_rhs(x.size, _p(x), _p(y), _p(z), _p(vort), _p(div), _p(surf), _p(area), m.ctypes.data_as(_i32),
                              radius, pse_eps, 0, x.size, *[_p(a) for a in o])
    return o


def pse_laplacian_sphere_at_points(x, y, z, f, area, mask, eps, tx, ty, tz, ftarget, sphere_radius=1.0):
    x, y, z, f, area, tx, ty, tz, ftarget = map(_f, (x, y, z, f, area, tx, ty, tz, ftarget))
    m = _m(mask)
    lib = get()
    lib.oracle_pse_laplacian_sphere_at_points.argtypes = [_n, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _n, _d, _d, _d, _d, _d]
    lib.oracle_pse_laplacian_sphere_at_points.restype = None
    out = np.zeros(tx.size)
    lib.oracle_pse_laplacian_sphere_at_points(x.size, _p(x), _p(y), _p(z), _p(f), _p(area), m.ctypes.data_as(_i32),
                                              eps, sphere_radius, tx.size, _p(tx), _p(ty), _p(tz), _p(ftarget), _p(out))
    return out
