"""numpy-facing wrappers of the C ABI (host buffers in, host buffers out).

These are the calls the reference's patched private kernels make through
ISO_C_BINDING (lpm_v2_b200/fortran/lpm_gpu.f90); the tests drive the same
entry points through ctypes.  Nothing here computes: every function forwards
to liblpmgpu.so and raises LpmError on a non-zero return.
"""
import ctypes as C

import numpy as np

from ._lib import lib, check, LpmError  # noqa: F401

_d = C.POINTER(C.c_double)
_i32 = C.POINTER(C.c_int32)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _pd(a):
    return a.ctypes.data_as(_d)


def _mask(m):
    """logical(klog) -> int32 (merge(1,0,mask) on the Fortran side)."""
    m = np.asarray(m)
    if m.dtype != np.int32:
        m = (m != 0).astype(np.int32)
    return np.ascontiguousarray(m)


def init(ndev=0):
    """lpm_gpu_init: claim `ndev` devices (0 = all) in single-process mode."""
    used = C.c_int(0)
    check(lib.lpm_gpu_init(int(ndev), C.byref(used)))
    return used.value


def init_rank(device):
    check(lib.lpm_gpu_init_rank(int(device)))


def finalize():
    check(lib.lpm_gpu_finalize())


def device_count():
    return lib.lpm_gpu_device_count()


def comm_unique_id():
    buf = C.create_string_buffer(128)
    check(lib.lpm_comm_unique_id(buf))
    return buf.raw


def comm_init_rank(world, rank, uid):
    check(lib.lpm_comm_init_rank(int(world), int(rank), C.create_string_buffer(uid, 128)))


def comm_alloc_shared(nbytes):
    """COLLECTIVE (rank mode): device memory every rank can store into over NVLink; returns the address."""
    p = C.c_void_p()
    check(lib.lpm_comm_alloc_shared(int(nbytes), C.byref(p)))
    return p.value


def comm_free_shared(ptr):
    check(lib.lpm_comm_free_shared(C.c_void_p(ptr)))


def comm_is_shared(ptr, nbytes):
    return bool(lib.lpm_comm_is_shared(C.c_void_p(ptr), int(nbytes)))


def load_balance(n_items, nprocs):
    """MPISetup%indexStart/indexEnd/messageLength (1-based, inclusive)."""
    s = np.zeros(nprocs, np.int64)
    e = np.zeros(nprocs, np.int64)
    m = np.zeros(nprocs, np.int64)
    p = C.POINTER(C.c_int64)
    check(lib.lpm_load_balance(int(n_items), int(nprocs), s.ctypes.data_as(p), e.ctypes.data_as(p), m.ctypes.data_as(p)))
    return s, e, m


def active_list(mask):
    m = _mask(mask)
    out = np.zeros(m.size, np.int32)
    cnt = C.c_int64(0)
    check(lib.lpm_active_list(m.size, m.ctypes.data_as(_i32), out.ctypes.data_as(_i32), C.byref(cnt)))
    return out[:cnt.value].copy()


def bve_velocity(x, y, z, relvort, area, mask, radius=1.0):
    x, y, z, relvort, area = map(_f64, (x, y, z, relvort, area))
    m = _mask(mask)
    n = x.size
    u, v, w = (np.empty(n) for _ in range(3))
    check(lib.lpm_bve_velocity(n, _pd(x), _pd(y), _pd(z), _pd(relvort), _pd(area), m.ctypes.data_as(_i32),
                               float(radius), _pd(u), _pd(v), _pd(w)))
    return u, v, w


def bve_stream(x, y, z, relvort, absvort, area, mask, radius=1.0):
    x, y, z, relvort, absvort, area = map(_f64, (x, y, z, relvort, absvort, area))
    m = _mask(mask)
    n = x.size
    rs, as_ = np.empty(n), np.empty(n)
    check(lib.lpm_bve_stream(n, _pd(x), _pd(y), _pd(z), _pd(relvort), _pd(absvort), _pd(area),
                             m.ctypes.data_as(_i32), float(radius), _pd(rs), _pd(as_)))
    return rs, as_


def plane_velocity(x, y, vort, area, mask):
    x, y, vort, area = map(_f64, (x, y, vort, area))
    m = _mask(mask)
    n = x.size
    u, v = np.empty(n), np.empty(n)
    check(lib.lpm_plane_velocity(n, _pd(x), _pd(y), _pd(vort), _pd(area), m.ctypes.data_as(_i32), _pd(u), _pd(v)))
    return u, v


def plane_stream(x, y, vort, area, mask):
    x, y, vort, area = map(_f64, (x, y, vort, area))
    m = _mask(mask)
    psi = np.empty(x.size)
    check(lib.lpm_plane_stream(x.size, _pd(x), _pd(y), _pd(vort), _pd(area), m.ctypes.data_as(_i32), _pd(psi)))
    return psi


def betaplane_velocity(x, y, relvort, area, mask):
    x, y, relvort, area = map(_f64, (x, y, relvort, area))
    m = _mask(mask)
    n = x.size
    u, v = np.empty(n), np.empty(n)
    check(lib.lpm_betaplane_velocity(n, _pd(x), _pd(y), _pd(relvort), _pd(area), m.ctypes.data_as(_i32), _pd(u), _pd(v)))
    return u, v


def betaplane_stream(x, y, relvort, absvort, area, mask):
    x, y, relvort, absvort, area = map(_f64, (x, y, relvort, absvort, area))
    m = _mask(mask)
    n = x.size
    rs, as_ = np.empty(n), np.empty(n)
    check(lib.lpm_betaplane_stream(n, _pd(x), _pd(y), _pd(relvort), _pd(absvort), _pd(area),
                                   m.ctypes.data_as(_i32), _pd(rs), _pd(as_)))
    return rs, as_


def pse_laplacian_sphere(x, y, z, f, area, mask, eps, sphere_radius=1.0):
    x, y, z, f, area = map(_f64, (x, y, z, f, area))
    m = _mask(mask)
    lap = np.empty(x.size)
    check(lib.lpm_pse_laplacian_sphere(x.size, _pd(x), _pd(y), _pd(z), _pd(f), _pd(area), m.ctypes.data_as(_i32),
                                       float(eps), float(sphere_radius), _pd(lap)))
    return lap


def pse_laplacian_plane(x, y, f, area, mask, eps):
    x, y, f, area = map(_f64, (x, y, f, area))
    m = _mask(mask)
    lap = np.empty(x.size)
    check(lib.lpm_pse_laplacian_plane(x.size, _pd(x), _pd(y), _pd(f), _pd(area), m.ctypes.data_as(_i32),
                                      float(eps), _pd(lap)))
    return lap


def fp64_peak_probe(iters=4096):
    tf, ms = C.c_double(0), C.c_double(0)
    check(lib.lpm_fp64_peak_probe(int(iters), C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def set_profiling(on):
    check(lib.lpm_set_profiling(1 if on else 0))


def set_symmetric(on):
    """Whole BVE evaluations of large particle sets pair-symmetrically (default) or always one-sided."""
    check(lib.lpm_set_symmetric(1 if on else 0))


def tune(key, value):
    """csrc/lpm_gpu_tuning.h: A/B and test knob, not part of the C ABI."""
    check(lib.lpm_tune(key.encode(), int(value)))


def set_pse_culling(mode):
    """0: reference order, no culling; 1 (default): cell order + tile culling; 2: cell order only."""
    check(lib.lpm_set_pse_culling(int(mode)))


def set_pse_series(on):
    """Sphere PSE kernels: series for theta^2 inside the cut-off (default) or atan2 always."""
    check(lib.lpm_set_pse_series(1 if on else 0))


def last_kernel_ms():
    ms = C.c_double(0)
    check(lib.lpm_last_kernel_ms(C.byref(ms)))
    return ms.value


def last_sum_ms():
    """Device time of the last one-shot direct sum as a whole (pack, sort, kernels; no host copies)."""
    ms = C.c_double(0)
    check(lib.lpm_last_sum_ms(C.byref(ms)))
    return ms.value


def launch_count(reset=False):
    return int(lib.lpm_launch_count(1 if reset else 0))


def profile_summary(reset=True):
    """(number of direct-sum main kernels, their total duration in ms) since the last reset."""
    k, ms = C.c_int64(0), C.c_double(0)
    check(lib.lpm_profile_summary(1 if reset else 0, C.byref(k), C.byref(ms)))
    return k.value, ms.value


PROFILE_TAGS = ("bve_velocity/one_sided", "bve_velocity/symmetric", "bve_stream/one_sided", "bve_stream/symmetric",
                "other/one_sided", "other/symmetric", "bve_velocity+stream/one_sided", "bve_velocity+stream/symmetric")


def profile_breakdown(reset=True):
    """{kernel family: (launches, total ms)} of the direct-sum main kernels since the last reset."""
    k = (C.c_int64 * 8)()
    ms = (C.c_double * 8)()
    check(lib.lpm_profile_breakdown(1 if reset else 0, k, ms))
    return {t: (int(k[i]), float(ms[i])) for i, t in enumerate(PROFILE_TAGS) if k[i]}


# ---- remaining PSE operators (src/PSEDirectSum.f90:128-456, 537-579) ----------------
def _mi(m):
    return m.ctypes.data_as(_i32)


def pse_interpolate_sphere(x, y, z, f, area, mask, eps, tx, ty, tz, sphere_radius=1.0):
    x, y, z, f, area, tx, ty, tz = map(_f64, (x, y, z, f, area, tx, ty, tz))
    m = _mask(mask)
    out = np.empty(tx.size)
    check(lib.lpm_pse_interpolate_sphere(x.size, _pd(x), _pd(y), _pd(z), _pd(f), _pd(area), _mi(m), float(eps),
                                         float(sphere_radius), tx.size, _pd(tx), _pd(ty), _pd(tz), _pd(out)))
    return out


def pse_interpolate_plane(x, y, f, area, mask, eps, tx, ty):
    x, y, f, area, tx, ty = map(_f64, (x, y, f, area, tx, ty))
    m = _mask(mask)
    out = np.empty(tx.size)
    check(lib.lpm_pse_interpolate_plane(x.size, _pd(x), _pd(y), _pd(f), _pd(area), _mi(m), float(eps),
                                        tx.size, _pd(tx), _pd(ty), _pd(out)))
    return out


def pse_gradient_sphere(x, y, z, f, area, mask, eps, sphere_radius=1.0):
    x, y, z, f, area = map(_f64, (x, y, z, f, area))
    m = _mask(mask)
    g = [np.empty(x.size) for _ in range(3)]
    check(lib.lpm_pse_gradient_sphere(x.size, _pd(x), _pd(y), _pd(z), _pd(f), _pd(area), _mi(m), float(eps),
                                      float(sphere_radius), *[_pd(a) for a in g]))
    return g


def pse_gradient_plane(x, y, f, area, mask, eps):
    x, y, f, area = map(_f64, (x, y, f, area))
    m = _mask(mask)
    g = [np.empty(x.size) for _ in range(2)]
    check(lib.lpm_pse_gradient_plane(x.size, _pd(x), _pd(y), _pd(f), _pd(area), _mi(m), float(eps), *[_pd(a) for a in g]))
    return g


def pse_second_partials_plane(x, y, gx, gy, area, mask, eps):
    x, y, gx, gy, area = map(_f64, (x, y, gx, gy, area))
    m = _mask(mask)
    o = [np.empty(x.size) for _ in range(3)]
    check(lib.lpm_pse_second_partials_plane(x.size, _pd(x), _pd(y), _pd(gx), _pd(gy), _pd(area), _mi(m), float(eps),
                                            *[_pd(a) for a in o]))
    return o


def pse_double_dot_plane(x, y, u, v, area, mask, eps):
    x, y, u, v, area = map(_f64, (x, y, u, v, area))
    m = _mask(mask)
    dd = np.empty(x.size)
    check(lib.lpm_pse_double_dot_plane(x.size, _pd(x), _pd(y), _pd(u), _pd(v), _pd(area), _mi(m), float(eps), _pd(dd)))
    return dd


def pse_double_dot_sphere(x, y, z, u, v, w, area, mask, eps, sphere_radius=1.0):
    x, y, z, u, v, w, area = map(_f64, (x, y, z, u, v, w, area))
    m = _mask(mask)
    dd = np.empty(x.size)
    check(lib.lpm_pse_double_dot_sphere(x.size, _pd(x), _pd(y), _pd(z), _pd(u), _pd(v), _pd(w), _pd(area), _mi(m),
                                        float(eps), float(sphere_radius), _pd(dd)))
    return dd


def pse_divergence_sphere(x, y, z, u, v, w, area, mask, eps, sphere_radius=1.0):
    x, y, z, u, v, w, area = map(_f64, (x, y, z, u, v, w, area))
    m = _mask(mask)
    div = np.empty(x.size)
    check(lib.lpm_pse_divergence_sphere(x.size, _pd(x), _pd(y), _pd(z), _pd(u), _pd(v), _pd(w), _pd(area), _mi(m),
                                        float(eps), float(sphere_radius), _pd(div)))
    return div


def swe_plane_rhs_integrals(x, y, vort, div, surf, area, mask, pse_eps):
    """SWEPlaneRHSIntegrals (src/SWEPlaneSolver.f90:457-560): returns u, v, doubleDot, lapSurf."""
    x, y, vort, div, surf, area = map(_f64, (x, y, vort, div, surf, area))
    m = _mask(mask)
    o = [np.empty(x.size) for _ in range(4)]
    check(lib.lpm_swe_plane_rhs_integrals(x.size, _pd(x), _pd(y), _pd(vort), _pd(div), _pd(surf), _pd(area), _mi(m),
                                          float(pse_eps), *[_pd(a) for a in o]))
    return o


def swe_plane_velocity(x, y, vort, div, area, mask):
    """SetVelocityFromFieldData (src/PlanarSWE.f90:469-494): returns u, v."""
    x, y, vort, div, area = map(_f64, (x, y, vort, div, area))
    m = _mask(mask)
    o = [np.empty(x.size) for _ in range(2)]
    check(lib.lpm_swe_plane_velocity(x.size, _pd(x), _pd(y), _pd(vort), _pd(div), _pd(area), _mi(m), *[_pd(a) for a in o]))
    return o


def swe_sphere_rhs_integrals(x, y, z, vort, div, surf, area, mask, radius, pse_eps):
    """SWESphereRHSIntegrals (src/SphereSWESolver.f90:296-375), as written: returns u, v, w, doubleDot (zero), lapSurf."""
    x, y, z, vort, div, surf, area = map(_f64, (x, y, z, vort, div, surf, area))
    m = _mask(mask)
    o = [np.empty(x.size) for _ in range(5)]
    check(lib.lpm_swe_sphere_rhs_integrals(x.size, _pd(x), _pd(y), _pd(z), _pd(vort), _pd(div), _pd(surf), _pd(area),
                                           _mi(m), float(radius), float(pse_eps), *[_pd(a) for a in o]))
    return o
