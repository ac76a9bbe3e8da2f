"""ctypes binding of liblpmgpu.so (the C ABI declared in include/lpm_gpu.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C lpm_v2_b200``.
There is no Python or CPU fallback: if the shared object is missing this module
raises at import, and every compute entry point returns an error without a
B200.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LPM_GPU_LIBRARY: load the library from another path (a build under test, e.g. tools/build_renamed.sh, or the
# emulator build of tests/cuda_emu).  It replaces the path only: a missing file still raises.
LIB_PATH = os.environ.get("LPM_GPU_LIBRARY") or os.path.join(_HERE, "liblpmgpu.so")


class LpmError(RuntimeError):
    """Non-zero return from liblpmgpu (message from lpm_gpu_last_error)."""

    def __init__(self, code, msg):
        super().__init__(f"liblpmgpu error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C lpm_v2_b200` (no CPU fallback exists)")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

_d = C.POINTER(C.c_double)
_i32 = C.POINTER(C.c_int32)
_i64 = C.POINTER(C.c_int64)
_vp = C.c_void_p
_n = C.c_int64
_dbl = C.c_double
_int = C.c_int

# name -> (restype, argtypes); must list every symbol of include/lpm_gpu.h
PROTOTYPES = {
    "lpm_gpu_init": (_int, [_int, C.POINTER(_int)]),
    "lpm_gpu_finalize": (_int, []),
    "lpm_gpu_last_error": (C.c_char_p, []),
    "lpm_gpu_device_count": (_int, []),
    "lpm_gpu_init_rank": (_int, [_int]),
    "lpm_comm_unique_id": (_int, [C.c_char_p]),
    "lpm_comm_init_rank": (_int, [_int, _int, C.c_char_p]),
    "lpm_comm_world_size": (_int, []),
    "lpm_comm_rank": (_int, []),
    "lpm_comm_allgather_slices_dev": (_int, [_int, C.POINTER(_vp), _n, _vp]),
    "lpm_gpu_pin": (_int, [_vp, _n]),
    "lpm_gpu_unpin": (_int, [_vp]),
    "lpm_load_balance": (_int, [_n, _int, _i64, _i64, _i64]),
    "lpm_active_list": (_int, [_n, _i32, _i32, _i64]),
    # host API
    "lpm_bve_velocity": (_int, [_n, _d, _d, _d, _d, _d, _i32, _dbl, _d, _d, _d]),
    "lpm_bve_stream": (_int, [_n, _d, _d, _d, _d, _d, _d, _i32, _dbl, _d, _d]),
    "lpm_plane_velocity": (_int, [_n, _d, _d, _d, _d, _i32, _d, _d]),
    "lpm_plane_stream": (_int, [_n, _d, _d, _d, _d, _i32, _d]),
    "lpm_betaplane_velocity": (_int, [_n, _d, _d, _d, _d, _i32, _d, _d]),
    "lpm_betaplane_stream": (_int, [_n, _d, _d, _d, _d, _d, _i32, _d, _d]),
    "lpm_pse_laplacian_sphere": (_int, [_n, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _d]),
    "lpm_pse_laplacian_plane": (_int, [_n, _d, _d, _d, _d, _i32, _dbl, _d]),
    "lpm_pse_interpolate_sphere": (_int, [_n, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _n, _d, _d, _d, _d]),
    "lpm_pse_interpolate_plane": (_int, [_n, _d, _d, _d, _d, _i32, _dbl, _n, _d, _d, _d]),
    "lpm_pse_gradient_sphere": (_int, [_n, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _d, _d, _d]),
    "lpm_pse_gradient_plane": (_int, [_n, _d, _d, _d, _d, _i32, _dbl, _d, _d]),
    "lpm_pse_second_partials_plane": (_int, [_n, _d, _d, _d, _d, _d, _i32, _dbl, _d, _d, _d]),
    "lpm_pse_double_dot_plane": (_int, [_n, _d, _d, _d, _d, _d, _i32, _dbl, _d]),
    "lpm_pse_double_dot_sphere": (_int, [_n, _d, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _d]),
    "lpm_pse_divergence_sphere": (_int, [_n, _d, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _d]),
    "lpm_swe_plane_rhs_integrals": (_int, [_n, _d, _d, _d, _d, _d, _d, _i32, _dbl, _d, _d, _d, _d]),
    "lpm_swe_plane_velocity": (_int, [_n, _d, _d, _d, _d, _d, _i32, _d, _d]),
    "lpm_swe_sphere_rhs_integrals": (_int, [_n, _d, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _d, _d, _d, _d, _d]),
    # device API (device pointers passed as integers)
    "lpm_bve_velocity_dev": (_int, [_n, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _n, _n, _vp, _vp, _vp, _vp]),
    "lpm_bve_stream_dev": (_int, [_n, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _n, _n, _vp, _vp, _vp]),
    "lpm_plane_velocity_dev": (_int, [_n, _vp, _vp, _vp, _vp, _vp, _n, _n, _vp, _vp, _vp]),
    "lpm_plane_stream_dev": (_int, [_n, _vp, _vp, _vp, _vp, _vp, _n, _n, _vp, _vp]),
    "lpm_betaplane_velocity_dev": (_int, [_n, _vp, _vp, _vp, _vp, _vp, _n, _n, _vp, _vp, _vp]),
    "lpm_betaplane_stream_dev": (_int, [_n, _vp, _vp, _vp, _vp, _vp, _vp, _n, _n, _vp, _vp, _vp]),
    "lpm_pse_laplacian_sphere_dev": (_int, [_n, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _n, _n, _vp, _vp]),
    "lpm_pse_laplacian_plane_dev": (_int, [_n, _vp, _vp, _vp, _vp, _vp, _dbl, _n, _n, _vp, _vp]),
    # resident solvers
    "lpm_bve_solver_new": (_int, [_n, _d, _d, _d, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, C.POINTER(_vp)]),
    "lpm_bve_solver_timestep": (_int, [_vp, _dbl, _int]),
    "lpm_bve_solver_get_state": (_int, [_vp, _d, _d, _d, _d, _d, _d, _d, _d, _d]),
    "lpm_bve_solver_diagnostics": (_int, [_vp, _d, _d]),
    "lpm_bve_solver_delete": (_int, [_vp]),
    "lpm_plane_solver_new": (_int, [_n, _d, _d, _d, _d, _d, _d, _i32, C.POINTER(_vp)]),
    "lpm_plane_solver_timestep": (_int, [_vp, _dbl, _int]),
    "lpm_plane_solver_get_state": (_int, [_vp, _d, _d, _d, _d, _d]),
    "lpm_plane_solver_delete": (_int, [_vp]),
    "lpm_betaplane_solver_new": (_int, [_n, _d, _d, _d, _d, _d, _d, _d, _i32, _dbl, C.POINTER(_vp)]),
    "lpm_betaplane_solver_timestep": (_int, [_vp, _dbl, _int]),
    "lpm_betaplane_solver_get_state": (_int, [_vp, _d, _d, _d, _d, _d, _d, _d]),
    "lpm_betaplane_solver_delete": (_int, [_vp]),
    "lpm_swe_plane_solver_new": (_int, [_n, _d, _d, _d, _d, _d, _d, _i32, _dbl, _dbl, _dbl, _dbl, _vp, _vp, C.POINTER(_vp)]),
    "lpm_swe_plane_solver_timestep": (_int, [_vp, _dbl]),
    "lpm_swe_plane_solver_get_state": (_int, [_vp, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d]),
    "lpm_swe_plane_solver_delete": (_int, [_vp]),
    # measurement
    "lpm_fp64_peak_probe": (_int, [_int, _d, _d]),
    "lpm_last_kernel_ms": (_int, [_d]),
    "lpm_last_sum_ms": (_int, [_d]),
    "lpm_profile_summary": (_int, [_int, _i64, _d]),
    "lpm_profile_breakdown": (_int, [_int, _i64, _d]),
    "lpm_launch_count": (C.c_int64, [_int]),
    "lpm_comm_alloc_shared": (_int, [C.c_int64, C.POINTER(_vp)]),
    "lpm_comm_free_shared": (_int, [_vp]),
    "lpm_comm_is_shared": (_int, [_vp, C.c_int64]),
    "lpm_set_profiling": (_int, [_int]),
    "lpm_set_symmetric": (_int, [_int]),
    "lpm_set_pse_culling": (_int, [_int]),
    "lpm_set_pse_series": (_int, [_int]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _f = getattr(lib, _name)      # AttributeError here == missing export
    _f.restype = _res
    _f.argtypes = _args


TOPOGRAPHY_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double, C.c_void_p)      # lpm_topography_fn

# not part of the C ABI (csrc/lpm_gpu_tuning.h): A/B and test knob
lib.lpm_tune.restype = _int
lib.lpm_tune.argtypes = [C.c_char_p, _int]


def last_error():
    return lib.lpm_gpu_last_error().decode("utf-8", "replace")


def check(rc):
    if rc != 0:
        raise LpmError(rc, last_error())
