"""ctypes binding of liblpmmesh.so (include/lpm_mesh.h): the host-only mesh generator and VTK writer.
Separate from _lib.py so that code which only needs meshes (bench.py's CPU reference arm, the CPU
test tier) never maps the GPU library."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LPM_MESH_LIBRARY") or os.path.join(_HERE, "liblpmmesh.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not found: build it with `make -C lpm_v2_b200`")
lib = C.CDLL(LIB_PATH)

_d = C.POINTER(C.c_double)
_i32 = C.POINTER(C.c_int32)
_vp = C.c_void_p
_dbl = C.c_double
_int = C.c_int

# name -> (restype, argtypes); must list every symbol of include/lpm_mesh.h
PROTOTYPES = {
    "lpm_mesh_create": (_int, [_int, _int, _dbl, C.POINTER(_vp)]),
    "lpm_mesh_destroy": (None, [_vp]),
    "lpm_mesh_num_particles": (C.c_int64, [_vp]),
    "lpm_mesh_num_faces": (C.c_int64, [_vp]),
    "lpm_mesh_num_edges": (C.c_int64, [_vp]),
    "lpm_mesh_num_leaf_faces": (C.c_int64, [_vp]),
    "lpm_mesh_num_leaf_edges": (C.c_int64, [_vp]),
    "lpm_mesh_max_edge_length": (_dbl, [_vp]),
    "lpm_mesh_get_particles": (_int, [_vp, _d, _d, _d, _d, _i32]),
    "lpm_mesh_get_leaf_faces": (_int, [_vp, _i32, _i32]),
    "lpm_mesh_write_vtk": (_int, [_vp, C.c_char_p, C.c_char_p, _d, _d, _d, _int, C.POINTER(C.c_char_p), C.POINTER(_int),
                                  C.POINTER(_d)]),
}
for _name, (_res, _args) in PROTOTYPES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


class MeshError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise MeshError(f"liblpmmesh error {rc} (invalid argument)")
