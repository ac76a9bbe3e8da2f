"""PolyMesh2d (uniform refinement) -- thin wrapper over the host mesh generator
in liblpmmesh.so (csrc/mesh.cpp, include/lpm_mesh.h; host only, no CUDA; reference: src/PolyMesh2d.f90:135-195)."""
import ctypes as C

import numpy as np

from ._meshlib import lib, check

TRI_HEX_SEED = 201
QUAD_RECT_SEED = 202
ICOS_TRI_SPHERE_SEED = 205
CUBED_SPHERE_SEED = 206
BETA_PLANE_SEED = 207

_d = C.POINTER(C.c_double)
_i32 = C.POINTER(C.c_int32)


def seed_from_face_kind(face_kind, sphere=True):
    """faceKind namelist value (3 = triangles, 4 = quadrilaterals) -> seed id,
    as the drivers do (examples/BVESingleGaussianVortex.f90:391-400)."""
    if sphere:
        return ICOS_TRI_SPHERE_SEED if face_kind == 3 else CUBED_SPHERE_SEED
    return TRI_HEX_SEED if face_kind == 3 else QUAD_RECT_SEED


class PolyMesh2d:
    """Particles of a uniformly refined mesh: x, y, z, area, is_active (numpy)."""

    def __init__(self, seed, init_nest, amp_factor=1.0):
        h = C.c_void_p()
        check(lib.lpm_mesh_create(int(seed), int(init_nest), float(amp_factor), C.byref(h)))
        try:
            self.seed = seed
            self.init_nest = init_nest
            self._amp = amp_factor
            n = lib.lpm_mesh_num_particles(h)
            self.n = int(n)
            self.n_faces_total = int(lib.lpm_mesh_num_faces(h))
            self.n_edges_total = int(lib.lpm_mesh_num_edges(h))
            self.n_leaf_faces = int(lib.lpm_mesh_num_leaf_faces(h))
            self.n_leaf_edges = int(lib.lpm_mesh_num_leaf_edges(h))
            self.max_edge_length = float(lib.lpm_mesh_max_edge_length(h))
            self.x, self.y, self.z, self.area = (np.empty(n) for _ in range(4))
            self.is_active = np.empty(n, np.int32)
            check(lib.lpm_mesh_get_particles(h, self.x.ctypes.data_as(_d), self.y.ctypes.data_as(_d),
                                             self.z.ctypes.data_as(_d), self.area.ctypes.data_as(_d),
                                             self.is_active.ctypes.data_as(_i32)))
            vpf = 3 if seed in (TRI_HEX_SEED, ICOS_TRI_SPHERE_SEED) else 4
            self.face_verts = np.empty((self.n_leaf_faces, vpf), np.int32)
            self.face_center = np.empty(self.n_leaf_faces, np.int32)
            check(lib.lpm_mesh_get_leaf_faces(h, self.face_verts.ctypes.data_as(_i32),
                                              self.face_center.ctypes.data_as(_i32)))
        finally:
            lib.lpm_mesh_destroy(h)

    @property
    def n_active(self):
        return int(self.is_active.sum())

    def write_vtk(self, filename, fields=(), title="", positions=None):
        """OutputToVTK (src/SphereBVE.f90:283-328): legacy ASCII .vtk PolyData of the mesh and point
        fields.  `fields` = [(name_units, array or tuple of 2-3 component arrays), ...];
        `positions` = current (x, y, z) of the particles, default the mesh's own."""
        h = C.c_void_p()
        check(lib.lpm_mesh_create(int(self.seed), int(self.init_nest), float(self._amp), C.byref(h)))
        try:
            names, ndim, blobs = [], [], []
            for name, val in fields:
                comps = [val] if isinstance(val, np.ndarray) else list(val)
                blob = np.ascontiguousarray(np.concatenate([np.asarray(c, np.float64).ravel() for c in comps]))
                assert blob.size == len(comps) * self.n, name
                names.append(name.encode()); ndim.append(len(comps)); blobs.append(blob)
            k = len(names)
            c_names = (C.c_char_p * max(k, 1))(*names)
            c_ndim = (C.c_int * max(k, 1))(*ndim)
            c_data = (_d * max(k, 1))(*[b.ctypes.data_as(_d) for b in blobs])
            pos = [None, None, None]
            if positions is not None:
                pos = [np.ascontiguousarray(p, np.float64) for p in positions]
                while len(pos) < 3:
                    pos.append(None)
            pp = [p.ctypes.data_as(_d) if p is not None else None for p in pos]
            check(lib.lpm_mesh_write_vtk(h, str(filename).encode(), title.encode(), pp[0], pp[1], pp[2], k, c_names,
                                         c_ndim, c_data))
        finally:
            lib.lpm_mesh_destroy(h)
