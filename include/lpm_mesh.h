/* lpm_mesh.h -- C ABI of liblpmmesh.so: the host-only uniform mesh generator and legacy-VTK writer on the
 * input / output side of the direct-sum path.  No CUDA, no dependency on liblpmgpu.so:
 * bench.py's --impl reference arm and the CPU test tier load this library alone.
 *
 * Reference: src/PolyMesh2d.f90:135-195 (New + uniform refinement), src/Faces.f90:529-858 (DivideTri / DivideQuad,
 * particle insertion order), src/Edges.f90:211-233, src/SphereGeometry.f90 (midpoints, centroids, areas), the
 * five *Seed.dat tables; OutputToVTK src/SphereBVE.f90:283-328. */
#ifndef LPM_MESH_H
#define LPM_MESH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* return codes shared with lpm_gpu.h */
#ifndef LPM_GPU_H
enum { LPM_OK = 0, LPM_ERR_INVALID = 1 };
#endif

/* mesh seed identifiers, src/TypeDefs.f90:66-73 */
enum {
    LPM_TRI_HEX_SEED = 201,
    LPM_QUAD_RECT_SEED = 202,
    LPM_ICOS_TRI_SPHERE_SEED = 205,
    LPM_CUBED_SPHERE_SEED = 206,
    LPM_BETA_PLANE_SEED = 207
};

/* Uniformly refined PolyMesh2d particle set in the reference's insertion
 * order: src/PolyMesh2d.f90:135-195, src/Faces.f90:529-858. */
typedef struct lpm_mesh lpm_mesh;
int lpm_mesh_create(int seed_kind, int init_nest, double amp_factor, lpm_mesh** out);
void lpm_mesh_destroy(lpm_mesh* m);
int64_t lpm_mesh_num_particles(const lpm_mesh* m);
int64_t lpm_mesh_num_faces(const lpm_mesh* m);        /* whole quadtree */
int64_t lpm_mesh_num_edges(const lpm_mesh* m);        /* whole binary tree */
int64_t lpm_mesh_num_leaf_faces(const lpm_mesh* m);
int64_t lpm_mesh_num_leaf_edges(const lpm_mesh* m);
double lpm_mesh_max_edge_length(const lpm_mesh* m);   /* src/Edges.f90:260-275 */
int lpm_mesh_get_particles(const lpm_mesh* m, double* x, double* y, double* z, double* area, int32_t* is_active);
int lpm_mesh_get_leaf_faces(const lpm_mesh* m, int32_t* verts, int32_t* center);
/* Legacy ASCII .vtk PolyData file of the mesh and `nfields` point fields, in the layout of
 * OutputToVTK (src/SphereBVE.f90:283-328): POINTS, POLYGONS (each leaf face as triangles around
 * its centre particle), POINT_DATA (lagParam, then the fields: names[f] = "name_units", ndim[f] in
 * 1..3, data[f] = ndim[f] component arrays of N doubles stored one after the other), CELL_DATA
 * faceArea.  x, y, z: current particle positions, or NULL for the mesh's own. */
int lpm_mesh_write_vtk(const lpm_mesh* m, const char* filename, const char* title, const double* x, const double* y,
                       const double* z, int nfields, const char* const* names, const int* ndim,
                       const double* const* data);

#ifdef __cplusplus
}
#endif
#endif /* LPM_MESH_H */
