/*
 * lpm_gpu.h -- C ABI of liblpmgpu.so: the B200 (sm_100a) replacement for
 * lpm-v2's O(N^2) direct-sum hot path.
 *
 * Every entry point is what the reference's Fortran would bind through
 * ISO_C_BINDING for this path (see lpm_v2_b200/fortran/lpm_gpu.f90 and
 * INTEGRATION.md).  Plain pointers and sizes only.  File:line citations are
 * into the lpm-v2 source tree.
 *
 * Conventions
 *   - all functions returning int: 0 (LPM_OK) on success, non-zero on error;
 *     lpm_gpu_last_error() gives the message (the Fortran shim forwards it to
 *     LogMessage(log, ERROR_LOGGING_LEVEL, ...), matching the reference's
 *     "log, never abort" behaviour, src/Logger.f90:158-160).
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with LPM_ERR_NO_DEVICE.
 *   - arrays are contiguous real(c_double) / integer(c_int32_t), length n.
 *   - mask: int32, non-zero = active (Fortran side: merge(1,0,activeMask)).
 *   - target ranges [ibeg, iend) are 0-based half-open; the reference slice
 *     indexStart(r)..indexEnd(r) (src/MPISetup.f90:132-146) is
 *     [indexStart-1, indexEnd).
 *   - "host" entry points take host pointers and are synchronous: outputs
 *     are complete on return.  "_dev" entry points take device pointers on
 *     the current CUDA device and enqueue on `stream` (a cudaStream_t passed
 *     as void*; NULL = default stream); the first evaluation for a mask
 *     synchronises that stream once (to read the active count).
 *     ONE STREAM PER DEVICE AT A TIME: every evaluation on a device packs its
 *     sources and partial sums into that device's single workspace, so two
 *     evaluations must not be in flight on different streams of the same
 *     device (the host entry points and the resident solvers use the
 *     library's own stream and are subject to the same rule: finish or
 *     synchronise one before starting the next on another stream).
 */
#ifndef LPM_GPU_H
#define LPM_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    LPM_OK = 0,
    LPM_ERR_INVALID = 1,      /* bad argument */
    LPM_ERR_NO_DEVICE = 2,    /* no CUDA device / not sm_100 / library not initialised */
    LPM_ERR_CUDA = 3,         /* CUDA runtime error (message in lpm_gpu_last_error) */
    LPM_ERR_COMM = 4,         /* NCCL / peer-access error */
    LPM_ERR_NOMEM = 5
};

/* ---------------------------------------------------------------- runtime */

/* Replaces MPI_INIT + MPI_COMM_SIZE for this path (examples/BVESingleGaussianVortex.f90:102-104).
 * Single-process mode: claims ndev_requested devices (0 = all visible), enables
 * peer access between them.  ndev_used may be NULL. */
int lpm_gpu_init(int ndev_requested, int* ndev_used);
int lpm_gpu_finalize(void);
const char* lpm_gpu_last_error(void);
int lpm_gpu_device_count(void);       /* devices claimed by lpm_gpu_init (0 before) */

/* Multi-process mode (one MPI rank / process per GPU): bind this process to
 * `device`, then build an NCCL communicator.  The 128-byte id comes from
 * lpm_comm_unique_id on rank 0 and is distributed by the host program
 * (MPI_BCAST in Fortran, torch.distributed in bench.py). */
int lpm_gpu_init_rank(int device);
int lpm_comm_unique_id(char id[128]);
int lpm_comm_init_rank(int world_size, int rank, const char id[128]);
int lpm_comm_world_size(void);
int lpm_comm_rank(void);

/* Replaces the MPI_BCAST loop that follows every direct sum
 * (src/SphereBVESolver.f90:422-429): every rank contributes
 * buf[start_r .. end_r) for each of the ncomp device arrays, ranges from
 * lpm_load_balance; in place; ragged last slice allowed. */
int lpm_comm_allgather_slices_dev(int ncomp, double* const* bufs, int64_t n, void* stream);

/* Rank mode: device memory that every rank of the communicator can store into over NVLink
 * (CUDA IPC mappings exchanged through the communicator).  COLLECTIVE: every rank calls with the
 * same size, in the same order.  When ALL output arrays of a `_dev` direct sum lie inside such
 * allocations (at the same offsets on every rank), the step that writes the results stores them
 * into every rank's copy -- the reference's MPI_BCAST loop (e.g. src/SphereBVESolver.f90:422-429)
 * fused into the sum as peer stores -- between two stream-ordered barriers; after the stream's
 * work every rank holds all n results and lpm_comm_allgather_slices_dev is not needed.
 * lpm_comm_is_shared tells whether [ptr, ptr + bytes) lies in a shared allocation (1 / 0). */
int lpm_comm_alloc_shared(int64_t bytes, void** ptr);
int lpm_comm_free_shared(void* ptr);
int lpm_comm_is_shared(const void* ptr, int64_t bytes);

/* Optional: page-lock host arrays the solver owns (at New), release at Delete. */
int lpm_gpu_pin(void* ptr, int64_t bytes);
int lpm_gpu_unpin(void* ptr);

/* src/MPISetup.f90:132-146 LoadBalance.  Outputs are 1-based inclusive like
 * the reference's indexStart/indexEnd/messageLength (arrays of nprocs). */
int lpm_load_balance(int64_t n_items, int nprocs, int64_t* index_start, int64_t* index_end, int64_t* message_length);

/* The active-source index list the device compaction builds == Fortran
 * pack([(j,j=1,n)], mask), 0-based; exposed for the bit-exact check. */
int lpm_active_list(int64_t n, const int32_t* mask, int32_t* list, int64_t* count);

/* ------------------------------------------------- direct sums, host API */

/* BVESphereVelocity, src/SphereBVESolver.f90:377-430; also
 * setVelocityFromVorticity, src/SphereBVE.f90:489-531. */
int lpm_bve_velocity(int64_t n, const double* x, const double* y, const double* z,
                     const double* relvort, const double* area, const int32_t* mask,
                     double radius, double* u, double* v, double* w);

/* SetStreamFunctionsOnMesh, src/SphereBVE.f90:445-485. */
int lpm_bve_stream(int64_t n, const double* x, const double* y, const double* z,
                   const double* relvort, const double* absvort, const double* area,
                   const int32_t* mask, double radius, double* relstream, double* absstream);

/* planarIncompressibleVelocity, src/PlaneIncompressibleSolver.f90:278-316;
 * setVelocityFromVorticity, src/PlanarIncompressible.f90:426-466. */
int lpm_plane_velocity(int64_t n, const double* x, const double* y, const double* vort,
                       const double* area, const int32_t* mask, double* u, double* v);

/* SetStreamFunctionOnMesh, src/PlanarIncompressible.f90:470-505. */
int lpm_plane_stream(int64_t n, const double* x, const double* y, const double* vort,
                     const double* area, const int32_t* mask, double* psi);

/* BetaPlaneVelocity, src/BetaPlaneSolver.f90:227-267;
 * setVelocityFromVorticity, src/BetaPlane.f90:359-397. */
int lpm_betaplane_velocity(int64_t n, const double* x, const double* y, const double* relvort,
                           const double* area, const int32_t* mask, double* u, double* v);

/* SetStreamFunctionsOnMesh, src/BetaPlane.f90:399-442. */
int lpm_betaplane_stream(int64_t n, const double* x, const double* y, const double* relvort,
                         const double* absvort, const double* area, const int32_t* mask,
                         double* relstream, double* absstream);

/* PSESphereLaplacianAtParticles, src/PSEDirectSum.f90:502-535.
 * sphere_radius is the module-global SphereRadius (src/TypeDefs.f90:74). */
int lpm_pse_laplacian_sphere(int64_t n, const double* x, const double* y, const double* z,
                             const double* f, const double* area, const int32_t* mask,
                             double eps, double sphere_radius, double* lap);

/* PSEPlaneLaplacianAtParticles, src/PSEDirectSum.f90:467-500. */
int lpm_pse_laplacian_plane(int64_t n, const double* x, const double* y,
                            const double* f, const double* area, const int32_t* mask,
                            double eps, double* lap);

/* Remaining PSE operators.  sphere_radius as above.
 * PSE{Sphere,Plane}InterpolateScalar, src/PSEDirectSum.f90:128-168, evaluated at m
 * arbitrary locations (tx, ty[, tz]) -- e.g. the lat-lon grid of tests/SpherePSEConvTest.f90:188-211. */
int lpm_pse_interpolate_sphere(int64_t n, const double* x, const double* y, const double* z,
                               const double* f, const double* area, const int32_t* mask,
                               double eps, double sphere_radius,
                               int64_t m, const double* tx, const double* ty, const double* tz, double* out);
int lpm_pse_interpolate_plane(int64_t n, const double* x, const double* y,
                              const double* f, const double* area, const int32_t* mask, double eps,
                              int64_t m, const double* tx, const double* ty, double* out);
/* PSESphereGradientAtParticles :221-267 / PSEPlaneGradientAtParticles :180-218 */
int lpm_pse_gradient_sphere(int64_t n, const double* x, const double* y, const double* z,
                            const double* f, const double* area, const int32_t* mask,
                            double eps, double sphere_radius, double* gx, double* gy, double* gz);
int lpm_pse_gradient_plane(int64_t n, const double* x, const double* y,
                           const double* f, const double* area, const int32_t* mask, double eps,
                           double* gx, double* gy);
/* PSEPlaneSecondPartialsAtParticles :269-320: outputs d_xx, (d_xy + d_yx)/2, d_yy */
int lpm_pse_second_partials_plane(int64_t n, const double* x, const double* y,
                                  const double* gx, const double* gy, const double* area,
                                  const int32_t* mask, double eps, double* dxx, double* dxy, double* dyy);
/* PSEPlaneDoubleDotProductAtParticles :322-365 / PSESphereDoubleDotProductAtParticles :367-420
 * (the sphere routine's w rows add yComp(i), :408-413; kept as written) */
int lpm_pse_double_dot_plane(int64_t n, const double* x, const double* y,
                             const double* u, const double* v, const double* area,
                             const int32_t* mask, double eps, double* dd);
int lpm_pse_double_dot_sphere(int64_t n, const double* x, const double* y, const double* z,
                              const double* u, const double* v, const double* w, const double* area,
                              const int32_t* mask, double eps, double sphere_radius, double* dd);
/* PSESphereDivergenceAtParticles :537-579 */
int lpm_pse_divergence_sphere(int64_t n, const double* x, const double* y, const double* z,
                              const double* u, const double* v, const double* w, const double* area,
                              const int32_t* mask, double eps, double sphere_radius, double* div);

/* SWEPlaneRHSIntegrals, src/SWEPlaneSolver.f90:457-560: velocity from vorticity and divergence,
 * the double dot product of the velocity gradient and the PSE Laplacian of the fluid surface in
 * one pass.  surf(j) = h(j) + topoFn(x(j), y(j)), evaluated by the caller (O(N)). */
int lpm_swe_plane_rhs_integrals(int64_t n, const double* x, const double* y, const double* vort,
                                const double* div, const double* surf, const double* area,
                                const int32_t* mask, double pse_eps,
                                double* u, double* v, double* double_dot, double* lap_surf);

/* SetVelocityFromFieldData, src/PlanarSWE.f90:469-494 (== SWEComputeVelocity :261-290): planar
 * velocity from relative vorticity and divergence, j /= i, active sources. */
int lpm_swe_plane_velocity(int64_t n, const double* x, const double* y, const double* vort, const double* div,
                           const double* area, const int32_t* mask, double* u, double* v);

/* SWESphereRHSIntegrals, src/SphereSWESolver.f90:296-375, as written: velocity from vorticity and
 * divergence on the sphere and the PSE Laplacian of the fluid surface `surf` = h + topography
 * (evaluated by the caller).  The reference routine is unfinished: double_dot is zeroed and never
 * accumulated, and lap_surf gets no trailing 1/eps^2 (the plane routine applies one); both are
 * reproduced.  SphereDistance's module-global SphereRadius is taken to be `radius`. */
int lpm_swe_sphere_rhs_integrals(int64_t n, const double* x, const double* y, const double* z, const double* vort,
                                 const double* div, const double* surf, const double* area, const int32_t* mask,
                                 double radius, double pse_eps, double* u, double* v, double* w, double* double_dot,
                                 double* lap_surf);

/* -------------------------- direct sums, device API (one rank's slice) */

int lpm_bve_velocity_dev(int64_t n, const double* x, const double* y, const double* z,
                         const double* relvort, const double* area, const int32_t* mask,
                         double radius, int64_t ibeg, int64_t iend,
                         double* u, double* v, double* w, void* stream);
int lpm_bve_stream_dev(int64_t n, const double* x, const double* y, const double* z,
                       const double* relvort, const double* absvort, const double* area,
                       const int32_t* mask, double radius, int64_t ibeg, int64_t iend,
                       double* relstream, double* absstream, void* stream);
int lpm_plane_velocity_dev(int64_t n, const double* x, const double* y, const double* vort,
                           const double* area, const int32_t* mask, int64_t ibeg, int64_t iend,
                           double* u, double* v, void* stream);
int lpm_plane_stream_dev(int64_t n, const double* x, const double* y, const double* vort,
                         const double* area, const int32_t* mask, int64_t ibeg, int64_t iend,
                         double* psi, void* stream);
int lpm_betaplane_velocity_dev(int64_t n, const double* x, const double* y, const double* relvort,
                               const double* area, const int32_t* mask, int64_t ibeg, int64_t iend,
                               double* u, double* v, void* stream);
int lpm_betaplane_stream_dev(int64_t n, const double* x, const double* y, const double* relvort,
                             const double* absvort, const double* area, const int32_t* mask,
                             int64_t ibeg, int64_t iend, double* relstream, double* absstream, void* stream);
int lpm_pse_laplacian_sphere_dev(int64_t n, const double* x, const double* y, const double* z,
                                 const double* f, const double* area, const int32_t* mask,
                                 double eps, double sphere_radius, int64_t ibeg, int64_t iend,
                                 double* lap, void* stream);
int lpm_pse_laplacian_plane_dev(int64_t n, const double* x, const double* y,
                                const double* f, const double* area, const int32_t* mask,
                                double eps, int64_t ibeg, int64_t iend, double* lap, void* stream);

/* ------------------------------------------- resident solvers (RK4 step) */

/* type BVESolver + New/Timestep/Delete, src/SphereBVESolver.f90:38-72,112-168,219-353.
 * All state lives on the device(s); Timestep runs the 4 stages, the final
 * velocity evaluation (:345-346) and, if with_stream != 0, the stream
 * functions (:352 -> src/SphereBVE.f90:445-485) without host round trips. */
typedef struct lpm_bve_solver lpm_bve_solver;
/* absvort (the materially conserved absVort field, src/SphereBVE.f90:346) may be NULL if
 * stream functions are never requested. */
int lpm_bve_solver_new(int64_t n, const double* x, const double* y, const double* z,
                       const double* relvort, const double* absvort,
                       const double* u, const double* v, const double* w,
                       const double* area, const int32_t* mask, double radius, double rotation_rate,
                       lpm_bve_solver** out);
int lpm_bve_solver_timestep(lpm_bve_solver* s, double dt, int with_stream);
/* copies back particles x,y,z, relVort, velocity (and stream functions; any pointer may be NULL) */
int lpm_bve_solver_get_state(lpm_bve_solver* s, double* x, double* y, double* z, double* relvort,
                             double* u, double* v, double* w, double* relstream, double* absstream);
/* TotalKE / TotalEnstrophy, src/SphereBVE.f90:410-441 (device reductions) */
int lpm_bve_solver_diagnostics(lpm_bve_solver* s, double* total_ke, double* total_enstrophy);
int lpm_bve_solver_delete(lpm_bve_solver* s);

/* type PlaneSolver, src/PlaneIncompressibleSolver.f90:37-59,171-259. */
typedef struct lpm_plane_solver lpm_plane_solver;
int lpm_plane_solver_new(int64_t n, const double* x, const double* y, const double* vort,
                         const double* u, const double* v, const double* area, const int32_t* mask,
                         lpm_plane_solver** out);
int lpm_plane_solver_timestep(lpm_plane_solver* s, double dt, int with_stream);
int lpm_plane_solver_get_state(lpm_plane_solver* s, double* x, double* y, double* u, double* v, double* psi);
int lpm_plane_solver_delete(lpm_plane_solver* s);

/* type BetaPlaneSolver, src/BetaPlaneSolver.f90:36-56,142-219. */
typedef struct lpm_betaplane_solver lpm_betaplane_solver;
int lpm_betaplane_solver_new(int64_t n, const double* x, const double* y, const double* relvort,
                             const double* absvort, const double* u, const double* v,
                             const double* area, const int32_t* mask, double beta,
                             lpm_betaplane_solver** out);
int lpm_betaplane_solver_timestep(lpm_betaplane_solver* s, double dt, int with_stream);
int lpm_betaplane_solver_get_state(lpm_betaplane_solver* s, double* x, double* y, double* relvort,
                                   double* u, double* v, double* relstream, double* absstream);
int lpm_betaplane_solver_delete(lpm_betaplane_solver* s);

/* type SWESolver (planar shallow water) + New/Timestep/Delete, src/SWEPlaneSolver.f90:47-110, 137-205, 298-429:
 * RK4 on x, y, relVort, divergence, area and depth, with one fused right-hand-side sum per stage
 * (SWEPlaneRHSIntegrals, :457-565).  f0, beta, g, pse_eps: plane%f0, plane%beta, plane%g, plane%pseEps.
 * topo(x, y, topo_user) is the bottom topography (the reference's topoFn argument; NULL = flat bottom): it is called
 * on the host for every particle at every stage, as the reference calls topoFn inside its loops.  New() evaluates
 * the right-hand side at the initial state (:202-204), so u, v, doubleDot, lapSurf are defined before the first step.
 * The step reproduces the reference AS WRITTEN, including the whole-array assignment of the stage-1 vorticity and
 * divergence tendencies (:312-315; every particle gets the last particle's value). */
typedef struct lpm_swe_plane_solver lpm_swe_plane_solver;
typedef double (*lpm_topography_fn)(double x, double y, void* user);
int lpm_swe_plane_solver_new(int64_t n, const double* x, const double* y, const double* relvort, const double* div,
                             const double* h, const double* area, const int32_t* mask, double f0, double beta,
                             double g, double pse_eps, lpm_topography_fn topo, void* topo_user,
                             lpm_swe_plane_solver** out);
int lpm_swe_plane_solver_timestep(lpm_swe_plane_solver* s, double dt);
/* copies back the particles, the prognostic fields and the last right-hand side (any pointer may be NULL) */
int lpm_swe_plane_solver_get_state(lpm_swe_plane_solver* s, double* x, double* y, double* relvort, double* div,
                                   double* h, double* area, double* u, double* v, double* double_dot, double* lap_surf);
int lpm_swe_plane_solver_delete(lpm_swe_plane_solver* s);

/* -------------------------------------------------------- measurement */

/* Dependent-free DFMA probe: runs `iters` rounds of independent FMA chains on
 * every SM and returns the achieved FP64 TFLOP/s (2 flop per DFMA) -- the
 * measured denominator for the FP64 roofline. */
int lpm_fp64_peak_probe(int iters, double* tflops, double* ms);

/* Time of the last direct-sum main kernel on this device (ms, CUDA events on
 * the launching stream), and the number of kernel launches the library has
 * issued since lpm_gpu_init / the last call with reset != 0. */
int lpm_last_kernel_ms(double* ms);
/* Device time (ms) of the last one-shot direct sum as a whole on this device: source packing,
 * the cell sort of the PSE kernels, the main kernel, finalize and scatter (no host copies). */
int lpm_last_sum_ms(double* ms);
/* Sum of the durations (ms) and count of the direct-sum main kernels recorded on the
 * current device since the last reset (profiling must be on). */
int lpm_profile_summary(int reset, int64_t* nkernels, double* total_ms);
/* The same per kernel family: entry 2 s + e holds the launches and the summed duration (ms) of sum s
 * (0 BVE velocity, 1 BVE stream functions, 2 every other sum, 3 the fused velocity + stream functions that end a
 * resident BVE time step) on engine e (0 the one-sided ds_kernel, 1 the pair-symmetric sym_kernel). */
int lpm_profile_breakdown(int reset, int64_t counts[8], double ms[8]);
int64_t lpm_launch_count(int reset);
/* 1: record events around each main kernel (adds a sync at query time only). */
int lpm_set_profiling(int enable);
/* Evaluation strategy of the BVE velocity and stream-function sums (sphere):
 *   1 (default)  whole evaluations of >= 200 000 active particles are done pair-symmetrically: the
 *                denominator R^2 - x_i.x_j and its reciprocal / logarithm serve both i <- j and j <- i
 *                (csrc/symmetric.cuh), with exact fixed-point accumulation, so results are bit-identical
 *                from run to run and for any number of ranks;
 *   0            always the one-sided engine (every ordered pair evaluated, as the reference does).
 * The two differ by summation order only (~1e-13 relative).  In rank mode every rank must set the same value. */
int lpm_set_symmetric(int enable);
/* Evaluation order of the compactly supported PSE kernels (tests and benchmarks):
 *   0  reference order (j = 1..N over the active particles), every source tile visited;
 *   1  sources and targets in cell (Morton) order, tiles out of reach skipped (default);
 *   2  cell order, every tile visited -- bit-identical to mode 1.
 * Modes 1/2 differ from mode 0 by summation order only (~1e-16 relative). */
int lpm_set_pse_culling(int mode);
/* Sphere PSE kernels: 1 (default) evaluates (d_ij / eps)^2 inside the cut-off from
 * tan^2(theta / 2) = |x_i cross x_j|^2 / (|x_i| |x_j| + x_i . x_j)^2 and a short series for
 * atan^2; 0 always calls sqrt + atan2 as the reference writes it.  Same value to ~1e-15. */
int lpm_set_pse_series(int enable);

#ifdef __cplusplus
}
#endif
#endif /* LPM_GPU_H */
