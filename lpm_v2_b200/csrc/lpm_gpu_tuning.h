/* lpm_gpu_tuning.h -- NOT part of the C ABI (include/lpm_gpu.h is).  One entry point for the A/B runs under
 * tools/ and for tests that need a small problem to take a large problem's code path.  A Fortran caller never
 * uses it; keys may disappear without notice.
 *   "max_chunks"        upper bound on source chunks per evaluation (every rank the same value)
 *   "chunk_min"         smallest source chunk (a multiple of 256; every rank the same value)
 *   "fuse_step_end"     0: the BVE RK4 step ends with separate velocity and stream-function sums (A/B; default 1: fused)
 *   "force_T"           targets per thread of the one-sided engine: 1, 2, 4, 8 (0 = automatic)
 *   "sym_panel_blocks"  target blocks per panel of the triangle kernel's launch order (default 256)
 *   "sym_chunk_tiles"   source tiles per CTA of the symmetric triangle kernel (default 16; every rank the same value)
 *   "sym_min_sources"   smallest active-particle count that takes the pair-symmetric path (default 200000)
 *   "sym_vel_order"     statement order of the symmetric velocity kernel, in builds with -DLPM_SYM_ORDER_SWEEP only
 * (Builds under A/B get a key here while a round measures them -- tools/ab_builds.py drove "sym_vel_build",
 * "sym_stream_build" and "sym_velstream_build" in round 2; none is pending.) */
#ifndef LPM_GPU_TUNING_H
#define LPM_GPU_TUNING_H
#ifdef __cplusplus
extern "C" {
#endif
int lpm_tune(const char* key, int value);
#ifdef __cplusplus
}
#endif
#endif
