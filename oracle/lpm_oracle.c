/*
 * lpm_oracle.c -- CPU restatement of lpm-v2's O(N^2) direct-sum hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lpm_v2_b200/, the
 * C-ABI library liblpmgpu.so) may include, link, import or call this file.
 * It is used by tests/, by __graft_entry__.smoke() as the checker, and by
 * bench.py's cpu_baseline / --impl reference legs.
 *
 * Every function restates, loop for loop and operation for operation, the
 * Fortran routine cited above it (paths relative to the lpm-v2 source tree).
 * Build for parity with:  gcc -O2 -ffp-contract=off -fno-fast-math
 * (no FMA contraction, no reassociation: what gfortran -O2/-O3 emits for the
 * reference on baseline x86-64).  The timing build (oracle/Makefile, target
 * liblpm_oracle_fast.so) uses -O3 -march=native and the *_mt entry points.
 *
 * PARITY PIN STATUS
 *   - PSE Laplacian (sphere): pinned by the reference's own regression
 *     thresholds, tests/SpherePSEConvTest.f90:373-390 (checked in
 *     tests/test_oracle_golden.py).
 *   - Every sum and RK4 step (BVE / planar / beta-plane velocity, the
 *     mesh-side twins, the stream functions, the PSE Laplacians and the other
 *     PSE operators, the shallow-water right-hand sides and the planar
 *     shallow-water step, TotalKE / TotalEnstrophy, LoadBalance): pinned BIT
 *     FOR BIT against the reference's own source text,
 *     executed by the Fortran-subset interpreter oracle/fortran_subset.py on
 *     small cases (oracle/make_refsrc_fixtures.py -> tests/golden/refsrc_*.npz,
 *     checked by tests/test_refsrc_golden.py).  That removes the hand-written
 *     restatement from the pin; it is not a compiled run of the reference (no
 *     Fortran compiler, no MPI in this environment): FMA contraction by a
 *     compiler and a different libm are outside what it can show.
 *   - In addition the BVE / planar / beta-plane sums are checked against the
 *     analytic solutions the reference's own examples log (solid-body
 *     rotation, examples/BVESolidBody.f90:231-243; Rossby-Haurwitz
 *     eigenfunction) to discretisation accuracy.
 *
 * Index convention: all ranges are 0-based half-open [ibeg, iend); the
 * Fortran slice indexStart(r)..indexEnd(r) is [indexStart-1, indexEnd).
 * mask is int32 (non-zero = active), the C view of logical(klog).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* src/TypeDefs.f90:31 */
static const double PI = 3.1415926535897932384626433832795027975;

/* ------------------------------------------------------------------ */
/* ------------------------------------------------------------------ */
/* Host threads for the target loops of the velocity / stream-function sums (tests on config-sized
 * meshes, bench.py's parity sample).  Each target's sum is computed by one thread in the reference's
 * own j order, so results are bit-identical for any thread count -- the same property the reference
 * has for any number of MPI ranks (src/MPISetup.f90:132-146: each rank owns whole targets).  Default 1.
 * A threaded call splits [ibeg, iend) evenly and re-enters the same function per sub-range.            */
static int g_threads = 1;
static __thread int tl_in_worker = 0;
void oracle_set_threads(int nthreads) { g_threads = nthreads < 1 ? 1 : nthreads; }
int oracle_get_threads(void) { return g_threads; }

typedef struct {
    int kind; int64_t n; const double *in[6]; const int32_t *mask; double R; int64_t ibeg, iend; double *out[3];
} range_job;
static void range_call(const range_job *j);
static void *range_worker(void *arg)
{
    tl_in_worker = 1;
    range_call((const range_job *)arg);
    return NULL;
}
/* returns 1 if the range was handled by worker threads */
static int range_split(int kind, int64_t n, const double *i0, const double *i1, const double *i2, const double *i3,
                       const double *i4, const double *i5, const int32_t *mask, double R, int64_t ibeg, int64_t iend,
                       double *o0, double *o1, double *o2)
{
    int nt = g_threads;
    if (nt <= 1 || tl_in_worker || iend - ibeg < 2 * (int64_t)nt) return 0;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
    range_job *jobs = (range_job *)malloc(sizeof(range_job) * nt);
    int64_t cnt = iend - ibeg;
    for (int t = 0; t < nt; ++t) {
        range_job j = { kind, n, { i0, i1, i2, i3, i4, i5 }, mask, R, ibeg + cnt * t / nt, ibeg + cnt * (t + 1) / nt, { o0, o1, o2 } };
        jobs[t] = j;
        pthread_create(&th[t], NULL, range_worker, &jobs[t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 1;
}
enum { K_BVE_VEL, K_BVE_VEL_MESH, K_BVE_STREAM, K_PLANE_VEL, K_PLANE_STREAM, K_BETA_VEL, K_BETA_STREAM };

/* src/MPISetup.f90:132-146  LoadBalance.  Outputs are the reference's
 * 1-based inclusive indexStart / indexEnd and messageLength.          */
void oracle_load_balance(int32_t nItems, int32_t nProcs, int32_t *indexStart,
                         int32_t *indexEnd, int32_t *messageLength)
{
    int32_t chunkSize = nItems / nProcs;
    for (int32_t i = 0; i < nProcs; ++i) {
        indexStart[i] = i * chunkSize + 1;
        indexEnd[i] = (i + 1) * chunkSize;
    }
    indexEnd[nProcs - 1] = nItems;
    for (int32_t i = 0; i < nProcs; ++i)
        messageLength[i] = indexEnd[i] - indexStart[i] + 1;
}

/* Fortran pack([(j,j=1,n)], mask): the active-source index list (0-based
 * here).  Returns the count.  This is the list the GPU compaction must
 * reproduce bit-exactly.                                               */
int64_t oracle_active_list(int64_t n, const int32_t *mask, int32_t *list)
{
    int64_t c = 0;
    for (int64_t j = 0; j < n; ++j)
        if (mask[j]) list[c++] = (int32_t)j;
    return c;
}

/* ------------------------------------------------------------------ */
/* src/SphereBVESolver.f90:377-430  BVESphereVelocity (loops :396-420).
 * Identical arithmetic to src/SphereBVE.f90:489-531.                   */
#define BVE_BODY                                                              \
    if (mask[j]) {                                                            \
        double strength = -relVort[j] * area[j] /                             \
            (4.0 * PI * R * (R * R - x[i] * x[j] - y[i] * y[j] - z[i] * z[j])); \
        u[i] = u[i] + (y[i] * z[j] - z[i] * y[j]) * strength;                 \
        v[i] = v[i] + (z[i] * x[j] - x[i] * z[j]) * strength;                 \
        w[i] = w[i] + (x[i] * y[j] - y[i] * x[j]) * strength;                 \
    }

void oracle_bve_velocity(int64_t n, const double *x, const double *y, const double *z,
                         const double *relVort, const double *area, const int32_t *mask,
                         double R, int64_t ibeg, int64_t iend,
                         double *u, double *v, double *w)
{
    if (range_split(K_BVE_VEL, n, x, y, z, relVort, area, NULL, mask, R, ibeg, iend, u, v, w)) return;
    for (int64_t i = ibeg; i < iend; ++i) {
        u[i] = 0.0; v[i] = 0.0; w[i] = 0.0;
        for (int64_t j = 0; j < i; ++j) BVE_BODY
        for (int64_t j = i + 1; j < n; ++j) BVE_BODY
    }
}

/* src/SphereBVE.f90:489-531  setVelocityFromVorticity: the mesh-side twin;
 * differs from the solver kernel only in forming sum(xi*xj) before the
 * subtraction from R*R (:505-506).                                      */
#define BVE_MESH_BODY                                                         \
    if (mask[j]) {                                                            \
        double strength = -relVort[j] * area[j] /                             \
            (4.0 * PI * R * (R * R - (x[i] * x[j] + y[i] * y[j] + z[i] * z[j]))); \
        u[i] = u[i] + (y[i] * z[j] - z[i] * y[j]) * strength;                 \
        v[i] = v[i] + (z[i] * x[j] - x[i] * z[j]) * strength;                 \
        w[i] = w[i] + (x[i] * y[j] - y[i] * x[j]) * strength;                 \
    }

void oracle_bve_velocity_mesh(int64_t n, const double *x, const double *y, const double *z,
                              const double *relVort, const double *area, const int32_t *mask,
                              double R, int64_t ibeg, int64_t iend,
                              double *u, double *v, double *w)
{
    if (range_split(K_BVE_VEL_MESH, n, x, y, z, relVort, area, NULL, mask, R, ibeg, iend, u, v, w)) return;
    for (int64_t i = ibeg; i < iend; ++i) {
        u[i] = 0.0; v[i] = 0.0; w[i] = 0.0;
        for (int64_t j = 0; j < i; ++j) BVE_MESH_BODY
        for (int64_t j = i + 1; j < n; ++j) BVE_MESH_BODY
    }
}

/* Extended-precision adjudicator for the same sum (not a restatement). */
void oracle_bve_velocity_ld(int64_t n, const double *x, const double *y, const double *z,
                            const double *relVort, const double *area, const int32_t *mask,
                            double R, int64_t ibeg, int64_t iend,
                            double *u, double *v, double *w)
{
    const long double PIl = 3.1415926535897932384626433832795027975L;
    for (int64_t i = ibeg; i < iend; ++i) {
        long double su = 0, sv = 0, sw = 0;
        long double xi = x[i], yi = y[i], zi = z[i], Rl = R;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i || !mask[j]) continue;
            long double xj = x[j], yj = y[j], zj = z[j];
            long double s = -(long double)relVort[j] * (long double)area[j] /
                (4.0L * PIl * Rl * (Rl * Rl - xi * xj - yi * yj - zi * zj));
            su += (yi * zj - zi * yj) * s;
            sv += (zi * xj - xi * zj) * s;
            sw += (xi * yj - yi * xj) * s;
        }
        u[i] = (double)su; v[i] = (double)sv; w[i] = (double)sw;
    }
}

/* src/SphereBVE.f90:445-485  SetStreamFunctionsOnMesh.                */
#define BVE_STREAM_BODY                                                       \
    if (mask[j]) {                                                            \
        double greensKernel = -log(R * R - (x[i] * x[j] + y[i] * y[j] + z[i] * z[j])) / (4.0 * PI); \
        relStream[i] = relStream[i] + greensKernel * relVort[j] * area[j];    \
        absStream[i] = absStream[i] + greensKernel * absVort[j] * area[j];    \
    }

void oracle_bve_stream(int64_t n, const double *x, const double *y, const double *z,
                       const double *relVort, const double *absVort, const double *area,
                       const int32_t *mask, double R, int64_t ibeg, int64_t iend,
                       double *relStream, double *absStream)
{
    if (range_split(K_BVE_STREAM, n, x, y, z, relVort, absVort, area, mask, R, ibeg, iend, relStream, absStream, NULL)) return;
    for (int64_t i = ibeg; i < iend; ++i) {
        relStream[i] = 0.0; absStream[i] = 0.0;
        for (int64_t j = 0; j < i; ++j) BVE_STREAM_BODY
        for (int64_t j = i + 1; j < n; ++j) BVE_STREAM_BODY
    }
}

void oracle_bve_stream_ld(int64_t n, const double *x, const double *y, const double *z,
                          const double *relVort, const double *absVort, const double *area,
                          const int32_t *mask, double R, int64_t ibeg, int64_t iend,
                          double *relStream, double *absStream)
{
    const long double PIl = 3.1415926535897932384626433832795027975L;
    for (int64_t i = ibeg; i < iend; ++i) {
        long double sr = 0, sa = 0, xi = x[i], yi = y[i], zi = z[i], Rl = R;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i || !mask[j]) continue;
            long double g = -logl(Rl * Rl - (xi * x[j] + yi * y[j] + zi * z[j])) / (4.0L * PIl);
            sr += g * (long double)relVort[j] * (long double)area[j];
            sa += g * (long double)absVort[j] * (long double)area[j];
        }
        relStream[i] = (double)sr; absStream[i] = (double)sa;
    }
}

/* ------------------------------------------------------------------ */
/* src/PlaneIncompressibleSolver.f90:278-316  planarIncompressibleVelocity
 * (loops :294-307); same arithmetic as src/PlanarIncompressible.f90:426-466.
 * Fortran (a)**2 is a*a.                                                */
#define PLANE_BODY                                                            \
    if (mask[j]) {                                                            \
        double strength = vort[j] * area[j] /                                 \
            (2.0 * PI * ((x[i] - x[j]) * (x[i] - x[j]) + (y[i] - y[j]) * (y[i] - y[j]))); \
        u[i] = u[i] - (y[i] - y[j]) * strength;                               \
        v[i] = v[i] + (x[i] - x[j]) * strength;                               \
    }

void oracle_plane_velocity(int64_t n, const double *x, const double *y, const double *vort,
                           const double *area, const int32_t *mask, int64_t ibeg, int64_t iend,
                           double *u, double *v)
{
    if (range_split(K_PLANE_VEL, n, x, y, vort, area, NULL, NULL, mask, 0.0, ibeg, iend, u, v, NULL)) return;
    for (int64_t i = ibeg; i < iend; ++i) {
        u[i] = 0.0; v[i] = 0.0;
        for (int64_t j = 0; j < i; ++j) PLANE_BODY
        for (int64_t j = i + 1; j < n; ++j) PLANE_BODY
    }
}

void oracle_plane_velocity_ld(int64_t n, const double *x, const double *y, const double *vort,
                              const double *area, const int32_t *mask, int64_t ibeg, int64_t iend,
                              double *u, double *v)
{
    const long double PIl = 3.1415926535897932384626433832795027975L;
    for (int64_t i = ibeg; i < iend; ++i) {
        long double su = 0, sv = 0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i || !mask[j]) continue;
            long double dx = (long double)x[i] - x[j], dy = (long double)y[i] - y[j];
            long double s = (long double)vort[j] * (long double)area[j] / (2.0L * PIl * (dx * dx + dy * dy));
            su -= dy * s; sv += dx * s;
        }
        u[i] = (double)su; v[i] = (double)sv;
    }
}

/* src/PlanarIncompressible.f90:470-505  SetStreamFunctionOnMesh.       */
#define PLANE_STREAM_BODY                                                     \
    if (mask[j]) {                                                            \
        double greensKernel = log(sqrt((x[i] - x[j]) * (x[i] - x[j]) +        \
                                       (y[i] - y[j]) * (y[i] - y[j]))) / (2.0 * PI); \
        psi[i] = psi[i] + greensKernel * vort[j] * area[j];                   \
    }

void oracle_plane_stream(int64_t n, const double *x, const double *y, const double *vort,
                         const double *area, const int32_t *mask, int64_t ibeg, int64_t iend,
                         double *psi)
{
    if (range_split(K_PLANE_STREAM, n, x, y, vort, area, NULL, NULL, mask, 0.0, ibeg, iend, psi, NULL, NULL)) return;
    for (int64_t i = ibeg; i < iend; ++i) {
        psi[i] = 0.0;
        for (int64_t j = 0; j < i; ++j) PLANE_STREAM_BODY
        for (int64_t j = i + 1; j < n; ++j) PLANE_STREAM_BODY
    }
}

void oracle_plane_stream_ld(int64_t n, const double *x, const double *y, const double *vort,
                            const double *area, const int32_t *mask, int64_t ibeg, int64_t iend,
                            double *psi)
{
    const long double PIl = 3.1415926535897932384626433832795027975L;
    for (int64_t i = ibeg; i < iend; ++i) {
        long double s = 0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i || !mask[j]) continue;
            long double dx = (long double)x[i] - x[j], dy = (long double)y[i] - y[j];
            s += logl(sqrtl(dx * dx + dy * dy)) / (2.0L * PIl) * (long double)vort[j] * (long double)area[j];
        }
        psi[i] = (double)s;
    }
}

/* ------------------------------------------------------------------ */
/* src/BetaPlaneSolver.f90:227-267  BetaPlaneVelocity (loops :243-258);
 * same arithmetic as src/BetaPlane.f90:359-397.                        */
#define BETA_BODY                                                             \
    if (mask[j]) {                                                            \
        double strength = 0.5 * relVort[j] * area[j] /                        \
            (cosh(2.0 * PI * (y[i] - y[j])) - cos(2.0 * PI * (x[i] - x[j]))); \
        u[i] = u[i] - sinh(2.0 * PI * (y[i] - y[j])) * strength;              \
        v[i] = v[i] + sin(2.0 * PI * (x[i] - x[j])) * strength;               \
    }

void oracle_betaplane_velocity(int64_t n, const double *x, const double *y, const double *relVort,
                               const double *area, const int32_t *mask, int64_t ibeg, int64_t iend,
                               double *u, double *v)
{
    if (range_split(K_BETA_VEL, n, x, y, relVort, area, NULL, NULL, mask, 0.0, ibeg, iend, u, v, NULL)) return;
    for (int64_t i = ibeg; i < iend; ++i) {
        u[i] = 0.0; v[i] = 0.0;
        for (int64_t j = 0; j < i; ++j) BETA_BODY
        for (int64_t j = i + 1; j < n; ++j) BETA_BODY
    }
}

/* Extended-precision adjudicator.  The reference expression cosh(a)-cos(b)
 * cancels catastrophically for near pairs ("Beta-plane kernel");
 * here the denominator is evaluated as 2 sinh^2(a/2) + 2 sin^2(b/2).     */
void oracle_betaplane_velocity_ld(int64_t n, const double *x, const double *y, const double *relVort,
                                  const double *area, const int32_t *mask, int64_t ibeg, int64_t iend,
                                  double *u, double *v)
{
    const long double PIl = 3.1415926535897932384626433832795027975L;
    for (int64_t i = ibeg; i < iend; ++i) {
        long double su = 0, sv = 0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i || !mask[j]) continue;
            long double a = 2.0L * PIl * ((long double)y[i] - y[j]);
            long double b = 2.0L * PIl * ((long double)x[i] - x[j]);
            long double sh = sinhl(0.5L * a), sn = sinl(0.5L * b);
            long double den = 2.0L * sh * sh + 2.0L * sn * sn;
            long double s = 0.5L * (long double)relVort[j] * (long double)area[j] / den;
            su -= sinhl(a) * s; sv += sinl(b) * s;
        }
        u[i] = (double)su; v[i] = (double)sv;
    }
}

/* src/BetaPlane.f90:399-442  SetStreamFunctionsOnMesh.                 */
#define BETA_STREAM_BODY                                                      \
    if (mask[j]) {                                                            \
        double greensKernel = log(cosh(2.0 * PI * (y[i] - y[j])) -            \
                                  cos(2.0 * PI * (x[i] - x[j]))) / (4.0 * PI); \
        relStream[i] = relStream[i] + greensKernel * relVort[j] * area[j];    \
        absStream[i] = absStream[i] + greensKernel * absVort[j] * area[j];    \
    }

void oracle_betaplane_stream(int64_t n, const double *x, const double *y, const double *relVort,
                             const double *absVort, const double *area, const int32_t *mask,
                             int64_t ibeg, int64_t iend, double *relStream, double *absStream)
{
    if (range_split(K_BETA_STREAM, n, x, y, relVort, absVort, area, NULL, mask, 0.0, ibeg, iend, relStream, absStream, NULL)) return;
    for (int64_t i = ibeg; i < iend; ++i) {
        absStream[i] = 0.0; relStream[i] = 0.0;
        for (int64_t j = 0; j < i; ++j) BETA_STREAM_BODY
        for (int64_t j = i + 1; j < n; ++j) BETA_STREAM_BODY
    }
}

void oracle_betaplane_stream_ld(int64_t n, const double *x, const double *y, const double *relVort,
                                const double *absVort, const double *area, const int32_t *mask,
                                int64_t ibeg, int64_t iend, double *relStream, double *absStream)
{
    const long double PIl = 3.1415926535897932384626433832795027975L;
    for (int64_t i = ibeg; i < iend; ++i) {
        long double sr = 0, sa = 0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i || !mask[j]) continue;
            long double a = 2.0L * PIl * ((long double)y[i] - y[j]);
            long double b = 2.0L * PIl * ((long double)x[i] - x[j]);
            long double sh = sinhl(0.5L * a), sn = sinl(0.5L * b);
            long double g = logl(2.0L * sh * sh + 2.0L * sn * sn) / (4.0L * PIl);
            sr += g * (long double)relVort[j] * (long double)area[j];
            sa += g * (long double)absVort[j] * (long double)area[j];
        }
        relStream[i] = (double)sr; absStream[i] = (double)sa;
    }
}

static void range_call(const range_job *j)
{
    const double *const *a = j->in;
    switch (j->kind) {
        case K_BVE_VEL: oracle_bve_velocity(j->n, a[0], a[1], a[2], a[3], a[4], j->mask, j->R, j->ibeg, j->iend, j->out[0], j->out[1], j->out[2]); break;
        case K_BVE_VEL_MESH: oracle_bve_velocity_mesh(j->n, a[0], a[1], a[2], a[3], a[4], j->mask, j->R, j->ibeg, j->iend, j->out[0], j->out[1], j->out[2]); break;
        case K_BVE_STREAM: oracle_bve_stream(j->n, a[0], a[1], a[2], a[3], a[4], a[5], j->mask, j->R, j->ibeg, j->iend, j->out[0], j->out[1]); break;
        case K_PLANE_VEL: oracle_plane_velocity(j->n, a[0], a[1], a[2], a[3], j->mask, j->ibeg, j->iend, j->out[0], j->out[1]); break;
        case K_PLANE_STREAM: oracle_plane_stream(j->n, a[0], a[1], a[2], a[3], j->mask, j->ibeg, j->iend, j->out[0]); break;
        case K_BETA_VEL: oracle_betaplane_velocity(j->n, a[0], a[1], a[2], a[3], j->mask, j->ibeg, j->iend, j->out[0], j->out[1]); break;
        case K_BETA_STREAM: oracle_betaplane_stream(j->n, a[0], a[1], a[2], a[3], a[4], j->mask, j->ibeg, j->iend, j->out[0], j->out[1]); break;
        default: break;
    }
}

/* ------------------------------------------------------------------ */
/* src/SphereGeometry.f90:67-73  ChordDistance                          */
static double ChordDistance(const double a[3], const double b[3])
{
    return sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) +
                (b[2] - a[2]) * (b[2] - a[2]));
}

/* src/SphereGeometry.f90:107-125  SphereDistanceVector.  sphereRadius is
 * the module-global SphereRadius (src/TypeDefs.f90:74), passed in.      */
static double SphereDistance(const double a[3], const double b[3], double sphereRadius)
{
    double cp0 = a[1] * b[2] - b[1] * a[2];
    double cp1 = b[0] * a[2] - a[0] * b[2];
    double cp2 = a[0] * b[1] - b[0] * a[1];
    double crossNorm = sqrt(cp0 * cp0 + cp1 * cp1 + cp2 * cp2);
    double dotProd = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    return atan2(crossNorm, dotProd) * sphereRadius;
}

/* src/PSEDirectSum.f90:622-627  bivariateLaplacianKernel8.
 * gfortran expands r**2, r**4, r**6 to repeated products.               */
static double bivariateLaplacianKernel8(double r)
{
    double r2 = r * r;
    double r4 = r2 * r2;
    double r6 = r4 * r2;
    return (40.0 - 40.0 * r2 + 10.0 * r4 - 2.0 * r6 / 3.0) * exp(-r * r) / PI;
}

double oracle_pse_laplacian_kernel8(double r) { return bivariateLaplacianKernel8(r); }

/* src/PSEDirectSum.f90:502-535  PSESphereLaplacianAtParticles (j = i is
 * included; the field is zeroed, summed, then scaled by 1/eps^2, :534). */
void oracle_pse_laplacian_sphere(int64_t n, const double *x, const double *y, const double *z,
                                 const double *f, const double *area, const int32_t *mask,
                                 double eps, double sphereRadius, int64_t ibeg, int64_t iend,
                                 double *lap)
{
    for (int64_t i = ibeg; i < iend; ++i) {
        double xi[3] = { x[i], y[i], z[i] };
        lap[i] = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double xj[3] = { x[j], y[j], z[j] };
                double kIn = SphereDistance(xi, xj, sphereRadius) / eps;
                lap[i] = lap[i] + (f[j] - f[i]) * bivariateLaplacianKernel8(kIn) / (eps * eps) * area[j];
            }
        }
    }
    for (int64_t i = ibeg; i < iend; ++i) lap[i] = (1.0 / (eps * eps)) * lap[i];
}

/* src/PSEDirectSum.f90:467-500  PSEPlaneLaplacianAtParticles (z == 0,
 * src/Particles.f90:663-670).                                           */
void oracle_pse_laplacian_plane(int64_t n, const double *x, const double *y,
                                const double *f, const double *area, const int32_t *mask,
                                double eps, int64_t ibeg, int64_t iend, double *lap)
{
    for (int64_t i = ibeg; i < iend; ++i) {
        double xi[3] = { x[i], y[i], 0.0 };
        lap[i] = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double xj[3] = { x[j], y[j], 0.0 };
                double kIn = ChordDistance(xi, xj) / eps;
                lap[i] = lap[i] + (f[j] - f[i]) * bivariateLaplacianKernel8(kIn) / (eps * eps) * area[j];
            }
        }
    }
    for (int64_t i = ibeg; i < iend; ++i) lap[i] = (1.0 / (eps * eps)) * lap[i];
}

void oracle_pse_laplacian_sphere_ld(int64_t n, const double *x, const double *y, const double *z,
                                    const double *f, const double *area, const int32_t *mask,
                                    double eps, double sphereRadius, int64_t ibeg, int64_t iend,
                                    double *lap)
{
    const long double PIl = 3.1415926535897932384626433832795027975L;
    long double e = eps;
    for (int64_t i = ibeg; i < iend; ++i) {
        long double s = 0, xi = x[i], yi = y[i], zi = z[i];
        for (int64_t j = 0; j < n; ++j) {
            if (!mask[j]) continue;
            long double xj = x[j], yj = y[j], zj = z[j];
            long double c0 = yi * zj - yj * zi, c1 = xj * zi - xi * zj, c2 = xi * yj - xj * yi;
            long double d = atan2l(sqrtl(c0 * c0 + c1 * c1 + c2 * c2), xi * xj + yi * yj + zi * zj) * sphereRadius;
            long double k = d / e, k2 = k * k;
            long double ker = (40.0L - 40.0L * k2 + 10.0L * k2 * k2 - 2.0L * k2 * k2 * k2 / 3.0L) * expl(-k2) / PIl;
            s += ((long double)f[j] - f[i]) * ker / (e * e) * area[j];
        }
        lap[i] = (double)(s / (e * e));
    }
}

/* ------------------------------------------------------------------ */
/* src/SphereBVESolver.f90:219-353  timestepPrivate (BVE RK4), with the
 * trailing SetStreamFunctionsOnMesh (:352) left to the caller.  State
 * arrays are updated in place; u,v,w enter as the velocity at the old
 * state and leave as the velocity at the new state (:345-350).          */
void oracle_bve_rk4_step(int64_t n, double *x, double *y, double *z, double *relVort,
                         double *u, double *v, double *w, const double *area,
                         const int32_t *mask, double R, double Omega, double dt)
{
    size_t nb = (size_t)n * sizeof(double);
    double *buf = (double *)malloc(nb * 20);
    double *xIn = buf, *yIn = buf + n, *zIn = buf + 2 * n, *vIn = buf + 3 * n;
    double *xS[4], *yS[4], *zS[4], *vS[4];
    for (int s = 0; s < 4; ++s) {
        xS[s] = buf + (4 + 4 * s) * n; yS[s] = xS[s] + n; zS[s] = yS[s] + n; vS[s] = zS[s] + n;
    }
    /* stage 1 (:239-244) */
    for (int64_t i = 0; i < n; ++i) {
        xS[0][i] = dt * u[i]; yS[0][i] = dt * v[i]; zS[0][i] = dt * w[i];
        vS[0][i] = -dt * 2.0 * Omega * w[i] / R;
    }
    for (int s = 1; s < 4; ++s) {
        /* stage inputs :250-255, :273-278, :296-301 */
        for (int64_t i = 0; i < n; ++i) {
            if (s < 3) {
                xIn[i] = x[i] + 0.5 * xS[s - 1][i]; yIn[i] = y[i] + 0.5 * yS[s - 1][i];
                zIn[i] = z[i] + 0.5 * zS[s - 1][i]; vIn[i] = relVort[i] + 0.5 * vS[s - 1][i];
            } else {
                xIn[i] = x[i] + xS[s - 1][i]; yIn[i] = y[i] + yS[s - 1][i];
                zIn[i] = z[i] + zS[s - 1][i]; vIn[i] = relVort[i] + vS[s - 1][i];
            }
        }
        oracle_bve_velocity(n, xIn, yIn, zIn, vIn, area, mask, R, 0, n, xS[s], yS[s], zS[s]);
        for (int64_t i = 0; i < n; ++i) {   /* :262-267 */
            vS[s][i] = -dt * 2.0 * Omega * zS[s][i] / R;
            xS[s][i] = dt * xS[s][i]; yS[s][i] = dt * yS[s][i]; zS[s][i] = dt * zS[s][i];
        }
    }
    for (int64_t i = 0; i < n; ++i) {       /* :320-329 */
        x[i] = x[i] + xS[0][i] / 6.0 + xS[1][i] / 3.0 + xS[2][i] / 3.0 + xS[3][i] / 6.0;
        y[i] = y[i] + yS[0][i] / 6.0 + yS[1][i] / 3.0 + yS[2][i] / 3.0 + yS[3][i] / 6.0;
        z[i] = z[i] + zS[0][i] / 6.0 + zS[1][i] / 3.0 + zS[2][i] / 3.0 + zS[3][i] / 6.0;
        relVort[i] = relVort[i] + vS[0][i] / 6.0 + vS[1][i] / 3.0 + vS[2][i] / 3.0 + vS[3][i] / 6.0;
    }
    oracle_bve_velocity(n, x, y, z, relVort, area, mask, R, 0, n, u, v, w);   /* :345-346 */
    free(buf);
}

/* src/PlaneIncompressibleSolver.f90:171-259 (stream function :258 left
 * to the caller).                                                       */
void oracle_plane_rk4_step(int64_t n, double *x, double *y, const double *vort,
                           double *u, double *v, const double *area, const int32_t *mask, double dt)
{
    double *buf = (double *)malloc((size_t)n * sizeof(double) * 10);
    double *xIn = buf, *yIn = buf + n, *xS[4], *yS[4];
    for (int s = 0; s < 4; ++s) { xS[s] = buf + (2 + 2 * s) * n; yS[s] = xS[s] + n; }
    for (int64_t i = 0; i < n; ++i) { xS[0][i] = dt * u[i]; yS[0][i] = dt * v[i]; }
    for (int s = 1; s < 4; ++s) {
        for (int64_t i = 0; i < n; ++i) {
            if (s < 3) { xIn[i] = x[i] + 0.5 * xS[s - 1][i]; yIn[i] = y[i] + 0.5 * yS[s - 1][i]; }
            else       { xIn[i] = x[i] + xS[s - 1][i];       yIn[i] = y[i] + yS[s - 1][i]; }
        }
        oracle_plane_velocity(n, xIn, yIn, vort, area, mask, 0, n, xS[s], yS[s]);
        for (int64_t i = 0; i < n; ++i) { xS[s][i] = dt * xS[s][i]; yS[s][i] = dt * yS[s][i]; }
    }
    for (int64_t i = 0; i < n; ++i) {
        x[i] = x[i] + xS[0][i] / 6.0 + xS[1][i] / 3.0 + xS[2][i] / 3.0 + xS[3][i] / 6.0;
        y[i] = y[i] + yS[0][i] / 6.0 + yS[1][i] / 3.0 + yS[2][i] / 3.0 + yS[3][i] / 6.0;
    }
    oracle_plane_velocity(n, x, y, vort, area, mask, 0, n, u, v);
    free(buf);
}

/* src/BetaPlaneSolver.f90:142-219 (stream functions :218 left to the
 * caller).  u,v enter as the mesh velocity, leave as SetVelocityOnMesh.  */
void oracle_betaplane_rk4_step(int64_t n, double *x, double *y, double *relVort,
                               double *u, double *v, const double *area, const int32_t *mask,
                               double beta, double dt)
{
    double *buf = (double *)malloc((size_t)n * sizeof(double) * 15);
    double *xIn = buf, *yIn = buf + n, *vIn = buf + 2 * n, *xS[4], *yS[4], *vS[4];
    for (int s = 0; s < 4; ++s) { xS[s] = buf + (3 + 3 * s) * n; yS[s] = xS[s] + n; vS[s] = yS[s] + n; }
    for (int64_t i = 0; i < n; ++i) {
        xS[0][i] = dt * u[i]; yS[0][i] = dt * v[i]; vS[0][i] = -dt * beta * v[i];
    }
    for (int s = 1; s < 4; ++s) {
        for (int64_t i = 0; i < n; ++i) {
            if (s < 3) {
                xIn[i] = x[i] + 0.5 * xS[s - 1][i]; yIn[i] = y[i] + 0.5 * yS[s - 1][i];
                vIn[i] = relVort[i] + 0.5 * vS[s - 1][i];
            } else {
                xIn[i] = x[i] + xS[s - 1][i]; yIn[i] = y[i] + yS[s - 1][i];
                vIn[i] = relVort[i] + vS[s - 1][i];
            }
        }
        oracle_betaplane_velocity(n, xIn, yIn, vIn, area, mask, 0, n, xS[s], yS[s]);
        for (int64_t i = 0; i < n; ++i) {
            vS[s][i] = -dt * beta * yS[s][i];
            xS[s][i] = dt * xS[s][i]; yS[s][i] = dt * yS[s][i];
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        x[i] = x[i] + xS[0][i] / 6.0 + xS[1][i] / 3.0 + xS[2][i] / 3.0 + xS[3][i] / 6.0;
        y[i] = y[i] + yS[0][i] / 6.0 + yS[1][i] / 3.0 + yS[2][i] / 3.0 + yS[3][i] / 6.0;
        relVort[i] = relVort[i] + vS[0][i] / 6.0 + vS[1][i] / 3.0 + vS[2][i] / 3.0 + vS[3][i] / 6.0;
    }
    oracle_betaplane_velocity(n, x, y, relVort, area, mask, 0, n, u, v);
    free(buf);
}

/* ------------------------------------------------------------------ */
/* src/SphereBVE.f90:410-441  TotalKE, TotalEnstrophy.                  */
double oracle_total_ke(int64_t n, const double *u, const double *v, const double *w,
                       const double *area, const int32_t *mask)
{
    double ke = 0.0;
    for (int64_t i = 0; i < n; ++i)
        if (mask[i]) {
            double magSq = u[i] * u[i] + v[i] * v[i] + w[i] * w[i];
            ke = ke + magSq * area[i];
        }
    return 0.5 * ke;
}

double oracle_total_enstrophy(int64_t n, const double *relVort, const double *area, const int32_t *mask)
{
    double e = 0.0;
    for (int64_t i = 0; i < n; ++i)
        if (mask[i]) e = e + relVort[i] * relVort[i] * area[i];
    return 0.5 * e;
}

/* ------------------------------------------------------------------ */
/* Multi-threaded driver used only for the CPU baseline timing: the
 * reference's replicated-data target split (LoadBalance above) with one
 * worker thread standing in for each MPI rank (mpirun -np nthreads).    */
typedef struct {
    int kind; int64_t n; const double *x, *y, *z, *q, *area; const int32_t *mask;
    double R; int64_t ibeg, iend; double *u, *v, *w;
} mt_job;

static void *mt_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    oracle_bve_velocity(j->n, j->x, j->y, j->z, j->q, j->area, j->mask, j->R, j->ibeg, j->iend, j->u, j->v, j->w);
    return NULL;
}

/* Evaluates targets [tbeg, tend) (a bounded sample of the full target
 * set) against all n sources, split over nthreads workers by LoadBalance. */
void oracle_bve_velocity_mt(int nthreads, int64_t n, const double *x, const double *y, const double *z,
                            const double *relVort, const double *area, const int32_t *mask,
                            double R, int64_t tbeg, int64_t tend, double *u, double *v, double *w)
{
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    mt_job *jobs = (mt_job *)malloc(sizeof(mt_job) * nthreads);
    int64_t cnt = tend - tbeg, chunk = cnt / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        mt_job j = { 0, n, x, y, z, relVort, area, mask, R,
                     tbeg + t * chunk, (t == nthreads - 1) ? tend : tbeg + (t + 1) * chunk, u, v, w };
        jobs[t] = j;
        pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* ================================================================== */
/* Remaining PSE operators.                    */

/* src/PSEDirectSum.f90:611-616  bivariateDeltaKernel8                 */
static double bivariateDeltaKernel8(double r)
{
    double r2 = r * r, r4 = r2 * r2, r6 = r4 * r2;
    return (4.0 - 6.0 * r2 + 2.0 * r4 - r6 / 6.0) * exp(-r * r) / PI;
}

/* src/PSEDirectSum.f90:629-634  bivariateFirstDerivativeKernel8       */
static double bivariateFirstDerivativeKernel8(double r)
{
    double r2 = r * r, r4 = r2 * r2, r6 = r4 * r2;
    return (-20.0 + 20.0 * r2 - 5.0 * r4 + r6 / 3.0) * exp(-r * r) / PI;
}

/* src/SphereGeometry.f90:578-590  SphereProjection, applied: P g      */
static void sphere_project(const double x[3], const double g[3], double out[3])
{
    double P[3][3];
    P[0][0] = 1.0 - x[0] * x[0]; P[1][0] = -x[1] * x[0]; P[2][0] = -x[2] * x[0];
    P[0][1] = -x[0] * x[1]; P[1][1] = 1.0 - x[1] * x[1]; P[2][1] = -x[2] * x[1];
    P[0][2] = -x[0] * x[2]; P[1][2] = -x[1] * x[2]; P[2][2] = 1.0 - x[2] * x[2];
    /* MATMUL as gfortran evaluates it (inline expansion and libgfortran alike): the result is zeroed, then
     * c(i) = c(i) + a(i,k) * b(k) for k = 1, 2, 3 -- the leading 0.0 only matters for the sign of a zero result
     * (found by tests/test_refsrc_golden.py against the interpreted reference text) */
    for (int r = 0; r < 3; ++r) {
        double c = 0.0;
        for (int k = 0; k < 3; ++k) c = c + P[r][k] * g[k];
        out[r] = c;
    }
}

/* src/PSEDirectSum.f90:128-168  PSE{Plane,Sphere}InterpolateScalar at m
 * arbitrary locations (tx,ty,tz); sphere != 0 selects SphereDistance.   */
void oracle_pse_interpolate(int sphere, int64_t n, const double *x, const double *y, const double *z,
                            const double *f, const double *area, const int32_t *mask, double eps,
                            double sphereRadius, int64_t m, const double *tx, const double *ty,
                            const double *tz, double *out)
{
    for (int64_t i = 0; i < m; ++i) {
        double loc[3] = { tx[i], ty[i], sphere ? tz[i] : 0.0 };
        double s = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double xj[3] = { x[j], y[j], sphere ? z[j] : 0.0 };
                double kIn = (sphere ? SphereDistance(xj, loc, sphereRadius) : ChordDistance(loc, xj)) / eps;
                s = s + f[j] * area[j] * bivariateDeltaKernel8(kIn) / (eps * eps);
            }
        }
        out[i] = s;
    }
}

/* src/PSEDirectSum.f90:180-218  PSEPlaneGradientAtParticles            */
void oracle_pse_gradient_plane(int64_t n, const double *x, const double *y, const double *f,
                               const double *area, const int32_t *mask, double eps,
                               int64_t ibeg, int64_t iend, double *gx, double *gy)
{
    for (int64_t i = ibeg; i < iend; ++i) {
        double xi[3] = { x[i], y[i], 0.0 };
        gx[i] = 0.0; gy[i] = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double xj[3] = { x[j], y[j], 0.0 };
                double kIn = ChordDistance(xi, xj) / eps;
                gx[i] = gx[i] + (f[j] + f[i]) * (xi[0] - xj[0]) * bivariateFirstDerivativeKernel8(kIn) / (eps * eps * eps) * area[j];
                gy[i] = gy[i] + (f[j] + f[i]) * (xi[1] - xj[1]) * bivariateFirstDerivativeKernel8(kIn) / (eps * eps * eps) * area[j];
            }
        }
    }
    for (int64_t i = ibeg; i < iend; ++i) { gx[i] = (1.0 / eps) * gx[i]; gy[i] = (1.0 / eps) * gy[i]; }
}

/* src/PSEDirectSum.f90:221-267  PSESphereGradientAtParticles           */
void oracle_pse_gradient_sphere(int64_t n, const double *x, const double *y, const double *z, const double *f,
                                const double *area, const int32_t *mask, double eps, double sphereRadius,
                                int64_t ibeg, int64_t iend, double *gx, double *gy, double *gz)
{
    for (int64_t i = ibeg; i < iend; ++i) {
        double xi[3] = { x[i], y[i], z[i] };
        double grad[3] = { 0.0, 0.0, 0.0 }, pg[3];
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double xj[3] = { x[j], y[j], z[j] };
                double kIn = SphereDistance(xj, xi, sphereRadius) / eps;
                double kOut = bivariateFirstDerivativeKernel8(kIn) / (eps * eps);
                grad[0] = grad[0] + (f[j] + f[i]) * (xi[0] - xj[0]) * kOut * area[j];
                grad[1] = grad[1] + (f[j] + f[i]) * (xi[1] - xj[1]) * kOut * area[j];
                grad[2] = grad[2] + (f[j] + f[i]) * (xi[2] - xj[2]) * kOut * area[j];
            }
        }
        sphere_project(xi, grad, pg);
        gx[i] = pg[0]; gy[i] = pg[1]; gz[i] = pg[2];
    }
    for (int64_t i = ibeg; i < iend; ++i) {
        gx[i] = (1.0 / (eps * eps)) * gx[i]; gy[i] = (1.0 / (eps * eps)) * gy[i]; gz[i] = (1.0 / (eps * eps)) * gz[i];
    }
}

/* src/PSEDirectSum.f90:269-320  PSEPlaneSecondPartialsAtParticles: outputs
 * (d_xx, (d_xy + d_yx)/2, d_yy) in xComp, yComp, zComp.                 */
void oracle_pse_second_partials_plane(int64_t n, const double *x, const double *y, const double *gxIn,
                                      const double *gyIn, const double *area, const int32_t *mask, double eps,
                                      int64_t ibeg, int64_t iend, double *oxx, double *oxy, double *oyy)
{
    for (int64_t i = ibeg; i < iend; ++i) {
        double xi[3] = { x[i], y[i], 0.0 };
        double dxx = 0.0, dxy = 0.0, dyx = 0.0, dyy = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double xj[3] = { x[j], y[j], 0.0 };
                double kIn = ChordDistance(xi, xj) / eps;
                double kOut = bivariateFirstDerivativeKernel8(kIn) / (eps * eps);
                dxx = dxx + (gxIn[j] + gxIn[i]) * (xi[0] - xj[0]) / eps * kOut * area[j];
                dxy = dxy + (gxIn[j] + gxIn[i]) * (xi[1] - xj[1]) / eps * kOut * area[j];
                dyx = dyx + (gyIn[j] + gyIn[i]) * (xi[0] - xj[0]) / eps * kOut * area[j];
                dyy = dyy + (gyIn[j] + gyIn[i]) * (xi[1] - xj[1]) / eps * kOut * area[j];
            }
        }
        oxx[i] = dxx; oxy[i] = 0.5 * (dxy + dyx); oyy[i] = dyy;
    }
    for (int64_t i = ibeg; i < iend; ++i) {
        oxx[i] = (1.0 / eps) * oxx[i]; oxy[i] = (1.0 / eps) * oxy[i]; oyy[i] = (1.0 / eps) * oyy[i];
    }
}

/* src/PSEDirectSum.f90:322-365  PSEPlaneDoubleDotProductAtParticles    */
void oracle_pse_double_dot_plane(int64_t n, const double *x, const double *y, const double *u, const double *v,
                                 const double *area, const int32_t *mask, double eps,
                                 int64_t ibeg, int64_t iend, double *dd)
{
    for (int64_t i = ibeg; i < iend; ++i) {
        double xi[3] = { x[i], y[i], 0.0 };
        double ux = 0.0, uy = 0.0, vx = 0.0, vy = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double xj[3] = { x[j], y[j], 0.0 };
                double kIn = ChordDistance(xi, xj) / eps;
                double kOut = bivariateFirstDerivativeKernel8(kIn) / (eps * eps);
                ux = ux + (u[j] + u[i]) * (xi[0] - xj[0]) / eps * kOut * area[j];
                uy = uy + (u[j] + u[i]) * (xi[1] - xj[1]) / eps * kOut * area[j];
                vx = vx + (v[j] + v[i]) * (xi[0] - xj[0]) / eps * kOut * area[j];
                vy = vy + (v[j] + v[i]) * (xi[1] - xj[1]) / eps * kOut * area[j];
            }
        }
        dd[i] = (ux * ux + 2.0 * uy * vx + vy * vy) / (eps * eps);
    }
}

/* src/PSEDirectSum.f90:367-420  PSESphereDoubleDotProductAtParticles.
 * The w rows add vectorField%yComp(i) (not zComp(i)) at :408-413 -- a latent
 * reference quirk, restated as written.                                 */
void oracle_pse_double_dot_sphere(int64_t n, const double *x, const double *y, const double *z,
                                  const double *u, const double *v, const double *w,
                                  const double *area, const int32_t *mask, double eps, double sphereRadius,
                                  int64_t ibeg, int64_t iend, double *dd)
{
    for (int64_t i = ibeg; i < iend; ++i) {
        double xi[3] = { x[i], y[i], z[i] };
        double ux = 0, uy = 0, uz = 0, vx = 0, vy = 0, vz = 0, wx = 0, wy = 0, wz = 0;
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double xj[3] = { x[j], y[j], z[j] };
                double kIn = SphereDistance(xi, xj, sphereRadius) / eps;
                double kOut = bivariateFirstDerivativeKernel8(kIn) / (eps * eps);
                ux = ux + (u[j] + u[i]) * (xi[0] - xj[0]) / eps * kOut * area[j];
                uy = uy + (u[j] + u[i]) * (xi[1] - xj[1]) / eps * kOut * area[j];
                uz = uz + (u[j] + u[i]) * (xi[2] - xj[2]) / eps * kOut * area[j];
                vx = vx + (v[j] + v[i]) * (xi[0] - xj[0]) / eps * kOut * area[j];
                vy = vy + (v[j] + v[i]) * (xi[1] - xj[1]) / eps * kOut * area[j];
                vz = vz + (v[j] + v[i]) * (xi[2] - xj[2]) / eps * kOut * area[j];
                wx = wx + (w[j] + v[i]) * (xi[0] - xj[0]) / eps * kOut * area[j];
                wy = wy + (w[j] + v[i]) * (xi[1] - xj[1]) / eps * kOut * area[j];
                wz = wz + (w[j] + v[i]) * (xi[2] - xj[2]) / eps * kOut * area[j];
            }
        }
        dd[i] = (ux * ux + vy * vy + wz * wz + 2.0 * (uy * vx + uz * wx + vz * wy)) / (eps * eps);
    }
}

/* src/PSEDirectSum.f90:537-579  PSESphereDivergenceAtParticles          */
void oracle_pse_divergence_sphere(int64_t n, const double *x, const double *y, const double *z,
                                  const double *u, const double *v, const double *w,
                                  const double *area, const int32_t *mask, double eps, double sphereRadius,
                                  int64_t ibeg, int64_t iend, double *div)
{
    double denom = eps * eps * eps;
    for (int64_t i = ibeg; i < iend; ++i) {
        double xi[3] = { x[i], y[i], z[i] };
        div[i] = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double xj[3] = { x[j], y[j], z[j] };
                double kIn = SphereDistance(xj, xi, sphereRadius) / eps;
                double g[3], pg[3];
                g[0] = (xi[0] - xj[0]) * bivariateFirstDerivativeKernel8(kIn) / denom;
                g[1] = (xi[1] - xj[1]) * bivariateFirstDerivativeKernel8(kIn) / denom;
                g[2] = (xi[2] - xj[2]) * bivariateFirstDerivativeKernel8(kIn) / denom;
                sphere_project(xi, g, pg);
                div[i] = div[i] + (pg[0] * (u[j] + u[i]) + pg[1] * (v[j] + v[i]) + pg[2] * (w[j] + w[i])) * area[j];
            }
        }
    }
    for (int64_t i = ibeg; i < iend; ++i) div[i] = (1.0 / eps) * div[i];
}

/* ================================================================== */
/* src/SWEPlaneSolver.f90:457-560  SWEPlaneRHSIntegrals: velocity from the
 * vorticity and divergence, the four velocity-gradient sums behind the
 * double dot product, and the PSE Laplacian of the fluid surface, fused
 * in one pair loop.  surf(j) = h(j) + topoFn(x(j), y(j)) (the reference
 * evaluates the topography function inside the loop, :483,:490).       */
void oracle_swe_plane_rhs(int64_t n, const double *x, const double *y, const double *vort, const double *div,
                          const double *surf, const double *area, const int32_t *mask, double pseEps,
                          int64_t ibeg, int64_t iend, double *u, double *v, double *doubleDot, double *lapSurf)
{
    for (int64_t i = ibeg; i < iend; ++i) {
        u[i] = 0.0; v[i] = 0.0; doubleDot[i] = 0.0; lapSurf[i] = 0.0;
        double surfHeightI = surf[i];
        double ux = 0.0, uy = 0.0, vx = 0.0, vy = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                double surfHeightJ = surf[j];
                double sqdist = (x[i] - x[j]) * (x[i] - x[j]) + (y[i] - y[j]) * (y[i] - y[j]);
                double pseKin = sqrt(sqdist) / pseEps;
                double lapKernel = bivariateLaplacianKernel8(pseKin) / (pseEps * pseEps);
                lapSurf[i] = lapSurf[i] + lapKernel * (surfHeightJ - surfHeightI) * area[j];
                if (i == j) continue;
                double denom = 2.0 * PI * sqdist;
                double denom2 = PI * sqdist * sqdist;
                double rotStrength = vort[j] * area[j] / denom;
                double potStrength = div[j] * area[j] / denom;
                u[i] = u[i] - (y[i] - y[j]) * rotStrength + (x[i] - x[j]) * potStrength;
                v[i] = v[i] + (x[i] - x[j]) * rotStrength + (y[i] - y[j]) * potStrength;
                ux = ux + potStrength - ((x[i] - x[j]) * ((x[i] - x[j]) * div[j] - (y[i] - y[j]) * vort[j])) * area[j] / denom2;
                uy = uy - rotStrength - ((y[i] - y[j]) * ((x[i] - x[j]) * div[j] - (y[i] - y[j]) * vort[j])) * area[j] / denom2;
                vx = vx + rotStrength - ((x[i] - x[j]) * ((y[i] - y[j]) * div[j] + (x[i] - x[j]) * vort[j])) * area[j] / denom2;
                vy = vy + potStrength - ((y[i] - y[j]) * ((y[i] - y[j]) * div[j] + (x[i] - x[j]) * vort[j])) * area[j] / denom2;
            }
        }
        lapSurf[i] = lapSurf[i] / (pseEps * pseEps);
        doubleDot[i] = ux * ux + 2.0 * uy * vx + vy * vy;
    }
}

/* src/SWEPlaneSolver.f90:298-429  timestepPrivate (planar shallow-water RK4), AS WRITTEN.  State arrays
 * x, y, relVort, div, h, area are updated in place; u, v, doubleDot, lapSurf enter as the right-hand-side
 * integrals at the old state (what New() leaves, :202-204) and leave as those at the new state (:416-417).
 * topo(x, y) is the bottom topography (NULL = flat: topoFn == 0); the reference evaluates it inside the pair
 * loop, here once per particle and stage (same values).
 * Latent reference bug preserved: in stage 1 the relative-vorticity and divergence tendencies are assigned to the
 * WHOLE arrays inside the particle loop (`self%relVortStage1 = dt * (...)`, `self%divStage1 = dt * (...)`,
 * :312-315, no `(i)`), so after the loop every entry holds the value computed for the LAST particle.        */
typedef double (*oracle_topo_fn)(double x, double y);
static void swe_surface(int64_t n, const double *x, const double *y, const double *h, oracle_topo_fn topo, double *surf)
{
    for (int64_t i = 0; i < n; ++i) surf[i] = h[i] + (topo ? topo(x[i], y[i]) : 0.0);
}
void oracle_swe_plane_rk4_step(int64_t n, double *x, double *y, double *relVort, double *div, double *h, double *area,
                               double *u, double *v, double *doubleDot, double *lapSurf, const int32_t *mask,
                               double f0, double beta, double g, double pseEps, double dt, oracle_topo_fn topo)
{
    double *buf = (double *)malloc((size_t)n * sizeof(double) * 31);
    double *in[6], *st[4][6], *surf = buf + 30 * n;          /* order: x, y, relVort, div, area, h */
    double *start[6] = { x, y, relVort, div, area, h };
    for (int k = 0; k < 6; ++k) in[k] = buf + k * n;
    for (int s = 0; s < 4; ++s)
        for (int k = 0; k < 6; ++k) st[s][k] = buf + (6 + 6 * s + k) * n;
    /* stage 1 (:309-320) */
    for (int64_t i = 0; i < n; ++i) {
        st[0][0][i] = dt * u[i];
        st[0][1][i] = dt * v[i];
        st[0][5][i] = dt * (-h[i] * div[i]);
        st[0][4][i] = dt * (area[i] * div[i]);
    }
    {   /* the whole-array assignments: the last particle's value everywhere */
        const int64_t i = n - 1;
        const double rvs = dt * (-(relVort[i] + f0 + beta * y[i]) * div[i] - beta * v[i]);
        const double dvs = dt * (-doubleDot[i] + (f0 + beta * y[i]) * relVort[i] - g * lapSurf[i]);
        for (int64_t k = 0; k < n; ++k) { st[0][2][k] = rvs; st[0][3][k] = dvs; }
    }
    for (int s = 1; s < 4; ++s) {
        const double c = s < 3 ? 0.5 : 1.0;                 /* :325-332, :350-357, :375-382 (stage 4 adds the whole stage) */
        for (int k = 0; k < 6; ++k)
            for (int64_t i = 0; i < n; ++i) in[k][i] = (s < 3) ? start[k][i] + c * st[s - 1][k][i] : start[k][i] + st[s - 1][k][i];
        swe_surface(n, in[0], in[1], in[5], topo, surf);
        oracle_swe_plane_rhs(n, in[0], in[1], in[2], in[3], surf, in[4], mask, pseEps, 0, n, u, v, doubleDot, lapSurf);
        for (int64_t i = 0; i < n; ++i) {                   /* :337-346 */
            st[s][0][i] = dt * u[i];
            st[s][1][i] = dt * v[i];
            st[s][2][i] = dt * (-(in[2][i] + f0 + beta * in[1][i]) * in[3][i] - beta * v[i]);
            st[s][3][i] = dt * (-doubleDot[i] + (f0 + beta * in[1][i]) * in[2][i] - g * lapSurf[i]);
            st[s][5][i] = dt * (-in[5][i] * in[3][i]);
            st[s][4][i] = dt * (in[4][i] * in[3][i]);
        }
    }
    for (int k = 0; k < 6; ++k)                             /* :400-413 */
        for (int64_t i = 0; i < n; ++i)
            start[k][i] = start[k][i] + st[0][k][i] / 6.0 + st[1][k][i] / 3.0 + st[2][k][i] / 3.0 + st[3][k][i] / 6.0;
    swe_surface(n, x, y, h, topo, surf);
    oracle_swe_plane_rhs(n, x, y, relVort, div, surf, area, mask, pseEps, 0, n, u, v, doubleDot, lapSurf);   /* :416-417 */
    free(buf);
}

/* src/PlanarSWE.f90:469-494  SetVelocityFromFieldData (private body; SWEComputeVelocity :261-290
 * is the same sum): planar velocity from relative vorticity and divergence,
 *   u_i = sum_{j /= i, active} [ -(y_i - y_j) rot + (x_i - x_j) pot ],  v_i = sum [ (x_i - x_j) rot + (y_i - y_j) pot ],
 *   rot = zeta_j A_j / (2 pi r^2), pot = delta_j A_j / (2 pi r^2). */
void oracle_swe_plane_velocity(int64_t n, const double *x, const double *y, const double *vort, const double *div,
                               const double *area, const int32_t *mask, int64_t ibeg, int64_t iend, double *u, double *v)
{
    for (int64_t i = ibeg; i < iend; ++i) {
        u[i] = 0.0; v[i] = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (i != j && mask[j]) {
                double sqDist = (x[i] - x[j]) * (x[i] - x[j]) + (y[i] - y[j]) * (y[i] - y[j]);
                double denom = 2.0 * PI * sqDist;
                double rotStrength = vort[j] * area[j] / denom;
                double potStrength = div[j] * area[j] / denom;
                u[i] = u[i] - (y[i] - y[j]) * rotStrength + (x[i] - x[j]) * potStrength;
                v[i] = v[i] + (x[i] - x[j]) * rotStrength + (y[i] - y[j]) * potStrength;
            }
        }
    }
}

/* src/SphereSWESolver.f90:296-375  SWESphereRHSIntegrals, as written: velocity from vorticity
 * and divergence on the sphere and the PSE Laplacian of the fluid surface (surf = h + topography,
 * evaluated by the caller).  The reference routine is unfinished: doubleDot is zeroed and never
 * accumulated (:328, the ux.. sums are dropped) and nothing is broadcast; this restates what
 * it does compute.  The Laplacian is NOT post-scaled by 1/eps^2 here (the plane twin does it
 * at SWEPlaneSolver.f90:556; the sphere routine has no such line). */
void oracle_swe_sphere_rhs(int64_t n, const double *x, const double *y, const double *z, const double *relVort,
                           const double *div, const double *surf, const double *area, const int32_t *mask,
                           double radius, double pseEps, int64_t ibeg, int64_t iend, double *u, double *v, double *w,
                           double *doubleDot, double *lapSurf)
{
    const double fourPiRsq = 4.0 * PI * radius * radius;
    for (int64_t i = ibeg; i < iend; ++i) {
        u[i] = 0.0; v[i] = 0.0; w[i] = 0.0; doubleDot[i] = 0.0; lapSurf[i] = 0.0;
        const double surfHeightI = surf[i];
        for (int64_t j = 0; j < n; ++j) {
            if (mask[j]) {
                const double surfHeightJ = surf[j];
                const double xi[3] = {x[i], y[i], z[i]}, xj[3] = {x[j], y[j], z[j]};
                const double pseKin = SphereDistance(xi, xj, radius) / pseEps;
                const double lapKernel = bivariateLaplacianKernel8(pseKin) / (pseEps * pseEps);
                lapSurf[i] = lapSurf[i] + lapKernel * (surfHeightJ - surfHeightI) * area[j];
                if (i == j) continue;
                const double dotProd = x[i] * x[j] + y[i] * y[j] + z[i] * z[j];
                const double denom = fourPiRsq * (radius * radius - dotProd);
                const double rotStrength = relVort[j] * area[j] / denom;
                const double potStrength = radius * div[j] * area[j] / denom;
                u[i] = u[i] - (y[i] * z[j] - z[i] * y[j]) * rotStrength - x[j] * potStrength;
                v[i] = v[i] - (z[i] * x[j] - x[i] * z[j]) * rotStrength - y[j] * potStrength;
                w[i] = w[i] - (x[i] * y[j] - y[i] * x[j]) * rotStrength - z[j] * potStrength;
            }
        }
    }
}

/* tests/SpherePSEConvTest.f90:196-207: the PSE Laplacian evaluated at arbitrary points of the
 * sphere (the test's 181 x 360 lat-lon grid), f_target = the field's exact value there:
 *   lap(t) = sum_{k active} eta(d(x_k, t)/eps)/eps^2 (f_k - f_target(t)) A_k / eps^2 .
 * Pins bivariateLaplacianKernel8 + SphereDistance at non-particle targets (unifLinfHarmLap). */
void oracle_pse_laplacian_sphere_at_points(int64_t n, const double *x, const double *y, const double *z,
                                           const double *f, const double *area, const int32_t *mask, double eps,
                                           double sphereRadius, int64_t m, const double *tx, const double *ty,
                                           const double *tz, const double *ftarget, double *lap)
{
    for (int64_t t = 0; t < m; ++t) {
        const double xt[3] = {tx[t], ty[t], tz[t]};
        double acc = 0.0;
        for (int64_t k = 0; k < n; ++k) {
            if (mask[k]) {
                const double xk[3] = {x[k], y[k], z[k]};
                double pseKin = SphereDistance(xk, xt, sphereRadius) / eps;
                double lapKernel = bivariateLaplacianKernel8(pseKin) / (eps * eps);
                acc = acc + lapKernel * (f[k] - ftarget[t]) * area[k] / (eps * eps);
            }
        }
        lap[t] = acc;
    }
}
