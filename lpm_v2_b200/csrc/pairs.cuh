// pairs.cuh -- the pair kernels of the hot path, as functors for ds_kernel,
// plus the kernels that pack the active particles into source records.
//
// Each functor cites the reference loop it replaces.  All arithmetic is FP64
// (the reference's real(kreal) = kind(0.d0), src/TypeDefs.f90:25).  Constant
// factors that the reference recomputes per pair (1/(4 pi R), 1/(2 pi), ...)
// are folded into the per-source strength at pack time.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>

namespace lpm {

// src/TypeDefs.f90:31
#define LPM_PI 3.1415926535897932384626433832795027975

constexpr int kMaxRep = 8;      // output replicas (peer GPUs) one finalize can write

// Output arrays of one direct sum, optionally replicated on peer devices: the
// finalize step stores each result to every replica, which is the all-gather
// of the reference's MPI_BCAST loop done with NVLink peer stores.
template <int NO>
struct Outs {
    int nrep;
    double* p[kMaxRep][NO];
    __device__ __forceinline__ void store(int o, int64_t i, double v) const
    {
#pragma unroll 1
        for (int r = 0; r < nrep; ++r) p[r][o][i] = v;
    }
};

// 1/d for d > 0 to <= 1 ulp: MUFU.RCP64H seed (>= 20 bits) and one cubic
// Newton step, r0 (1 + e + e^2) with e = 1 - d r0: 3 DFMA, relative error
// e^3 < 2^-57 before rounding.  (An IEEE divide costs ~10 DFMA-pipe slots.)
__device__ __forceinline__ double rcp_fast(double d)
{
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
    double e = fma(-d, r0, 1.0);
    double t = fma(e, e, e);
    return fma(r0, t, r0);
}


// Reciprocals of T positive numbers behind ONE MUFU per group of (up to) four:
// 1/d_k = (prod_{m != k} d_m) / (d_0 d_1 d_2 d_3), by a product tree.  FP64-pipe cost
// stays 3 ops per reciprocal (12 per four), but the XU instruction -- which costs the
// FP64 pipe dispatch slots (ncu: math-pipe throttle with the pipe 80 % busy) -- is
// issued a quarter as often.  A few ulp, far inside the 1e-12 budget.
template <int T>
__device__ __forceinline__ void rcp_batch(const double (&d)[T], double (&r)[T])
{
    if constexpr (T == 1) {
        r[0] = rcp_fast(d[0]);
    } else if constexpr (T == 2) {
        double q = rcp_fast(d[0] * d[1]);
        r[0] = q * d[1]; r[1] = q * d[0];
    } else if constexpr (T == 3) {
        double p01 = d[0] * d[1];
        double q = rcp_fast(p01 * d[2]);
        double q01 = q * d[2];
        r[2] = q * p01; r[0] = q01 * d[1]; r[1] = q01 * d[0];
    } else if constexpr (T == 4) {
        double p01 = d[0] * d[1], p23 = d[2] * d[3];
        double q = rcp_fast(p01 * p23);
        double q01 = q * p23, q23 = q * p01;
        r[0] = q01 * d[1]; r[1] = q01 * d[0]; r[2] = q23 * d[3]; r[3] = q23 * d[2];
    } else {
        static_assert(T % 4 == 0 || T == 6, "rcp_batch: unsupported group size");
        constexpr int H = (T == 6) ? 3 : 4;
        double dd[H], rr[H];
#pragma unroll
        for (int b = 0; b < T; b += H) {
#pragma unroll
            for (int k = 0; k < H; ++k) dd[k] = d[b + k];
            rcp_batch<H>(dd, rr);
#pragma unroll
            for (int k = 0; k < H; ++k) r[b + k] = rr[k];
        }
    }
}

// Kernels without a per-CTA shared table.
struct NoSharedTable {
    static constexpr int KS = 0;
    __device__ static __forceinline__ void init_shared(double*, int, int) {}
};

// ln(d) for finite normal d > 0: d = 2^e m, bin k = top 7 mantissa bits of m, q_k ~ 1/c_k from
// MUFU.RCP64H (c_k = 1 + (k + 1/2)/128), r = m q_k - 1 with |r| <= 2^-8, then
//   ln d = e ln2 - ln q_k + ln(1 + r)
// with -ln q_k from a 128-entry table in shared memory and a degree-6 polynomial for ln(1 + r):
// 9 FP64-pipe ops, one MUFU and one LDS.64 instead of the ~28 FP64 ops + branches of log().
// (An earlier version kept {1/c_k, ln c_k} pairs in the table; its LDS.128 with a different
// index per lane was bank-conflict bound -- 59 cycles per pair-warp measured against 28 of FP64 work.)
// Absolute error ~2e-16 + 1 ulp of the result.  Anything else (d <= 0, subnormal, inf,
// NaN) takes the library log() so the reference's -inf / NaN behaviour is kept.
constexpr int kLogTabDoubles = 128;
__device__ double g_log_table[kLogTabDoubles];

// The bin-centre reciprocal used by log_tab: q_k = MUFU.RCP64H(c_k), c_k = 1 + (k + 1/2)/128.
// It has ~20 significant bits, so m q_k - 1 is exact to an FMA rounding and |.| <= 2^-8 + 2^-20.
__device__ __forceinline__ double log_bin_rcp(int hi_of_m)
{
    const double c = __hiloint2double((hi_of_m & 0x000fe000) | 0x3ff01000, 0);
    double q;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(c));
    return q;
}
__global__ void log_table_seed_kernel(double* q)      // q[k] for the host to turn into -ln q[k]
{
    const int k = threadIdx.x;
    if (k < kLogTabDoubles) q[k] = log_bin_rcp(0x3ff00000 | (k << 13));
}

struct LogSharedTable {
    static constexpr int KS = kLogTabDoubles;
    __device__ static __forceinline__ void init_shared(double* ks, int tid, int nthreads)
    {
        for (int q = tid; q < kLogTabDoubles; q += nthreads) ks[q] = g_log_table[q];
    }
};

__device__ __noinline__ double log_slow_path(double d) { return log(d); }

// -1/6, 1/5, -1/4, 1/3, -1/2 (ln(1+r) = r - r^2/2 + ... - r^6/6), ln 2, 2^52 + 2^31
__constant__ double kLogC[7] = {-1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5, 0.693147180559945309417232121458,
                                4503601774854144.0};

__device__ __forceinline__ bool log_needs_slow_path(double d)     // d <= 0, subnormal, inf or NaN
{
    return (unsigned)(__double2hiint(d) - 0x00100000) >= 0x7fe00000u;
}

// Branch-free fast path (the caller has excluded the special arguments).
__device__ __forceinline__ double log_tab(double d, const double* tab)
{
    const int hi = __double2hiint(d), lo = __double2loint(d);
    const int e = (hi >> 20) - 1023;
    const int k = (hi >> 13) & 0x7f;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double r = fma(m, log_bin_rcp(hi), -1.0);        // m / c_k - 1
    // Constants come from the constant bank (operands of the DFMAs themselves): as immediates
    // ptxas rebuilt each 64-bit literal with two moves per use, and the kernel was issue-bound
    // (ncu: 35 instructions per pair, issue 68 %, FP64 pipe 55 %).
    const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - kLogC[6];   // (double)e
    double p = fma(r, kLogC[0], kLogC[1]);
    p = fma(r, p, kLogC[2]);
    p = fma(r, p, kLogC[3]);
    p = fma(r, p, kLogC[4]);
    p = fma(r, p, 1.0);
    return fma(r, p, fma(ed, kLogC[5], tab[k]));   // e ln2 - ln q_k + ln(1 + r)
}

// Logs of a thread's T arguments.  ONE branch per group: the fast block is straight-line
// code whose T dependent chains interleave (a branch per argument serialised them -- the
// first version ran at 59 cycles per pair-warp for 28 cycles of FP64 work).
template <int T>
__device__ __forceinline__ void log_group(const double (&d)[T], double (&l)[T], const double* tab)
{
    bool slow = false;
#pragma unroll
    for (int k = 0; k < T; ++k) slow |= log_needs_slow_path(d[k]);
    if (__builtin_expect(slow, 0)) {
#pragma unroll 1
        for (int k = 0; k < T; ++k) l[k] = log_slow_path(d[k]);
    } else {
#pragma unroll
        for (int k = 0; k < T; ++k) l[k] = log_tab(d[k], tab);
    }
}

// Default group(): one source against the thread's T targets, one pair at a time.
#define LPM_DEFAULT_GROUP()                                                                              \
    template <int T, bool CHECK>                                                                         \
    __device__ static __forceinline__ void group(const Params& p, const Tgt (&t)[T], const double (&s)[NS], \
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], \
                                                 const double* ks)                                       \
    {                                                                                                    \
        _Pragma("unroll") for (int k = 0; k < T; ++k) pair<CHECK>(p, t[k], s, acc[k], CHECK && (j == self[k]), ks); \
    }

// =============================================================================
// BVE velocity.  src/SphereBVESolver.f90:396-420 (== src/SphereBVE.f90:497-521):
//   strength = -zeta_j A_j / (4 pi R (R^2 - x_i.x_j));  u_i += (x_i cross x_j) strength
// Factored as u_i = x_i cross a_i,  a_i = sum_j P_j / (R^2 - x_i.x_j),
// P_j = -zeta_j A_j/(4 pi R) x_j : 3 DFMA (denominator) + 3 (reciprocal) +
// 3 (accumulate) + 1 MUFU per pair instead of the ~25 flops + divide as written.
// Source record: x, y, z, Px, Py, Pz.
struct BveVelParams {
    const double *x, *y, *z;
    double R2;
    Outs<3> out;
};
struct BveVelTgt { double x, y, z; };
template <int RG>      // RG: reciprocals sharing one MUFU (1 = none, 2, 4)
struct BveVelT : NoSharedTable {
    static constexpr int NS = 6, NA = 3;
    static constexpr bool SKIP_SELF = true;
    using Params = BveVelParams;
    using Tgt = BveVelTgt;
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        return Tgt{p.x[i], p.y[i], p.z[i]};
    }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params& p, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], const double*)
    {
        double d[T], r[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            d[k] = fma(-t[k].x, s[0], p.R2);
            d[k] = fma(-t[k].y, s[1], d[k]);
            d[k] = fma(-t[k].z, s[2], d[k]);
            if (CHECK) d[k] = (j == self[k]) ? 1.0 : d[k];     // keep the self pair out of the shared product
        }
        if constexpr (RG >= 4 || T < 2) {
            rcp_batch<T>(d, r);
        } else if constexpr (RG == 2 && T % 2 == 0) {
#pragma unroll
            for (int k = 0; k < T; k += 2) {
                double dd[2] = {d[k], d[k + 1]}, rr[2];
                rcp_batch<2>(dd, rr);
                r[k] = rr[0]; r[k + 1] = rr[1];
            }
        } else {
#pragma unroll
            for (int k = 0; k < T; ++k) r[k] = rcp_fast(d[k]);
        }
#pragma unroll
        for (int k = 0; k < T; ++k) {
            if (CHECK) r[k] = (j == self[k]) ? 0.0 : r[k];
            acc[k][0] = fma(r[k], s[3], acc[k][0]);
            acc[k][1] = fma(r[k], s[4], acc[k][1]);
            acc[k][2] = fma(r[k], s[5], acc[k][2]);
        }
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt& t, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, fma(t.y, a[2], -(t.z * a[1])));
        p.out.store(1, i, fma(t.z, a[0], -(t.x * a[2])));
        p.out.store(2, i, fma(t.x, a[1], -(t.y * a[0])));
    }
};
using BveVel = BveVelT<4>;

__global__ void pack_bve_vel(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                             const double* __restrict__ x, const double* __restrict__ y,
                             const double* __restrict__ z, const double* __restrict__ zeta,
                             const double* __restrict__ area, double R, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};     // null source: d = R^2, P = 0
    if (c < nsrc) {
        int32_t j = active[c];
        double s = -zeta[j] * area[j] / (4.0 * LPM_PI * R);
        r[0] = x[j]; r[1] = y[j]; r[2] = z[j];
        r[3] = s * r[0]; r[4] = s * r[1]; r[5] = s * r[2];
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 6);
    o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]); o[2] = make_double2(r[4], r[5]);
}

// =============================================================================
// BVE stream functions.  src/SphereBVE.f90:454-475:
//   g = -log(R^2 - x_i.x_j)/(4 pi);  relStream_i += g zeta_j A_j;  absStream_i += g omega_j A_j
// Source record: x, y, z, -zeta A/(4 pi), -omega A/(4 pi), 0.
struct BveStream : LogSharedTable {
    static constexpr int NS = 6, NA = 2;
    static constexpr bool SKIP_SELF = true;
    struct Params {
        const double *x, *y, *z;
        double R2;
        Outs<2> out;
    };
    struct Tgt { double x, y, z; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        return Tgt{p.x[i], p.y[i], p.z[i]};
    }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params& p, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T],
                                                 const double* ks)
    {
        double d[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            d[k] = fma(-t[k].x, s[0], p.R2);
            d[k] = fma(-t[k].y, s[1], d[k]);
            d[k] = fma(-t[k].z, s[2], d[k]);
            if (CHECK) d[k] = (j == self[k]) ? 1.0 : d[k];      // log(1) = 0 removes the pair
        }
        log_group<T>(d, l, ks);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            acc[k][0] = fma(l[k], s[3], acc[k][0]);
            acc[k][1] = fma(l[k], s[4], acc[k][1]);
        }
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0]);
        p.out.store(1, i, a[1]);
    }
};

__global__ void pack_bve_stream(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                                const double* __restrict__ x, const double* __restrict__ y,
                                const double* __restrict__ z, const double* __restrict__ zeta,
                                const double* __restrict__ omega, const double* __restrict__ area,
                                double R, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};     // null source: log(R^2) * 0
    if (c < nsrc) {
        int32_t j = active[c];
        r[0] = x[j]; r[1] = y[j]; r[2] = z[j];
        r[3] = -zeta[j] * area[j] / (4.0 * LPM_PI);
        r[4] = -omega[j] * area[j] / (4.0 * LPM_PI);
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 6);
    o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]); o[2] = make_double2(r[4], r[5]);
}

// =============================================================================
// Planar Biot-Savart (singular).  src/PlaneIncompressibleSolver.f90:294-307
// (== src/PlanarIncompressible.f90:437-456):
//   strength = omega_j A_j / (2 pi ((x_i-x_j)^2 + (y_i-y_j)^2))
//   u_i -= (y_i-y_j) strength;  v_i += (x_i-x_j) strength
// Source record: x, y, omega A/(2 pi), 0.
struct PlaneVel : NoSharedTable {
    static constexpr int NS = 4, NA = 2;
    static constexpr bool SKIP_SELF = true;
    struct Params {
        const double *x, *y;
        Outs<2> out;
    };
    struct Tgt { double x, y; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i]}; }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], const double*)
    {
        double dx[T], dy[T], r2[T], r[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            dx[k] = t[k].x - s[0]; dy[k] = t[k].y - s[1];
            r2[k] = fma(dx[k], dx[k], dy[k] * dy[k]);
            if (CHECK) r2[k] = (j == self[k]) ? 1.0 : r2[k];
        }
        rcp_batch<T>(r2, r);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double w = r[k] * s[2];
            if (CHECK) w = (j == self[k]) ? 0.0 : w;
            acc[k][0] = fma(-dy[k], w, acc[k][0]);
            acc[k][1] = fma(dx[k], w, acc[k][1]);
        }
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0]);
        p.out.store(1, i, a[1]);
    }
};

// Null source for the planar kernels: far away with zero strength.  r^2 ~ 2e74, so the
// product of four of them in rcp_batch stays finite.
#define LPM_PLANE_FAR 1.0e37

__global__ void pack_plane(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                           const double* __restrict__ x, const double* __restrict__ y,
                           const double* __restrict__ vort, const double* __restrict__ area,
                           double inv_norm, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[4] = {LPM_PLANE_FAR, LPM_PLANE_FAR, 0.0, 0.0};
    if (c < nsrc) {
        int32_t j = active[c];
        r[0] = x[j]; r[1] = y[j];
        r[2] = vort[j] * area[j] * inv_norm;      // 1/(2 pi) velocity, 1/(4 pi) stream fn
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 4);
    o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]);
}

// Planar stream function.  src/PlanarIncompressible.f90:481-497:
//   psi_i += log(sqrt(r^2))/(2 pi) omega_j A_j  ==  log(r^2) omega_j A_j/(4 pi)
struct PlaneStream : LogSharedTable {
    static constexpr int NS = 4, NA = 1;
    static constexpr bool SKIP_SELF = true;
    struct Params {
        const double *x, *y;
        Outs<1> out;
    };
    struct Tgt { double x, y; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i]}; }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T],
                                                 const double* ks)
    {
        double r2[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double dx = t[k].x - s[0], dy = t[k].y - s[1];
            r2[k] = fma(dx, dx, dy * dy);
            if (CHECK) r2[k] = (j == self[k]) ? 1.0 : r2[k];
        }
        log_group<T>(r2, l, ks);
#pragma unroll
        for (int k = 0; k < T; ++k) acc[k][0] = fma(l[k], s[2], acc[k][0]);
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0]);
    }
};

// =============================================================================
// Beta-plane (x-periodic, period 1) Biot-Savart.  src/BetaPlaneSolver.f90:243-258
// (== src/BetaPlane.f90:369-388):
//   strength = 0.5 zeta_j A_j / (cosh(2 pi dy) - cos(2 pi dx))
//   u_i -= sinh(2 pi dy) strength;  v_i += sin(2 pi dx) strength
// With S = sinh(pi dy), C = cosh(pi dy), s = sin(pi dx), c = cos(pi dx):
//   cosh(2 pi dy) - cos(2 pi dx) = 2 (S^2 + s^2)   (no cancellation for near pairs),
//   sinh(2 pi dy) = 2 S C,  sin(2 pi dx) = 2 s c,
// and S, C, s, c follow from per-particle sinh/cosh(pi y), sin/cos(pi x) by
// the addition formulas, so no transcendental is evaluated per pair.
// Source record: sh, ch, sn, cs, zeta A / 2, 0.
struct BetaVel : NoSharedTable {
    static constexpr int NS = 6, NA = 2;
    static constexpr bool SKIP_SELF = true;
    struct Params {
        const double *x, *y;
        Outs<2> out;
    };
    struct Tgt { double sh, ch, sn, cs; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        Tgt t;
        double a = LPM_PI * p.y[i];
        t.sh = sinh(a); t.ch = cosh(a);
        sincospi(p.x[i], &t.sn, &t.cs);
        return t;
    }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], const double*)
    {
        double SC[T], sc[T], den[T], r[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double S = fma(t[k].sh, s[1], -(t[k].ch * s[0]));
            double C = fma(t[k].ch, s[1], -(t[k].sh * s[0]));
            double sn = fma(t[k].sn, s[3], -(t[k].cs * s[2]));
            double cs = fma(t[k].cs, s[3], t[k].sn * s[2]);
            den[k] = fma(S, S, sn * sn);
            if (CHECK) den[k] = (j == self[k]) ? 1.0 : den[k];
            SC[k] = S * C; sc[k] = sn * cs;
        }
        rcp_batch<T>(den, r);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double w = r[k] * s[4];
            if (CHECK) w = (j == self[k]) ? 0.0 : w;
            acc[k][0] = fma(-SC[k], w, acc[k][0]);
            acc[k][1] = fma(sc[k], w, acc[k][1]);
        }
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0]);
        p.out.store(1, i, a[1]);
    }
};

// Beta-plane stream functions.  src/BetaPlane.f90:409-431:
//   g = log(cosh(2 pi dy) - cos(2 pi dx))/(4 pi) = log(2 (S^2 + s^2))/(4 pi)
// Source record: sh, ch, sn, cs, zeta A/(4 pi), omega A/(4 pi).
struct BetaStream : LogSharedTable {
    static constexpr int NS = 6, NA = 2;
    static constexpr bool SKIP_SELF = true;
    using Params = BetaVel::Params;
    using Tgt = BetaVel::Tgt;
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return BetaVel::load_target(p, i); }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T],
                                                 const double* ks)
    {
        double den[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double S = fma(t[k].sh, s[1], -(t[k].ch * s[0]));
            double sn = fma(t[k].sn, s[3], -(t[k].cs * s[2]));
            den[k] = 2.0 * fma(S, S, sn * sn);
            if (CHECK) den[k] = (j == self[k]) ? 1.0 : den[k];
        }
        log_group<T>(den, l, ks);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            acc[k][0] = fma(l[k], s[4], acc[k][0]);
            acc[k][1] = fma(l[k], s[5], acc[k][1]);
        }
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0]);
        p.out.store(1, i, a[1]);
    }
};

// mode 0: velocity record (zeta A/2, 0); mode 1: stream record (zeta A/(4 pi), omega A/(4 pi)).
// Null source: sh = ch = 1e10 gives S = 1e10 (sh_i - ch_i) /= 0 for every target; zero strength.
__global__ void pack_beta(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                          const double* __restrict__ x, const double* __restrict__ y,
                          const double* __restrict__ zeta, const double* __restrict__ omega,
                          const double* __restrict__ area, int mode, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[6] = {1.0e10, 1.0e10, 0.0, 1.0, 0.0, 0.0};
    if (c < nsrc) {
        int32_t j = active[c];
        double a = LPM_PI * y[j];
        r[0] = sinh(a); r[1] = cosh(a);
        sincospi(x[j], &r[2], &r[3]);
        if (mode == 0) {
            r[4] = 0.5 * zeta[j] * area[j];
        } else {
            r[4] = zeta[j] * area[j] / (4.0 * LPM_PI);
            r[5] = omega[j] * area[j] / (4.0 * LPM_PI);
        }
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 6);
    o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]); o[2] = make_double2(r[4], r[5]);
}

// =============================================================================
// PSE Laplacian.  src/PSEDirectSum.f90:467-535, kernel :622-627:
//   lap_i = eps^-2 sum_{j active} (f_j - f_i) eta(d_ij/eps)/eps^2 A_j,
//   eta(k) = (40 - 40 k^2 + 10 k^4 - 2/3 k^6) exp(-k^2)/pi
// j == i is included (the term is exactly zero).  eta decays like exp(-k^2):
// pairs with k > kPseCut contribute < 1e-23 of eta(0) and are skipped.
constexpr double kPseCut = 8.0;

__device__ __forceinline__ double pse_eta_pi(double k2)    // pi * eta
{
    double poly = fma(fma(fma(-2.0 / 3.0, k2, 10.0), k2, -40.0), k2, 40.0);
    return poly * exp(-k2);
}

// Sphere: d_ij = atan2(|x_i cross x_j|, x_i.x_j) * SphereRadius  (src/SphereGeometry.f90:107-125).
// Source record: x, y, z, f, A/(pi eps^2), |x_j|.
struct PseSphere : NoSharedTable {
    static constexpr int NS = 6, NA = 1;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *z, *f;
        double rad_over_eps;     // SphereRadius / eps
        double cos_cut;          // cos(kPseCut eps / SphereRadius), or -2 if the cut-off exceeds pi
        double inv_eps2;
        Outs<1> out;
    };
    struct Tgt { double x, y, z, f, thr; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        Tgt t{p.x[i], p.y[i], p.z[i], p.f[i], 0.0};
        t.thr = p.cos_cut * sqrt(t.x * t.x + t.y * t.y + t.z * t.z);
        return t;
    }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const double*)
    {
        double dot = fma(t.x, s[0], fma(t.y, s[1], t.z * s[2]));
        if (dot < t.thr * s[5]) return;           // angle beyond the cut-off
        double c0 = fma(t.y, s[2], -(s[1] * t.z));
        double c1 = fma(s[0], t.z, -(t.x * s[2]));
        double c2 = fma(t.x, s[1], -(s[0] * t.y));
        double cn = sqrt(fma(c0, c0, fma(c1, c1, c2 * c2)));
        double k = atan2(cn, dot) * p.rad_over_eps;
        acc[0] = fma((s[3] - t.f) * pse_eta_pi(k * k), s[4], acc[0]);
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0] * p.inv_eps2);
    }
};

__global__ void pack_pse_sphere(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                                const double* __restrict__ x, const double* __restrict__ y,
                                const double* __restrict__ z, const double* __restrict__ f,
                                const double* __restrict__ area, double eps, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};     // null source: zero area
    if (c < nsrc) {
        int32_t j = active[c];
        r[0] = x[j]; r[1] = y[j]; r[2] = z[j]; r[3] = f[j];
        r[4] = area[j] / (LPM_PI * eps * eps);
        r[5] = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 6);
    o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]); o[2] = make_double2(r[4], r[5]);
}

// Plane: d_ij = ChordDistance with z = 0 (src/SphereGeometry.f90:67-73, src/Particles.f90:663-670).
// Source record: x, y, f, A/(pi eps^2).
struct PsePlane : NoSharedTable {
    static constexpr int NS = 4, NA = 1;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *f;
        double inv_eps2;
        Outs<1> out;
    };
    struct Tgt { double x, y, f; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i], p.f[i]}; }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const double*)
    {
        double dx = s[0] - t.x, dy = s[1] - t.y;
        double k2 = fma(dx, dx, dy * dy) * p.inv_eps2;
        if (k2 > kPseCut * kPseCut) return;
        acc[0] = fma((s[2] - t.f) * pse_eta_pi(k2), s[3], acc[0]);
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0] * p.inv_eps2);
    }
};

__global__ void pack_pse_plane(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                               const double* __restrict__ x, const double* __restrict__ y,
                               const double* __restrict__ f, const double* __restrict__ area,
                               double eps, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[4] = {0.0, 0.0, 0.0, 0.0};
    if (c < nsrc) {
        int32_t j = active[c];
        r[0] = x[j]; r[1] = y[j]; r[2] = f[j];
        r[3] = area[j] / (LPM_PI * eps * eps);
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 4);
    o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]);
}

}  // namespace lpm
