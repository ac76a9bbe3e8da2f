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

#include "directsum.cuh"

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
    const double r0 = rcp_approx_f64(d);
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

// What the functors call.  The unchecked tile loop (CHECK = false) shares one MUFU between the targets of a thread;
// the careful loop (CHECK = true: tiles that contain self pairs, and tiles ds_kernel re-runs because a sum came out
// non-finite) takes each reciprocal on its own, so that a zero or non-finite denominator -- coincident particles --
// reaches only its own target, as in the reference's loops (e.g. src/SphereBVESolver.f90:403-407).
template <int T, bool CHECK>
__device__ __forceinline__ void rcp_group(const double (&d)[T], double (&r)[T])
{
    if constexpr (CHECK) {
#pragma unroll
        for (int k = 0; k < T; ++k) r[k] = rcp_fast(d[k]);
    } else {
        rcp_batch<T>(d, r);
    }
}

// Tile culling (directsum.cuh) is for the compactly supported PSE kernels only; they
// derive from CullSphere / CullPlane below and provide cull_dist(Params).
struct NoCull {
    static constexpr bool CULL = false;
    static constexpr int CULL_GEOM = 0;
    static constexpr bool RETRY = false;     // no group_fast(); see LogSharedTable
    static constexpr bool BATCHED_RCP = false;   // true: ds_kernel re-runs a tile whose sums are non-finite (rcp_group)
};

// Kernels without a per-CTA shared table.
struct NoSharedTable : NoCull {
    static constexpr int KS = 0;
    template <class Params>
    __device__ static __forceinline__ int32_t init_shared(double*, const Params&, int, int) { return 0; }
};

// ln(d) for d > 0 by table: the high word of d -- exponent and the top kLogBits mantissa bits --
// names a bin [c - w, c + w); Q = MUFU.RCP64H(c) is a ~20-bit reciprocal of the bin centre that is
// EXACT as a double (high word only), so r = fma(d, Q, -1) is exact to one rounding, |r| <= 2^-9 + 2^-20, and
//   ln d = -ln Q + ln(1 + r),
// with -ln Q from a table indexed by that same high word and a degree-5 polynomial for ln(1 + r)
// (truncation r^6/6 < 1e-17).  No exponent extraction, no int-to-double conversion: 6 FP64-pipe
// instructions, one MUFU and one LDS.64 per logarithm (the library log() is ~28 FP64 instructions
// plus branches; a first table version with a 128-entry mantissa table and e ln2 added separately
// took 9 and ran issue-bound).  Absolute error <= 1.5 ulp of max(|ln d|, 1/2) (tests/test_kernel_math.py).
//
// The full table (every binade, 4 MB, built once per device from the device's own MUFU
// results and long-double logarithms on the host, runtime.cuh) lives in global memory; each
// launch copies the WINDOW of binades its arguments can fall in to shared memory.  The window
// origin is computed on the device from a bound on d (log_window_kernel).  An argument outside
// the window, or d <= 0 / subnormal / inf / NaN, takes the library log(), which also keeps the
// reference's -inf / NaN behaviour.
constexpr int kLogBits = 8;
constexpr int kLogBin = 1 << kLogBits;                 // table entries per binade
constexpr int kLogFull = 2048 * kLogBin;               // entries of the full table
constexpr int kLogBinadeMin = 2, kLogBinadeMax = 2044; // biased exponents the fast path may see
__device__ double g_log_full[kLogFull];

// Q for the bin that holds a double with high word `hi`
__device__ __forceinline__ double log_bin_rcp(int hi, int half = 0x00000800)
{
    // (hi & 0xfffff000) | half as ONE LOP3: `half` arrives in a register (SharedCtx), because
    // ptxas splits the expression in two when both masks are immediates
    const double c = __hiloint2double(lop3_and_or(hi, half), 0);
    return rcp_approx_f64(c);
}
__global__ void log_table_seed_kernel(double* q)      // q[idx] for the host to turn into -ln q[idx]
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < kLogFull) q[idx] = log_bin_rcp(idx << (20 - kLogBits));
}

// Parameters every log kernel carries (filled by direct_sum from the Device).
struct LogParams {
    const double* logtab;     // g_log_full of this device
    const int32_t* win;       // window origin (entry index) written by log_window_kernel
};

template <int WB>     // WB: binades in the shared-memory window
struct LogSharedTable : NoCull {
    static constexpr int KS = WB * kLogBin;
    static constexpr int WINDOW_BINADES = WB;
    // The unchecked tile loop calls group_fast(), which has no slow path (so no call, no
    // branch): it records the worst table offset it saw, and ds_kernel re-runs the tile
    // through group<T, true>() -- which has the library-log() path -- if any was outside the window.
    static constexpr bool RETRY = true;
    static_assert((KS & (KS - 1)) == 0, "the fast path wraps table offsets with a mask");
    __device__ static __forceinline__ bool needs_retry(unsigned worst) { return worst >= (unsigned)(KS * 8); }
    template <class Params>
    __device__ static __forceinline__ int32_t init_shared(double* ks, const Params& p, int tid, int nthreads)
    {
        const int32_t i0 = *p.win;
        const double2* g = reinterpret_cast<const double2*>(p.logtab + i0);     // i0 is a multiple of 256
        double2* s2 = reinterpret_cast<double2*>(ks);
        for (int q = tid; q < KS / 2; q += nthreads) s2[q] = g[q];
        return i0;
    }
};

// Window origin from an upper bound on the arguments.  mode 0: dmax = bound (BVE: 2 R^2);
// mode 1: plane, bound = high word of max(|x|, |y|) over all particles, r^2 <= 8 m^2;
// mode 2: beta-plane, bound = high word of max |y|, den <= 2 (4 cosh(pi y)^4 + 1).
__global__ void log_window_kernel(int mode, double bound, const int32_t* __restrict__ maxhi, int wb, int32_t* win)
{
    double dmax = bound;
    if (mode != 0) {
        const double m = __hiloint2double(*maxhi, (int)0xffffffff);     // >= every |coordinate|
        if (mode == 1) {
            dmax = 8.0 * m * m;
        } else {
            const double c = cosh(LPM_PI * m);
            dmax = 2.0 * (4.0 * c * c * c * c + 1.0);
        }
    }
    dmax *= 1.000001;
    int top = kLogBinadeMax;
    if (dmax > 0.0 && dmax < CUDART_INF) top = __double2hiint(dmax) >> 20;
    int lo = top - wb + 1;
    if (lo > kLogBinadeMax - wb + 1) lo = kLogBinadeMax - wb + 1;
    if (lo < kLogBinadeMin) lo = kLogBinadeMin;
    *win = lo << kLogBits;
}

// max over i of the high word of |a[i]| (and |b[i]|): non-negative doubles order like their high words
__global__ void __launch_bounds__(256)
absmax_hi_kernel(int64_t n, const double* __restrict__ a, const double* __restrict__ b, int32_t* out)
{
    int32_t m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        m = max(m, __double2hiint(a[i]) & 0x7fffffff);
        if (b) m = max(m, __double2hiint(b[i]) & 0x7fffffff);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__device__ __noinline__ double log_slow_path(double d) { return log(d); }

// 1/5, -1/4, 1/3, -1/2 : ln(1 + r) = r (1 + r (-1/2 + r (1/3 + r (-1/4 + r/5))))
__constant__ double kLogC[4] = {0.2, -0.25, 1.0 / 3.0, -0.5};

// Fast path; idx = (high word >> 12) - window origin, already known to be inside the window.
// Constants come from the constant bank (operands of the DFMAs themselves): as immediates
// ptxas rebuilt each 64-bit literal with two moves per use.
__device__ __forceinline__ double log_tab(double d, int idx, const SharedCtx& sc)
{
    const double* tab = sc.ks;
    const double r = fma(d, log_bin_rcp(__double2hiint(d), sc.i1), -1.0);
    double p = fma(r, kLogC[0], kLogC[1]);
    p = fma(r, p, kLogC[2]);
    p = fma(r, p, kLogC[3]);
    p = fma(r, p, 1.0);
    return fma(r, p, tab[idx]);
}

// Logs of a thread's T arguments.  ONE branch per group: the fast block is straight-line
// code whose T dependent chains interleave (a branch per argument serialised them).  A
// negative, zero, NaN or out-of-window argument has an index outside [0, KS) as unsigned.
template <int KS, int T>
__device__ __forceinline__ void log_group(const double (&d)[T], double (&l)[T], const SharedCtx& sc)
{
    int idx[T];
    unsigned worst = 0;
#pragma unroll
    for (int k = 0; k < T; ++k) {
        idx[k] = (__double2hiint(d[k]) >> (20 - kLogBits)) - sc.i0;
        worst = max(worst, (unsigned)idx[k]);
    }
    if (__builtin_expect(worst >= (unsigned)KS, 0)) {
#pragma unroll
        for (int k = 0; k < T; ++k) l[k] = log_slow_path(d[k]);
    } else {
#pragma unroll
        for (int k = 0; k < T; ++k) l[k] = log_tab(d[k], idx[k], sc);
    }
}

// The same without the branch: `worst` collects the largest byte offset into the window
// (as unsigned: negative, zero, NaN and out-of-window arguments all give >= 8 KS); offsets
// are wrapped into the table so that the load is always legal, and the caller discards the
// tile's sums when needs_retry(worst).
template <int KS, int T>
__device__ __forceinline__ void log_group_fast(const double (&d)[T], double (&l)[T], unsigned& worst,
                                               const SharedCtx& sc)
{
    // Written "vertically" (one Horner step of all T chains, then the next) because ptxas keeps
    // close to source order: a chain-by-chain version came out with back-to-back dependent
    // DFMAs and stalled on their latency (ncu: `wait` the top stall, FP64 pipe 62 %).
    double r[T], p[T], tk[T];
#pragma unroll
    for (int k = 0; k < T; ++k) {
        const int hi = __double2hiint(d[k]);
        const unsigned u = (unsigned)((hi >> (17 - kLogBits)) - sc.i0 * 8);
        worst = max(worst, u);
        tk[k] = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(sc.ks) + (u & (KS * 8 - 8)));
        r[k] = fma(d[k], log_bin_rcp(hi, sc.i1), -1.0);
    }
    // -1/4 and -1/2 are written as literals: their low words are zero, so they fit the DFMA's
    // 32-bit immediate and the first Horner step reads r and one constant register instead of two
    // (the kernel is bound by register operand delivery, see BveVelT).
#pragma unroll
    for (int k = 0; k < T; ++k) p[k] = fma(r[k], kLogC[0], -0.25);
#pragma unroll
    for (int k = 0; k < T; ++k) p[k] = fma(r[k], p[k], kLogC[2]);
#pragma unroll
    for (int k = 0; k < T; ++k) p[k] = fma(r[k], p[k], -0.5);
#pragma unroll
    for (int k = 0; k < T; ++k) p[k] = fma(r[k], p[k], 1.0);
#pragma unroll
    for (int k = 0; k < T; ++k) l[k] = fma(r[k], p[k], tk[k]);
}

// Default group(): one source against the thread's T targets, one pair at a time.
#define LPM_DEFAULT_GROUP()                                                                              \
    template <int T, bool CHECK>                                                                         \
    __device__ static __forceinline__ void group(const Params& p, const Tgt (&t)[T], const double (&s)[NS], \
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], \
                                                 const SharedCtx& ks)                                    \
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
// Compile-time permutation of {0..3} number `idx` (0..23), factoradic.
struct Perm4 { int v[4]; };
__host__ __device__ constexpr Perm4 perm4(int idx)
{
    int pool[4] = {0, 1, 2, 3};
    Perm4 p{};
    int n = 4, div = 6;
    for (int k = 0; k < 4; ++k) {
        const int q = idx / div;
        idx %= div;
        p.v[k] = pool[q];
        for (int m = q; m + 1 < n; ++m) pool[m] = pool[m + 1];
        --n;
        if (n > 1) div /= (n);
        else div = 1;
    }
    return p;
}

// RG: reciprocals sharing one MUFU (1 = none, 2, 4).
//
// ORDER permutes INDEPENDENT statements of group() -- the arithmetic per pair is the same --
// and selects how the four denominators are paired in the reciprocal's product tree.  It exists
// because the FP64 pipe is limited by register-bank conflicts, not by the instruction count:
// over 23 builds with identical instruction mixes, time = (FP64-bound time) + 0.5 cycle per DFMA
// with two fresh operands in one bank + 1 cycle with three (R^2 = 0.62, intercept = the roofline;
// tools/sass_banks.py, profiles/README.md), and ptxas' allocation changes with statement order.
// tools/search_order.py scores orders with that model; the sweep on the GPU picks among the best.
//   bits 0-4   target order of the denominator phase (perm4 index)
//   bits 5-9   target order of the accumulation phase
//   bit  10    denominator phase coordinate-major
//   bits 11-12 accumulation phase: 0 target-major, 1 component-major, 2 target-major with components reversed
//   bits 13-14 product-tree pairing: (0 1)(2 3), (0 2)(1 3), (0 3)(1 2)        [T = 4 only]
template <int RG, int ORDER = 0>
struct BveVelT : NoSharedTable {
    static constexpr int NS = 6, NA = 3;
    static constexpr bool SKIP_SELF = true;
    static constexpr bool BATCHED_RCP = RG >= 2;
    using Params = BveVelParams;
    using Tgt = BveVelTgt;
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        return Tgt{p.x[i], p.y[i], p.z[i]};
    }
    // k-th target of a phase: permuted within each aligned group of four
    __host__ __device__ static constexpr int pick(int pidx, int k) { return (k & ~3) | perm4(pidx).v[k & 3]; }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params& p, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], const SharedCtx&)
    {
        constexpr int PD = (T % 4 == 0) ? (ORDER & 31) % 24 : 0, PA = (T % 4 == 0) ? ((ORDER >> 5) & 31) % 24 : 0;
        constexpr int DN = (ORDER >> 10) & 1, AN = (ORDER >> 11) & 3, PT = ((ORDER >> 13) & 3) % 3;
        double d[T], r[T];
        if constexpr (DN == 0) {
#pragma unroll
            for (int q = 0; q < T; ++q) {
                const int k = pick(PD, q);
                d[k] = fma(-t[k].x, s[0], p.R2);
                d[k] = fma(-t[k].y, s[1], d[k]);
                d[k] = fma(-t[k].z, s[2], d[k]);
            }
        } else {        // coordinate by coordinate
#pragma unroll
            for (int q = 0; q < T; ++q) { const int k = pick(PD, q); d[k] = fma(-t[k].x, s[0], p.R2); }
#pragma unroll
            for (int q = 0; q < T; ++q) { const int k = pick(PD, q); d[k] = fma(-t[k].y, s[1], d[k]); }
#pragma unroll
            for (int q = 0; q < T; ++q) { const int k = pick(PD, q); d[k] = fma(-t[k].z, s[2], d[k]); }
        }
        if (CHECK) {
#pragma unroll
            for (int k = 0; k < T; ++k) d[k] = (j == self[k]) ? 1.0 : d[k];     // keep the self pair out of the shared product
        }
        if constexpr (CHECK) {
            rcp_group<T, true>(d, r);
        } else if constexpr (RG >= 4 && T == 4) {
            // rcp_batch<4> with the pairing chosen by PT
            constexpr int a0 = 0, a1 = (PT == 0) ? 1 : (PT == 1) ? 2 : 3;
            constexpr int b0 = (PT == 0) ? 2 : 1, b1 = (PT == 2) ? 2 : 3;
            const double pa = d[a0] * d[a1], pb = d[b0] * d[b1];
            const double q = rcp_fast(pa * pb);
            const double qa = q * pb, qb = q * pa;
            r[a0] = qa * d[a1]; r[a1] = qa * d[a0]; r[b0] = qb * d[b1]; r[b1] = qb * d[b0];
        } else if constexpr (RG >= 4 || T < 2) {
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
        if (CHECK) {
#pragma unroll
            for (int k = 0; k < T; ++k) r[k] = (j == self[k]) ? 0.0 : r[k];
        }
        if constexpr (AN == 1) {      // component by component
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int q = 0; q < T; ++q) { const int k = pick(PA, q); acc[k][a] = fma(r[k], s[3 + a], acc[k][a]); }
        } else if constexpr (AN == 2) {
#pragma unroll
            for (int q = 0; q < T; ++q) {
                const int k = pick(PA, q);
                acc[k][2] = fma(r[k], s[5], acc[k][2]);
                acc[k][1] = fma(r[k], s[4], acc[k][1]);
                acc[k][0] = fma(r[k], s[3], acc[k][0]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < T; ++q) {
                const int k = pick(PA, q);
                acc[k][0] = fma(r[k], s[3], acc[k][0]);
                acc[k][1] = fma(r[k], s[4], acc[k][1]);
                acc[k][2] = fma(r[k], s[5], acc[k][2]);
            }
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
struct BveStream : LogSharedTable<32> {
    static constexpr int NS = 6, NA = 2;
    static constexpr bool SKIP_SELF = true;
    struct Params : LogParams {
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
                                                 const SharedCtx& sc)
    {
        double d[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            d[k] = fma(-t[k].x, s[0], p.R2);
            d[k] = fma(-t[k].y, s[1], d[k]);
            d[k] = fma(-t[k].z, s[2], d[k]);
            if (CHECK) d[k] = (j == self[k]) ? p.R2 : d[k];     // any in-window value; the pair is zeroed below
        }
        log_group<KS, T>(d, l, sc);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            if (CHECK) l[k] = (j == self[k]) ? 0.0 : l[k];      // exactly, whatever the table gives for log(1)
            acc[k][0] = fma(l[k], s[3], acc[k][0]);
            acc[k][1] = fma(l[k], s[4], acc[k][1]);
        }
    }
    template <int T>
    __device__ static __forceinline__ void group_fast(const Params& p, const Tgt (&t)[T], const double (&s)[NS],
                                                      double (&acc)[T][NA], unsigned& worst, const SharedCtx& sc)
    {
        double d[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            d[k] = fma(-t[k].x, s[0], p.R2);
            d[k] = fma(-t[k].y, s[1], d[k]);
            d[k] = fma(-t[k].z, s[2], d[k]);
        }
        log_group_fast<KS, T>(d, l, worst, sc);
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
// BVE velocity AND stream functions in one pass: the end of every RK4 step (src/SphereBVESolver.f90:345-352: the
// velocity at the new state, then SetStreamFunctionsOnMesh on the same particles).  Both sums depend on the pair
// through d = R^2 - x_i.x_j only, so the denominator (3 DFMA) is computed once: 3 + 3 (reciprocal) + 3 + 6 (logarithm)
// + 2 = 17 FP64 instructions per interaction instead of 9 + 11 = 20 for the two kernels.
// Source record: x, y, z, Px, Py, Pz (P = -zeta A/(4 pi R) x), -zeta A/(4 pi), -omega A/(4 pi).
struct BveVelStream : LogSharedTable<32> {
    static constexpr int NS = 8, NA = 5;
    static constexpr bool SKIP_SELF = true;
    static constexpr bool BATCHED_RCP = true;
    struct Params : LogParams {
        const double *x, *y, *z;
        double R2;
        Outs<5> out;
    };
    struct Tgt { double x, y, z; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        return Tgt{p.x[i], p.y[i], p.z[i]};
    }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params& p, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T],
                                                 const SharedCtx& sc)
    {
        double d[T], r[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            d[k] = fma(-t[k].x, s[0], p.R2);
            d[k] = fma(-t[k].y, s[1], d[k]);
            d[k] = fma(-t[k].z, s[2], d[k]);
            if (CHECK) d[k] = (j == self[k]) ? p.R2 : d[k];     // any in-window value; the pair is zeroed below
        }
        rcp_group<T, CHECK>(d, r);
        log_group<KS, T>(d, l, sc);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            if (CHECK) {
                r[k] = (j == self[k]) ? 0.0 : r[k];
                l[k] = (j == self[k]) ? 0.0 : l[k];
            }
            acc[k][0] = fma(r[k], s[3], acc[k][0]);
            acc[k][1] = fma(r[k], s[4], acc[k][1]);
            acc[k][2] = fma(r[k], s[5], acc[k][2]);
            acc[k][3] = fma(l[k], s[6], acc[k][3]);
            acc[k][4] = fma(l[k], s[7], acc[k][4]);
        }
    }
    template <int T>
    __device__ static __forceinline__ void group_fast(const Params& p, const Tgt (&t)[T], const double (&s)[NS],
                                                      double (&acc)[T][NA], unsigned& worst, const SharedCtx& sc)
    {
        double d[T], r[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) d[k] = fma(-t[k].x, s[0], p.R2);
#pragma unroll
        for (int k = 0; k < T; ++k) d[k] = fma(-t[k].y, s[1], d[k]);
#pragma unroll
        for (int k = 0; k < T; ++k) d[k] = fma(-t[k].z, s[2], d[k]);
        rcp_batch<T>(d, r);
        log_group_fast<KS, T>(d, l, worst, sc);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            acc[k][0] = fma(r[k], s[3], acc[k][0]);
            acc[k][1] = fma(r[k], s[4], acc[k][1]);
            acc[k][2] = fma(r[k], s[5], acc[k][2]);
            acc[k][3] = fma(l[k], s[6], acc[k][3]);
            acc[k][4] = fma(l[k], s[7], acc[k][4]);
        }
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt& t, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, fma(t.y, a[2], -(t.z * a[1])));
        p.out.store(1, i, fma(t.z, a[0], -(t.x * a[2])));
        p.out.store(2, i, fma(t.x, a[1], -(t.y * a[0])));
        p.out.store(3, i, a[3]);
        p.out.store(4, i, a[4]);
    }
};

__global__ void pack_bve_velstream(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                                   const double* __restrict__ x, const double* __restrict__ y,
                                   const double* __restrict__ z, const double* __restrict__ zeta,
                                   const double* __restrict__ omega, const double* __restrict__ area,
                                   double R, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};     // null source: d = R^2, all weights 0
    if (c < nsrc) {
        int32_t j = active[c];
        // the same expressions as pack_bve_vel / pack_bve_stream, so the fused sums see the same records
        double s = -zeta[j] * area[j] / (4.0 * LPM_PI * R);
        r[0] = x[j]; r[1] = y[j]; r[2] = z[j];
        r[3] = s * r[0]; r[4] = s * r[1]; r[5] = s * r[2];
        r[6] = -zeta[j] * area[j] / (4.0 * LPM_PI);
        r[7] = -omega[j] * area[j] / (4.0 * LPM_PI);
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 8);
    o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]);
    o[2] = make_double2(r[4], r[5]); o[3] = make_double2(r[6], r[7]);
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
    static constexpr bool BATCHED_RCP = true;
    struct Params {
        const double *x, *y;
        Outs<2> out;
    };
    struct Tgt { double x, y; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i]}; }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], const SharedCtx&)
    {
        double dx[T], dy[T], r2[T], r[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            dx[k] = t[k].x - s[0]; dy[k] = t[k].y - s[1];
            r2[k] = fma(dx[k], dx[k], dy[k] * dy[k]);
            if (CHECK) r2[k] = (j == self[k]) ? 1.0 : r2[k];
        }
        rcp_group<T, CHECK>(r2, r);
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
struct PlaneStream : LogSharedTable<32> {
    static constexpr int NS = 4, NA = 1;
    static constexpr bool SKIP_SELF = true;
    struct Params : LogParams {
        const double *x, *y;
        Outs<1> out;
    };
    struct Tgt { double x, y; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i]}; }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T],
                                                 const SharedCtx& sc)
    {
        double r2[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double dx = t[k].x - s[0], dy = t[k].y - s[1];
            r2[k] = fma(dx, dx, dy * dy);
            if (CHECK) r2[k] = (j == self[k]) ? 1.0 : r2[k];
        }
        log_group<KS, T>(r2, l, sc);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            if (CHECK) l[k] = (j == self[k]) ? 0.0 : l[k];
            acc[k][0] = fma(l[k], s[2], acc[k][0]);
        }
    }
    template <int T>
    __device__ static __forceinline__ void group_fast(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                      double (&acc)[T][NA], unsigned& worst, const SharedCtx& sc)
    {
        double r2[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double dx = t[k].x - s[0], dy = t[k].y - s[1];
            r2[k] = fma(dx, dx, dy * dy);
        }
        log_group_fast<KS, T>(r2, l, worst, sc);
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
    static constexpr bool BATCHED_RCP = true;
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
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], const SharedCtx&)
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
        rcp_group<T, CHECK>(den, r);
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
struct BetaStream : LogSharedTable<32> {
    static constexpr int NS = 6, NA = 2;
    static constexpr bool SKIP_SELF = true;
    struct Params : LogParams {
        const double *x, *y;
        Outs<2> out;
    };
    using Tgt = BetaVel::Tgt;
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        BetaVel::Params q{};
        q.x = p.x; q.y = p.y;
        return BetaVel::load_target(q, i);
    }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T],
                                                 const SharedCtx& sc)
    {
        double den[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double S = fma(t[k].sh, s[1], -(t[k].ch * s[0]));
            double sn = fma(t[k].sn, s[3], -(t[k].cs * s[2]));
            den[k] = 2.0 * fma(S, S, sn * sn);
            if (CHECK) den[k] = (j == self[k]) ? 1.0 : den[k];
        }
        log_group<KS, T>(den, l, sc);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            if (CHECK) l[k] = (j == self[k]) ? 0.0 : l[k];
            acc[k][0] = fma(l[k], s[4], acc[k][0]);
            acc[k][1] = fma(l[k], s[5], acc[k][1]);
        }
    }
    template <int T>
    __device__ static __forceinline__ void group_fast(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                      double (&acc)[T][NA], unsigned& worst, const SharedCtx& sc)
    {
        double den[T], l[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double S = fma(t[k].sh, s[1], -(t[k].ch * s[0]));
            double sn = fma(t[k].sn, s[3], -(t[k].cs * s[2]));
            den[k] = 2.0 * fma(S, S, sn * sn);
        }
        log_group_fast<KS, T>(den, l, worst, sc);
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
constexpr double kNullNorm = 1.0e300;     // |x| of a padding source in the sphere PSE records

// Culling geometry.  Sphere kernels reject a pair by its ANGLE (dot < cos_cut |x_i| |x_j|), so
// points are compared as unit vectors: an accepted pair has |u_i - u_j| <= 2 sin(theta_cut / 2)
// whatever the particles' norms, and chords of unit vectors obey the triangle inequality.
// Plane kernels reject by Euclidean distance > kPseCut eps.  Record fields 0..DIM-1 are the position.
struct CullSphere : NoSharedTable {
    static constexpr bool CULL = true;
    static constexpr int CULL_GEOM = 3;
    template <class Tgt>
    __device__ static __forceinline__ void tgt_point(const Tgt& t, double (&p)[3])
    {
        const double inv = 1.0 / sqrt(t.x * t.x + t.y * t.y + t.z * t.z);
        p[0] = t.x * inv; p[1] = t.y * inv; p[2] = t.z * inv;
    }
};
struct CullPlane : NoSharedTable {
    static constexpr bool CULL = true;
    static constexpr int CULL_GEOM = 2;
    template <class Tgt>
    __device__ static __forceinline__ void tgt_point(const Tgt& t, double (&p)[3])
    {
        p[0] = t.x; p[1] = t.y; p[2] = 0.0;
    }
};
// 2 sin(theta_cut / 2), or 1e301 when the cut-off reaches the antipode (no culling)
inline double sphere_chord_cut(double eps, double sphere_radius)
{
    const double theta_cut = kPseCut * eps / sphere_radius;
    return (theta_cut < LPM_PI) ? 2.0 * sin(0.5 * theta_cut) : 1.0e301;      // >= 1e300: no culling
}

// One bounding ball (centre, radius) per source tile of 256 (= kTile) records; an all-padding tile
// gets radius -1e300 so that it is never visited.
template <int NS, int GEOM>
__global__ void __launch_bounds__(256)
tile_bounds_kernel(int32_t nsrc, const double* __restrict__ src, double* __restrict__ bounds)
{
    __shared__ double red[4][8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t c = blockIdx.x * 256 + tid;
    const bool real = c < nsrc;
    double p[3] = {0.0, 0.0, 0.0};
    if (real) {
        p[0] = src[(size_t)c * NS]; p[1] = src[(size_t)c * NS + 1];
        if (GEOM == 3) {
            p[2] = src[(size_t)c * NS + 2];
            const double inv = 1.0 / sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
            p[0] *= inv; p[1] *= inv; p[2] *= inv;
        }
    }
    double s[4] = {p[0], p[1], p[2], real ? 1.0 : 0.0};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < 4; ++q) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < 4; ++q) red[q][wid] = s[q];
    __syncthreads();
    double c4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int w = 0; w < 8; ++w)
#pragma unroll
        for (int q = 0; q < 4; ++q) c4[q] += red[q][w];
    const double cnt = c4[3];
    const double inv = cnt > 0.0 ? 1.0 / cnt : 0.0;
    const double cx = c4[0] * inv, cy = c4[1] * inv, cz = c4[2] * inv;
    double r2 = 0.0;
    if (real) {
        const double dx = p[0] - cx, dy = p[1] - cy, dz = p[2] - cz;
        r2 = dx * dx + dy * dy + dz * dz;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    __syncthreads();
    if (lane == 0) red[0][wid] = r2;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) m = fmax(m, red[0][w]);
        double* b = bounds + (size_t)blockIdx.x * 4;
        b[0] = cx; b[1] = cy; b[2] = cz;
        b[3] = cnt > 0.0 ? sqrt(m) * (1.0 + 1.0e-12) : -1.0e300;
    }
}

// exp(-t) for the PSE kernels, 0 <= t <= 708 (they call it for t = k^2 <= kPseCut^2 = 64).  The library exp() is ~28
// FP64-pipe instructions with its range checks; here m = rint(64 t / ln 2) comes from a magic-number add,
// r = t - m ln2/64 (two steps: m < 2^16 times a 37-bit constant is exact), exp(-r) for |r| <= ln2/128 is a degree-5
// polynomial (truncation r^6/720 < 4e-17), 2^(-(m mod 64)/64) comes from a 64-entry table filled at init
// (g_exp2_tab, read through the read-only path), and 2^(-(m div 64)) is subtracted from the exponent field:
// 10 FP64 instructions + a load + 5 integer instructions, <= 2.5 ulp (tests/test_kernel_math.py).
__device__ double g_exp2_tab[64];
__device__ __forceinline__ double pse_exp_neg(double t)
{
    constexpr double kMagic = 6755399441055744.0;                      // 1.5 * 2^52
    const double u = fma(t, 92.33248261689366 /* 64 / ln 2 */, kMagic);
    const int m = __double2loint(u);
    const double mf = u - kMagic;
    double r = fma(mf, -0x1.62e42fefa0000p-7, t);                      // ln2/64, high 37 bits
    r = fma(mf, -0x1.cf79abc9e3b3ap-46, r);                            // ... and the rest
    double p = fma(r, -1.0 / 120.0, 1.0 / 24.0);
    p = fma(r, p, -1.0 / 6.0);
    p = fma(r, p, 0.5);
    p = fma(r, p, -1.0);
    p = fma(r, p, 1.0);
#ifdef LPM_CUDA_EMU
    const double v = g_exp2_tab[m & 63] * p;
#else
    const double v = __ldg(&g_exp2_tab[m & 63]) * p;
#endif
    return __hiloint2double(__double2hiint(v) - ((m >> 6) << 20), __double2loint(v));
}

__device__ __forceinline__ double pse_eta_pi(double k2)    // pi * eta
{
    double poly = fma(fma(fma(-2.0 / 3.0, k2, 10.0), k2, -40.0), k2, 40.0);
    return poly * pse_exp_neg(k2);
}

// (d_ij / eps)^2 on the sphere.  The reference evaluates d_ij = R atan2(|x_i cross x_j|, x_i . x_j)
// (src/SphereGeometry.f90:107-125).  Inside the cut-off the angle is small, and
//   theta^2 = 4 atan^2(tan(theta / 2)),  tan^2(theta / 2) = |x_i cross x_j|^2 / (|x_i| |x_j| + x_i . x_j)^2 =: t
// has no cancellation (the cross product carries the small angle), so theta^2 = 4 t A(t) with
// A(t) = atan^2(sqrt t) / t = sum_k (-1)^(k-1) / k (1 + 1/3 + ... + 1/(2k-1)) t^(k-1): one reciprocal and
// a short polynomial (nterms from the cut-off angle, relative truncation < 1e-17) instead of the
// library sqrt + atan2 (~75 FP64 instructions).  nterms == 0 (cut-off angle > 0.9 rad) keeps atan2.
constexpr int kAtanSqMaxTerms = 28;
__constant__ double kAtanSq[kAtanSqMaxTerms] = {1.0, -0.6666666666666666, 0.5111111111111111, -0.41904761904761906, 0.3574603174603175, -0.31303511303511306, 0.279304822161965, -0.25272505272505275, 0.23118043902357627, -0.2133255530159555, 0.19826132525260023, -0.18536273655401397, 0.1741809875883206, -0.16438499112037178, 0.15572484228705963, -0.14800816867637648, 0.1410843370073561, -0.1348336198720268, 0.12915958866965838, -0.12398366051822675, 0.1192411168698559, -0.11487814855547557, 0.11084963001924716, -0.10711742025780689, 0.10364904997810687, -0.10041669586884333, 0.09739637100437883, -0.09456727983214452};

struct PseSphereConsts {
    double rad_over_eps;     // SphereRadius / eps
    double cos_cut;          // cos(kPseCut eps / SphereRadius), or -2 if the cut-off exceeds pi
    double chord_cut;        // sphere_chord_cut(eps, SphereRadius), for tile culling
    double scale;            // post-scaling of the summed field (the trailing MultiplyFieldByScalar)
    double k2_scale;         // 4 (SphereRadius / eps)^2
    int nterms;              // terms of A(t); 0: use atan2
};

// nn = |x_i| |x_j|
__device__ __forceinline__ double sphere_k2(double tx, double ty, double tz, double sx, double sy, double sz,
                                            double dot, double nn, const PseSphereConsts& c)
{
    double c0 = fma(ty, sz, -(sy * tz));
    double c1 = fma(sx, tz, -(tx * sz));
    double c2 = fma(tx, sy, -(sx * ty));
    double s2 = fma(c0, c0, fma(c1, c1, c2 * c2));
    if (c.nterms == 0) {
        double k = atan2(sqrt(s2), dot) * c.rad_over_eps;
        return k * k;
    }
    const double den = nn + dot;
    const double t = s2 * rcp_fast(den * den);
    double a = kAtanSq[c.nterms - 1];
#pragma unroll 1
    for (int n = c.nterms - 2; n >= 0; --n) a = fma(a, t, kAtanSq[n]);
    return c.k2_scale * (t * a);
}

// host: number of terms of A(t) for a cut-off angle, 0 if atan2 is the better choice
inline int atan_sq_terms(double theta_cut)
{
    if (!(theta_cut < 0.9)) return 0;
    const double t = tan(0.5 * theta_cut) * tan(0.5 * theta_cut) * 1.02;       // margin for the cut-off's rounding
    int n = 2;
    double tn = t;                      // t^(n-1)
    while (n < kAtanSqMaxTerms && tn * t > 1.0e-18) { tn *= t; ++n; }
    return tn * t > 1.0e-18 ? 0 : n;
}

// Sphere: d_ij = atan2(|x_i cross x_j|, x_i.x_j) * SphereRadius  (src/SphereGeometry.f90:107-125).
// Source record: x, y, z, f, A/(pi eps^2), |x_j|.
struct PseSphere : CullSphere {
    static constexpr int NS = 6, NA = 1;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *z, *f;
        PseSphereConsts c;       // c.scale = 1 / eps^2
        Outs<1> out;
    };
    __device__ static __forceinline__ double cull_dist(const Params& p) { return p.c.chord_cut; }
    struct Tgt { double x, y, z, f, nrm, thr; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        Tgt t{p.x[i], p.y[i], p.z[i], p.f[i], 0.0, 0.0};
        t.nrm = sqrt(t.x * t.x + t.y * t.y + t.z * t.z);
        t.thr = p.c.cos_cut * t.nrm;
        return t;
    }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const SharedCtx&)
    {
        double dot = fma(t.x, s[0], fma(t.y, s[1], t.z * s[2]));
        if (dot < t.thr * s[5]) return;           // angle beyond the cut-off
        const double k2 = sphere_k2(t.x, t.y, t.z, s[0], s[1], s[2], dot, t.nrm * s[5], p.c);
        acc[0] = fma((s[3] - t.f) * pse_eta_pi(k2), s[4], acc[0]);
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0] * p.c.scale);
    }
};

__global__ void pack_pse_sphere(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                                const double* __restrict__ x, const double* __restrict__ y,
                                const double* __restrict__ z, const double* __restrict__ f,
                                const double* __restrict__ area, double eps, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    // null source: zero area, and a norm so large that the cut-off test (dot < thr |x_j|) rejects it
    // before the distance is evaluated (0 / 0 in the series form of sphere_k2)
    double r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, kNullNorm};
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
struct PsePlane : CullPlane {
    static constexpr int NS = 4, NA = 1;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *f;
        double inv_eps2;
        Outs<1> out;
    };
    __device__ static __forceinline__ double cull_dist(const Params& p) { return kPseCut * rsqrt(p.inv_eps2); }
    struct Tgt { double x, y, f; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i], p.f[i]}; }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const SharedCtx&)
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
