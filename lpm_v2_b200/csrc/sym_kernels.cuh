// sym_kernels.cuh -- device code of the pair-symmetric BVE sums (see symmetric.cuh for the design and
// the host side).  Kept apart so that a kernel can be compiled alone for SASS inspection.
#pragma once
#include "directsum.cuh"
#include "pairs.cuh"

namespace lpm {

struct SymGeom {
    int32_t nsrc;           // active particles F
    int32_t nsrc_pad;       // padded to whole tiles (null records)
    int32_t ntiles;         // nsrc_pad / (the kernel's tile size TS; launch_sym converts from tiles of kTile)
    int32_t nblocks;        // target blocks of BLOCK*T compact indices
    int32_t chunk_tiles;    // source tiles per chunk
    int32_t nchunks;
    int32_t panel_blocks;   // target blocks per panel of the launch order (see sym_kernel)
    int32_t world, rank;    // target blocks are dealt round-robin to ranks (sums joined by the caller)
    int32_t half_bin;       // as DsGeom::half_bin (log kernels)
};

// Window of the fixed-point accumulators of one evaluation (sym_fx_scale_kernel).
struct FxWindow {
    double unit;            // 2^E0: what one count of limb 0 is worth
    int32_t fxe;            // 1075 + E0: biased exponent of a double whose 53-bit mantissa, read as an integer, counts units of 2^E0
    int32_t pad;
};

// kernel-wide constants of the symmetric functors
struct SymParams : LogParams {
    double R2;
    const FxWindow* fx;     // fx[w]: the window of component class w (K::window)
};

// ---- order-independent accumulation ------------------------------------------------------------
// A floating-point RED would make the result depend on the order in which the CTAs' contributions land.
// Every value that reaches an accumulator here is itself deterministic (a fixed (block, tile) computes it
// in a fixed order), so adding those values EXACTLY makes the total independent of the order -- and of how
// the blocks were dealt to ranks.  An accumulator is kFxLimbs signed 64-bit limbs, limb k counting units of
// 2^(E0 + 40 k), plus one overflow counter: a value's 53-bit mantissa is shifted to its place in the window
// and split into (at most three non-zero) 40-bit pieces, each added with an integer atomic; what lies below
// 2^E0 -- 240 bits under the top of the window, chosen per evaluation by sym_fx_scale_kernel -- is dropped
// (toward zero).  All of it is integer arithmetic on the otherwise idle ALU pipe: the first version split
// the value with 12 FP64 operations per RED and cost 4-16 % of a sum (profiles/r02_ab_sym.log).
// A value above the window (only a non-finite one, or coincident particles, can be) bumps the overflow
// counter, and the accumulator then reads as NaN -- what the reference's 0 * Inf gives for such a pair.
constexpr int kFxLimbs = 6;
constexpr int kFxWords = kFxLimbs + 1;      // limbs + overflow counter
constexpr int kFxLimbBits = 40;
__device__ __forceinline__ void sym_red_add(double* acc, size_t idx, double v, int fxe)
{
    unsigned long long* a = reinterpret_cast<unsigned long long*>(acc) + idx * kFxWords;
    const long long bits = __double_as_longlong(v);
    const int e = (int)((bits >> 52) & 0x7ff);
    int shift = e - fxe;                                    // v = +-m 2^shift units of 2^E0
    if (e == 0 || shift < -52) return;                      // zero / subnormal / wholly below the window
    if (shift > kFxLimbBits * kFxLimbs - 56) {              // above the window (or Inf / NaN)
        atomicAdd(a + kFxLimbs, 1ULL);
        return;
    }
    unsigned long long m = ((unsigned long long)bits & 0x000fffffffffffffULL) | 0x0010000000000000ULL;
    if (shift < 0) {
        m >>= -shift;
        shift = 0;
    }
    const int k = shift / kFxLimbBits, o = shift - kFxLimbBits * k;     // m 2^o spans limbs k .. k + 2
    const unsigned long long lo = m << o, hi = o ? (m >> (64 - o)) : 0ULL;
    constexpr unsigned long long MASK = (1ULL << kFxLimbBits) - 1;
    unsigned long long p0 = lo & MASK, p1 = ((lo >> kFxLimbBits) | (hi << (64 - kFxLimbBits))) & MASK,
                       p2 = hi >> (2 * kFxLimbBits - 64);
    if (bits < 0) {                                         // limbs are signed (two's complement)
        p0 = 0ULL - p0; p1 = 0ULL - p1; p2 = 0ULL - p2;
    }
    if (p0) atomicAdd(a + k, p0);
    if (p1) atomicAdd(a + k + 1, p1);
    if (p2) atomicAdd(a + k + 2, p2);
}

// Warp reduction of cb[s][a] (thread-local sums for SB sources, NC components: <= 3 with batches of 8, <= 7 with
// batches of 4) by recursive halving, then one add per (source, component) from the lane that ends up owning it;
// fxe[w]: the fixed-point window of component class w = WIN(q) (the fused kernel's velocity and stream sums differ).
// At each halving level a lane keeps the sums of one half of the sources it still holds and sends the other half to
// its partner (lane ^ 16, ^ 8, ^ 4).  Which half a lane keeps follows from its lane bits, so the lanes take the batch's
// sources in ROTATED order -- lane l reads source u ^ sym_rot<SB>(l) as its u-th (the functors' batch()) -- and slot k
// of cb always holds a source of the half this lane keeps: every lane keeps the low slots and sends the high ones, with
// no selects (round 2's first version selected keep / send per value: 84 FSEL per batch of 8 sources x 3 components,
// 8 % of the velocity kernel's instructions; 765 -> 730 ms at icosTri 8, profiles/r02m_ab_builds.log).  After the
// halving levels lane l holds source sym_rot<SB>(l) summed over the lanes that differ from it in the high bits; a
// butterfly over the remaining low bits finishes the sum.
template <int SB>
__device__ __forceinline__ int sym_rot(int lane) { return SB == 8 ? (lane >> 2) & 7 : (lane >> 3) & 3; }
// v[q] for a run-time q with static register indexing
template <int NC>
__device__ __forceinline__ double sym_pick(const double (&v)[NC], int q)
{
    double val = v[0];
#pragma unroll
    for (int a = 1; a < NC; ++a) val = (q == a) ? v[a] : val;
    return val;
}
template <class K, int SB, int NC, bool COMBINE = false>
__device__ __forceinline__ void sym_reduce_red(double (&cb)[SB][NC], int lane, double* __restrict__ acc, size_t idx0,
                                               const int (&fxe)[2], double* __restrict__ slot = nullptr)
{
    static_assert(SB == 8 || SB == 4, "source batch");
    static_assert(NC >= 1 && NC <= (SB == 8 ? 3 : 7), "components per source");
    constexpr unsigned FULL = 0xffffffffu;
    double v[NC];
    if constexpr (SB == 8) {
        double v4[4][NC], v2[2][NC];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int a = 0; a < NC; ++a) v4[k][a] = cb[k][a] + __shfl_xor_sync(FULL, cb[k + 4][a], 16);
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int a = 0; a < NC; ++a) v2[k][a] = v4[k][a] + __shfl_xor_sync(FULL, v4[k + 2][a], 8);
#pragma unroll
        for (int a = 0; a < NC; ++a) v[a] = v2[0][a] + __shfl_xor_sync(FULL, v2[1][a], 4);
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            v[a] += __shfl_xor_sync(FULL, v[a], 2);
            v[a] += __shfl_xor_sync(FULL, v[a], 1);
        }
    } else {
        double v2[2][NC];
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int a = 0; a < NC; ++a) v2[k][a] = cb[k][a] + __shfl_xor_sync(FULL, cb[k + 2][a], 16);
#pragma unroll
        for (int a = 0; a < NC; ++a) v[a] = v2[0][a] + __shfl_xor_sync(FULL, v2[1][a], 8);
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            v[a] += __shfl_xor_sync(FULL, v[a], 4);
            v[a] += __shfl_xor_sync(FULL, v[a], 2);
            v[a] += __shfl_xor_sync(FULL, v[a], 1);
        }
    }
    const int sidx = sym_rot<SB>(lane);
    const int q = lane & (SB == 8 ? 3 : 7);
    if (q < NC) {
        const double val = sym_pick<NC>(v, q);
        if constexpr (COMBINE) slot[sidx * NC + q] = val;      // this warp's sum; joined with the other warps' after the tile
        else sym_red_add(acc, idx0 + sidx * NC + q, val, fxe[K::window(q)]);
    }
}

// =============================================================================
// Functors of the symmetric kernels.  K provides
//   NS, NA, NC        doubles per record / sums per target / sums per source (the transposed direction)
//   KS                per-CTA shared table (doubles), init_shared() as in directsum.cuh
//   Tgt, from_record  what a thread keeps of a target (read from its packed record), null()
//   batch<T,SB,ORDER> SB sources of a tile above the diagonal against the thread's T targets, both
//                     directions: a[t] += (source's weight) g(d) and cb[u] = sum_t (target's weight) g(d)
//   diag<T>           one source of a diagonal tile, one-sided, pair skipped where isself[t]
// =============================================================================

// BVE velocity (BveVelT in pairs.cuh): a_i = sum_j P_j / (R^2 - x_i.x_j); record x, y, z, Px, Py, Pz.
// ORDER permutes INDEPENDENT statements only (as BveVelT's ORDER does): the kernel is bound by register
// operand delivery, and what ptxas allocates and where it can set .reuse follows the statement order.
//   bit 0     denominators coordinate by coordinate (else target by target)
//   bit 1     a-phase component by component (else target by target)
//   bits 2-3  sources handled together, phase by phase: 1, 2, 4, SB
//   bit 4     cb-phase nest (target, component, source) -- P_t stays in the reuse cache -- else
//             (source, target, component) -- 1/d stays
//   bit 5     a scheduling fence after every source group (sched_fence below)
struct SymBveVel : NoSharedTable {
    static constexpr int NS = 6, NA = 3, NC = 3;
    __host__ __device__ static constexpr int window(int) { return 0; }
    struct Tgt { double x, y, z, px, py, pz; };
    __device__ static __forceinline__ Tgt null() { return Tgt{0.0, 0.0, 0.0, 0.0, 0.0, 0.0}; }
    __device__ static __forceinline__ Tgt from_record(const double2* p2)
    {
        const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
        return Tgt{v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
    }
    template <int T, int SB, int ORDER>
    __device__ static __forceinline__ void batch(const SymParams& p, const Tgt (&tg)[T], double (&a)[T][NA],
                                                 const double* __restrict__ sm, double (&cb)[SB][NC], const SharedCtx&,
                                                 const int rot)
    {
        constexpr int DN = ORDER & 1, AN = (ORDER >> 1) & 1, GS = (ORDER >> 2) & 3, CU = (ORDER >> 4) & 1, FENCE = (ORDER >> 5) & 1;
        constexpr int G = GS == 0 ? 1 : GS == 1 ? 2 : GS == 2 ? 4 : SB;
        static_assert(SB % G == 0, "source group");
#pragma unroll
        for (int g0 = 0; g0 < SB; g0 += G) {
            double s[G][NS], d[G][T], r[G][T];
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const double2* p2 = reinterpret_cast<const double2*>(sm + ((g0 + u) ^ rot) * NS);     // this lane's u-th source (sym_reduce_red)
#pragma unroll
                for (int q = 0; q < NS / 2; ++q) {
                    const double2 v = p2[q];
                    s[u][2 * q] = v.x; s[u][2 * q + 1] = v.y;
                }
            }
            if constexpr (DN == 0) {
#pragma unroll
                for (int u = 0; u < G; ++u)
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        d[u][t] = fma(-tg[t].x, s[u][0], p.R2);
                        d[u][t] = fma(-tg[t].y, s[u][1], d[u][t]);
                        d[u][t] = fma(-tg[t].z, s[u][2], d[u][t]);
                    }
            } else {
#pragma unroll
                for (int u = 0; u < G; ++u)
#pragma unroll
                    for (int t = 0; t < T; ++t) d[u][t] = fma(-tg[t].x, s[u][0], p.R2);
#pragma unroll
                for (int u = 0; u < G; ++u)
#pragma unroll
                    for (int t = 0; t < T; ++t) d[u][t] = fma(-tg[t].y, s[u][1], d[u][t]);
#pragma unroll
                for (int u = 0; u < G; ++u)
#pragma unroll
                    for (int t = 0; t < T; ++t) d[u][t] = fma(-tg[t].z, s[u][2], d[u][t]);
            }
#pragma unroll
            for (int u = 0; u < G; ++u) rcp_batch<T>(d[u], r[u]);
            if constexpr (AN == 0) {
#pragma unroll
                for (int u = 0; u < G; ++u)
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        a[t][0] = fma(r[u][t], s[u][3], a[t][0]);
                        a[t][1] = fma(r[u][t], s[u][4], a[t][1]);
                        a[t][2] = fma(r[u][t], s[u][5], a[t][2]);
                    }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int u = 0; u < G; ++u)
#pragma unroll
                        for (int t = 0; t < T; ++t) a[t][c] = fma(r[u][t], s[u][3 + c], a[t][c]);
            }
            if constexpr (CU == 0) {
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    cb[g0 + u][0] = r[u][0] * tg[0].px; cb[g0 + u][1] = r[u][0] * tg[0].py; cb[g0 + u][2] = r[u][0] * tg[0].pz;
#pragma unroll
                    for (int t = 1; t < T; ++t) {
                        cb[g0 + u][0] = fma(r[u][t], tg[t].px, cb[g0 + u][0]);
                        cb[g0 + u][1] = fma(r[u][t], tg[t].py, cb[g0 + u][1]);
                        cb[g0 + u][2] = fma(r[u][t], tg[t].pz, cb[g0 + u][2]);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < G; ++u) cb[g0 + u][0] = r[u][0] * tg[0].px;
#pragma unroll
                for (int u = 0; u < G; ++u) cb[g0 + u][1] = r[u][0] * tg[0].py;
#pragma unroll
                for (int u = 0; u < G; ++u) cb[g0 + u][2] = r[u][0] * tg[0].pz;
#pragma unroll
                for (int t = 1; t < T; ++t) {
#pragma unroll
                    for (int u = 0; u < G; ++u) cb[g0 + u][0] = fma(r[u][t], tg[t].px, cb[g0 + u][0]);
#pragma unroll
                    for (int u = 0; u < G; ++u) cb[g0 + u][1] = fma(r[u][t], tg[t].py, cb[g0 + u][1]);
#pragma unroll
                    for (int u = 0; u < G; ++u) cb[g0 + u][2] = fma(r[u][t], tg[t].pz, cb[g0 + u][2]);
                }
            }
            // (the lanes' records differ, and a branch on lane-dependent data would make ptxas guard the shuffles that
            // follow against divergence -- the fence tests a warp-uniform load)
            if constexpr (FENCE != 0) sched_fence(sm[g0 * NS]);
        }
    }
    // as BveVelT::group<T, true>
    template <int T>
    __device__ static __forceinline__ void diag(const SymParams& p, const Tgt (&tg)[T], double (&a)[T][NA],
                                                const double (&s)[NS], const bool (&isself)[T], const SharedCtx&)
    {
        double d[T], r[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            d[t] = fma(-tg[t].x, s[0], p.R2);
            d[t] = fma(-tg[t].y, s[1], d[t]);
            d[t] = fma(-tg[t].z, s[2], d[t]);
            d[t] = isself[t] ? 1.0 : d[t];          // the self pair is zeroed below
        }
        rcp_group<T, true>(d, r);                   // one by one: a coincident pair stays with its own targets
#pragma unroll
        for (int t = 0; t < T; ++t) {
            r[t] = isself[t] ? 0.0 : r[t];
            a[t][0] = fma(r[t], s[3], a[t][0]);
            a[t][1] = fma(r[t], s[4], a[t][1]);
            a[t][2] = fma(r[t], s[5], a[t][2]);
        }
    }
};

// BVE stream functions (BveStream in pairs.cuh): psi_i = sum_j w_j ln(R^2 - x_i.x_j) for two weights;
// record x, y, z, w_rel, w_abs, 0.  One dot product and ONE logarithm serve both directions of a pair:
// 3 + 6 + 2 + 2 = 13 FP64 instructions for two interactions instead of 22.
// The branch-free table logarithm (log_group_fast) is evaluated for a source's T pairs; if any argument
// of this thread was outside the table window those logarithms are recomputed with the library log()
// before anything is accumulated (the one-sided kernel redoes a whole tile instead).
struct SymBveStream : LogSharedTable<32> {
    static constexpr int NS = 6, NA = 2, NC = 2;
    __host__ __device__ static constexpr int window(int) { return 0; }
    struct Tgt { double x, y, z, w0, w1; };
    __device__ static __forceinline__ Tgt null() { return Tgt{0.0, 0.0, 0.0, 0.0, 0.0}; }
    __device__ static __forceinline__ Tgt from_record(const double2* p2)
    {
        const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
        return Tgt{v0.x, v0.y, v1.x, v1.y, v2.x};
    }
    // The retry is per source: a branch after every source's logarithms also keeps ptxas from interleaving the whole
    // batch (see sched_fence); a retry per batch measured 5 % slower (profiles/r02b_ab_paths.log).
    template <int T, int SB, int ORDER>
    __device__ static __forceinline__ void batch(const SymParams& p, const Tgt (&tg)[T], double (&a)[T][NA],
                                                 const double* __restrict__ sm, double (&cb)[SB][NC], const SharedCtx& sc,
                                                 const int rot)
    {
#pragma unroll
        for (int u = 0; u < SB; ++u) {
            const double2* p2 = reinterpret_cast<const double2*>(sm + (u ^ rot) * NS);     // this lane's u-th source (sym_reduce_red)
            const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
            double d[T], l[T];
            unsigned worst = 0;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                d[t] = fma(-tg[t].x, v0.x, p.R2);
                d[t] = fma(-tg[t].y, v0.y, d[t]);
                d[t] = fma(-tg[t].z, v1.x, d[t]);
            }
            log_group_fast<KS, T>(d, l, worst, sc);
            if (__builtin_expect(needs_retry(worst), 0)) {
#pragma unroll
                for (int t = 0; t < T; ++t) l[t] = log_slow_path(d[t]);
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                a[t][0] = fma(l[t], v1.y, a[t][0]);
                a[t][1] = fma(l[t], v2.x, a[t][1]);
            }
            cb[u][0] = l[0] * tg[0].w0; cb[u][1] = l[0] * tg[0].w1;
#pragma unroll
            for (int t = 1; t < T; ++t) {
                cb[u][0] = fma(l[t], tg[t].w0, cb[u][0]);
                cb[u][1] = fma(l[t], tg[t].w1, cb[u][1]);
            }
        }
    }
    // as BveStream::group<T, true>
    template <int T>
    __device__ static __forceinline__ void diag(const SymParams& p, const Tgt (&tg)[T], double (&a)[T][NA],
                                                const double (&s)[NS], const bool (&isself)[T], const SharedCtx& sc)
    {
        double d[T], l[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            d[t] = fma(-tg[t].x, s[0], p.R2);
            d[t] = fma(-tg[t].y, s[1], d[t]);
            d[t] = fma(-tg[t].z, s[2], d[t]);
            d[t] = isself[t] ? p.R2 : d[t];         // any in-window value; the pair is zeroed below
        }
        log_group<KS, T>(d, l, sc);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            l[t] = isself[t] ? 0.0 : l[t];
            a[t][0] = fma(l[t], s[3], a[t][0]);
            a[t][1] = fma(l[t], s[4], a[t][1]);
        }
    }
};

// BVE velocity and stream functions fused (BveVelStream in pairs.cuh; the end of every RK4 step): one denominator, one
// reciprocal and one logarithm serve a_c += P_c' / d, a_c' += P_c / d, psi_c += w_c' ln d and psi_c' += w_c ln d:
// 3 + 3 + 6 + (3 + 2) + (3 + 2) = 22 FP64 instructions for two interactions of each kind, against 12 + 13 in the two
// separate kernels.  Record x, y, z, Px, Py, Pz, w_rel, w_abs.  Components 0-2 (velocity) and 3-4 (stream functions)
// have their own fixed-point windows.
struct SymBveVelStream : LogSharedTable<32> {
    static constexpr int NS = 8, NA = 5, NC = 5;
    __host__ __device__ static constexpr int window(int q) { return q >= 3 ? 1 : 0; }
    struct Tgt { double x, y, z, px, py, pz, w0, w1; };
    __device__ static __forceinline__ Tgt null() { return Tgt{0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}; }
    __device__ static __forceinline__ Tgt from_record(const double2* p2)
    {
        const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2], v3 = p2[3];
        return Tgt{v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, v3.x, v3.y};
    }
    template <int T, int SB, int ORDER>
    __device__ static __forceinline__ void batch(const SymParams& p, const Tgt (&tg)[T], double (&a)[T][NA],
                                                 const double* __restrict__ sm, double (&cb)[SB][NC], const SharedCtx& sc,
                                                 const int rot)
    {
#pragma unroll
        for (int u = 0; u < SB; ++u) {
            double s[NS], d[T], r[T], l[T];
            const double2* p2 = reinterpret_cast<const double2*>(sm + (u ^ rot) * NS);     // this lane's u-th source (sym_reduce_red)
#pragma unroll
            for (int q = 0; q < NS / 2; ++q) {
                const double2 v = p2[q];
                s[2 * q] = v.x; s[2 * q + 1] = v.y;
            }
            unsigned worst = 0;
#pragma unroll
            for (int t = 0; t < T; ++t) d[t] = fma(-tg[t].x, s[0], p.R2);
#pragma unroll
            for (int t = 0; t < T; ++t) d[t] = fma(-tg[t].y, s[1], d[t]);
#pragma unroll
            for (int t = 0; t < T; ++t) d[t] = fma(-tg[t].z, s[2], d[t]);
            rcp_batch<T>(d, r);
            log_group_fast<KS, T>(d, l, worst, sc);
            if (__builtin_expect(needs_retry(worst), 0)) {
#pragma unroll
                for (int t = 0; t < T; ++t) l[t] = log_slow_path(d[t]);
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                a[t][0] = fma(r[t], s[3], a[t][0]);
                a[t][1] = fma(r[t], s[4], a[t][1]);
                a[t][2] = fma(r[t], s[5], a[t][2]);
                a[t][3] = fma(l[t], s[6], a[t][3]);
                a[t][4] = fma(l[t], s[7], a[t][4]);
            }
            cb[u][0] = r[0] * tg[0].px; cb[u][1] = r[0] * tg[0].py; cb[u][2] = r[0] * tg[0].pz;
            cb[u][3] = l[0] * tg[0].w0; cb[u][4] = l[0] * tg[0].w1;
#pragma unroll
            for (int t = 1; t < T; ++t) {
                cb[u][0] = fma(r[t], tg[t].px, cb[u][0]);
                cb[u][1] = fma(r[t], tg[t].py, cb[u][1]);
                cb[u][2] = fma(r[t], tg[t].pz, cb[u][2]);
                cb[u][3] = fma(l[t], tg[t].w0, cb[u][3]);
                cb[u][4] = fma(l[t], tg[t].w1, cb[u][4]);
            }
        }
    }
    // as BveVelStream::group<T, true>
    template <int T>
    __device__ static __forceinline__ void diag(const SymParams& p, const Tgt (&tg)[T], double (&a)[T][NA],
                                                const double (&s)[NS], const bool (&isself)[T], const SharedCtx& sc)
    {
        double d[T], r[T], l[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            d[t] = fma(-tg[t].x, s[0], p.R2);
            d[t] = fma(-tg[t].y, s[1], d[t]);
            d[t] = fma(-tg[t].z, s[2], d[t]);
            d[t] = isself[t] ? p.R2 : d[t];         // any in-window value; the pair is zeroed below
        }
        rcp_group<T, true>(d, r);
        log_group<KS, T>(d, l, sc);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            r[t] = isself[t] ? 0.0 : r[t];
            l[t] = isself[t] ? 0.0 : l[t];
            a[t][0] = fma(r[t], s[3], a[t][0]);
            a[t][1] = fma(r[t], s[4], a[t][1]);
            a[t][2] = fma(r[t], s[5], a[t][2]);
            a[t][3] = fma(l[t], s[6], a[t][3]);
            a[t][4] = fma(l[t], s[7], a[t][4]);
        }
    }
};

// ---- the kernel ---------------------------------------------------------------
// acc: [nsrc_pad][NC][kFxWords] 64-bit words (fixed-point limbs + overflow counter, see sym_red_add), zeroed by the caller
// (NA == NC: both directions feed the same sums).
// dynamic shared memory: [2 tiles][K::KS table][2 mbarriers] and, with COMBINE, [2][warps][TS][NC] per-warp source sums:
// the warps' sums for a tile are added in warp order after the tile and ONE RED per (CTA, source, component) is
// issued instead of one per warp (a quarter of the atomics with 128 threads)
// TS: sources per shared-memory tile (SymGeom's ntiles and chunk_tiles count tiles of TS); the kernels that carry the
// 64 KB log table take tiles shorter than kTile so that the combine buffers fit beside it with two CTAs per SM.
template <class K, int T, int BLOCK, int SB, int MINB, int ORDER = 0, bool COMBINE = false, int TS = kTile>
__global__ void __launch_bounds__(BLOCK, MINB)
sym_kernel(const SymParams prm, const SymGeom g, const double* __restrict__ src, double* __restrict__ acc)
{
    constexpr int NS = K::NS, NA = K::NA, NC = K::NC, TB = BLOCK * T, DT = TB / TS;
    static_assert(kTile % TS == 0, "records are padded to whole tiles of kTile");
    static_assert(NS % 2 == 0, "records are read as double2");
    static_assert(NA == NC, "one accumulator array for both directions");
    static_assert(TB % TS == 0, "a target block must be whole source tiles");
    static_assert(TS % SB == 0, "source batch");
    constexpr uint32_t kTileBytes = TS * NS * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double(*tile)[TS * NS] = reinterpret_cast<double(*)[TS * NS]>(smem_raw);
    double* ks = reinterpret_cast<double*>(smem_raw + 2 * kTileBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(ks + K::KS);
    constexpr int NW = BLOCK / 32;
    double* cT = reinterpret_cast<double*>(full + 2);        // COMBINE only

    const int tid = threadIdx.x, lane = tid & 31;
    // Launch order: panels of `panel_blocks` target blocks; inside a panel the chunk is the slow index, so the CTAs
    // in flight read the same few source tiles (and hit the same few accumulators) while the panel's target
    // records stay in L2.  With one panel over all blocks (round 2's first version) every chunk re-read all target
    // records -- 63 MB at icosTri 8, which the L2 does not hold beside 220 MB of accumulators: ncu showed 58 GB of
    // DRAM traffic per launch (harmless at 75 GB/s, but 370 x the algorithmic bytes).  A panel of 256 blocks keeps
    // 12.5 MB of targets hot and passes over the sources and accumulators once per panel.
    const int per_panel = g.panel_blocks * g.nchunks;
    const int pnl = blockIdx.x / per_panel, rem = blockIdx.x % per_panel;
    const int ck = rem / g.panel_blocks;
    const int I = pnl * g.panel_blocks + rem % g.panel_blocks;
    if (I >= g.nblocks) return;
    // Blocks are dealt to ranks in serpentine order (0 .. w-1, w-1 .. 0, ...): a block's work falls linearly with
    // its index (the triangle), so plain round-robin would give rank 0 the longest block of every round.
    if (g.world > 1) {
        const int q = I / g.world, p = I % g.world;
        if (((q & 1) ? g.world - 1 - p : p) != g.rank) return;
    }
    const int kdiag = I * DT;                        // first of this block's DT diagonal tiles
    int k0 = ck * g.chunk_tiles;
    const int k1 = min(k0 + g.chunk_tiles, g.ntiles);
    if (k0 < kdiag) k0 = kdiag;
    if (k0 >= k1) return;                            // chunk entirely below the diagonal (whole CTA)

    typename K::Tgt tg[T];
    double a[T][NA];
    int32_t cidx[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int32_t c = I * TB + t * BLOCK + tid;
        cidx[t] = c;
        tg[t] = K::null();                           // past the active list: a null TARGET (K::null() never coincides
        if (c < g.nsrc)                              // with a null source record, see SymPlaneVel)
            tg[t] = K::from_record(reinterpret_cast<const double2*>(src + (size_t)c * NS));
#pragma unroll
        for (int q = 0; q < NA; ++q) a[t][q] = 0.0;
    }
    const SharedCtx sctx{ks, K::init_shared(ks, prm, tid, BLOCK), g.half_bin};
    const int fxe[2] = {prm.fx[0].fxe, prm.fx[K::window(NC - 1)].fxe};      // one window per component class
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto load_tile = [&](const int k, const int st) {      // thread 0 only
        mbar_expect_tx(&full[st], kTileBytes);
        tma_bulk_g2s(tile[st], src + (size_t)k * TS * NS, kTileBytes, &full[st]);
    };
    // the block against itself: one-sided, self pair excluded
    auto diag_tile = [&](const int k, const int st) {
        const double* sm = tile[st];
        const int32_t j0 = k * TS;
#pragma unroll 1
        for (int j = 0; j < TS; ++j) {
            double s[NS];
            const double2* p2 = reinterpret_cast<const double2*>(sm + j * NS);
#pragma unroll
            for (int q = 0; q < NS / 2; ++q) {
                const double2 v = p2[q];
                s[2 * q] = v.x; s[2 * q + 1] = v.y;
            }
            bool isself[T];
#pragma unroll
            for (int t = 0; t < T; ++t) isself[t] = (j0 + j == cidx[t]);
            K::template diag<T>(prm, tg, a, s, isself, sctx);
        }
    };
    // a tile above the diagonal: every pair once, both directions
    auto sym_tile = [&](const int k, const int st, const int parity) {
        const double* sm = tile[st];
        double* mine = cT + ((size_t)parity * NW + (tid >> 5)) * TS * NC;        // this warp's slots (COMBINE)
#pragma unroll 1
        for (int jb = 0; jb < TS; jb += SB) {
            double cb[SB][NC];
            K::template batch<T, SB, ORDER>(prm, tg, a, sm + jb * NS, cb, sctx, sym_rot<SB>(lane));
            sym_reduce_red<K, SB, NC, COMBINE>(cb, lane, acc, ((size_t)k * TS + jb) * NC, fxe, mine + jb * NC);
        }
    };
    // COMBINE: after the barrier that ends tile k, add the warps' sums in warp order and issue the REDs.  The slots are
    // double-buffered by tile parity: a warp that is already in the next tile writes the other half.
    auto flush = [&](const int k, const int parity) {
        const double* base = cT + (size_t)parity * NW * TS * NC;
        for (int idx = tid; idx < TS * NC; idx += BLOCK) {
            double sum = base[idx];
#pragma unroll
            for (int w = 1; w < NW; ++w) sum += base[(size_t)w * TS * NC + idx];
            sym_red_add(acc, (size_t)k * TS * NC + idx, sum, fxe[K::window(idx % NC)]);
        }
    };

    const int nt = k1 - k0;
    if (tid == 0) {
        load_tile(k0, 0);
        if (nt > 1) load_tile(k0 + 1, 1);
    }
    for (int it = 0; it < nt; ++it) {
        const int st = it & 1, k = k0 + it;
        mbar_wait(&full[st], (it >> 1) & 1);
        if (k < kdiag + DT) diag_tile(k, st);
        else sym_tile(k, st, it & 1);
        __syncthreads();        // everyone is done with tile[st]
        if (tid == 0 && it + 2 < nt) load_tile(k + 2, st);
        if constexpr (COMBINE) {
            if (k >= kdiag + DT) flush(k, it & 1);
        }
    }
    // this CTA's own sums join the accumulators
#pragma unroll
    for (int t = 0; t < T; ++t)
        if (cidx[t] < g.nsrc) {
#pragma unroll
            for (int q = 0; q < NA; ++q) sym_red_add(acc, (size_t)cidx[t] * NA + q, a[t][q], fxe[K::window(q)]);
        }
}

template <class K, int T, int BLOCK, bool COMBINE = false, int TS = kTile>
constexpr size_t sym_smem_bytes()
{
    return 2 * size_t(TS) * K::NS * sizeof(double) + sizeof(double) * K::KS + 2 * sizeof(uint64_t) +
           (COMBINE ? 2 * size_t(BLOCK / 32) * TS * K::NC * sizeof(double) : 0);
}

// Window of the fixed-point accumulators for one evaluation.  maxhi: high word of the largest |entry| of the source
// records (coordinates and strengths; absmax_hi_kernel).  mode 0 (velocity): |sum| <= F M / d_min with
// d_min >= R^2 2^-110 (two FP64 points cannot be closer); mode 1 (stream functions): |ln| < 2^10.  The top of the
// window sits 8 bits above that bound, E0 = top - 240.
__global__ void sym_fx_scale_kernel(int mode, double R2, int32_t nsrc, const int32_t* __restrict__ maxhi, FxWindow* __restrict__ fx)
{
    const int eM = ((*maxhi >> 20) & 0x7ff) - 1023 + 1;
    int eF = 1;
    while ((1 << eF) < nsrc && eF < 31) ++eF;
    const int eR = ((__double2hiint(R2) >> 20) & 0x7ff) - 1023;
    const int top = eM + eF + (mode == 0 ? 110 - eR : 12) + 8;
    int e0 = top - kFxLimbBits * kFxLimbs;
    if (e0 > 700) e0 = 700;
    if (e0 < -900) e0 = -900;
    fx->unit = __hiloint2double((1023 + e0) << 20, 0);
    fx->fxe = 1075 + e0;
    fx->pad = 0;
}

// limbs -> doubles (values are [sum][component], nc components per sum): carries first (so that every limb but the top one is below 2^40), then the sum from the top;
// an accumulator whose overflow counter is set reads as NaN (see sym_red_add)
__global__ void __launch_bounds__(256)
sym_fx_to_double_kernel(int64_t nvalues, const long long* __restrict__ words, const FxWindow* __restrict__ fx,
                        double* __restrict__ out, int nc, int split)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvalues) return;
    long long l[kFxLimbs];
#pragma unroll
    for (int k = 0; k < kFxLimbs; ++k) l[k] = words[i * kFxWords + k];
    const long long over = words[i * kFxWords + kFxLimbs];
#pragma unroll
    for (int k = 0; k + 1 < kFxLimbs; ++k) {
        const long long carry = l[k] >> kFxLimbBits;        // arithmetic shift: floor
        l[k] -= carry << kFxLimbBits;
        l[k + 1] += carry;
    }
    constexpr double unit[kFxLimbs] = {0x1p0, 0x1p40, 0x1p80, 0x1p120, 0x1p160, 0x1p200};
    double s = 0.0;
#pragma unroll
    for (int k = kFxLimbs - 1; k >= 0; --k) s = fma((double)l[k], unit[k], s);
    const double win = fx[(int)(i % nc) >= split ? 1 : 0].unit;       // components >= split use the second window
    out[i] = over != 0 ? __longlong_as_double(0x7ff8000000000000LL) : s * win;
}

// u_i = x_i cross a_i for the active particles (BveVelT::finalize)
__global__ void __launch_bounds__(256)
sym_bve_finalize(int32_t nsrc, const int32_t* __restrict__ active, const double* __restrict__ src,
                 const double* __restrict__ acc, Outs<3> out)
{
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc) return;
    const int64_t i = active[c];
    const double x = src[(size_t)c * 6], y = src[(size_t)c * 6 + 1], z = src[(size_t)c * 6 + 2];
    const double a0 = acc[(size_t)c * 3], a1 = acc[(size_t)c * 3 + 1], a2 = acc[(size_t)c * 3 + 2];
    out.store(0, i, fma(y, a2, -(z * a1)));
    out.store(1, i, fma(z, a0, -(x * a2)));
    out.store(2, i, fma(x, a1, -(y * a0)));
}

// fused sums: u_i = x_i cross a_i, and the two stream functions copied out
__global__ void __launch_bounds__(256)
sym_bve_velstream_finalize(int32_t nsrc, const int32_t* __restrict__ active, const double* __restrict__ src,
                           const double* __restrict__ acc, Outs<5> out)
{
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc) return;
    const int64_t i = active[c];
    const double x = src[(size_t)c * 8], y = src[(size_t)c * 8 + 1], z = src[(size_t)c * 8 + 2];
    const double* a = acc + (size_t)c * 5;
    out.store(0, i, fma(y, a[2], -(z * a[1])));
    out.store(1, i, fma(z, a[0], -(x * a[2])));
    out.store(2, i, fma(x, a[1], -(y * a[0])));
    out.store(3, i, a[3]);
    out.store(4, i, a[4]);
}

// two sums per active particle, copied out (BveStream::finalize)
__global__ void __launch_bounds__(256)
sym_stream_finalize(int32_t nsrc, const int32_t* __restrict__ active, const double* __restrict__ acc, Outs<2> out)
{
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc) return;
    const int64_t i = active[c];
    out.store(0, i, acc[(size_t)c * 2]);
    out.store(1, i, acc[(size_t)c * 2 + 1]);
}

// passive[i - scan[i]] = i for every particle with mask 0 (stable, like the active list)
__global__ void __launch_bounds__(256)
passive_list_kernel(int64_t n, const int32_t* __restrict__ scan, int32_t* __restrict__ passive)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && scan[i + 1] == scan[i]) passive[i - scan[i]] = (int32_t)i;
}

}  // namespace lpm
