// sym_kernels.cuh -- device code of the EXPERIMENTAL pair-symmetric paths (see symmetric.cuh for the
// design and the host side).  Kept apart so that tools/sym_score.py can compile a kernel alone.
#pragma once
#include "directsum.cuh"
#include "pairs.cuh"

namespace lpm {

struct SymGeom {
    int32_t nsrc;           // active particles F
    int32_t nsrc_pad;       // padded to whole tiles (null records)
    int32_t ntiles;         // nsrc_pad / kTile
    int32_t nblocks;        // target blocks of BLOCK*T compact indices
    int32_t chunk_tiles;    // source tiles per chunk
    int32_t nchunks;
    int32_t world, rank;    // target blocks are dealt round-robin to ranks (sums joined by the caller)
    int32_t half_bin;       // as DsGeom::half_bin (log kernels)
};

// kernel-wide constants of the symmetric functors
struct SymParams : LogParams {
    double R2;
    const double* fx;       // fixed-point accumulation (kernels built with FX): fx[0] = 2^-E0, see sym_red_add
};

// ---- order-independent accumulation ------------------------------------------------------------
// RED.ADD.F64 makes the result depend on the order in which the CTAs' contributions land.  Every value that
// reaches a RED here is itself deterministic (a fixed (block, tile, warp) computes it in a fixed order), so adding
// those values EXACTLY makes the total independent of the order -- and of how the blocks were dealt to ranks.
// In a kernel built with FX, an accumulator is kFxLimbs signed 64-bit limbs, limb k counting units of 2^(E0 + 40 k): a value is
// split into (at most three non-zero) 40-bit pieces and each is added with an integer atomic; what lies below
// 2^E0 -- 240 bits under the top of the window, chosen per evaluation by sym_fx_scale_kernel -- is dropped.
constexpr int kFxLimbs = 6;
template <bool FX>
__device__ __forceinline__ void sym_red_add(double* acc, size_t idx, double v, const double* __restrict__ fx)
{
    if constexpr (!FX) {
        atomicAdd(acc + idx, v);
        return;
    }
    unsigned long long* a = reinterpret_cast<unsigned long long*>(acc) + idx * kFxLimbs;
    constexpr double unit[kFxLimbs] = {0x1p0, 0x1p40, 0x1p80, 0x1p120, 0x1p160, 0x1p200};
    constexpr double inv[kFxLimbs] = {0x1p0, 0x1p-40, 0x1p-80, 0x1p-120, 0x1p-160, 0x1p-200};
    double r = v * fx[0];                   // a power of two: exact
#pragma unroll
    for (int k = kFxLimbs - 1; k >= 0; --k) {
        const double t = trunc(r * inv[k]);
        r = fma(-t, unit[k], r);            // exact: t unit[k] is the part of r at and above 2^(40 k)
        if (t != 0.0) atomicAdd(a + k, (unsigned long long)(long long)t);
    }
}

// Warp reduction of cb[s][a] (thread-local sums for SB sources, NC <= 3 components) by recursive
// halving, then one RED per (source, component) from the lane that ends up owning it.
// After the halving levels lane l holds source ((l >> (5 - LV)) & (SB - 1)) summed over the lanes
// that differ from it in the high LV bits; a butterfly over the remaining low bits finishes the sum.
template <int SB, int NC, bool FX, bool COMBINE = false>
__device__ __forceinline__ void sym_reduce_red(double (&cb)[SB][NC], int lane, double* __restrict__ acc, size_t idx0,
                                               const double* __restrict__ fx, double* __restrict__ slot = nullptr)
{
    static_assert(SB == 8 || SB == 4, "source batch");
    static_assert(NC >= 1 && NC <= 3, "components per source");
    constexpr unsigned FULL = 0xffffffffu;
    double v[NC];
    int sidx;
    if constexpr (SB == 8) {
        const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
        double v4[4][NC], v2[2][NC];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int a = 0; a < NC; ++a) {
                const double lo = cb[k][a], hi = cb[k + 4][a];
                const double keep = b4 ? hi : lo, send = b4 ? lo : hi;
                v4[k][a] = keep + __shfl_xor_sync(FULL, send, 16);
            }
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int a = 0; a < NC; ++a) {
                const double lo = v4[k][a], hi = v4[k + 2][a];
                const double keep = b3 ? hi : lo, send = b3 ? lo : hi;
                v2[k][a] = keep + __shfl_xor_sync(FULL, send, 8);
            }
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            const double lo = v2[0][a], hi = v2[1][a];
            const double keep = b2 ? hi : lo, send = b2 ? lo : hi;
            v[a] = keep + __shfl_xor_sync(FULL, send, 4);
        }
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            v[a] += __shfl_xor_sync(FULL, v[a], 2);
            v[a] += __shfl_xor_sync(FULL, v[a], 1);
        }
        sidx = (lane >> 2) & 7;
        const int q = lane & 3;
        if (q < NC) {
            const double val = q == 0 ? v[0] : (q == 1 ? v[NC > 1 ? 1 : 0] : v[NC > 2 ? 2 : 0]);
            if constexpr (COMBINE) slot[sidx * NC + q] = val;      // this warp's sum; joined with the other warps' after the tile
            else sym_red_add<FX>(acc, idx0 + sidx * NC + q, val, fx);
        }
    } else {
        const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
        double v2[2][NC];
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int a = 0; a < NC; ++a) {
                const double lo = cb[k][a], hi = cb[k + 2][a];
                const double keep = b4 ? hi : lo, send = b4 ? lo : hi;
                v2[k][a] = keep + __shfl_xor_sync(FULL, send, 16);
            }
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            const double lo = v2[0][a], hi = v2[1][a];
            const double keep = b3 ? hi : lo, send = b3 ? lo : hi;
            v[a] = keep + __shfl_xor_sync(FULL, send, 8);
        }
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            v[a] += __shfl_xor_sync(FULL, v[a], 4);
            v[a] += __shfl_xor_sync(FULL, v[a], 2);
            v[a] += __shfl_xor_sync(FULL, v[a], 1);
        }
        sidx = (lane >> 3) & 3;
        const int q = lane & 7;
        if (q < NC) {
            const double val = q == 0 ? v[0] : (q == 1 ? v[NC > 1 ? 1 : 0] : v[NC > 2 ? 2 : 0]);
            if constexpr (COMBINE) slot[sidx * NC + q] = val;      // this warp's sum; joined with the other warps' after the tile
            else sym_red_add<FX>(acc, idx0 + sidx * NC + q, val, fx);
        }
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
    struct Tgt { double x, y, z, px, py, pz; };
    __device__ static __forceinline__ Tgt null() { return Tgt{0.0, 0.0, 0.0, 0.0, 0.0, 0.0}; }
    __device__ static __forceinline__ Tgt from_record(const double2* p2)
    {
        const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
        return Tgt{v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
    }
    template <int T, int SB, int ORDER>
    __device__ static __forceinline__ void batch(const SymParams& p, const Tgt (&tg)[T], double (&a)[T][NA],
                                                 const double* __restrict__ sm, double (&cb)[SB][NC], const SharedCtx&)
    {
        constexpr int DN = ORDER & 1, AN = (ORDER >> 1) & 1, GS = (ORDER >> 2) & 3, CU = (ORDER >> 4) & 1, FENCE = (ORDER >> 5) & 1;
        constexpr int G = GS == 0 ? 1 : GS == 1 ? 2 : GS == 2 ? 4 : SB;
        static_assert(SB % G == 0, "source group");
#pragma unroll
        for (int g0 = 0; g0 < SB; g0 += G) {
            double s[G][NS], d[G][T], r[G][T];
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const double2* p2 = reinterpret_cast<const double2*>(sm + (g0 + u) * NS);
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
            if constexpr (FENCE != 0) sched_fence(s[0][0]);
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
            d[t] = isself[t] ? 1.0 : d[t];          // keep the self pair out of the shared product
        }
        rcp_batch<T>(d, r);
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
// The branch-free table logarithm (log_group_fast) is evaluated for the whole batch first; if any
// argument of this thread was outside the table window the batch's logarithms are recomputed with
// the library log() before anything is accumulated (the one-sided kernel redoes a whole tile instead).
struct SymBveStream : LogSharedTable<32> {
    static constexpr int NS = 6, NA = 2, NC = 2;
    struct Tgt { double x, y, z, w0, w1; };
    __device__ static __forceinline__ Tgt null() { return Tgt{0.0, 0.0, 0.0, 0.0, 0.0}; }
    __device__ static __forceinline__ Tgt from_record(const double2* p2)
    {
        const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
        return Tgt{v0.x, v0.y, v1.x, v1.y, v2.x};
    }
    // ORDER bit 0: retry per source instead of per batch -- a branch after every source's logarithms, which also
    // keeps ptxas from interleaving the whole batch (see sched_fence)
    template <int T, int SB, int ORDER>
    __device__ static __forceinline__ void batch(const SymParams& p, const Tgt (&tg)[T], double (&a)[T][NA],
                                                 const double* __restrict__ sm, double (&cb)[SB][NC], const SharedCtx& sc)
    {
        if constexpr ((ORDER & 1) != 0) {
#pragma unroll
            for (int u = 0; u < SB; ++u) {
                const double2* p2 = reinterpret_cast<const double2*>(sm + u * NS);
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
            return;
        }
        double w[SB][2], l[SB][T];
        unsigned worst = 0;
#pragma unroll
        for (int u = 0; u < SB; ++u) {
            const double2* p2 = reinterpret_cast<const double2*>(sm + u * NS);
            const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
            w[u][0] = v1.y; w[u][1] = v2.x;
            double d[T];
#pragma unroll
            for (int t = 0; t < T; ++t) {
                d[t] = fma(-tg[t].x, v0.x, p.R2);
                d[t] = fma(-tg[t].y, v0.y, d[t]);
                d[t] = fma(-tg[t].z, v1.x, d[t]);
            }
            log_group_fast<KS, T>(d, l[u], worst, sc);
        }
        if (__builtin_expect(needs_retry(worst), 0)) {
#pragma unroll 1
            for (int u = 0; u < SB; ++u) {
                const double2* p2 = reinterpret_cast<const double2*>(sm + u * NS);
                const double2 v0 = p2[0], v1 = p2[1];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    double d = fma(-tg[t].x, v0.x, p.R2);
                    d = fma(-tg[t].y, v0.y, d);
                    d = fma(-tg[t].z, v1.x, d);
                    const double lv = log_slow_path(d);
#pragma unroll
                    for (int uu = 0; uu < SB; ++uu)         // static indexing keeps l in registers
                        if (uu == u) l[uu][t] = lv;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < SB; ++u) {
#pragma unroll
            for (int t = 0; t < T; ++t) {
                a[t][0] = fma(l[u][t], w[u][0], a[t][0]);
                a[t][1] = fma(l[u][t], w[u][1], a[t][1]);
            }
            cb[u][0] = l[u][0] * tg[0].w0; cb[u][1] = l[u][0] * tg[0].w1;
#pragma unroll
            for (int t = 1; t < T; ++t) {
                cb[u][0] = fma(l[u][t], tg[t].w0, cb[u][0]);
                cb[u][1] = fma(l[u][t], tg[t].w1, cb[u][1]);
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

// Planar Biot-Savart (PlaneVel in pairs.cuh): u_i -= dy s_j / r^2, v_i += dx s_j / r^2 with dx = x_i - x_j,
// dy = y_i - y_j, s_j = omega_j A_j / (2 pi); record x, y, s, 0.  The pair shares dx, dy, r^2 and 1 / r^2:
// 2 + 2 + 3 + (1 + 2) + (1 + 2) = 13 FP64 instructions for two interactions instead of 20.  The transposed
// direction sees -dx, -dy:  u_j += dy s_i / r^2,  v_j -= dx s_i / r^2.
// Null SOURCE records sit at (+1e37, +1e37) (pack_plane); a null TARGET is put at (-1e37, -1e37), so that no pair
// has r^2 = 0 -- a zero would poison the reciprocal the four targets of a thread share.
//   ORDER bit 0: a scheduling fence after every source (sched_fence)
struct SymPlaneVel : NoSharedTable {
    static constexpr int NS = 4, NA = 2, NC = 2;
    struct Tgt { double x, y, s; };
    __device__ static __forceinline__ Tgt null() { return Tgt{-LPM_PLANE_FAR, -LPM_PLANE_FAR, 0.0}; }
    __device__ static __forceinline__ Tgt from_record(const double2* p2)
    {
        const double2 v0 = p2[0], v1 = p2[1];
        return Tgt{v0.x, v0.y, v1.x};
    }
    template <int T, int SB, int ORDER>
    __device__ static __forceinline__ void batch(const SymParams&, const Tgt (&tg)[T], double (&a)[T][NA],
                                                 const double* __restrict__ sm, double (&cb)[SB][NC], const SharedCtx&)
    {
#pragma unroll
        for (int u = 0; u < SB; ++u) {
            const double2* p2 = reinterpret_cast<const double2*>(sm + u * NS);
            const double2 v0 = p2[0], v1 = p2[1];
            double dx[T], dy[T], r2[T], r[T];
#pragma unroll
            for (int t = 0; t < T; ++t) {
                dx[t] = tg[t].x - v0.x; dy[t] = tg[t].y - v0.y;
                r2[t] = fma(dx[t], dx[t], dy[t] * dy[t]);
            }
            rcp_batch<T>(r2, r);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const double w = r[t] * v1.x;
                a[t][0] = fma(-dy[t], w, a[t][0]);
                a[t][1] = fma(dx[t], w, a[t][1]);
            }
            {
                const double w = r[0] * tg[0].s;
                cb[u][0] = dy[0] * w; cb[u][1] = -dx[0] * w;
            }
#pragma unroll
            for (int t = 1; t < T; ++t) {
                const double w = r[t] * tg[t].s;
                cb[u][0] = fma(dy[t], w, cb[u][0]);
                cb[u][1] = fma(-dx[t], w, cb[u][1]);
            }
            if constexpr ((ORDER & 1) != 0) sched_fence(v0.x);
        }
    }
    // as PlaneVel::group<T, true>
    template <int T>
    __device__ static __forceinline__ void diag(const SymParams&, const Tgt (&tg)[T], double (&a)[T][NA],
                                                const double (&s)[NS], const bool (&isself)[T], const SharedCtx&)
    {
        double dx[T], dy[T], r2[T], r[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            dx[t] = tg[t].x - s[0]; dy[t] = tg[t].y - s[1];
            r2[t] = fma(dx[t], dx[t], dy[t] * dy[t]);
            r2[t] = isself[t] ? 1.0 : r2[t];
        }
        rcp_batch<T>(r2, r);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            double w = r[t] * s[2];
            w = isself[t] ? 0.0 : w;
            a[t][0] = fma(-dy[t], w, a[t][0]);
            a[t][1] = fma(dx[t], w, a[t][1]);
        }
    }
};

// Beta-plane Biot-Savart (BetaVel in pairs.cuh); record sinh(pi y), cosh(pi y), sin(pi x), cos(pi x), zeta A / 2, 0.
// With S = sinh(pi dy), C = cosh(pi dy), s = sin(pi dx), c = cos(pi dx) from the addition formulas:
// u_i -= S C w_j / (S^2 + s^2),  v_i += s c w_j / (S^2 + s^2).  Swapping the pair flips the signs of S and s only, so
// the eight addition-formula operations, the denominator, S C, s c and the reciprocal (15 of 18 FP64 instructions)
// serve both directions:  u_j += S C w_i / (..),  v_j -= s c w_i / (..):  21 instructions for two interactions.
// A null TARGET is (-1e10, 1e10, 0, 1) against the null source record's (+1e10, 1e10, 0, 1): no pair has S = s = 0.
//   ORDER bit 0: a scheduling fence after every source (sched_fence)
struct SymBetaVel : NoSharedTable {
    static constexpr int NS = 6, NA = 2, NC = 2;
    struct Tgt { double sh, ch, sn, cs, w; };
    __device__ static __forceinline__ Tgt null() { return Tgt{-1.0e10, 1.0e10, 0.0, 1.0, 0.0}; }
    __device__ static __forceinline__ Tgt from_record(const double2* p2)
    {
        const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
        return Tgt{v0.x, v0.y, v1.x, v1.y, v2.x};
    }
    template <int T>
    __device__ static __forceinline__ void pair_terms(const Tgt (&tg)[T], const double (&s)[NS], double (&SC)[T],
                                                      double (&sc)[T], double (&den)[T])
    {
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const double S = fma(tg[t].sh, s[1], -(tg[t].ch * s[0]));
            const double C = fma(tg[t].ch, s[1], -(tg[t].sh * s[0]));
            const double sn = fma(tg[t].sn, s[3], -(tg[t].cs * s[2]));
            const double cs = fma(tg[t].cs, s[3], tg[t].sn * s[2]);
            den[t] = fma(S, S, sn * sn);
            SC[t] = S * C; sc[t] = sn * cs;
        }
    }
    template <int T, int SB, int ORDER>
    __device__ static __forceinline__ void batch(const SymParams&, const Tgt (&tg)[T], double (&a)[T][NA],
                                                 const double* __restrict__ sm, double (&cb)[SB][NC], const SharedCtx&)
    {
#pragma unroll
        for (int u = 0; u < SB; ++u) {
            double s[NS], SC[T], sc[T], den[T], r[T];
            const double2* p2 = reinterpret_cast<const double2*>(sm + u * NS);
#pragma unroll
            for (int q = 0; q < NS / 2; ++q) {
                const double2 v = p2[q];
                s[2 * q] = v.x; s[2 * q + 1] = v.y;
            }
            pair_terms<T>(tg, s, SC, sc, den);
            rcp_batch<T>(den, r);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const double w = r[t] * s[4];
                a[t][0] = fma(-SC[t], w, a[t][0]);
                a[t][1] = fma(sc[t], w, a[t][1]);
            }
            {
                const double w = r[0] * tg[0].w;
                cb[u][0] = SC[0] * w; cb[u][1] = -sc[0] * w;
            }
#pragma unroll
            for (int t = 1; t < T; ++t) {
                const double w = r[t] * tg[t].w;
                cb[u][0] = fma(SC[t], w, cb[u][0]);
                cb[u][1] = fma(-sc[t], w, cb[u][1]);
            }
            if constexpr ((ORDER & 1) != 0) sched_fence(s[1]);
        }
    }
    // as BetaVel::group<T, true>
    template <int T>
    __device__ static __forceinline__ void diag(const SymParams&, const Tgt (&tg)[T], double (&a)[T][NA],
                                                const double (&s)[NS], const bool (&isself)[T], const SharedCtx&)
    {
        double SC[T], sc[T], den[T], r[T];
        pair_terms<T>(tg, s, SC, sc, den);
#pragma unroll
        for (int t = 0; t < T; ++t) den[t] = isself[t] ? 1.0 : den[t];
        rcp_batch<T>(den, r);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            double w = r[t] * s[4];
            w = isself[t] ? 0.0 : w;
            a[t][0] = fma(-SC[t], w, a[t][0]);
            a[t][1] = fma(sc[t], w, a[t][1]);
        }
    }
};

// Planar and beta-plane stream functions (PlaneStream, BetaStream in pairs.cuh): psi_i = sum_j w_j ln(arg_ij) with a
// symmetric argument (r^2, resp. 2 (S^2 + s^2)), so the argument and its logarithm serve both directions; as
// SymBveStream with the retry branch per source.  G supplies the geometry:
//   NS, NW            doubles per record, weights per particle (1 or 2)
//   Tgt, null(), from_record()
//   arg(tgt, s)       the logarithm's argument for one pair;  tw(tgt, k), sw(s, k)  the k-th weight of target / source
template <class G>
struct SymLogStream : LogSharedTable<32> {
    static constexpr int NS = G::NS, NA = G::NW, NC = G::NW;
    using Tgt = typename G::Tgt;
    __device__ static __forceinline__ Tgt null() { return G::null(); }
    __device__ static __forceinline__ Tgt from_record(const double2* p2) { return G::from_record(p2); }
    template <int T, int SB, int ORDER>
    __device__ static __forceinline__ void batch(const SymParams&, const Tgt (&tg)[T], double (&a)[T][NA],
                                                 const double* __restrict__ sm, double (&cb)[SB][NC], const SharedCtx& sc)
    {
#pragma unroll
        for (int u = 0; u < SB; ++u) {
            double s[NS], d[T], l[T];
            const double2* p2 = reinterpret_cast<const double2*>(sm + u * NS);
#pragma unroll
            for (int q = 0; q < NS / 2; ++q) {
                const double2 v = p2[q];
                s[2 * q] = v.x; s[2 * q + 1] = v.y;
            }
            unsigned worst = 0;
#pragma unroll
            for (int t = 0; t < T; ++t) d[t] = G::arg(tg[t], s);
            log_group_fast<KS, T>(d, l, worst, sc);
            if (__builtin_expect(needs_retry(worst), 0)) {
#pragma unroll
                for (int t = 0; t < T; ++t) l[t] = log_slow_path(d[t]);
            }
#pragma unroll
            for (int k = 0; k < NA; ++k) {
#pragma unroll
                for (int t = 0; t < T; ++t) a[t][k] = fma(l[t], G::sw(s, k), a[t][k]);
                cb[u][k] = l[0] * G::tw(tg[0], k);
#pragma unroll
                for (int t = 1; t < T; ++t) cb[u][k] = fma(l[t], G::tw(tg[t], k), cb[u][k]);
            }
        }
    }
    template <int T>
    __device__ static __forceinline__ void diag(const SymParams&, const Tgt (&tg)[T], double (&a)[T][NA],
                                                const double (&s)[NS], const bool (&isself)[T], const SharedCtx& sc)
    {
        double d[T], l[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            d[t] = G::arg(tg[t], s);
            d[t] = isself[t] ? 1.0 : d[t];          // any positive value; the pair is zeroed below
        }
        log_group<KS, T>(d, l, sc);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            l[t] = isself[t] ? 0.0 : l[t];
#pragma unroll
            for (int k = 0; k < NA; ++k) a[t][k] = fma(l[t], G::sw(s, k), a[t][k]);
        }
    }
};
// record x, y, omega A / (4 pi), 0;  arg = r^2
struct PlaneStreamGeom {
    static constexpr int NS = 4, NW = 1;
    struct Tgt { double x, y, w; };
    __device__ static __forceinline__ Tgt null() { return Tgt{-LPM_PLANE_FAR, -LPM_PLANE_FAR, 0.0}; }
    __device__ static __forceinline__ Tgt from_record(const double2* p2)
    {
        const double2 v0 = p2[0], v1 = p2[1];
        return Tgt{v0.x, v0.y, v1.x};
    }
    __device__ static __forceinline__ double arg(const Tgt& t, const double (&s)[NS])
    {
        const double dx = t.x - s[0], dy = t.y - s[1];
        return fma(dx, dx, dy * dy);
    }
    __device__ static __forceinline__ double tw(const Tgt& t, int) { return t.w; }
    __device__ static __forceinline__ double sw(const double (&s)[NS], int) { return s[2]; }
};
// record sinh(pi y), cosh(pi y), sin(pi x), cos(pi x), zeta A / (4 pi), omega A / (4 pi);  arg = 2 (S^2 + s^2)
struct BetaStreamGeom {
    static constexpr int NS = 6, NW = 2;
    struct Tgt { double sh, ch, sn, cs, w0, w1; };
    __device__ static __forceinline__ Tgt null() { return Tgt{-1.0e10, 1.0e10, 0.0, 1.0, 0.0, 0.0}; }
    __device__ static __forceinline__ Tgt from_record(const double2* p2)
    {
        const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
        return Tgt{v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
    }
    __device__ static __forceinline__ double arg(const Tgt& t, const double (&s)[NS])
    {
        const double S = fma(t.sh, s[1], -(t.ch * s[0]));
        const double sn = fma(t.sn, s[3], -(t.cs * s[2]));
        return 2.0 * fma(S, S, sn * sn);
    }
    __device__ static __forceinline__ double tw(const Tgt& t, int k) { return k == 0 ? t.w0 : t.w1; }
    __device__ static __forceinline__ double sw(const double (&s)[NS], int k) { return s[4 + k]; }
};
using SymPlaneStream = SymLogStream<PlaneStreamGeom>;
using SymBetaStream = SymLogStream<BetaStreamGeom>;

// ---- the kernel ---------------------------------------------------------------
// acc: [nsrc_pad][NC] doubles -- or, with prm.fx, [nsrc_pad][NC][kFxLimbs] 64-bit limbs -- zeroed by the caller
// (NA == NC: both directions feed the same sums).
// dynamic shared memory: [2 tiles][K::KS table][2 mbarriers] and, with COMBINE, [2][warps][TS][NC] per-warp source sums:
// the warps' sums for a tile are added in warp order after the tile and ONE RED per (CTA, source, component) is
// issued instead of one per warp (a quarter of the atomics with 128 threads)
template <class K, int T, int BLOCK, int SB, int MINB, int ORDER = 0, bool FX = false, bool COMBINE = false>
__global__ void __launch_bounds__(BLOCK, MINB)
sym_kernel(const SymParams prm, const SymGeom g, const double* __restrict__ src, double* __restrict__ acc)
{
    constexpr int NS = K::NS, NA = K::NA, NC = K::NC, TS = kTile, TB = BLOCK * T, DT = TB / TS;
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
    const int I = blockIdx.x % g.nblocks;           // chunk is the slow index, as in ds_kernel
    const int ck = blockIdx.x / g.nblocks;
    if (g.world > 1 && (I % g.world) != g.rank) return;
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
            K::template batch<T, SB, ORDER>(prm, tg, a, sm + jb * NS, cb, sctx);
            sym_reduce_red<SB, NC, FX, COMBINE>(cb, lane, acc, ((size_t)k * TS + jb) * NC, prm.fx, mine + jb * NC);
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
            sym_red_add<FX>(acc, (size_t)k * TS * NC + idx, sum, prm.fx);
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
            for (int q = 0; q < NA; ++q) sym_red_add<FX>(acc, (size_t)cidx[t] * NA + q, a[t][q], prm.fx);
        }
}

template <class K, int T, int BLOCK, bool COMBINE = false>
constexpr size_t sym_smem_bytes()
{
    return 2 * size_t(kTile) * K::NS * sizeof(double) + sizeof(double) * K::KS + 2 * sizeof(uint64_t) +
           (COMBINE ? 2 * size_t(BLOCK / 32) * kTile * K::NC * sizeof(double) : 0);
}

// Window of the fixed-point accumulators for one evaluation.  maxhi: high word of the largest |entry| of the source
// records (coordinates and strengths; absmax_hi_kernel).  mode 0 (velocity): |sum| <= F M / d_min with
// d_min >= R^2 2^-110 (two FP64 points cannot be closer); mode 1 (stream functions): |ln| < 2^10.  The top of the
// window sits there, E0 240 bits below.  fx[0] = 2^-E0, fx[1] = 2^E0.
__global__ void sym_fx_scale_kernel(int mode, double R2, int32_t nsrc, const int32_t* __restrict__ maxhi, double* __restrict__ fx)
{
    const int eM = ((*maxhi >> 20) & 0x7ff) - 1023 + 1;
    int eF = 1;
    while ((1 << eF) < nsrc && eF < 31) ++eF;
    const int eR = ((__double2hiint(R2) >> 20) & 0x7ff) - 1023;
    int top = eM + eF + (mode == 0 ? 110 - eR : 12);
    int e0 = top - 40 * kFxLimbs;
    if (e0 > 700) e0 = 700;
    if (e0 < -900) e0 = -900;
    fx[0] = __hiloint2double((1023 - e0) << 20, 0);
    fx[1] = __hiloint2double((1023 + e0) << 20, 0);
}

// limbs -> doubles: carries first (so that every limb but the top one is below 2^40), then the sum from the top
__global__ void __launch_bounds__(256)
sym_fx_to_double_kernel(int64_t nvalues, const long long* __restrict__ limbs, const double* __restrict__ fx,
                        double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvalues) return;
    long long l[kFxLimbs];
#pragma unroll
    for (int k = 0; k < kFxLimbs; ++k) l[k] = limbs[i * kFxLimbs + k];
#pragma unroll
    for (int k = 0; k + 1 < kFxLimbs; ++k) {
        const long long carry = l[k] >> 40;         // arithmetic shift: floor
        l[k] -= carry << 40;
        l[k + 1] += carry;
    }
    constexpr double unit[kFxLimbs] = {0x1p0, 0x1p40, 0x1p80, 0x1p120, 0x1p160, 0x1p200};
    double s = 0.0;
#pragma unroll
    for (int k = kFxLimbs - 1; k >= 0; --k) s = fma((double)l[k], unit[k], s);
    out[i] = s * fx[1];
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

// two sums per active particle, copied out (BveStream::finalize, PlaneVel::finalize)
__global__ void __launch_bounds__(256)
sym_stream_finalize(int32_t nsrc, const int32_t* __restrict__ active, const double* __restrict__ acc, Outs<2> out)
{
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc) return;
    const int64_t i = active[c];
    out.store(0, i, acc[(size_t)c * 2]);
    out.store(1, i, acc[(size_t)c * 2 + 1]);
}

// one sum per active particle, copied out (PlaneStream::finalize)
__global__ void __launch_bounds__(256)
sym_copy1_finalize(int32_t nsrc, const int32_t* __restrict__ active, const double* __restrict__ acc, Outs<1> out)
{
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc) return;
    out.store(0, active[c], acc[c]);
}

// passive[i - scan[i]] = i for every particle with mask 0 (stable, like the active list)
__global__ void __launch_bounds__(256)
passive_list_kernel(int64_t n, const int32_t* __restrict__ scan, int32_t* __restrict__ passive)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && scan[i + 1] == scan[i]) passive[i - scan[i]] = (int32_t)i;
}

}  // namespace lpm
