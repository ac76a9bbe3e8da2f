// sym_kernels.cuh -- device code of the EXPERIMENTAL pair-symmetric BVE velocity path (see symmetric.cuh
// for the design and the host side).  Kept apart so that tools/sym_score.py can compile the kernel alone.
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
    double R2;
};

// Warp reduction of cb[s][a] (thread-local sums for SB sources, 3 components) by recursive
// halving, then one RED per (source, component) from the lane that ends up owning it.
// After the halving levels lane l holds source ((l >> (5 - LV)) & (SB - 1)) summed over the lanes
// that differ from it in the high LV bits; a butterfly over the remaining low bits finishes the sum.
template <int SB>
__device__ __forceinline__ void sym_reduce_red(double (&cb)[SB][3], int lane, double* __restrict__ accj)
{
    static_assert(SB == 8 || SB == 4, "source batch");
    constexpr unsigned FULL = 0xffffffffu;
    double v[3];
    int sidx;
    if constexpr (SB == 8) {
        const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
        double v4[4][3], v2[2][3];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double lo = cb[k][a], hi = cb[k + 4][a];
                const double keep = b4 ? hi : lo, send = b4 ? lo : hi;
                v4[k][a] = keep + __shfl_xor_sync(FULL, send, 16);
            }
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double lo = v4[k][a], hi = v4[k + 2][a];
                const double keep = b3 ? hi : lo, send = b3 ? lo : hi;
                v2[k][a] = keep + __shfl_xor_sync(FULL, send, 8);
            }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double lo = v2[0][a], hi = v2[1][a];
            const double keep = b2 ? hi : lo, send = b2 ? lo : hi;
            v[a] = keep + __shfl_xor_sync(FULL, send, 4);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            v[a] += __shfl_xor_sync(FULL, v[a], 2);
            v[a] += __shfl_xor_sync(FULL, v[a], 1);
        }
        sidx = (lane >> 2) & 7;
        const int q = lane & 3;
        if (q < 3) atomicAdd(accj + sidx * 3 + q, q == 0 ? v[0] : (q == 1 ? v[1] : v[2]));
    } else {
        const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
        double v2[2][3];
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double lo = cb[k][a], hi = cb[k + 2][a];
                const double keep = b4 ? hi : lo, send = b4 ? lo : hi;
                v2[k][a] = keep + __shfl_xor_sync(FULL, send, 16);
            }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double lo = v2[0][a], hi = v2[1][a];
            const double keep = b3 ? hi : lo, send = b3 ? lo : hi;
            v[a] = keep + __shfl_xor_sync(FULL, send, 8);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            v[a] += __shfl_xor_sync(FULL, v[a], 4);
            v[a] += __shfl_xor_sync(FULL, v[a], 2);
            v[a] += __shfl_xor_sync(FULL, v[a], 1);
        }
        sidx = (lane >> 3) & 3;
        const int q = lane & 7;
        if (q < 3) atomicAdd(accj + sidx * 3 + q, q == 0 ? v[0] : (q == 1 ? v[1] : v[2]));
    }
}

// SB sources of a tile above the diagonal against the thread's T targets, both directions:
// a[t] += P_j / d (the targets' sums) and cb[u] = sum_t P_t / d (this thread's share of source u's sum).
// ORDER permutes INDEPENDENT statements only (as BveVelT's ORDER does): the kernel is bound by register
// operand delivery, and what ptxas allocates and where it can set .reuse follows the statement order.
//   bit 0     denominators coordinate by coordinate (else target by target)
//   bit 1     a-phase component by component (else target by target)
//   bits 2-3  sources handled together, phase by phase: 1, 2, 4, SB
//   bit 4     cb-phase nest (target, component, source) -- P_t stays in the reuse cache -- else
//             (source, target, component) -- 1/d stays
template <int T, int SB, int ORDER>
__device__ __forceinline__ void sym_batch(const double (&tx)[T], const double (&ty)[T], const double (&tz)[T],
                                          const double (&px)[T], const double (&py)[T], const double (&pz)[T],
                                          double (&a)[T][3], const double* __restrict__ sm, double R2, double (&cb)[SB][3])
{
    constexpr int NS = 6;
    constexpr int DN = ORDER & 1, AN = (ORDER >> 1) & 1, GS = (ORDER >> 2) & 3, CU = (ORDER >> 4) & 1;
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
                    d[u][t] = fma(-tx[t], s[u][0], R2);
                    d[u][t] = fma(-ty[t], s[u][1], d[u][t]);
                    d[u][t] = fma(-tz[t], s[u][2], d[u][t]);
                }
        } else {
#pragma unroll
            for (int u = 0; u < G; ++u)
#pragma unroll
                for (int t = 0; t < T; ++t) d[u][t] = fma(-tx[t], s[u][0], R2);
#pragma unroll
            for (int u = 0; u < G; ++u)
#pragma unroll
                for (int t = 0; t < T; ++t) d[u][t] = fma(-ty[t], s[u][1], d[u][t]);
#pragma unroll
            for (int u = 0; u < G; ++u)
#pragma unroll
                for (int t = 0; t < T; ++t) d[u][t] = fma(-tz[t], s[u][2], d[u][t]);
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
                cb[g0 + u][0] = r[u][0] * px[0]; cb[g0 + u][1] = r[u][0] * py[0]; cb[g0 + u][2] = r[u][0] * pz[0];
#pragma unroll
                for (int t = 1; t < T; ++t) {
                    cb[g0 + u][0] = fma(r[u][t], px[t], cb[g0 + u][0]);
                    cb[g0 + u][1] = fma(r[u][t], py[t], cb[g0 + u][1]);
                    cb[g0 + u][2] = fma(r[u][t], pz[t], cb[g0 + u][2]);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < G; ++u) cb[g0 + u][0] = r[u][0] * px[0];
#pragma unroll
            for (int u = 0; u < G; ++u) cb[g0 + u][1] = r[u][0] * py[0];
#pragma unroll
            for (int u = 0; u < G; ++u) cb[g0 + u][2] = r[u][0] * pz[0];
#pragma unroll
            for (int t = 1; t < T; ++t) {
#pragma unroll
                for (int u = 0; u < G; ++u) cb[g0 + u][0] = fma(r[u][t], px[t], cb[g0 + u][0]);
#pragma unroll
                for (int u = 0; u < G; ++u) cb[g0 + u][1] = fma(r[u][t], py[t], cb[g0 + u][1]);
#pragma unroll
                for (int u = 0; u < G; ++u) cb[g0 + u][2] = fma(r[u][t], pz[t], cb[g0 + u][2]);
            }
        }
    }
}

// acc: [nsrc_pad][3] doubles, zeroed by the caller.
template <int T, int BLOCK, int SB, int MINB, int ORDER = 0>
__global__ void __launch_bounds__(BLOCK, MINB)
sym_bve_kernel(const SymGeom g, const double* __restrict__ src, double* __restrict__ acc)
{
    constexpr int NS = 6, TS = kTile, TB = BLOCK * T, DT = TB / TS;
    static_assert(TB % TS == 0, "a target block must be whole source tiles");
    static_assert(TS % SB == 0, "source batch");
    constexpr uint32_t kTileBytes = TS * NS * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double(*tile)[TS * NS] = reinterpret_cast<double(*)[TS * NS]>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 2 * kTileBytes);

    const int tid = threadIdx.x, lane = tid & 31;
    const int I = blockIdx.x % g.nblocks;           // chunk is the slow index, as in ds_kernel
    const int ck = blockIdx.x / g.nblocks;
    if (g.world > 1 && (I % g.world) != g.rank) return;
    const int kdiag = I * DT;                        // first of this block's DT diagonal tiles
    int k0 = ck * g.chunk_tiles;
    const int k1 = min(k0 + g.chunk_tiles, g.ntiles);
    if (k0 < kdiag) k0 = kdiag;
    if (k0 >= k1) return;                            // chunk entirely below the diagonal (whole CTA)

    double tx[T], ty[T], tz[T], px[T], py[T], pz[T], a[T][3];
    int32_t cidx[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int32_t c = I * TB + t * BLOCK + tid;
        cidx[t] = c;
        tx[t] = ty[t] = tz[t] = px[t] = py[t] = pz[t] = 0.0;     // past the padded list: a null particle
        if (c < g.nsrc_pad) {
            const double2* p2 = reinterpret_cast<const double2*>(src + (size_t)c * NS);
            const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2];
            tx[t] = v0.x; ty[t] = v0.y; tz[t] = v1.x; px[t] = v1.y; py[t] = v2.x; pz[t] = v2.y;
        }
        a[t][0] = a[t][1] = a[t][2] = 0.0;
    }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto load_tile = [&](const int k, const int st) {      // thread 0 only
        mbar_expect_tx(&full[st], kTileBytes);
        tma_bulk_g2s(tile[st], src + (size_t)k * TS * NS, kTileBytes, &full[st]);
    };
    auto load_source = [&](const double* sm, int j, double (&s)[NS]) {
        const double2* p2 = reinterpret_cast<const double2*>(sm + j * NS);
#pragma unroll
        for (int q = 0; q < NS / 2; ++q) {
            const double2 v = p2[q];
            s[2 * q] = v.x; s[2 * q + 1] = v.y;
        }
    };
    // the block against itself: one-sided, self pair excluded (as BveVelT::group<T, true>)
    auto diag_tile = [&](const int k, const int st) {
        const double* sm = tile[st];
        const int32_t j0 = k * TS;
#pragma unroll 1
        for (int j = 0; j < TS; ++j) {
            double s[NS], d[T], r[T];
            load_source(sm, j, s);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                d[t] = fma(-tx[t], s[0], g.R2);
                d[t] = fma(-ty[t], s[1], d[t]);
                d[t] = fma(-tz[t], s[2], d[t]);
                d[t] = (j0 + j == cidx[t]) ? 1.0 : d[t];
            }
            rcp_batch<T>(d, r);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                r[t] = (j0 + j == cidx[t]) ? 0.0 : r[t];
                a[t][0] = fma(r[t], s[3], a[t][0]);
                a[t][1] = fma(r[t], s[4], a[t][1]);
                a[t][2] = fma(r[t], s[5], a[t][2]);
            }
        }
    };
    // a tile above the diagonal: every pair once, both directions
    auto sym_tile = [&](const int k, const int st) {
        const double* sm = tile[st];
        double* accj = acc + (size_t)k * TS * 3;
#pragma unroll 1
        for (int jb = 0; jb < TS; jb += SB) {
            double cb[SB][3];
            sym_batch<T, SB, ORDER>(tx, ty, tz, px, py, pz, a, sm + jb * NS, g.R2, cb);
            sym_reduce_red<SB>(cb, lane, accj + jb * 3);
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
        else sym_tile(k, st);
        __syncthreads();        // everyone is done with tile[st]
        if (tid == 0 && it + 2 < nt) load_tile(k + 2, st);
    }
    // this CTA's own sums join the accumulators
#pragma unroll
    for (int t = 0; t < T; ++t)
        if (cidx[t] < g.nsrc) {
            atomicAdd(acc + (size_t)cidx[t] * 3 + 0, a[t][0]);
            atomicAdd(acc + (size_t)cidx[t] * 3 + 1, a[t][1]);
            atomicAdd(acc + (size_t)cidx[t] * 3 + 2, a[t][2]);
        }
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

// passive[i - scan[i]] = i for every particle with mask 0 (stable, like the active list)
__global__ void __launch_bounds__(256)
passive_list_kernel(int64_t n, const int32_t* __restrict__ scan, int32_t* __restrict__ passive)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && scan[i + 1] == scan[i]) passive[i - scan[i]] = (int32_t)i;
}

}  // namespace lpm
