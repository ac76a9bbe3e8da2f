// directsum.cuh -- the tiled O(N*F) direct-sum engine for sm_100a.
//
// One kernel template drives every pair kernel of the hot path (BVE velocity,
// planar / beta-plane Biot-Savart, PSE Laplacians, stream functions).  The
// reference evaluates, for each target i of a rank's slice, the masked loop
// over j /= i (e.g. src/SphereBVESolver.f90:396-420).  Here:
//
//   * sources are the compacted active particles (stable order == Fortran
//     pack(), see scan.cuh) stored AoS, NS doubles per source, padded with
//     null sources to a whole number of tiles;
//   * work item = (target block, source chunk).  A CTA holds BLOCK*T targets
//     in registers (T per thread, strided by BLOCK so loads coalesce) and
//     streams the chunk's sources through shared memory in TS-source tiles
//     with a two-stage TMA bulk-copy (cp.async.bulk + mbarrier) pipeline;
//     every lane reads the same source (LDS.128 broadcast);
//   * the self interaction j == i is excluded only in the (at most
//     BLOCK*T/TS + 1) tiles that overlap the CTA's own compact index range --
//     all other tiles run the unchecked loop;
//   * chunk partial sums go to scratch and are added in chunk order by the
//     finalize kernel, so a target's result does not depend on the launch
//     geometry, the slice it belongs to, or the number of GPUs.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lpm {

// Upper bound on source chunks per evaluation (lpm_set_max_chunks).  64 rather than 16: at icosTri 8
// on 8 GPUs a rank's slice is only ~9 waves of CTAs with 16 chunks and the tail costs 3 %
// (184.1 -> 178.7 ms per evaluation); the price is partial-sum traffic (3 GB at level 8, < 1 ms).
constexpr int kMaxChunksDefault = 64;
inline int& max_chunks_ref()
{
    static int v = kMaxChunksDefault;
    return v;
}
// Do not split the source list finer than this.  1024 rather than round 1's 8192: a rank of an 8-GPU run at icosTri 6
// launches only 30 target blocks, and with 10 chunks of 8192 sources its 300 CTAs filled less than one wave (kernel at
// 0.54 of its FP64 bound); 64 chunks of 1280 sources with 4 targets per thread reach 0.84 (profiles/r02d_slice_sweep_L6.log).
// From icosTri 8 upwards the cap of 64 chunks was binding already, so nothing changes there.
constexpr int kChunkMinDefault = 1024;
inline int& chunk_min_ref()
{
    static int v = kChunkMinDefault;
    return v;
}
constexpr int kTile = 256;          // sources per shared-memory tile (all kernels)
constexpr int kMaxNeedWords = 256;  // tile culling: bitmap words per CTA (<= 8192 tiles per chunk)

struct DsGeom {
    int64_t tbeg, tend;     // target range, global particle indices [tbeg, tend)
    int32_t nsrc;           // active sources
    int32_t nsrc_pad;       // padded to a multiple of kTile
    int32_t chunk;          // sources per chunk (multiple of kTile)
    int32_t nchunks;
    int32_t ntblocks;       // target blocks = ceil((tend-tblk0) / (BLOCK*T))
    int64_t ntgt;           // tend - tbeg
    int64_t tblk0;          // tbeg rounded down to a multiple of BLOCK*T
    int64_t nall;           // all particles (launch shape depends on this, not on the slice)
    const double* bounds;   // tile culling: (cx, cy, cz, radius) per source tile, or nullptr
    int32_t half_bin;       // 0x800 (half a log-table bin in the high word), passed at run time
};

// Source chunking depends on the number of active sources only.
// `nall` (all targets of the evaluation, NOT the slice) caps the partial-sum scratch for very large runs.
inline void ds_chunks(int64_t nsrc, int32_t* nsrc_pad, int32_t* chunk, int32_t* nchunks, int64_t nall = 0)
{
    int64_t pad = (nsrc + kTile - 1) / kTile * kTile;
    if (pad == 0) pad = kTile;
    const int64_t cmin = chunk_min_ref();
    int64_t nc = (pad + cmin - 1) / cmin;
    int cap = max_chunks_ref();
    if (nall > (16LL << 20)) cap = cap < 16 ? cap : 16;
    else if (nall > (8LL << 20)) cap = cap < 32 ? cap : 32;
    if (nc > cap) nc = cap;
    if (nc < 1) nc = 1;
    int64_t ch = (pad + nc - 1) / nc;
    ch = (ch + kTile - 1) / kTile * kTile;
    nc = (pad + ch - 1) / ch;
    *nsrc_pad = (int32_t)pad; *chunk = (int32_t)ch; *nchunks = (int32_t)nc;
}

// ---- mbarrier / TMA bulk copy / MUFU (PTX) ------------------------------------
// LPM_CUDA_EMU is defined only by tests/cuda_emu, which compiles these kernel sources with g++ to
// run them on CPU threads (a test of the kernels' logic, never part of liblpmgpu.so); it supplies
// its own versions of the few primitives that are inline PTX here.
#ifdef LPM_CUDA_EMU
#include "emu_primitives.h"
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// MUFU.RCP64H: a >= 20-bit reciprocal whose low word is zero
__device__ __forceinline__ double rcp_approx_f64(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
}
// (hi & 0xfffff000) | half as ONE LOP3 (pairs.cuh, log_bin_rcp)
__device__ __forceinline__ int lop3_and_or(int hi, int half)
{
    int r;
    asm("lop3.b32 %0, %1, 0xfffff000, %2, 0xEA;" : "=r"(r) : "r"(hi), "r"(half));
    return r;
}
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    // Bounded: a lost copy traps instead of hanging the device.
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 26); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

// A never-taken branch on loaded data: it ends the basic block, so ptxas schedules the sources (or source groups)
// one after the other instead of interleaving them, which loses the operand reuse between neighbouring
// instructions of a phase (the kernels are bound by register operand delivery).  The payload is a signalling-NaN
// pattern no coordinate carries.  Used by the symmetric velocity kernel (sym_kernels.cuh), where it measured
// -5 %; in the one-sided kernels it measured +1.5 % (velocity) / -2.5 % (stream functions) and was dropped.
__device__ __forceinline__ void sched_fence(double v)
{
    if (__builtin_expect(__double2hiint(v) == 0x7ff4dead, 0)) __trap();
}

// What ds_kernel hands to group(): the functor's per-CTA shared table and one integer of
// context that init_shared() returns (the log kernels' window origin).
struct SharedCtx {
    const double* ks;
    int32_t i0;
    int32_t i1;     // = 0x800, but opaque to the compiler (see log_bin_rcp in pairs.cuh)
};

// ---- the kernel ---------------------------------------------------------------
//
// K (pair functor) provides:
//   NS, NA            doubles per source record / accumulators per target
//   SKIP_SELF         exclude j == i (velocity, stream fn) or not (PSE)
//   Params            kernel-wide constants + target array pointers
//   Tgt               per-target registers
//   load_target(p,i)  -> Tgt
//   group<T,CHECK>(p, tgt[T], s[NS], acc[T][NA], j, self[T], sctx)   one source against the
//                     thread's T targets; with CHECK, target t skips the pair when j == self[t]
//   KS, init_shared(ks, p, tid, nthreads) -> int   optional per-CTA shared table (KS doubles)
//                     and one integer of context, handed to group() as SharedCtx
//   CULL, cull_dist(p), tgt_point(tgt, xyz)   tile culling for compactly supported kernels
//   finalize(p, tgt, acc, i)   writes the outputs of target i
//
// partial layout: [(chunk*NA + a) * ntgt + local_target]
template <class K, int T, int BLOCK, int U, int MINB = 1>
__global__ void __launch_bounds__(BLOCK, MINB)
ds_kernel(const typename K::Params prm, const DsGeom g, const double* __restrict__ src,
          const int32_t* __restrict__ scan, double* __restrict__ partial)
{
    constexpr int NS = K::NS, NA = K::NA, TS = kTile;
    constexpr uint32_t kTileBytes = TS * NS * sizeof(double);
    // dynamic shared memory: [2 tiles][BLOCK*T*NA running sums][K::KS table][2 mbarriers]
    // and, for culling kernels, [kMaxNeedWords bitmap][4 * BLOCK/32 reduction scratch]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double(*tile)[TS * NS] = reinterpret_cast<double(*)[TS * NS]>(smem_raw);
    double* run = reinterpret_cast<double*>(smem_raw + 2 * kTileBytes);
    double* ks = run + BLOCK * T * NA;              // kernel-specific shared data (K::KS doubles)
    uint64_t* full = reinterpret_cast<uint64_t*>(ks + K::KS);

    const int tid = threadIdx.x;
    const int tb = blockIdx.x % g.ntblocks;         // chunk is the slow index: CTAs that run
    const int ck = blockIdx.x / g.ntblocks;         // together read the same sources (L2)
    // Target blocks are aligned to GLOBAL particle indices (tblk0 is a multiple of BLOCK*T)
    // and a CTA always evaluates its whole block, even the part outside [tbeg, tend): every
    // particle is resident on every GPU, so the targets that share a thread -- and hence a
    // batched reciprocal -- are the same whatever the partition.  Only stores are sliced.
    const int64_t blk0 = g.tblk0 + (int64_t)tb * (BLOCK * T);
    const int64_t blk1 = (blk0 + BLOCK * T < g.nall) ? blk0 + BLOCK * T : g.nall;

    typename K::Tgt tg[T];
    double acc[T][NA];
    int32_t self[T];
    bool live[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        int64_t i = blk0 + t * BLOCK + tid;
        live[t] = i >= g.tbeg && i < g.tend;
        if (i >= g.nall) i = g.nall - 1;            // past the last particle: shadow it, never stored
        tg[t] = K::load_target(prm, i);
        if (K::SKIP_SELF) {
            int32_t a = scan[i], b = scan[i + 1];
            self[t] = (b > a) ? a : -1;
        } else {
            self[t] = -1;
        }
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            acc[t][a] = 0.0;
            run[(t * NA + a) * BLOCK + tid] = 0.0;
        }
    }
    // compact indices of this CTA's own targets: [selflo, selfhi)
    int32_t selflo = 0, selfhi = 0;
    if (K::SKIP_SELF) {
        selflo = scan[blk0];
        selfhi = scan[blk1];
    }

    const int32_t s0 = ck * g.chunk;
    int32_t s1 = s0 + g.chunk;
    if (s1 > g.nsrc_pad) s1 = g.nsrc_pad;
    const int ntiles = (s1 - s0) / TS;

    const SharedCtx sctx{ks, K::init_shared(ks, prm, tid, BLOCK), g.half_bin};
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }

    // ---- tile culling (compactly supported kernels only) ------------------------------
    // A tile is visited only if its bounding ball can come within K::cull_dist of the ball
    // around this CTA's targets; `need` is a bitmap over the chunk's tiles.  Skipped pairs
    // are exactly the ones the per-pair cut-off test rejects (pairs.cuh, kPseCut), so results
    // are unchanged bit for bit.
    uint32_t* need = reinterpret_cast<uint32_t*>(full + 2);
    bool cull = false;
    if constexpr (K::CULL) {
        const double D = K::cull_dist(prm);
        cull = g.bounds != nullptr && D < 1.0e300 && ntiles <= 32 * kMaxNeedWords;
        if (cull) {
            double* red = reinterpret_cast<double*>(need + kMaxNeedWords);       // [4][BLOCK/32]
            const int lane = tid & 31, wid = tid >> 5;
            constexpr int NW = BLOCK / 32;
            double px[T], py[T], pz[T], sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                double p[3];
                K::tgt_point(tg[t], p);
                px[t] = p[0]; py[t] = p[1]; pz[t] = p[2];
                sx += p[0]; sy += p[1]; sz += p[2];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                sz += __shfl_xor_sync(0xffffffffu, sz, o);
            }
            if (lane == 0) { red[wid] = sx; red[NW + wid] = sy; red[2 * NW + wid] = sz; }
            __syncthreads();
            double cx = 0.0, cy = 0.0, cz = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { cx += red[w]; cy += red[NW + w]; cz += red[2 * NW + w]; }
            cx /= BLOCK * T; cy /= BLOCK * T; cz /= BLOCK * T;
            double r2 = 0.0;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const double dx = px[t] - cx, dy = py[t] - cy, dz = pz[t] - cz;
                r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
            if (lane == 0) red[3 * NW + wid] = r2;
            __syncthreads();
            if (wid == 0) {
                double rt2 = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) rt2 = fmax(rt2, red[3 * NW + w]);
                const double reach = (D + sqrt(rt2)) * (1.0 + 1.0e-6) + 1.0e-14;
                const double* b = g.bounds + (size_t)(s0 / TS) * 4;
                for (int base = 0; base < ntiles; base += 32) {
                    const int k = base + lane;
                    bool nd = false;
                    if (k < ntiles) {
                        const double dx = b[4 * k] - cx, dy = b[4 * k + 1] - cy, dz = b[4 * k + 2] - cz;
                        const double lim = reach + b[4 * k + 3];       // an all-padding tile has radius -1e300
                        nd = lim > 0.0 && dx * dx + dy * dy + dz * dz <= lim * lim;
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, nd);
                    if (lane == 0) need[base >> 5] = m;
                }
            }
        }
    }
    __syncthreads();

    // next tile to visit at or after `from` (ntiles if none)
    auto next_tile = [&](int from) -> int {
        while (from < ntiles) {
            const uint32_t w = need[from >> 5] >> (from & 31);
            if (w) return from + __ffs(w) - 1;
            from = (from | 31) + 1;
        }
        return ntiles;
    };

    // One tile against the thread's targets; the tile's sums then join the running sums.
    auto do_tile = [&](const int k, const int st) {
        const double* sm = tile[st];
        const int32_t j0 = s0 + k * TS;
        const bool check = K::SKIP_SELF && (j0 < selfhi) && (j0 + TS > selflo);
        bool careful = check;      // run the tile through group<T, true>()
        if (!check) {
            unsigned worst = 0;
#pragma unroll U
            for (int j = 0; j < TS; ++j) {
                double s[NS];
                const double2* p2 = reinterpret_cast<const double2*>(sm + j * NS);
#pragma unroll
                for (int q = 0; q < NS / 2; ++q) {
                    double2 v = p2[q];
                    s[2 * q] = v.x; s[2 * q + 1] = v.y;
                }
                if constexpr (K::RETRY) K::template group_fast<T>(prm, tg, s, acc, worst, sctx);
                else K::template group<T, false>(prm, tg, s, acc, 0, self, sctx);
            }
            if constexpr (K::RETRY) {
                // some argument of this thread was outside the fast path's domain: drop the
                // tile's sums and redo the tile with the general code (per thread; no barrier inside)
                careful = K::needs_retry(worst);
            }
            if constexpr (K::BATCHED_RCP) {
                // The targets of a thread share one MUFU per source here, so a zero or non-finite denominator
                // (coincident particles) poisons its thread-mates' sums too.  Any non-finite tile sum sends the
                // thread through the careful loop, whose reciprocals are taken one by one: the Inf / NaN then
                // stays with the target that owns it, as in the reference.  One DADD chain per tile.
                double chk = 0.0;
#pragma unroll
                for (int t = 0; t < T; ++t)
#pragma unroll
                    for (int a = 0; a < NA; ++a) chk += acc[t][a];
                careful = careful || !(fabs(chk) <= 1.7976931348623157e308);
            }
            if (careful) {
#pragma unroll
                for (int t = 0; t < T; ++t)
#pragma unroll
                    for (int a = 0; a < NA; ++a) acc[t][a] = 0.0;
            }
        }
        if (careful) {
#pragma unroll 1
            for (int j = 0; j < TS; ++j) {
                double s[NS];
                const double2* p2 = reinterpret_cast<const double2*>(sm + j * NS);
#pragma unroll
                for (int q = 0; q < NS / 2; ++q) {
                    double2 v = p2[q];
                    s[2 * q] = v.x; s[2 * q + 1] = v.y;
                }
                K::template group<T, true>(prm, tg, s, acc, j0 + j, self, sctx);
            }
        }
        // Two-level summation: the tile's sum joins the running sum (kept in shared memory,
        // one column per thread), so rounding error grows like sqrt(TS) + sqrt(ntiles)
        // instead of sqrt(TS * ntiles).
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                double* r = &run[(t * NA + a) * BLOCK + tid];
                *r = __dadd_rn(*r, acc[t][a]);
                acc[t][a] = 0.0;
            }
    };
    auto load_tile = [&](const int k, const int st) {      // thread 0 only
        mbar_expect_tx(&full[st], kTileBytes);
        tma_bulk_g2s(tile[st], src + (size_t)(s0 + k * TS) * NS, kTileBytes, &full[st]);
    };

    if (!cull) {
        // every tile, in order: two-stage pipeline
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < 2; ++s)
                if (s < ntiles) load_tile(s, s);
        }
        for (int k = 0; k < ntiles; ++k) {
            const int st = k & 1;
            mbar_wait(&full[st], (k >> 1) & 1);
            do_tile(k, st);
            __syncthreads();    // everyone is done with tile[st]
            if (tid == 0 && k + 2 < ntiles) load_tile(k + 2, st);
        }
    } else {
        // only the tiles named in `need`, same pipeline
        int cur = next_tile(0);
        int nxt = next_tile(cur + 1);
        if (tid == 0) {
            if (cur < ntiles) load_tile(cur, 0);
            if (nxt < ntiles) load_tile(nxt, 1);
        }
        for (int it = 0; cur < ntiles; ++it) {
            const int st = it & 1;
            mbar_wait(&full[st], (it >> 1) & 1);
            do_tile(cur, st);
            __syncthreads();
            const int after = next_tile(nxt + 1);       // refill this stage with the tile after `nxt`
            if (tid == 0 && nxt < ntiles && after < ntiles) load_tile(after, st);
            cur = nxt;
            nxt = (nxt < ntiles) ? after : ntiles;
        }
    }

#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int a = 0; a < NA; ++a) acc[t][a] = run[(t * NA + a) * BLOCK + tid];

    if (g.nchunks == 1) {
#pragma unroll
        for (int t = 0; t < T; ++t)
            if (live[t]) K::finalize(prm, tg[t], acc[t], blk0 + t * BLOCK + tid);
    } else {
#pragma unroll
        for (int t = 0; t < T; ++t) {
            if (!live[t]) continue;
            const int64_t li = blk0 + t * BLOCK + tid - g.tbeg;
#pragma unroll
            for (int a = 0; a < NA; ++a) partial[((size_t)ck * NA + a) * g.ntgt + li] = acc[t][a];
        }
    }
}

template <class K, int T, int BLOCK>
constexpr size_t ds_smem_bytes()
{
    return 2 * size_t(kTile) * K::NS * sizeof(double) + sizeof(double) * (BLOCK * T * K::NA + K::KS) +
           2 * sizeof(uint64_t) +
           (K::CULL ? kMaxNeedWords * sizeof(uint32_t) + 4 * (BLOCK / 32) * sizeof(double) : 0);
}

// Adds the chunk partials in chunk order and writes the outputs.
template <class K>
__global__ void __launch_bounds__(256)
ds_finalize_kernel(const typename K::Params prm, const DsGeom g, const double* __restrict__ partial)
{
    const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= g.ntgt) return;
    double acc[K::NA];
#pragma unroll
    for (int a = 0; a < K::NA; ++a) acc[a] = 0.0;
    for (int c = 0; c < g.nchunks; ++c)
#pragma unroll
        for (int a = 0; a < K::NA; ++a) acc[a] = __dadd_rn(acc[a], partial[((size_t)c * K::NA + a) * g.ntgt + li]);
    const int64_t i = g.tbeg + li;
    typename K::Tgt tg = K::load_target(prm, i);
    K::finalize(prm, tg, acc, i);
}

}  // namespace lpm
