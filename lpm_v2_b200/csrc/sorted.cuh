// sorted.cuh -- spatially sorted evaluation of the compactly supported (PSE) sums.
//
// The PSE kernels (src/PSEDirectSum.f90) weight a pair by eta(d_ij / eps), which is < 1e-23 of
// its peak beyond d = kPseCut eps, so almost all of the reference's O(N F) pairs contribute
// nothing.  Tile culling (directsum.cuh) skips a source tile when its bounding ball cannot
// reach the target block's; that only pays when tiles and blocks are spatially compact, and
// the reference's particle order (mesh-refinement order) is not: at icosTri 7 the median
// 512-particle block has radius 0.19 against a cut-off of 0.26, and 16 % of the (block, tile)
// pairs survive where 1.7 % of the particle pairs do.
//
// So, for these kernels only, the active sources are packed in Morton order of their cell
// (a stable radix sort of (cell key, particle index), so the order is deterministic and
// the same on every GPU), the slice's targets are evaluated in Morton order as well (their
// input arrays are gathered, the results scattered back to particle order), and the engine
// runs unchanged on the reordered arrays.  The sum over sources is then taken in a different
// order than the reference's j = 1..N loop: a ~1e-16 relative difference, inside the 1e-12
// parity budget; which pairs are evaluated does not depend on the order.
//
// The sort is CUB's DeviceRadixSort (library code, off the O(N F) path).
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include <type_traits>

#include "ops.cuh"

namespace lpm {

// interleave the low 10 bits of x, y, z / the low 16 bits of x, y
__device__ __forceinline__ uint32_t spread3(uint32_t v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t spread2(uint32_t v)
{
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// Cell key of a point.  GEOM 3: direction on the unit sphere, 1024^3 cells over [-1, 1]^3.
// GEOM 2: plane, square cells of size `cell` (wrapping every 65536 cells: a far-away particle can
// share a key with a near one, which costs compactness, never correctness).
template <int GEOM>
__device__ __forceinline__ uint32_t cell_key(double x, double y, double z, double inv_cell)
{
    if (GEOM == 3) {
        const double inv = 511.5 / sqrt(x * x + y * y + z * z);
        const int cx = (int)(x * inv + 511.5), cy = (int)(y * inv + 511.5), cz = (int)(z * inv + 511.5);
        return spread3((uint32_t)cx) | (spread3((uint32_t)cy) << 1) | (spread3((uint32_t)cz) << 2);
    } else {
        const long long cx = (long long)floor(x * inv_cell), cy = (long long)floor(y * inv_cell);
        return spread2((uint32_t)(cx & 0xffff)) | (spread2((uint32_t)(cy & 0xffff)) << 1);
    }
}

// keys of the entries idx[c] (or first + c when idx is null), c < count
template <int GEOM>
__global__ void cell_keys_kernel(int64_t count, const int32_t* __restrict__ idx, int64_t first,
                                 const double* __restrict__ x, const double* __restrict__ y,
                                 const double* __restrict__ z, double inv_cell, uint32_t* __restrict__ keys,
                                 int32_t* __restrict__ vals)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    const int64_t i = idx ? idx[c] : first + c;
    keys[c] = cell_key<GEOM>(x[i], y[i], GEOM == 3 ? z[i] : 0.0, inv_cell);
    vals[c] = (int32_t)i;
}

__global__ void gather_kernel(int64_t count, const int32_t* __restrict__ perm, const double* __restrict__ in,
                              double* __restrict__ out)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < count) out[c] = in[perm[c]];
}

// out[r][perm[c]] = in[c] on every replica r (the slice exchange of the peer-store mode)
struct ScatterDst { int nrep; double* p[kMaxRep]; };
__global__ void scatter_kernel(int64_t count, const int32_t* __restrict__ perm, const double* __restrict__ in,
                               ScatterDst dst)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    const double v = in[c];
    const int64_t i = perm[c];
    for (int r = 0; r < dst.nrep; ++r) dst.p[r][i] = v;
}

// Pair-symmetric evaluation (symmetric.cuh, included at the end of this file): runs Op's sum that way and sets
// *taken when Op has a symmetric form, the particle set is large enough and the call covers this device's /
// rank's whole share of the targets (sym_applicable).  out[k]: where the results go (this device's copy; in rank mode every rank
// ends up with all n results, so nothing is left to exchange).
template <class Op>
inline int sym_try(Device& dev, cudaStream_t st, MaskPlan& mp, const Args& a, double* const* out, int64_t tbeg,
                   int64_t tend, int64_t nt, int nrep, bool* taken);

// (key, value) pairs sorted by key into ws.sort_vals[1]; returns that pointer
inline int sort_by_cell(Device& dev, cudaStream_t st, int64_t count, int bits, uint32_t* keys_in, int32_t* vals_in,
                        uint32_t* keys_out, int32_t* vals_out)
{
    size_t need = 0;
    LPM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, vals_in, vals_out, (int)count, 0, bits, st));
    LPM_TRY(dev.ws.sort_tmp.reserve(need));
    LPM_CUDA(cub::DeviceRadixSort::SortPairs(dev.ws.sort_tmp.p, need, keys_in, keys_out, vals_in, vals_out, (int)count, 0,
                                             bits, st));
    count_launch(4);     // histogram + onesweep passes (CUB internals; an estimate)
    return LPM_OK;
}

// Evaluate Op for targets [tbeg, tend) of nt into out[k][i] (device arrays in particle / target
// order).  Plain path: reference order.  Sorted path: see above.  In rank mode, when every
// out[k] lies in a shared slab (lpm_comm_alloc_shared), the results are stored into every
// rank's copy over NVLink, bracketed by barriers (runtime.cuh), and *exchanged is set: the
// caller then skips the NCCL slice exchange.
template <class Op>
inline int evaluate_impl(Device& dev, cudaStream_t st, MaskPlan& mp, const Args& a, int64_t tbeg, int64_t tend, int64_t nt,
                         double* const* out, bool* exchanged);

template <class Op>
inline int evaluate(Device& dev, cudaStream_t st, MaskPlan& mp, const Args& a, int64_t tbeg, int64_t tend, int64_t nt,
                    double* const* out, bool* exchanged = nullptr)
{
    bool ex = false;
    if (!exchanged) exchanged = &ex;
    const bool prof = rt().profiling;
    if (prof) {
        for (auto& e : dev.ev_sum)
            if (!e) LPM_CUDA(cudaEventCreate(&e));
        LPM_CUDA(cudaEventRecord(dev.ev_sum[0], st));
    }
    LPM_TRY(evaluate_impl<Op>(dev, st, mp, a, tbeg, tend, nt, out, exchanged));
    if (prof) LPM_CUDA(cudaEventRecord(dev.ev_sum[1], st));
    return LPM_OK;
}

template <class Op>
inline int evaluate_impl(Device& dev, cudaStream_t st, MaskPlan& mp, const Args& a, int64_t tbeg, int64_t tend, int64_t nt,
                         double* const* out, bool* exchanged)
{
    using K = typename Op::K;
    const int mode = rt().pse_culling;      // 0 plain, 1 sort + cull, 2 sort only
    bool sorted = false;
    if constexpr (K::CULL) sorted = mode != 0 && mp.nsrc > 2 * kTile && tend - tbeg > 0;
    if (!sorted) {
        bool taken = false;
        LPM_TRY(sym_try<Op>(dev, st, mp, a, out, tbeg, tend, nt, 1, &taken));
        if (taken) {
            *exchanged = rt().rank_mode && rt().world > 1;
            return LPM_OK;
        }
        LPM_TRY(Op::pack(dev, st, mp, a));
        typename K::Params prm = Op::params(a);
        *exchanged = set_outs_shared(prm.out, out, nt);
        if (tend <= tbeg && *exchanged) {      // empty slice: still take part in the barriers
            LPM_TRY(comm_barrier(dev, st));
            LPM_TRY(comm_barrier(dev, st));
        }
        return direct_sum<K>(dev, st, mp, tbeg, tend, prm, nt);
    }
    if constexpr (K::CULL) {
        constexpr int GEOM = K::CULL_GEOM;
        constexpr int BITS = GEOM == 3 ? 30 : 32;
        const int64_t ns = tend - tbeg;
        Workspace& ws = dev.ws;
        // plane cells: a quarter of the cut-off radius (sc[0] = eps for every plane operator)
        const double inv_cell = GEOM == 2 ? 1.0 / (0.25 * kPseCut * a.sc[0]) : 0.0;
        const int64_t nmax = std::max<int64_t>(mp.nsrc, ns);
        for (int q = 0; q < 2; ++q) {
            LPM_TRY(ws.sort_keys[q].reserve((size_t)nmax * sizeof(uint32_t)));
            LPM_TRY(ws.sort_vals[q].reserve((size_t)nmax * sizeof(int32_t)));
        }
        LPM_TRY(ws.sorted_active.reserve((size_t)mp.nsrc * sizeof(int32_t)));
        LPM_TRY(ws.sorted_targets.reserve((size_t)ns * sizeof(int32_t)));
        uint32_t* k0 = ws.sort_keys[0].as<uint32_t>();
        uint32_t* k1 = ws.sort_keys[1].as<uint32_t>();
        int32_t* v0 = ws.sort_vals[0].as<int32_t>();
        // ---- sources: the active particles in cell order
        const double* sx = a.in[0];
        const double* sy = a.in[1];
        const double* sz = GEOM == 3 ? a.in[2] : nullptr;
        cell_keys_kernel<GEOM><<<(unsigned)((mp.nsrc + 255) / 256), 256, 0, st>>>(mp.nsrc, mp.active.as<int32_t>(), 0, sx, sy,
                                                                                sz, inv_cell, k0, v0);
        count_launch();
        LPM_TRY(sort_by_cell(dev, st, mp.nsrc, BITS, k0, v0, k1, ws.sorted_active.as<int32_t>()));
        MaskPlan mps;                       // a view: same counts, the sorted list (buffers are not owned)
        mps.n = mp.n; mps.nsrc = mp.nsrc;
        mps.active.p = ws.sorted_active.p; mps.active.cap = ws.sorted_active.cap;
        mps.scan.p = mp.scan.p; mps.scan.cap = mp.scan.cap;        // unused: no PSE kernel skips the self pair
        static_assert(!K::SKIP_SELF, "the sorted path has no scan of the reordered mask");
        LPM_TRY(Op::pack(dev, st, mps, a));
        // ---- targets of the slice in cell order
        const double* tx = Op::NTGT ? a.tgt[0] : a.in[0];
        const double* ty = Op::NTGT ? a.tgt[1] : a.in[1];
        const double* tz = GEOM == 3 ? (Op::NTGT ? a.tgt[2] : a.in[2]) : nullptr;
        cell_keys_kernel<GEOM><<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(ns, nullptr, tbeg, tx, ty, tz, inv_cell, k0, v0);
        count_launch();
        LPM_TRY(sort_by_cell(dev, st, ns, BITS, k0, v0, k1, ws.sorted_targets.as<int32_t>()));
        const int32_t* tperm = ws.sorted_targets.as<int32_t>();
        // gather the target-side arrays; the source side (pack) has already read the originals
        Args ag = a;
        constexpr int NG = Op::NTGT ? Op::NTGT : Op::NIN;
        for (int k = 0; k < NG; ++k) {
            LPM_TRY(ws.gathered[k].reserve((size_t)ns * sizeof(double)));
            const double* in = Op::NTGT ? a.tgt[k] : a.in[k];
            gather_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(ns, tperm, in, ws.gathered[k].as<double>());
            if (Op::NTGT) ag.tgt[k] = ws.gathered[k].as<double>();
            else ag.in[k] = ws.gathered[k].as<double>();
        }
        count_launch(NG);
        typename K::Params prm = Op::params(ag);
        prm.out.nrep = 1;
        for (int k = 0; k < Op::NOUT; ++k) {
            LPM_TRY(ws.sorted_out[k].reserve((size_t)ns * sizeof(double)));
            prm.out.p[0][k] = ws.sorted_out[k].as<double>();
        }
        LPM_TRY(direct_sum<K>(dev, st, mps, 0, ns, prm, ns));
        Outs<Op::NOUT> o{};
        *exchanged = set_outs_shared(o, out, nt);      // the scatter is the storing kernel here
        if (*exchanged) LPM_TRY(comm_barrier(dev, st));
        for (int k = 0; k < Op::NOUT; ++k) {
            ScatterDst dst{};
            dst.nrep = o.nrep;
            for (int r = 0; r < o.nrep; ++r) dst.p[r] = o.p[r][k];
            scatter_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(ns, tperm, ws.sorted_out[k].as<double>(), dst);
        }
        count_launch(Op::NOUT);
        if (*exchanged) LPM_TRY(comm_barrier(dev, st));
        LPM_CUDA(cudaGetLastError());
    }
    return LPM_OK;
}

}  // namespace lpm

#include "symmetric.cuh"
