// runtime.cuh -- device claiming, NCCL (loaded lazily with dlopen), the mask
// plan and the direct-sum launcher.
#pragma once
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "directsum.cuh"
#include "pairs.cuh"
#include "scan.cuh"

namespace lpm {

// ---------------------------------------------------------------- devices
inline int init_device(Device& d, int id)
{
    d.id = id;
    LPM_CUDA(cudaSetDevice(id));
    cudaDeviceProp prop;
    LPM_CUDA(cudaGetDeviceProperties(&prop, id));
    if (prop.major != 10)
        return set_error(LPM_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; liblpmgpu is built for sm_100a only", id,
                         prop.name, prop.major, prop.minor);
    d.sm_count = prop.multiProcessorCount;
    {   // table for log_tab(): -ln Q for every (binade, bin), Q = MUFU.RCP64H(bin centre) read back
        // from THIS device and evaluated in long double on the host
        double* dq = nullptr;
        std::vector<double> tab(kLogFull);
        LPM_CUDA(cudaMalloc(&dq, sizeof(double) * kLogFull));
        log_table_seed_kernel<<<kLogFull / 256, 256>>>(dq);
        LPM_CUDA(cudaMemcpy(tab.data(), dq, sizeof(double) * kLogFull, cudaMemcpyDeviceToHost));
        LPM_CUDA(cudaFree(dq));
        for (int idx = 0; idx < kLogFull; ++idx) {
            const int binade = idx >> kLogBits;
            if (binade < kLogBinadeMin || binade > kLogBinadeMax) { tab[idx] = 0.0; continue; }
            // the bin centre is c = 2^(binade-1023) (1 + (k + 1/2)/256); Q must be its reciprocal to ~2^-20
            const double c = std::ldexp(1.0 + ((idx & (kLogBin - 1)) + 0.5) / kLogBin, binade - 1023);
            if (!(std::fabs(tab[idx] * c - 1.0) < 1.0e-5))
                return set_error(LPM_ERR_CUDA, "log table seed %d out of range (Q c - 1 = %g)", idx, tab[idx] * c - 1.0);
            tab[idx] = (double)(-logl((long double)tab[idx]));
        }
        LPM_CUDA(cudaMemcpyToSymbol(g_log_full, tab.data(), sizeof(double) * kLogFull));
        void* sym = nullptr;
        LPM_CUDA(cudaGetSymbolAddress(&sym, g_log_full));
        d.logtab = (const double*)sym;
    }
    {   // table for pse_exp_neg(): 2^(-j/64), correctly rounded from long double
        double tab[64];
        for (int j = 0; j < 64; ++j) tab[j] = (double)exp2l(-(long double)j / 64.0L);
        LPM_CUDA(cudaMemcpyToSymbol(g_exp2_tab, tab, sizeof(tab)));
    }
    LPM_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    LPM_CUDA(cudaEventCreateWithFlags(&d.ev_done, cudaEventDisableTiming));
    return LPM_OK;
}

inline int require_init()
{
    if (!rt().initialised)
        return set_error(LPM_ERR_NO_DEVICE, "liblpmgpu: not initialised (call lpm_gpu_init or lpm_gpu_init_rank); "
                                            "there is no CPU fallback");
    return LPM_OK;
}

// the Device entry for the CUDA device that is current on this thread
inline int current_device(Device** out)
{
    LPM_TRY(require_init());
    int id = -1;
    LPM_CUDA(cudaGetDevice(&id));
    for (auto& d : rt().devs)
        if (d.id == id) { *out = &d; return LPM_OK; }
    return set_error(LPM_ERR_NO_DEVICE, "current CUDA device %d was not claimed by lpm_gpu_init", id);
}

// ---------------------------------------------------------------- NCCL
// Only the handful of entry points the slice exchange needs, resolved at
// run time so the library loads (and single-GPU runs work) without NCCL.
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool loaded = false;
};
inline NcclApi& nccl()
{
    static NcclApi a;
    return a;
}
inline int load_nccl()
{
    NcclApi& a = nccl();
    if (a.loaded) return LPM_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy torch already mapped, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(LPM_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
    rt().nccl_lib = h;
#define LPM_SYM(field, name)                                                            \
    *(void**)(&a.field) = dlsym(h, name);                                               \
    if (!a.field) return set_error(LPM_ERR_COMM, "libnccl: missing symbol %s", name)
    LPM_SYM(GetUniqueId, "ncclGetUniqueId");
    LPM_SYM(CommInitRank, "ncclCommInitRank");
    LPM_SYM(CommDestroy, "ncclCommDestroy");
    LPM_SYM(Broadcast, "ncclBroadcast");
    LPM_SYM(AllGather, "ncclAllGather");
    LPM_SYM(AllReduce, "ncclAllReduce");
    LPM_SYM(GroupStart, "ncclGroupStart");
    LPM_SYM(GroupEnd, "ncclGroupEnd");
    LPM_SYM(GetErrorString, "ncclGetErrorString");
#undef LPM_SYM
    a.loaded = true;
    return LPM_OK;
}
#define LPM_NCCL(expr)                                                                              \
    do {                                                                                            \
        int _r = (expr);                                                                            \
        if (_r != 0)                                                                                \
            return ::lpm::set_error(LPM_ERR_COMM, "%s failed: %s", #expr, nccl().GetErrorString(_r)); \
    } while (0)

// src/MPISetup.f90:132-146 LoadBalance, 0-based half-open form
inline void load_balance0(int64_t n, int nprocs, int r, int64_t* beg, int64_t* end)
{
    int64_t chunk = n / nprocs;
    *beg = (int64_t)r * chunk;
    *end = (r == nprocs - 1) ? n : (int64_t)(r + 1) * chunk;
}

// In-place exchange of every rank's slice (the reference's loop of MPI_BCAST
// rooted at each rank in turn, src/SphereBVESolver.f90:422-429), as one NCCL group.
inline int allgather_slices(int ncomp, double* const* bufs, int64_t n, cudaStream_t st)
{
    Runtime& R = rt();
    if (R.world <= 1) return LPM_OK;
    if (!R.comm) return set_error(LPM_ERR_COMM, "world size %d but no communicator (lpm_comm_init_rank)", R.world);
    LPM_NCCL(nccl().GroupStart());
    for (int r = 0; r < R.world; ++r) {
        int64_t b, e;
        load_balance0(n, R.world, r, &b, &e);
        if (e <= b) continue;
        for (int c = 0; c < ncomp; ++c)
            LPM_NCCL(nccl().Broadcast(bufs[c] + b, bufs[c] + b, (size_t)(e - b), /*ncclDouble*/ 8, r, R.comm, st));
    }
    LPM_NCCL(nccl().GroupEnd());
    return LPM_OK;
}

// ---------------------------------------------------------------- shared slabs (rank mode)
// Stream-ordered barrier across the ranks: a 4-byte all-reduce.  Work enqueued on `st` after it
// starts only when every rank's work before its own call has finished.
inline int comm_barrier(Device& dev, cudaStream_t st)
{
    Runtime& R = rt();
    if (R.world <= 1) return LPM_OK;
    if (!R.comm) return set_error(LPM_ERR_COMM, "world size %d but no communicator (lpm_comm_init_rank)", R.world);
    LPM_TRY(dev.ws.barrier.reserve(64));
    int32_t* b = dev.ws.barrier.as<int32_t>();
    LPM_NCCL(nccl().AllReduce(b, b + 8, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, R.comm, st));
    return LPM_OK;
}

// the slab that holds [p, p + bytes), or nullptr
inline const SharedSlab* find_slab(const void* p, size_t bytes)
{
    const char* c = (const char*)p;
    for (const auto& s : rt().slabs)
        if (c >= s.local && c + bytes <= s.local + s.bytes) return &s;
    return nullptr;
}

inline int alloc_shared(size_t bytes, void** out)
{
    Runtime& R = rt();
    LPM_TRY(require_init());
    if (!R.rank_mode) return set_error(LPM_ERR_INVALID, "lpm_comm_alloc_shared needs rank mode (lpm_gpu_init_rank)");
    if (bytes == 0) return set_error(LPM_ERR_INVALID, "lpm_comm_alloc_shared(0)");
    if (R.world > 8) return set_error(LPM_ERR_INVALID, "at most 8 ranks");
    if (R.world > 1 && !R.comm) return set_error(LPM_ERR_COMM, "lpm_comm_alloc_shared before lpm_comm_init_rank");
    Device& dev = R.devs[0];
    LPM_CUDA(cudaSetDevice(dev.id));
    SharedSlab s;
    s.bytes = bytes;
    cudaError_t e = cudaMalloc((void**)&s.local, bytes);
    if (e != cudaSuccess) return set_error(LPM_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    s.peer[R.rank] = s.local;
    // on any failure below: close what was opened, free the buffer (the other ranks fail the same call)
    auto body = [&]() -> int {
        if (R.world <= 1) return LPM_OK;
        // exchange the IPC handles with an all-gather of 64 bytes per rank
        cudaIpcMemHandle_t mine, all[8];
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        LPM_CUDA(cudaIpcGetMemHandle(&mine, s.local));
        LPM_TRY(dev.ws.barrier.reserve(64 + 9 * 64));
        char* scratch = dev.ws.barrier.as<char>() + 64;
        LPM_CUDA(cudaMemcpyAsync(scratch, &mine, 64, cudaMemcpyHostToDevice, dev.stream));
        LPM_NCCL(nccl().AllGather(scratch, scratch + 64, 64, /*ncclInt8*/ 0, R.comm, dev.stream));
        LPM_CUDA(cudaMemcpyAsync(all, scratch + 64, 64 * (size_t)R.world, cudaMemcpyDeviceToHost, dev.stream));
        LPM_CUDA(cudaStreamSynchronize(dev.stream));
        for (int r = 0; r < R.world; ++r) {
            if (r == R.rank) continue;
            void* p = nullptr;
            cudaError_t oe = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
            if (oe != cudaSuccess)
                return set_error(LPM_ERR_COMM, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(oe));
            s.peer[r] = (char*)p;
        }
        return LPM_OK;
    };
    const int rc = body();
    if (rc != LPM_OK) {
        for (int r = 0; r < R.world; ++r)
            if (r != R.rank && s.peer[r]) cudaIpcCloseMemHandle(s.peer[r]);
        cudaFree(s.local);
        cudaGetLastError();
        return rc;
    }
    R.slabs.push_back(s);
    *out = s.local;
    return LPM_OK;
}

inline int free_shared(void* ptr)
{
    Runtime& R = rt();
    for (size_t k = 0; k < R.slabs.size(); ++k) {
        if (R.slabs[k].local != ptr) continue;
        SharedSlab s = R.slabs[k];
        // The teardown is unconditional: whatever fails below, the slab leaves the registry and its
        // memory is freed exactly once; the first error is reported after that.
        R.slabs.erase(R.slabs.begin() + k);
        Device& dev = R.devs[0];
        int rc = LPM_OK;
        cudaSetDevice(dev.id);
        cudaStreamSynchronize(dev.stream);
        cudaDeviceSynchronize();
        for (int r = 0; r < R.world; ++r)
            if (r != R.rank && s.peer[r]) {
                cudaIpcCloseMemHandle(s.peer[r]);
                s.peer[r] = nullptr;
            }
        // nobody should free while a peer still has stores in flight or the mapping open
        if (R.world > 1 && R.comm) {
            rc = comm_barrier(dev, dev.stream);
            if (rc == LPM_OK && cudaStreamSynchronize(dev.stream) != cudaSuccess)
                rc = set_error(LPM_ERR_CUDA, "lpm_comm_free_shared: barrier did not complete: %s",
                               cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(s.local);
        cudaGetLastError();
        return rc;
    }
    return set_error(LPM_ERR_INVALID, "lpm_comm_free_shared: not a shared allocation");
}

// Fill `o` so that output k goes to every rank's copy of out[k] (own copy first) when all
// outputs of this evaluation live in shared slabs; otherwise to out[k] only.  Returns whether
// the peer-store exchange is in effect.
template <int NO>
inline bool set_outs_shared(Outs<NO>& o, double* const* out, int64_t n)
{
    Runtime& R = rt();
    o.nrep = 1;
    for (int k = 0; k < NO; ++k) o.p[0][k] = out[k];
    if (!R.rank_mode || R.world <= 1) return false;
    const SharedSlab* sl[NO];
    for (int k = 0; k < NO; ++k) {
        sl[k] = find_slab(out[k], (size_t)n * sizeof(double));
        if (!sl[k]) return false;
    }
    o.nrep = R.world;
    for (int q = 0; q < R.world; ++q) {
        const int r = (R.rank + q) % R.world;
        for (int k = 0; k < NO; ++k) o.p[q][k] = (double*)(sl[k]->peer[r] + ((char*)out[k] - sl[k]->local));
    }
    return true;
}

// ---------------------------------------------------------------- mask plan
// Scans the device mask; synchronises `st` once to read the active count.
inline int build_mask_plan(cudaStream_t st, int64_t n, const int32_t* mask_dev, MaskPlan& mp)
{
    if (n <= 0 || n > 0x7ffffff0LL) return set_error(LPM_ERR_INVALID, "particle count %lld out of range", (long long)n);
    const int nblocks = (int)((n + kScanBlock - 1) / kScanBlock);
    LPM_TRY(mp.scan.reserve((size_t)(n + 1) * sizeof(int32_t)));
    LPM_TRY(mp.active.reserve((size_t)n * sizeof(int32_t)));
    LPM_TRY(mp.blocksums.reserve((size_t)(nblocks + 1) * sizeof(int32_t)));
    scan_count<<<nblocks, kScanBlock, 0, st>>>(n, mask_dev, mp.blocksums.as<int32_t>());
    scan_blocksums<<<1, kScanBlock, 0, st>>>(nblocks, mp.blocksums.as<int32_t>());
    scan_scatter<<<nblocks, kScanBlock, 0, st>>>(n, mask_dev, mp.blocksums.as<int32_t>(), nblocks,
                                                 mp.scan.as<int32_t>(), mp.active.as<int32_t>());
    count_launch(3);
    LPM_CUDA(cudaGetLastError());
    int32_t total = 0;
    LPM_CUDA(cudaMemcpyAsync(&total, mp.scan.as<int32_t>() + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    LPM_CUDA(cudaStreamSynchronize(st));
    mp.n = n;
    mp.nsrc = total;
    return LPM_OK;
}

// ---------------------------------------------------------------- launcher
template <class K, int T, int BLOCK, int U, int MINB = 1>
inline void launch_ds(cudaStream_t st, const typename K::Params& prm, DsGeom g, const double* src,
                      const int32_t* scan, double* partial)
{
    g.tblk0 = g.tbeg / (BLOCK * T) * (BLOCK * T);
    g.half_bin = 1 << (19 - kLogBits);
    g.ntblocks = (int32_t)((g.tend - g.tblk0 + BLOCK * T - 1) / (BLOCK * T));
    const int64_t grid = (int64_t)g.ntblocks * g.nchunks;
    constexpr size_t smem = ds_smem_bytes<K, T, BLOCK>();
    if (smem > 48 * 1024) {
        // the attribute is per device: remember which devices this instantiation has been configured on
        static bool configured[64] = {};
        int devid = 0;
        cudaGetDevice(&devid);
        if (devid < 0 || devid >= 64 || !configured[devid]) {
            cudaFuncSetAttribute(ds_kernel<K, T, BLOCK, U, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (devid >= 0 && devid < 64) configured[devid] = true;
        }
    }
    ds_kernel<K, T, BLOCK, U, MINB><<<(unsigned)grid, BLOCK, smem, st>>>(prm, g, src, scan, partial);
}

// Chooses the targets-per-thread shape by problem size (never by the slice, see auto_T) and launches.
template <class K>
inline int launch_auto(cudaStream_t st, const typename K::Params& prm, const DsGeom& g,
                       const double* src, const int32_t* scan, double* partial, int sm_count);

// lpm_profile_breakdown: which sum a functor belongs to (0 BVE velocity, 1 BVE stream functions, 2 other, 3 the
// fused velocity + stream functions of an RK4 step's end)
template <class K> struct prof_sum_of { static constexpr int value = 2; };
template <int RG, int ORDER> struct prof_sum_of<BveVelT<RG, ORDER>> { static constexpr int value = 0; };
template <> struct prof_sum_of<BveStream> { static constexpr int value = 1; };
template <> struct prof_sum_of<BveVelStream> { static constexpr int value = 3; };

// Packs nothing: the caller has already written the source records for this
// evaluation into ws.sources (see the api functions).  Runs targets
// [tbeg, tend) against all sources and writes the outputs named in prm.out.
template <class K>
inline int direct_sum(Device& dev, cudaStream_t st, const MaskPlan& mp, int64_t tbeg, int64_t tend,
                      const typename K::Params& prm, int64_t ntargets_all = -1)
{
    if (tend <= tbeg) return LPM_OK;
    DsGeom g{};
    g.tbeg = tbeg; g.tend = tend; g.ntgt = tend - tbeg;
    g.nsrc = mp.nsrc;
    g.nall = ntargets_all >= 0 ? ntargets_all : mp.n;     // targets are the particles unless told otherwise
    ds_chunks(mp.nsrc, &g.nsrc_pad, &g.chunk, &g.nchunks, g.nall);
    double* partial = nullptr;
    if (g.nchunks > 1) {
        LPM_TRY(dev.ws.partial.reserve((size_t)g.nchunks * K::NA * g.ntgt * sizeof(double)));
        partial = dev.ws.partial.as<double>();
    }
    if constexpr (K::CULL) {
        // bounding ball per source tile, from the records the caller has just packed
        static_assert(kTile == 256, "tile_bounds_kernel uses one 256-thread CTA per tile");
        const int ntiles = g.nsrc_pad / kTile;
        LPM_TRY(dev.ws.bounds.reserve((size_t)ntiles * 4 * sizeof(double)));
        tile_bounds_kernel<K::NS, K::CULL_GEOM><<<ntiles, 256, 0, st>>>(g.nsrc, dev.ws.sources.as<double>(),
                                                                       dev.ws.bounds.as<double>());
        count_launch();
        g.bounds = rt().pse_culling == 1 ? dev.ws.bounds.as<double>() : nullptr;
    }
    typename K::Params prm2 = prm;
    if constexpr (K::KS > 0) {      // log kernels: this device's table, and the window the caller's pack() computed
        prm2.logtab = dev.logtab;
        prm2.win = dev.ws.logwin.as<int32_t>() + 1;
    }
    const bool prof = rt().profiling;
    cudaEvent_t pb = nullptr, pe = nullptr;
    if (prof) {
        if (dev.next_prof(&pb, &pe, 2 * prof_sum_of<K>::value) != LPM_OK) return set_error(LPM_ERR_CUDA, "cannot create profiling events");
        LPM_CUDA(cudaEventRecord(pb, st));
    }
    // Rank mode with outputs in shared slabs: the storing kernel writes every rank's copy over
    // NVLink.  A barrier BEFORE it (every rank has finished reading the previous contents of its
    // copy) and one AFTER it (every rank's stores have landed) replace the NCCL exchange.
    const bool peer_exchange = rt().rank_mode && prm2.out.nrep > 1;
    if (peer_exchange && g.nchunks == 1) LPM_TRY(comm_barrier(dev, st));
    LPM_TRY(launch_auto<K>(st, prm2, g, dev.ws.sources.as<double>(), mp.scan.as<int32_t>(), partial, dev.sm_count));
    if (prof) LPM_CUDA(cudaEventRecord(pe, st));
    count_launch();
    if (g.nchunks > 1) {
        if (peer_exchange) LPM_TRY(comm_barrier(dev, st));
        const unsigned nb = (unsigned)((g.ntgt + 255) / 256);
        ds_finalize_kernel<K><<<nb, 256, 0, st>>>(prm2, g, partial);
        count_launch();
    }
    if (peer_exchange) LPM_TRY(comm_barrier(dev, st));
    LPM_CUDA(cudaGetLastError());
    return LPM_OK;
}

// Log kernels: compute the table window for this evaluation on the device (pairs.cuh).
// mode 0: bound = upper bound on the argument; mode 1 / 2: bound from max |a|, |b| over n entries.
template <class K>
inline int log_window(Device& dev, cudaStream_t st, int mode, double bound, int64_t n, const double* a, const double* b)
{
    LPM_TRY(dev.ws.logwin.reserve(2 * sizeof(int32_t)));
    int32_t* w = dev.ws.logwin.as<int32_t>();
    if (mode != 0) {
        LPM_CUDA(cudaMemsetAsync(w, 0, sizeof(int32_t), st));
        const unsigned nb = (unsigned)std::min<int64_t>((n + 255) / 256, 4 * (int64_t)dev.sm_count);
        absmax_hi_kernel<<<nb, 256, 0, st>>>(n, a, b, w);
        count_launch();
    }
    log_window_kernel<<<1, 1, 0, st>>>(mode, bound, w, K::WINDOW_BINADES, w + 1);
    count_launch();
    return LPM_OK;
}

// number of source-record doubles the pack kernels must fill for this plan
template <class K>
inline int reserve_sources(Device& dev, const MaskPlan& mp, int32_t* nsrc_pad)
{
    int32_t chunk, nchunks;
    ds_chunks(mp.nsrc, nsrc_pad, &chunk, &nchunks);
    return dev.ws.sources.reserve((size_t)(*nsrc_pad) * K::NS * sizeof(double));
}

// Heuristic shared by all kernels: largest T that still yields >= 2 CTAs per SM for the
// WHOLE particle set (not the slice), so the launch shape -- and with it the grouping of
// batched reciprocals -- is the same on 1 and on 8 GPUs.
inline int auto_T(int64_t ntgt, int nchunks, int sm_count, int block)
{
    if (rt().force_T == 1 || rt().force_T == 2 || rt().force_T == 4) return rt().force_T;
    const int64_t want = 2LL * sm_count;
    for (int T : {4, 2}) {
        int64_t items = (ntgt + (int64_t)block * T - 1) / ((int64_t)block * T) * nchunks;
        if (items >= want) return T;
    }
    return 1;
}

template <class K>
inline int launch_auto(cudaStream_t st, const typename K::Params& prm, const DsGeom& g,
                       const double* src, const int32_t* scan, double* partial, int sm_count)
{
    if constexpr (K::KS > 0) {
        // Log kernels carry a 64 KB table per CTA: 256 threads share it and two CTAs fit an SM.
        // __launch_bounds__(256, 2) tells ptxas it may spend up to 128 registers; it then keeps the
        // T x U independent log chains interleaved (>= 4 FP64 instructions between producer and
        // consumer).  With its default target of ~80 registers it emitted each chain back to back
        // and the kernel stalled on DFMA latency (ncu: FP64 pipe 62 %, top stall `wait`).
        switch (auto_T(g.nall, g.nchunks, sm_count, 256)) {
            case 4: launch_ds<K, 4, 256, 2, 2>(st, prm, g, src, scan, partial); break;
            case 2: launch_ds<K, 2, 256, 2, 2>(st, prm, g, src, scan, partial); break;
            default: launch_ds<K, 1, 256, 2, 2>(st, prm, g, src, scan, partial); break;
        }
    } else {
        switch (auto_T(g.nall, g.nchunks, sm_count, 128)) {
            case 4: launch_ds<K, 4, 128, 1>(st, prm, g, src, scan, partial); break;
            case 2: launch_ds<K, 2, 128, 2>(st, prm, g, src, scan, partial); break;
            default: launch_ds<K, 1, 128, 4>(st, prm, g, src, scan, partial); break;
        }
    }
    return LPM_OK;
}

// The fused velocity + stream kernel: 64 KB table, 2 x 16 KB of tiles and BLOCK*T*5 running sums -- one CTA of 256
// threads per SM (136 KB with 4 targets per thread), so no register cap: __launch_bounds__(256, 1).
template <>
inline int launch_auto<BveVelStream>(cudaStream_t st, const BveVelStream::Params& prm, const DsGeom& g,
                                     const double* src, const int32_t* scan, double* partial, int sm_count)
{
    using K = BveVelStream;
    switch (auto_T(g.nall, g.nchunks, sm_count, 256)) {
        case 4: launch_ds<K, 4, 256, 2, 1>(st, prm, g, src, scan, partial); break;
        case 2: launch_ds<K, 2, 256, 2, 1>(st, prm, g, src, scan, partial); break;
        default: launch_ds<K, 1, 256, 2, 1>(st, prm, g, src, scan, partial); break;
    }
    return LPM_OK;
}

// The BVE velocity kernel (the headline path of round 1; now the passive-target and partial-range path).
// 8 targets per thread (3 CTAs of 128 threads per SM) for large sets, else 4 (5 CTAs per SM), 2, 1.  The statement orders (BveVelT<4, ORDER>) are the ones tools/search_order.py
// short-listed and the GPU sweeps measured fastest (profiles/r01b_sweep_L{7,8}_orders.log): at icosTri 8,
// T = 8 / ORDER 3680: 1412 ms;  T = 4 / ORDER 10765: 1441 ms;  T = 4 / ORDER 0: 1499 ms.  Every shape that
// lost a sweep (other block sizes and unrolls, register caps, one MUFU per pair or per two pairs,
// scheduling fences: profiles/r01c_sweep_L7_caps.log, profiles/r02_ab_sym.log) is no longer built.
template <>
inline int launch_auto<BveVel>(cudaStream_t st, const BveVel::Params& prm, const DsGeom& g,
                               const double* src, const int32_t* scan, double* partial, int sm_count)
{
    using K = BveVel;
    // 8 targets per thread only when the WHOLE particle set gives >= 32 waves of such CTAs (3 per SM), so that a rank
    // of an 8-GPU run still has >= 4 waves: icosTri 7 and up; icosTri 6 takes 4 targets per thread (0.84 against 0.80 of
    // the FP64 bound on a rank's slice).  Decided from the whole set, never the slice: the bits must not depend on G.
    const int64_t items8 = (g.nall + 128 * 8 - 1) / (128 * 8) * g.nchunks;
    if (rt().force_T == 8 || (rt().force_T == 0 && items8 >= 32LL * 3 * sm_count)) {
        launch_ds<BveVelT<4, 3680>, 8, 128, 2>(st, prm, g, src, scan, partial);
        return LPM_OK;
    }
    switch (auto_T(g.nall, g.nchunks, sm_count, 128)) {
        case 4: launch_ds<BveVelT<4, 10765>, 4, 128, 2>(st, prm, g, src, scan, partial); break;
        case 2: launch_ds<K, 2, 128, 4>(st, prm, g, src, scan, partial); break;
        default: launch_ds<K, 1, 128, 4>(st, prm, g, src, scan, partial); break;
    }
    return LPM_OK;
}

}  // namespace lpm
