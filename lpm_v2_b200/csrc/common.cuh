// common.cuh -- runtime plumbing shared by every entry point of liblpmgpu.so:
// error reporting, per-device context (stream, workspace, profiling events).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "lpm_gpu.h"

namespace lpm {

// ---- error reporting (src/Logger.f90:158-160: log, never abort) -------------
inline std::string& last_error_ref()
{
    static thread_local std::string s;
    return s;
}
inline int set_error(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

#define LPM_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::lpm::set_error(LPM_ERR_CUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__,  \
                                    #expr, cudaGetErrorString(_e));                            \
    } while (0)

#define LPM_TRY(expr)                 \
    do {                              \
        int _r = (expr);              \
        if (_r != LPM_OK) return _r;  \
    } while (0)

// ---- growable device buffer --------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return LPM_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return set_error(LPM_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return LPM_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Scratch a direct sum needs on one device.  Reused across calls.
// Result of scanning a mask (scan.cuh): scan[n+1], active[nsrc] and the active count.
struct MaskPlan {
    int64_t n = 0;
    int32_t nsrc = 0;
    DevBuf scan, active, blocksums;
    void release() { scan.release(); active.release(); blocksums.release(); }
};

struct SharedOut { void* p = nullptr; size_t cap = 0; };     // a shared slab used as output staging (rank mode)

struct Workspace {
    MaskPlan plan;      // mask scan of the one-shot entry points (rebuilt per call, buffers reused)
    DevBuf sources;     // double [nsrc_pad * NS]
    DevBuf partial;     // double [nchunks * NA * ntgt]
    DevBuf bounds;      // double [nsrc_pad / kTile][4]: bounding ball per source tile (PSE tile culling)
    DevBuf staging[16]; // host API: device copies of the caller's arrays (0-7 in, 8 mask, 9-11 out, 12-14 targets)
    DevBuf reduce;      // small reduction scratch
    DevBuf barrier;     // rank-mode barrier / IPC handle exchange scratch
    SharedOut shared_out[4];   // host API in rank mode: outputs staged where the peers can store (freed with the slabs)
    DevBuf logwin;      // int32 [2]: max |coordinate| high word, window origin of the log table (pairs.cuh)
    // spatially sorted PSE evaluation (sorted.cuh)
    DevBuf sort_tmp, sort_keys[2], sort_vals[2], sorted_active, sorted_targets, gathered[8], sorted_out[8];
    DevBuf sym_acc;     // fixed-point accumulators of the symmetric paths (symmetric.cuh): kFxWords 64-bit words per sum
    DevBuf sym_acc2, sym_fx;    // the same sums as doubles; FxWindow + max high word of the records
    void release()
    {
        plan.release(); sources.release(); partial.release(); bounds.release();
        reduce.release(); logwin.release(); barrier.release(); sym_acc.release(); sym_acc2.release(); sym_fx.release();
        for (auto& so : shared_out) so = SharedOut{};      // the slabs themselves are freed by the runtime
        sort_tmp.release(); sorted_active.release(); sorted_targets.release();
        for (auto& s : sort_keys) s.release();
        for (auto& s : sort_vals) s.release();
        for (auto& s : gathered) s.release();
        for (auto& s : sorted_out) s.release();
        for (auto& s : staging) s.release();
    }
};

struct Device {
    int id = -1;
    int sm_count = 0;
    const double* logtab = nullptr;      // this device's g_log_full (pairs.cuh)
    cudaStream_t stream = nullptr;       // library-owned stream (host API, resident solvers)
    cudaEvent_t ev_done = nullptr;       // cross-device barrier (resident solvers)
    cudaEvent_t ev_sum[2] = {nullptr, nullptr};   // profiling: around the last whole direct sum (pack, sort, kernels)
    cudaStream_t comm_stream = nullptr;  // rank mode: collectives that overlap a sum's compute (symmetric.cuh)
    cudaEvent_t ev_comm[2] = {nullptr, nullptr};
    int ensure_comm_stream()
    {
        if (comm_stream) return LPM_OK;
        if (cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking) != cudaSuccess) return LPM_ERR_CUDA;
        for (auto& e : ev_comm)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return LPM_ERR_CUDA;
        return LPM_OK;
    }
    // profiling: one event pair per direct-sum main kernel since the last reset
    // and what it timed (lpm_profile_breakdown): tag = 2 * sum + engine; sum 0 BVE velocity, 1 BVE stream
    // functions, 2 any other, 3 the fused velocity + stream functions; engine 0 one-sided ds_kernel, 1 pair-symmetric sym_kernel
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof;
    std::vector<int> prof_tag;
    size_t prof_used = 0;
    Workspace ws;
    int next_prof(cudaEvent_t* b, cudaEvent_t* e, int tag)
    {
        if (prof_used == prof.size()) {
            cudaEvent_t x, y;
            if (cudaEventCreate(&x) != cudaSuccess || cudaEventCreate(&y) != cudaSuccess) return LPM_ERR_CUDA;
            prof.emplace_back(x, y);
            prof_tag.push_back(0);
        }
        *b = prof[prof_used].first; *e = prof[prof_used].second;
        prof_tag[prof_used] = tag;
        ++prof_used;
        return LPM_OK;
    }
};

// Rank mode: a device buffer of this rank that every other rank has mapped with CUDA IPC
// (lpm_comm_alloc_shared).  peer[r] is rank r's buffer as seen from this process; peer[rank] == local.
struct SharedSlab {
    char* local = nullptr;
    size_t bytes = 0;
    char* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

struct Runtime {
    bool initialised = false;
    bool rank_mode = false;              // one process per GPU (lpm_gpu_init_rank)
    std::vector<Device> devs;            // devices this process drives
    bool profiling = false;
    int64_t launches = 0;
    bool symmetric = true;               // pair-symmetric evaluation of whole BVE sums (lpm_set_symmetric)
    int sym_vel_order = 11;              // LPM_SYM_ORDER_SWEEP builds only
    bool fuse_step_end = true;           // BVE RK4 step: final velocity + stream functions in one pass (A/B: lpm_tune)
    int force_T = 0;                     // A/B: targets per thread of the one-sided engine (0 = automatic)
    int32_t sym_min_sources = 200000;
    int32_t sym_panel_blocks = 256;      // ... launched in panels of this many target blocks (kSymPanelBlocks)
    int32_t sym_chunk_tiles = 16;        // ... in chunks of this many source tiles per CTA (symmetric.cuh, kSymChunkTiles)    // ... for at least this many active particles (symmetric.cuh)
    bool pse_series = true;              // sphere PSE kernels: theta^2 by series inside the cut-off (false: atan2 always)
    int pse_culling = 1;                 // PSE kernels: 0 reference order, every tile; 1 cell order + tile culling; 2 cell order only
    // NCCL (rank mode)
    void* nccl_lib = nullptr;
    void* comm = nullptr;
    int world = 1, rank = 0;
    std::vector<SharedSlab> slabs;
    // live resident solvers: handle -> how to destroy it.  lpm_gpu_finalize destroys what the caller left behind
    // (their device memory, streams and events go with the runtime); a later *_delete of such a handle is a no-op
    // and any other use of it an error, instead of a dereference of freed memory.
    struct LiveSolver { void* handle; void (*destroy)(void*); };
    std::vector<LiveSolver> live_solvers;
};

inline Runtime& rt()
{
    static Runtime r;
    return r;
}

inline void count_launch(int k = 1) { rt().launches += k; }

}  // namespace lpm
