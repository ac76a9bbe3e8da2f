// lpm_gpu.cu -- extern "C" entry points of liblpmgpu.so (see include/lpm_gpu.h).
//
// Host API:   stage the caller's arrays on every claimed device, run each
//             device's LoadBalance slice, copy the slices back.
// Device API: one slice on the current device / given stream (rank mode).
#include <string>

#include "lpm_gpu_tuning.h"
#include "runtime.cuh"
#include "ops.cuh"
#include "sorted.cuh"
#include "solvers.cuh"

using namespace lpm;

// ============================================================== runtime
extern "C" const char* lpm_gpu_last_error(void) { return last_error_ref().c_str(); }

extern "C" int lpm_gpu_device_count(void) { return rt().initialised ? (int)rt().devs.size() : 0; }

static int claim_devices(const std::vector<int>& ids)
{
    Runtime& R = rt();
    R.devs.clear();
    R.devs.resize(ids.size());
    for (size_t k = 0; k < ids.size(); ++k) LPM_TRY(init_device(R.devs[k], ids[k]));
    // peer access between every pair (single-process multi-GPU all-gather by peer stores)
    for (size_t a = 0; a < ids.size(); ++a)
        for (size_t b = 0; b < ids.size(); ++b) {
            if (a == b) continue;
            int can = 0;
            LPM_CUDA(cudaDeviceCanAccessPeer(&can, ids[a], ids[b]));
            if (!can)
                return set_error(LPM_ERR_COMM, "device %d cannot access peer %d; multi-GPU mode needs NVLink/P2P",
                                 ids[a], ids[b]);
            LPM_CUDA(cudaSetDevice(ids[a]));
            cudaError_t e = cudaDeviceEnablePeerAccess(ids[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return set_error(LPM_ERR_COMM, "cudaDeviceEnablePeerAccess(%d->%d): %s", ids[a], ids[b],
                                 cudaGetErrorString(e));
            cudaGetLastError();
        }
    LPM_CUDA(cudaSetDevice(ids[0]));
    R.initialised = true;
    R.launches = 0;
    return LPM_OK;
}

extern "C" int lpm_gpu_init(int ndev_requested, int* ndev_used)
{
    Runtime& R = rt();
    if (R.initialised) {
        if (ndev_used) *ndev_used = (int)R.devs.size();
        return LPM_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(LPM_ERR_NO_DEVICE, "no CUDA device visible (%s); liblpmgpu has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (ndev_requested < 0 || ndev_requested > count)
        return set_error(LPM_ERR_INVALID, "requested %d devices, %d visible", ndev_requested, count);
    int use = ndev_requested == 0 ? count : ndev_requested;
    if (use > kMaxRep) use = kMaxRep;
    std::vector<int> ids(use);
    for (int k = 0; k < use; ++k) ids[k] = k;
    R.rank_mode = false;
    R.world = 1; R.rank = 0;
    LPM_TRY(claim_devices(ids));
    if (ndev_used) *ndev_used = use;
    return LPM_OK;
}

extern "C" int lpm_gpu_init_rank(int device)
{
    Runtime& R = rt();
    if (R.initialised) {
        if (R.rank_mode && R.devs.size() == 1 && R.devs[0].id == device) return LPM_OK;
        return set_error(LPM_ERR_INVALID, "liblpmgpu already initialised with a different configuration");
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(LPM_ERR_NO_DEVICE, "no CUDA device visible (%s); liblpmgpu has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= count) return set_error(LPM_ERR_INVALID, "device %d of %d", device, count);
    R.rank_mode = true;
    R.world = 1; R.rank = 0;
    return claim_devices({device});
}

extern "C" int lpm_gpu_finalize(void)
{
    Runtime& R = rt();
    if (!R.initialised) return LPM_OK;
    while (!R.live_solvers.empty()) {        // solvers the caller never deleted: their memory lives on these devices
        auto l = R.live_solvers.back();
        R.live_solvers.pop_back();
        l.destroy(l.handle);
    }
    while (!R.slabs.empty()) free_shared(R.slabs.back().local);
    if (R.comm && nccl().loaded) nccl().CommDestroy(R.comm);
    R.comm = nullptr;
    for (auto& d : R.devs) {
        cudaSetDevice(d.id);
        cudaStreamSynchronize(d.stream);
        d.ws.release();
        cudaEventDestroy(d.ev_done);
        for (auto& e : d.ev_sum) { if (e) cudaEventDestroy(e); e = nullptr; }
        for (auto& e : d.ev_comm) { if (e) cudaEventDestroy(e); e = nullptr; }
        if (d.comm_stream) { cudaStreamSynchronize(d.comm_stream); cudaStreamDestroy(d.comm_stream); d.comm_stream = nullptr; }
        for (auto& pr : d.prof) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        d.prof.clear(); d.prof_used = 0;
        cudaStreamDestroy(d.stream);
    }
    R.devs.clear();
    R.initialised = false;
    R.world = 1; R.rank = 0;
    return LPM_OK;
}

extern "C" int lpm_comm_unique_id(char id[128])
{
    LPM_TRY(load_nccl());
    NcclApi::UniqueId u;
    LPM_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return LPM_OK;
}

extern "C" int lpm_comm_init_rank(int world_size, int rank, const char id[128])
{
    Runtime& R = rt();
    LPM_TRY(require_init());
    if (!R.rank_mode) return set_error(LPM_ERR_INVALID, "lpm_comm_init_rank needs lpm_gpu_init_rank (one process per GPU)");
    if (world_size < 1 || rank < 0 || rank >= world_size) return set_error(LPM_ERR_INVALID, "bad rank %d/%d", rank, world_size);
    R.world = world_size; R.rank = rank;
    if (world_size == 1) return LPM_OK;
    LPM_TRY(load_nccl());
    NcclApi::UniqueId u;
    memcpy(u.internal, id, 128);
    LPM_CUDA(cudaSetDevice(R.devs[0].id));
    LPM_NCCL(nccl().CommInitRank(&R.comm, world_size, u, rank));
    return LPM_OK;
}

extern "C" int lpm_comm_world_size(void) { return rt().world; }
extern "C" int lpm_comm_rank(void) { return rt().rank; }

extern "C" int lpm_comm_allgather_slices_dev(int ncomp, double* const* bufs, int64_t n, void* stream)
{
    LPM_TRY(require_init());
    return allgather_slices(ncomp, bufs, n, (cudaStream_t)stream);
}

extern "C" int lpm_comm_alloc_shared(int64_t bytes, void** ptr)
{
    if (!ptr || bytes <= 0) return set_error(LPM_ERR_INVALID, "lpm_comm_alloc_shared(%lld)", (long long)bytes);
    return alloc_shared((size_t)bytes, ptr);
}
extern "C" int lpm_comm_free_shared(void* ptr)
{
    LPM_TRY(require_init());
    return free_shared(ptr);
}
extern "C" int lpm_comm_is_shared(const void* ptr, int64_t bytes)
{
    return rt().initialised && ptr && bytes > 0 && find_slab(ptr, (size_t)bytes) != nullptr ? 1 : 0;
}

extern "C" int lpm_gpu_pin(void* ptr, int64_t bytes)
{
    LPM_TRY(require_init());
    LPM_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return LPM_OK;
}
extern "C" int lpm_gpu_unpin(void* ptr)
{
    LPM_TRY(require_init());
    LPM_CUDA(cudaHostUnregister(ptr));
    return LPM_OK;
}

extern "C" int lpm_load_balance(int64_t n_items, int nprocs, int64_t* index_start, int64_t* index_end,
                                int64_t* message_length)
{
    if (nprocs < 1 || n_items < 0) return set_error(LPM_ERR_INVALID, "lpm_load_balance(%lld, %d)", (long long)n_items, nprocs);
    for (int r = 0; r < nprocs; ++r) {
        int64_t b, e;
        load_balance0(n_items, nprocs, r, &b, &e);
        index_start[r] = b + 1;
        index_end[r] = e;
        if (message_length) message_length[r] = e - b;
    }
    return LPM_OK;
}

extern "C" int lpm_set_profiling(int enable) { rt().profiling = enable != 0; return LPM_OK; }
extern "C" int lpm_set_symmetric(int enable) { rt().symmetric = enable != 0; return LPM_OK; }
// Not part of the C ABI (not in include/lpm_gpu.h): knobs for the A/B runs under tools/ and for tests that need a
// small problem to take a large problem's path.  Declared in csrc/lpm_gpu_tuning.h.
extern "C" int lpm_tune(const char* key, int value)
{
    const std::string k = key ? key : "";
    if (k == "max_chunks") {
        if (value < 1 || value > 256) return set_error(LPM_ERR_INVALID, "lpm_tune(max_chunks, %d): 1..256", value);
        max_chunks_ref() = value;
    } else if (k == "chunk_min") {
        if (value < kTile || value % kTile) return set_error(LPM_ERR_INVALID, "lpm_tune(chunk_min, %d): a multiple of %d", value, kTile);
        chunk_min_ref() = value;
    } else if (k == "sym_vel_order") {
        rt().sym_vel_order = value;
    } else if (k == "fuse_step_end") {
        rt().fuse_step_end = value != 0;
    } else if (k == "force_T") {
        rt().force_T = value;
    } else if (k == "sym_panel_blocks") {
        if (value < 1) return set_error(LPM_ERR_INVALID, "lpm_tune(sym_panel_blocks, %d)", value);
        rt().sym_panel_blocks = value;
    } else if (k == "sym_chunk_tiles") {
        if (value < 1) return set_error(LPM_ERR_INVALID, "lpm_tune(sym_chunk_tiles, %d)", value);
        rt().sym_chunk_tiles = value;
    } else if (k == "sym_min_sources") {
        rt().sym_min_sources = value;
    } else {
        return set_error(LPM_ERR_INVALID, "lpm_tune: unknown key '%s'", k.c_str());
    }
    return LPM_OK;
}
extern "C" int lpm_set_pse_series(int enable) { rt().pse_series = enable != 0; return LPM_OK; }
extern "C" int lpm_set_pse_culling(int mode)
{
    if (mode < 0 || mode > 2) return set_error(LPM_ERR_INVALID, "lpm_set_pse_culling(%d): mode must be 0, 1 or 2", mode);
    rt().pse_culling = mode;
    return LPM_OK;
}
extern "C" int64_t lpm_launch_count(int reset)
{
    int64_t v = rt().launches;
    if (reset) rt().launches = 0;
    return v;
}
extern "C" int lpm_last_kernel_ms(double* ms)
{
    Device* d = nullptr;
    LPM_TRY(current_device(&d));
    if (d->prof_used == 0) return set_error(LPM_ERR_INVALID, "no timed kernel (lpm_set_profiling(1) first)");
    auto& pr = d->prof[d->prof_used - 1];
    LPM_CUDA(cudaEventSynchronize(pr.second));
    float f = 0;
    LPM_CUDA(cudaEventElapsedTime(&f, pr.first, pr.second));
    *ms = f;
    return LPM_OK;
}
extern "C" int lpm_last_sum_ms(double* ms)
{
    Device* d = nullptr;
    LPM_TRY(current_device(&d));
    if (!d->ev_sum[1]) return set_error(LPM_ERR_INVALID, "no timed direct sum (lpm_set_profiling(1) first)");
    LPM_CUDA(cudaEventSynchronize(d->ev_sum[1]));
    float f = 0;
    LPM_CUDA(cudaEventElapsedTime(&f, d->ev_sum[0], d->ev_sum[1]));
    *ms = f;
    return LPM_OK;
}
extern "C" int lpm_profile_summary(int reset, int64_t* nkernels, double* total_ms)
{
    Device* d = nullptr;
    LPM_TRY(current_device(&d));
    double tot = 0.0;
    for (size_t k = 0; k < d->prof_used; ++k) {
        LPM_CUDA(cudaEventSynchronize(d->prof[k].second));
        float f = 0;
        LPM_CUDA(cudaEventElapsedTime(&f, d->prof[k].first, d->prof[k].second));
        tot += f;
    }
    if (nkernels) *nkernels = (int64_t)d->prof_used;
    if (total_ms) *total_ms = tot;
    if (reset) d->prof_used = 0;
    return LPM_OK;
}

extern "C" int lpm_profile_breakdown(int reset, int64_t counts[8], double ms[8])
{
    Device* d = nullptr;
    LPM_TRY(current_device(&d));
    for (int t = 0; t < 8; ++t) { counts[t] = 0; ms[t] = 0.0; }
    for (size_t k = 0; k < d->prof_used; ++k) {
        LPM_CUDA(cudaEventSynchronize(d->prof[k].second));
        float f = 0;
        LPM_CUDA(cudaEventElapsedTime(&f, d->prof[k].first, d->prof[k].second));
        const int t = d->prof_tag[k] >= 0 && d->prof_tag[k] < 8 ? d->prof_tag[k] : 4;
        counts[t] += 1;
        ms[t] += f;
    }
    if (reset) d->prof_used = 0;
    return LPM_OK;
}

// ---- FP64 peak probe ----------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double seed, double* sink)
{
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + k * 1e-3 + threadIdx.x * 1e-6;
    const double m = 0.999999, c = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456) sink[0] = s;
}

extern "C" int lpm_fp64_peak_probe(int iters, double* tflops, double* ms)
{
    Device* d = nullptr;
    LPM_TRY(current_device(&d));
    if (iters < 1) iters = 1;
    LPM_TRY(d->ws.reduce.reserve(256));
    const int blocks = d->sm_count * 8;
    cudaEvent_t e0, e1;
    LPM_CUDA(cudaEventCreate(&e0));
    LPM_CUDA(cudaEventCreate(&e1));
    dfma_probe_kernel<<<blocks, 256, 0, d->stream>>>(iters / 8 + 1, 1.0, d->ws.reduce.as<double>());   // warm-up
    LPM_CUDA(cudaEventRecord(e0, d->stream));
    dfma_probe_kernel<<<blocks, 256, 0, d->stream>>>(iters, 1.0, d->ws.reduce.as<double>());
    LPM_CUDA(cudaEventRecord(e1, d->stream));
    LPM_CUDA(cudaEventSynchronize(e1));
    count_launch(2);
    float f = 0;
    LPM_CUDA(cudaEventElapsedTime(&f, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    double flop = 2.0 * 64.0 * (double)iters * 256.0 * blocks;
    if (tflops) *tflops = flop / (f * 1e-3) / 1e12;
    if (ms) *ms = f;
    return LPM_OK;
}

// ============================================================== active list
extern "C" int lpm_active_list(int64_t n, const int32_t* mask, int32_t* list, int64_t* count)
{
    LPM_TRY(require_init());
    Device& d = rt().devs[0];
    LPM_CUDA(cudaSetDevice(d.id));
    LPM_TRY(d.ws.staging[0].reserve((size_t)n * sizeof(int32_t)));
    LPM_CUDA(cudaMemcpyAsync(d.ws.staging[0].p, mask, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, d.stream));
    MaskPlan mp;
    int rc = build_mask_plan(d.stream, n, d.ws.staging[0].as<int32_t>(), mp);
    if (rc == LPM_OK) {
        if (count) *count = mp.nsrc;
        if (list && mp.nsrc > 0) {
            cudaError_t e = cudaMemcpy(list, mp.active.p, (size_t)mp.nsrc * sizeof(int32_t), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = set_error(LPM_ERR_CUDA, "copy of active list failed: %s", cudaGetErrorString(e));
        }
    }
    mp.release();
    return rc;
}

// ============================================================== direct sums
//
// One description per kernel: how to pack sources and fill Params from the
// staged device arrays.  `in[]` are the input arrays in header order (after
// n), `out[]` the outputs; all device pointers.
namespace {

// Device API body: one slice on the current device.
template <class Op>
int run_dev(const Args& a, int64_t ibeg, int64_t iend, double* const* out, void* stream)
{
    Device* dev = nullptr;
    LPM_TRY(current_device(&dev));
    if (a.n <= 0) return set_error(LPM_ERR_INVALID, "n = %lld", (long long)a.n);
    if (ibeg < 0 || iend > a.n || ibeg > iend)
        return set_error(LPM_ERR_INVALID, "target range [%lld, %lld) outside [0, %lld)", (long long)ibeg, (long long)iend, (long long)a.n);
    cudaStream_t st = (cudaStream_t)stream;
    MaskPlan& mp = dev->ws.plan;        // rebuilt per call; buffers reused
    LPM_TRY(build_mask_plan(st, a.n, a.mask, mp));
    return evaluate<Op>(*dev, st, mp, a, ibeg, iend, a.n, out);
}

// Host API body: every claimed device (single-process mode) or this rank's
// device (rank mode) evaluates its LoadBalance slice.
template <class Op>
int run_host(const Args& host, double* const* out_host)
{
    Runtime& R = rt();
    LPM_TRY(require_init());
    const int64_t n = host.n;
    if (n <= 0) return set_error(LPM_ERR_INVALID, "n = %lld", (long long)n);
    for (int k = 0; k < Op::NIN; ++k)
        if (!host.in[k]) return set_error(LPM_ERR_INVALID, "null input array %d", k);
    for (int k = 0; k < Op::NOUT; ++k)
        if (!out_host[k]) return set_error(LPM_ERR_INVALID, "null output array %d", k);
    if (!host.mask) return set_error(LPM_ERR_INVALID, "null mask");
    const size_t nb = (size_t)n * sizeof(double);
    const int64_t nt = Op::NTGT ? host.m : n;       // targets: the particles, or separate locations
    if (nt <= 0) return set_error(LPM_ERR_INVALID, "number of targets = %lld", (long long)nt);
    for (int k = 0; k < Op::NTGT; ++k)
        if (!host.tgt[k]) return set_error(LPM_ERR_INVALID, "null target array %d", k);
    const size_t ntb = (size_t)nt * sizeof(double);
    const int nparts = R.rank_mode ? R.world : (int)R.devs.size();
    int rc = LPM_OK;
    for (size_t g = 0; g < R.devs.size() && rc == LPM_OK; ++g) {
        Device& dev = R.devs[g];
        auto body = [&]() -> int {
            LPM_CUDA(cudaSetDevice(dev.id));
            Args a = host;
            for (int k = 0; k < Op::NIN; ++k) {
                LPM_TRY(dev.ws.staging[k].reserve(nb));
                LPM_CUDA(cudaMemcpyAsync(dev.ws.staging[k].p, host.in[k], nb, cudaMemcpyHostToDevice, dev.stream));
                a.in[k] = dev.ws.staging[k].as<double>();
            }
            DevBuf& mbuf = dev.ws.staging[8];
            LPM_TRY(mbuf.reserve((size_t)n * sizeof(int32_t)));
            LPM_CUDA(cudaMemcpyAsync(mbuf.p, host.mask, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, dev.stream));
            a.mask = mbuf.as<int32_t>();
            for (int k = 0; k < Op::NTGT; ++k) {
                LPM_TRY(dev.ws.staging[12 + k].reserve(ntb));
                LPM_CUDA(cudaMemcpyAsync(dev.ws.staging[12 + k].p, host.tgt[k], ntb, cudaMemcpyHostToDevice, dev.stream));
                a.tgt[k] = dev.ws.staging[12 + k].as<double>();
            }
            double* out[4] = {nullptr, nullptr, nullptr, nullptr};
            static_assert(Op::NOUT <= 4 && Op::NIN <= 8, "staging layout");
            for (int k = 0; k < Op::NOUT; ++k) {      // outputs: staging 9-11, a 4th shares slot 15
                if (R.rank_mode && R.world > 1) {
                    // rank mode: output staging lives in shared slabs (same sizes, same order on every
                    // rank -- the host API is collective), so the slices are exchanged by peer stores
                    SharedOut& so = dev.ws.shared_out[k];
                    if (so.cap < ntb) {
                        if (so.p) LPM_TRY(free_shared(so.p));
                        so.p = nullptr; so.cap = 0;
                        const size_t want = ntb + ntb / 8 + 256;
                        LPM_TRY(alloc_shared(want, &so.p));
                        so.cap = want;
                    }
                    out[k] = (double*)so.p;
                } else {
                    DevBuf& ob = dev.ws.staging[k < 3 ? 9 + k : 15];
                    LPM_TRY(ob.reserve(ntb));
                    out[k] = ob.as<double>();
                }
            }
            LPM_TRY(build_mask_plan(dev.stream, n, a.mask, dev.ws.plan));
            int64_t b, e;
            load_balance0(nt, nparts, R.rank_mode ? R.rank : (int)g, &b, &e);
            bool exchanged = false;
            LPM_TRY(evaluate<Op>(dev, dev.stream, dev.ws.plan, a, b, e, nt, out, &exchanged));
            if (R.rank_mode) {
                if (!exchanged) LPM_TRY(allgather_slices(Op::NOUT, out, nt, dev.stream));
                b = 0; e = nt;
            }
            for (int k = 0; k < Op::NOUT; ++k)
                if (e > b)
                    LPM_CUDA(cudaMemcpyAsync(out_host[k] + b, out[k] + b, (size_t)(e - b) * sizeof(double),
                                             cudaMemcpyDeviceToHost, dev.stream));
            return LPM_OK;
        };
        rc = body();
    }
    for (auto& dev : R.devs) {
        cudaSetDevice(dev.id);
        cudaError_t e = cudaStreamSynchronize(dev.stream);
        if (e != cudaSuccess && rc == LPM_OK) rc = set_error(LPM_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
    }
    cudaSetDevice(R.devs[0].id);
    return rc;
}

}  // namespace

// ---- BVE ----------------------------------------------------------------------
extern "C" int lpm_bve_velocity(int64_t n, const double* x, const double* y, const double* z, const double* relvort,
                                const double* area, const int32_t* mask, double radius, double* u, double* v, double* w)
{
    Args a{n, {x, y, z, relvort, area}, mask, {radius}};
    double* out[3] = {u, v, w};
    return run_host<OpBveVel>(a, out);
}
extern "C" int lpm_bve_velocity_dev(int64_t n, const double* x, const double* y, const double* z, const double* relvort,
                                    const double* area, const int32_t* mask, double radius, int64_t ibeg, int64_t iend,
                                    double* u, double* v, double* w, void* stream)
{
    Args a{n, {x, y, z, relvort, area}, mask, {radius}};
    double* out[3] = {u, v, w};
    return run_dev<OpBveVel>(a, ibeg, iend, out, stream);
}
extern "C" int lpm_bve_stream(int64_t n, const double* x, const double* y, const double* z, const double* relvort,
                              const double* absvort, const double* area, const int32_t* mask, double radius,
                              double* relstream, double* absstream)
{
    Args a{n, {x, y, z, relvort, absvort, area}, mask, {radius}};
    double* out[2] = {relstream, absstream};
    return run_host<OpBveStream>(a, out);
}
extern "C" int lpm_bve_stream_dev(int64_t n, const double* x, const double* y, const double* z, const double* relvort,
                                  const double* absvort, const double* area, const int32_t* mask, double radius,
                                  int64_t ibeg, int64_t iend, double* relstream, double* absstream, void* stream)
{
    Args a{n, {x, y, z, relvort, absvort, area}, mask, {radius}};
    double* out[2] = {relstream, absstream};
    return run_dev<OpBveStream>(a, ibeg, iend, out, stream);
}

// ---- plane --------------------------------------------------------------------
extern "C" int lpm_plane_velocity(int64_t n, const double* x, const double* y, const double* vort, const double* area,
                                  const int32_t* mask, double* u, double* v)
{
    Args a{n, {x, y, vort, area}, mask, {}};
    double* out[2] = {u, v};
    return run_host<OpPlaneVel>(a, out);
}
extern "C" int lpm_plane_velocity_dev(int64_t n, const double* x, const double* y, const double* vort,
                                      const double* area, const int32_t* mask, int64_t ibeg, int64_t iend, double* u,
                                      double* v, void* stream)
{
    Args a{n, {x, y, vort, area}, mask, {}};
    double* out[2] = {u, v};
    return run_dev<OpPlaneVel>(a, ibeg, iend, out, stream);
}
extern "C" int lpm_plane_stream(int64_t n, const double* x, const double* y, const double* vort, const double* area,
                                const int32_t* mask, double* psi)
{
    Args a{n, {x, y, vort, area}, mask, {}};
    double* out[1] = {psi};
    return run_host<OpPlaneStream>(a, out);
}
extern "C" int lpm_plane_stream_dev(int64_t n, const double* x, const double* y, const double* vort, const double* area,
                                    const int32_t* mask, int64_t ibeg, int64_t iend, double* psi, void* stream)
{
    Args a{n, {x, y, vort, area}, mask, {}};
    double* out[1] = {psi};
    return run_dev<OpPlaneStream>(a, ibeg, iend, out, stream);
}

// ---- beta plane ---------------------------------------------------------------
extern "C" int lpm_betaplane_velocity(int64_t n, const double* x, const double* y, const double* relvort,
                                      const double* area, const int32_t* mask, double* u, double* v)
{
    Args a{n, {x, y, relvort, area}, mask, {}};
    double* out[2] = {u, v};
    return run_host<OpBetaVel>(a, out);
}
extern "C" int lpm_betaplane_velocity_dev(int64_t n, const double* x, const double* y, const double* relvort,
                                          const double* area, const int32_t* mask, int64_t ibeg, int64_t iend,
                                          double* u, double* v, void* stream)
{
    Args a{n, {x, y, relvort, area}, mask, {}};
    double* out[2] = {u, v};
    return run_dev<OpBetaVel>(a, ibeg, iend, out, stream);
}
extern "C" int lpm_betaplane_stream(int64_t n, const double* x, const double* y, const double* relvort,
                                    const double* absvort, const double* area, const int32_t* mask, double* relstream,
                                    double* absstream)
{
    Args a{n, {x, y, relvort, absvort, area}, mask, {}};
    double* out[2] = {relstream, absstream};
    return run_host<OpBetaStream>(a, out);
}
extern "C" int lpm_betaplane_stream_dev(int64_t n, const double* x, const double* y, const double* relvort,
                                        const double* absvort, const double* area, const int32_t* mask, int64_t ibeg,
                                        int64_t iend, double* relstream, double* absstream, void* stream)
{
    Args a{n, {x, y, relvort, absvort, area}, mask, {}};
    double* out[2] = {relstream, absstream};
    return run_dev<OpBetaStream>(a, ibeg, iend, out, stream);
}

// ---- PSE ----------------------------------------------------------------------
extern "C" int lpm_pse_laplacian_sphere(int64_t n, const double* x, const double* y, const double* z, const double* f,
                                        const double* area, const int32_t* mask, double eps, double sphere_radius,
                                        double* lap)
{
    if (!(eps > 0.0) || !(sphere_radius > 0.0)) return set_error(LPM_ERR_INVALID, "eps and sphere_radius must be positive");
    Args a{n, {x, y, z, f, area}, mask, {eps, sphere_radius}};
    double* out[1] = {lap};
    return run_host<OpPseSphere>(a, out);
}
extern "C" int lpm_pse_laplacian_sphere_dev(int64_t n, const double* x, const double* y, const double* z,
                                            const double* f, const double* area, const int32_t* mask, double eps,
                                            double sphere_radius, int64_t ibeg, int64_t iend, double* lap, void* stream)
{
    if (!(eps > 0.0) || !(sphere_radius > 0.0)) return set_error(LPM_ERR_INVALID, "eps and sphere_radius must be positive");
    Args a{n, {x, y, z, f, area}, mask, {eps, sphere_radius}};
    double* out[1] = {lap};
    return run_dev<OpPseSphere>(a, ibeg, iend, out, stream);
}
extern "C" int lpm_pse_laplacian_plane(int64_t n, const double* x, const double* y, const double* f, const double* area,
                                       const int32_t* mask, double eps, double* lap)
{
    if (!(eps > 0.0)) return set_error(LPM_ERR_INVALID, "eps must be positive");
    Args a{n, {x, y, f, area}, mask, {eps}};
    double* out[1] = {lap};
    return run_host<OpPsePlane>(a, out);
}
extern "C" int lpm_pse_laplacian_plane_dev(int64_t n, const double* x, const double* y, const double* f,
                                           const double* area, const int32_t* mask, double eps, int64_t ibeg,
                                           int64_t iend, double* lap, void* stream)
{
    if (!(eps > 0.0)) return set_error(LPM_ERR_INVALID, "eps must be positive");
    Args a{n, {x, y, f, area}, mask, {eps}};
    double* out[1] = {lap};
    return run_dev<OpPsePlane>(a, ibeg, iend, out, stream);
}

// ---- remaining PSE operators (src/PSEDirectSum.f90:128-456, 537-579) -------------
static int check_eps(double eps, double sr)
{
    if (!(eps > 0.0) || !(sr > 0.0)) return set_error(LPM_ERR_INVALID, "eps and sphere_radius must be positive");
    return LPM_OK;
}
extern "C" int lpm_pse_interpolate_sphere(int64_t n, const double* x, const double* y, const double* z, const double* f,
                                          const double* area, const int32_t* mask, double eps, double sphere_radius,
                                          int64_t m, const double* tx, const double* ty, const double* tz, double* out)
{
    LPM_TRY(check_eps(eps, sphere_radius));
    Args a{n, {x, y, z, f, area}, mask, {eps, sphere_radius}};
    a.m = m; a.tgt[0] = tx; a.tgt[1] = ty; a.tgt[2] = tz;
    double* o[1] = {out};
    return run_host<OpPseInterpSphere>(a, o);
}
extern "C" int lpm_pse_interpolate_plane(int64_t n, const double* x, const double* y, const double* f, const double* area,
                                         const int32_t* mask, double eps, int64_t m, const double* tx, const double* ty,
                                         double* out)
{
    LPM_TRY(check_eps(eps, 1.0));
    Args a{n, {x, y, f, area}, mask, {eps}};
    a.m = m; a.tgt[0] = tx; a.tgt[1] = ty;
    double* o[1] = {out};
    return run_host<OpPseInterpPlane>(a, o);
}
extern "C" int lpm_pse_gradient_sphere(int64_t n, const double* x, const double* y, const double* z, const double* f,
                                       const double* area, const int32_t* mask, double eps, double sphere_radius,
                                       double* gx, double* gy, double* gz)
{
    LPM_TRY(check_eps(eps, sphere_radius));
    Args a{n, {x, y, z, f, area}, mask, {eps, sphere_radius}};
    double* o[3] = {gx, gy, gz};
    return run_host<OpPseGradSphere>(a, o);
}
extern "C" int lpm_pse_gradient_plane(int64_t n, const double* x, const double* y, const double* f, const double* area,
                                      const int32_t* mask, double eps, double* gx, double* gy)
{
    LPM_TRY(check_eps(eps, 1.0));
    Args a{n, {x, y, f, area}, mask, {eps}};
    double* o[2] = {gx, gy};
    return run_host<OpPseGradPlane>(a, o);
}
extern "C" int lpm_pse_second_partials_plane(int64_t n, const double* x, const double* y, const double* gx,
                                             const double* gy, const double* area, const int32_t* mask, double eps,
                                             double* dxx, double* dxy, double* dyy)
{
    LPM_TRY(check_eps(eps, 1.0));
    Args a{n, {x, y, gx, gy, area}, mask, {eps}};
    double* o[3] = {dxx, dxy, dyy};
    return run_host<OpPseTensorPlane<0>>(a, o);
}
extern "C" int lpm_pse_double_dot_plane(int64_t n, const double* x, const double* y, const double* u, const double* v,
                                        const double* area, const int32_t* mask, double eps, double* dd)
{
    LPM_TRY(check_eps(eps, 1.0));
    Args a{n, {x, y, u, v, area}, mask, {eps}};
    double* o[1] = {dd};
    return run_host<OpPseTensorPlane<1>>(a, o);
}
extern "C" int lpm_pse_double_dot_sphere(int64_t n, const double* x, const double* y, const double* z, const double* u,
                                         const double* v, const double* w, const double* area, const int32_t* mask,
                                         double eps, double sphere_radius, double* dd)
{
    LPM_TRY(check_eps(eps, sphere_radius));
    Args a{n, {x, y, z, u, v, w, area}, mask, {eps, sphere_radius}};
    double* o[1] = {dd};
    return run_host<OpPseDoubleDotSphere>(a, o);
}
extern "C" int lpm_pse_divergence_sphere(int64_t n, const double* x, const double* y, const double* z, const double* u,
                                         const double* v, const double* w, const double* area, const int32_t* mask,
                                         double eps, double sphere_radius, double* div)
{
    LPM_TRY(check_eps(eps, sphere_radius));
    Args a{n, {x, y, z, u, v, w, area}, mask, {eps, sphere_radius}};
    double* o[1] = {div};
    return run_host<OpPseDivSphere>(a, o);
}

// ---- planar shallow water: SWEPlaneRHSIntegrals (src/SWEPlaneSolver.f90:457-560) ----
extern "C" int lpm_swe_plane_rhs_integrals(int64_t n, const double* x, const double* y, const double* vort,
                                           const double* div, const double* surf, const double* area,
                                           const int32_t* mask, double pse_eps, double* u, double* v,
                                           double* double_dot, double* lap_surf)
{
    LPM_TRY(check_eps(pse_eps, 1.0));
    Args a{n, {x, y, vort, div, surf, area}, mask, {pse_eps}};
    double* o[4] = {u, v, double_dot, lap_surf};
    return run_host<OpSweRhsPlane>(a, o);
}

// ---- planar SWE velocity: SetVelocityFromFieldData (src/PlanarSWE.f90:469-494) ----
extern "C" int lpm_swe_plane_velocity(int64_t n, const double* x, const double* y, const double* vort, const double* div,
                                      const double* area, const int32_t* mask, double* u, double* v)
{
    Args a{n, {x, y, vort, div, area}, mask, {}};
    double* o[2] = {u, v};
    return run_host<OpSwePlaneVel>(a, o);
}

// ---- spherical shallow water: SWESphereRHSIntegrals (src/SphereSWESolver.f90:296-375), as written ----
extern "C" int lpm_swe_sphere_rhs_integrals(int64_t n, const double* x, const double* y, const double* z,
                                            const double* vort, const double* div, const double* surf,
                                            const double* area, const int32_t* mask, double radius, double pse_eps,
                                            double* u, double* v, double* w, double* double_dot, double* lap_surf)
{
    LPM_TRY(check_eps(pse_eps, radius));
    if (!double_dot) return set_error(LPM_ERR_INVALID, "null output array 3");
    Args a{n, {x, y, z, vort, div, surf, area}, mask, {radius, pse_eps}};
    double* o[4] = {u, v, w, lap_surf};
    LPM_TRY(run_host<OpSweRhsSphere>(a, o));
    memset(double_dot, 0, (size_t)n * sizeof(double));      // the reference zeroes it and never accumulates it (:328)
    return LPM_OK;
}

// ============================================================== resident solvers
#include "solvers_api.inc"
