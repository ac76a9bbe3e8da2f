// solvers.cuh -- device-resident mirrors of the reference's RK4 solver types:
//   BVESolver        src/SphereBVESolver.f90:38-72, New :112-168, Timestep :219-353
//   PlaneSolver      src/PlaneIncompressibleSolver.f90:37-59, Timestep :171-259
//   BetaPlaneSolver  src/BetaPlaneSolver.f90:36-56, Timestep :142-219
// All state stays in HBM for the life of the solver; one Timestep is 4 direct
// sums (3 stage evaluations + the velocity at the new state) and optionally
// the stream-function sum, with the O(N) stage arithmetic between them done
// by the small kernels below.  Those kernels use the round-to-nearest
// intrinsics in the reference's operation order (no FMA contraction), so given
// the same stage velocities the RK4 update is bit-identical to the Fortran.
//
// Multi-GPU.  Every device holds a full replica of the state (the reference's
// replicated-data model, src/MPISetup.f90).  A direct sum is evaluated for the
// device's LoadBalance slice only; the results reach the other replicas
//   - single-process mode: by peer stores from the finalize step (Outs::nrep
//     replicas over NVLink), followed by an event barrier between the streams;
//   - rank mode: the arrays of every rank live in one CUDA-IPC shared slab each, and the
//     finalize step stores to all of them over NVLink between two stream-ordered barriers
//     (runtime.cuh); the grouped NCCL broadcast of allgather_slices() is the fallback.
// The O(N) stage arithmetic is then repeated identically on every replica.
#pragma once
#include <memory>

#include "ops.cuh"

namespace lpm {

// ---------------------------------------------------------------- O(N) kernels
#define LPM_GRID_STRIDE(i, n) \
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

// stage = dt * vel                                   (e.g. SphereBVESolver.f90:240-242, :264-266)
__global__ void rk_scale(int64_t n, double dt, double* __restrict__ a)
{
    LPM_GRID_STRIDE(i, n) a[i] = __dmul_rn(dt, a[i]);
}
__global__ void rk_scale_to(int64_t n, double dt, const double* __restrict__ v, double* __restrict__ a)
{
    LPM_GRID_STRIDE(i, n) a[i] = __dmul_rn(dt, v[i]);
}
// BVE vorticity tendency: - dt * 2 * Omega * w / R     (SphereBVESolver.f90:243, :263)
__global__ void rk_bve_vort(int64_t n, double dt, double Omega, double R, const double* __restrict__ w,
                            double* __restrict__ vs)
{
    const double c = __dmul_rn(__dmul_rn(-dt, 2.0), Omega);
    LPM_GRID_STRIDE(i, n) vs[i] = __ddiv_rn(__dmul_rn(c, w[i]), R);
}
// beta-plane vorticity tendency: - dt * beta * v        (BetaPlaneSolver.f90:154, :170)
__global__ void rk_beta_vort(int64_t n, double dt, double beta, const double* __restrict__ v, double* __restrict__ vs)
{
    const double c = __dmul_rn(-dt, beta);
    LPM_GRID_STRIDE(i, n) vs[i] = __dmul_rn(c, v[i]);
}
// in = start + c * stage, c = 0.5 or 1                  (SphereBVESolver.f90:250-255, :296-301)
__global__ void rk_input(int64_t n, double c, const double* __restrict__ start, const double* __restrict__ stage,
                         double* __restrict__ in)
{
    LPM_GRID_STRIDE(i, n) in[i] = __dadd_rn(start[i], __dmul_rn(c, stage[i]));
}
// start += s1/6 + s2/3 + s3/3 + s4/6, left to right    (SphereBVESolver.f90:320-329)
__global__ void rk_update(int64_t n, double* __restrict__ start, const double* __restrict__ s1,
                          const double* __restrict__ s2, const double* __restrict__ s3, const double* __restrict__ s4)
{
    LPM_GRID_STRIDE(i, n)
    {
        double v = start[i];
        v = __dadd_rn(v, __ddiv_rn(s1[i], 6.0));
        v = __dadd_rn(v, __ddiv_rn(s2[i], 3.0));
        v = __dadd_rn(v, __ddiv_rn(s3[i], 3.0));
        v = __dadd_rn(v, __ddiv_rn(s4[i], 6.0));
        start[i] = v;
    }
}

// TotalKE / TotalEnstrophy partial sums (src/SphereBVE.f90:410-441); fixed-shape
// tree, so the value does not depend on the launch.
__global__ void __launch_bounds__(256) diag_partial(int64_t n, const double* __restrict__ u, const double* __restrict__ v,
                                                    const double* __restrict__ w, const double* __restrict__ zeta,
                                                    const double* __restrict__ area, const int32_t* __restrict__ mask,
                                                    double* __restrict__ part)
{
    __shared__ double sk[256], se[256];
    double ke = 0.0, en = 0.0;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t b = (int64_t)blockIdx.x * per;
    int64_t e = b + per;
    if (e > n) e = n;
    for (int64_t i = b + threadIdx.x; i < e; i += 256)
        if (mask[i]) {
            ke += (u[i] * u[i] + v[i] * v[i] + w[i] * w[i]) * area[i];
            en += zeta[i] * zeta[i] * area[i];
        }
    sk[threadIdx.x] = ke; se[threadIdx.x] = en;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sk[threadIdx.x] += sk[threadIdx.x + s]; se[threadIdx.x] += se[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = sk[0]; part[2 * blockIdx.x + 1] = se[0]; }
}
__global__ void diag_final(int nb, const double* __restrict__ part, double* __restrict__ out)
{
    double ke = 0.0, en = 0.0;
    for (int b = 0; b < nb; ++b) { ke += part[2 * b]; en += part[2 * b + 1]; }
    out[0] = 0.5 * ke; out[1] = 0.5 * en;
}

inline unsigned ew_grid(int64_t n, int sm_count)
{
    int64_t nb = (n + 255) / 256;
    int64_t cap = (int64_t)sm_count * 16;
    return (unsigned)std::max<int64_t>(1, std::min(nb, cap));
}

// ---------------------------------------------------------------- replicas
// The per-device copy of a solver's state: `narr` arrays of n doubles + mask.
struct Replica {
    Device* dev = nullptr;
    std::vector<DevBuf> a;
    DevBuf mask;
    MaskPlan mp;
    int64_t sb = 0, se = 0;       // this replica's LoadBalance slice [sb, se)
    double* A(int k) const { return a[k].as<double>(); }
};

struct SolverBase {
    int64_t n = 0;
    std::vector<Replica> reps;
    void* slab = nullptr;         // rank mode: all arrays live in one shared slab (peer-store exchange)

    int alloc(int64_t n_, int narr, const int32_t* mask_host)
    {
        Runtime& R = rt();
        LPM_TRY(require_init());
        if (n_ <= 0) return set_error(LPM_ERR_INVALID, "n = %lld", (long long)n_);
        n = n_;
        const int nparts = R.rank_mode ? R.world : (int)R.devs.size();
        reps.resize(R.devs.size());
        for (size_t g = 0; g < R.devs.size(); ++g) {
            Replica& r = reps[g];
            r.dev = &R.devs[g];
            LPM_CUDA(cudaSetDevice(r.dev->id));
            r.a.resize(narr);
            if (R.rank_mode && R.world > 1) {
                // one slab every rank can store into (COLLECTIVE: solvers are created by all ranks together)
                const size_t stride = ((size_t)n * sizeof(double) + 255) / 256 * 256;
                LPM_TRY(alloc_shared(stride * narr, &slab));
                LPM_CUDA(cudaMemsetAsync(slab, 0, stride * narr, r.dev->stream));
                for (int k = 0; k < narr; ++k) { r.a[k].p = (char*)slab + stride * k; r.a[k].cap = 0; }
            } else {
                for (auto& b : r.a) {
                    LPM_TRY(b.reserve((size_t)n * sizeof(double)));
                    LPM_CUDA(cudaMemsetAsync(b.p, 0, (size_t)n * sizeof(double), r.dev->stream));
                }
            }
            LPM_TRY(r.mask.reserve((size_t)n * sizeof(int32_t)));
            LPM_CUDA(cudaMemcpyAsync(r.mask.p, mask_host, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, r.dev->stream));
            LPM_TRY(build_mask_plan(r.dev->stream, n, r.mask.as<int32_t>(), r.mp));
            load_balance0(n, nparts, R.rank_mode ? R.rank : (int)g, &r.sb, &r.se);
        }
        return LPM_OK;
    }
    int upload(int k, const double* host)
    {
        for (auto& r : reps) {
            LPM_CUDA(cudaSetDevice(r.dev->id));
            if (host)
                LPM_CUDA(cudaMemcpyAsync(r.A(k), host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, r.dev->stream));
        }
        return LPM_OK;
    }
    int download(int k, double* host)
    {
        if (!host) return LPM_OK;
        Replica& r = reps[0];
        LPM_CUDA(cudaSetDevice(r.dev->id));
        LPM_CUDA(cudaMemcpyAsync(host, r.A(k), (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, r.dev->stream));
        return LPM_OK;
    }
    int sync()
    {
        int rc = LPM_OK;
        for (auto& r : reps) {
            cudaSetDevice(r.dev->id);
            cudaError_t e = cudaStreamSynchronize(r.dev->stream);
            if (e != cudaSuccess && rc == LPM_OK) rc = set_error(LPM_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
        }
        if (!reps.empty()) cudaSetDevice(reps[0].dev->id);
        return rc;
    }
    void free_all()
    {
        for (auto& r : reps) {
            cudaSetDevice(r.dev->id);
            cudaStreamSynchronize(r.dev->stream);
            if (slab)
                for (auto& b : r.a) b.p = nullptr;      // views into the slab
            for (auto& b : r.a) b.release();
            r.mask.release();
            r.mp.release();
        }
        reps.clear();
        if (slab) { free_shared(slab); slab = nullptr; }
    }

    // One direct sum over all replicas: inputs are array indices `in` (Op order),
    // outputs array indices `out`; afterwards every replica holds all n results.
    template <class Op>
    int eval(const int* in, const int* out, const double* sc)
    {
        Runtime& R = rt();
        const int nrep = (int)reps.size();
        bool exchanged = false;
        for (int g = 0; g < nrep; ++g) {
            Replica& r = reps[g];
            LPM_CUDA(cudaSetDevice(r.dev->id));
            Args a{};
            a.n = n;
            for (int k = 0; k < Op::NIN; ++k) a.in[k] = r.A(in[k]);
            a.mask = r.mask.as<int32_t>();
            for (int k = 0; k < 3; ++k) a.sc[k] = sc ? sc[k] : 0.0;
            {   // large BVE sums: the whole sum in pair-symmetric form (symmetric.cuh)
                double* o[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
                static_assert(Op::NOUT <= 8, "outputs per sum");
                for (int k = 0; k < Op::NOUT; ++k) o[k] = r.A(out[k]);
                bool taken = false;
                LPM_TRY(sym_try<Op>(*r.dev, r.dev->stream, r.mp, a, o, r.sb, r.se, n, nrep, &taken));
                if (taken) {
                    exchanged = true;       // rank mode: every rank already holds all n results
                    continue;
                }
            }
            LPM_TRY(Op::pack(*r.dev, r.dev->stream, r.mp, a));
            typename Op::K::Params prm = Op::params(a);
            if (R.rank_mode) {                   // one replica per process: the peers' copies through the shared slab
                double* o[8];
                for (int k = 0; k < Op::NOUT; ++k) o[k] = r.A(out[k]);
                exchanged = set_outs_shared(prm.out, o, n);
                if (exchanged && r.se <= r.sb) {
                    LPM_TRY(comm_barrier(*r.dev, r.dev->stream));
                    LPM_TRY(comm_barrier(*r.dev, r.dev->stream));
                }
            } else {
                prm.out.nrep = nrep;
                for (int q = 0; q < nrep; ++q)       // own replica first, then the peers
                    for (int k = 0; k < Op::NOUT; ++k) prm.out.p[q][k] = reps[(g + q) % nrep].A(out[k]);
            }
            LPM_TRY(direct_sum<typename Op::K>(*r.dev, r.dev->stream, r.mp, r.sb, r.se, prm));
        }
        if (nrep > 1) {
            for (auto& r : reps) {
                LPM_CUDA(cudaSetDevice(r.dev->id));
                LPM_CUDA(cudaEventRecord(r.dev->ev_done, r.dev->stream));
            }
            for (int g = 0; g < nrep; ++g)
                for (int h = 0; h < nrep; ++h)
                    if (h != g) LPM_CUDA(cudaStreamWaitEvent(reps[g].dev->stream, reps[h].dev->ev_done, 0));
        } else if (R.rank_mode && R.world > 1 && !exchanged) {
            double* bufs[8];
            for (int k = 0; k < Op::NOUT; ++k) bufs[k] = reps[0].A(out[k]);
            LPM_TRY(allgather_slices(Op::NOUT, bufs, n, reps[0].dev->stream));
        }
        return LPM_OK;
    }

    // run an O(N) kernel on every replica
    template <class F>
    int each(F&& f)
    {
        for (auto& r : reps) {
            LPM_CUDA(cudaSetDevice(r.dev->id));
            count_launch(f(r, ew_grid(n, r.dev->sm_count), r.dev->stream));
        }
        LPM_CUDA(cudaGetLastError());
        return LPM_OK;
    }
};

// ---------------------------------------------------------------- BVE
struct BveSolver : SolverBase {
    enum { X, Y, Z, ZETA, U, V, W, AREA, ABSV, XIN, YIN, ZIN, VIN,
           XS1, YS1, ZS1, VS1, XS2, YS2, ZS2, VS2, XS3, YS3, ZS3, VS3, XS4, YS4, ZS4, VS4, RELS, ABSS, NARR };
    double R = 1.0, Omega = 0.0;
    bool has_absvort = false;

    int velocity(int xi, int yi, int zi, int vi, int ox, int oy, int oz)
    {
        const int in[5] = {xi, yi, zi, vi, AREA};
        const int out[3] = {ox, oy, oz};
        const double sc[3] = {R, 0, 0};
        return eval<OpBveVel>(in, out, sc);
    }

    // src/SphereBVESolver.f90:219-353
    int timestep(double dt, bool with_stream)
    {
        // stage 1 (:239-244)
        LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
            rk_scale_to<<<g, 256, 0, s>>>(n, dt, r.A(U), r.A(XS1));
            rk_scale_to<<<g, 256, 0, s>>>(n, dt, r.A(V), r.A(YS1));
            rk_scale_to<<<g, 256, 0, s>>>(n, dt, r.A(W), r.A(ZS1));
            rk_bve_vort<<<g, 256, 0, s>>>(n, dt, Omega, R, r.A(W), r.A(VS1));
            return 4;
        }));
        for (int st = 1; st < 4; ++st) {            // stages 2-4 (:250-313)
            const double c = st < 3 ? 0.5 : 1.0;
            const int prev = XS1 + 4 * (st - 1), cur = XS1 + 4 * st;
            LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
                rk_input<<<g, 256, 0, s>>>(n, c, r.A(X), r.A(prev + 0), r.A(XIN));
                rk_input<<<g, 256, 0, s>>>(n, c, r.A(Y), r.A(prev + 1), r.A(YIN));
                rk_input<<<g, 256, 0, s>>>(n, c, r.A(Z), r.A(prev + 2), r.A(ZIN));
                rk_input<<<g, 256, 0, s>>>(n, c, r.A(ZETA), r.A(prev + 3), r.A(VIN));
                return 4;
            }));
            LPM_TRY(velocity(XIN, YIN, ZIN, VIN, cur + 0, cur + 1, cur + 2));
            LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
                rk_bve_vort<<<g, 256, 0, s>>>(n, dt, Omega, R, r.A(cur + 2), r.A(cur + 3));
                rk_scale<<<g, 256, 0, s>>>(n, dt, r.A(cur + 0));
                rk_scale<<<g, 256, 0, s>>>(n, dt, r.A(cur + 1));
                rk_scale<<<g, 256, 0, s>>>(n, dt, r.A(cur + 2));
                return 4;
            }));
        }
        LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {      // :320-329
            for (int c = 0; c < 4; ++c)
                rk_update<<<g, 256, 0, s>>>(n, r.A(X + c), r.A(XS1 + c), r.A(XS2 + c), r.A(XS3 + c), r.A(XS4 + c));
            return 4;
        }));
        if (with_stream && rt().fuse_step_end) {
            // :345-346 and :352 in one pass over the pairs: the velocity and the stream functions of the new state
            // share the denominator R^2 - x_i.x_j (BveVelStream / SymBveVelStream)
            if (!has_absvort) return set_error(LPM_ERR_INVALID, "stream functions need absvort (pass it to lpm_bve_solver_new)");
            const int in[6] = {X, Y, Z, ZETA, ABSV, AREA};
            const int out[5] = {U, V, W, RELS, ABSS};
            const double sc[3] = {R, 0, 0};
            LPM_TRY(eval<OpBveVelStream>(in, out, sc));
            return sync();
        }
        LPM_TRY(velocity(X, Y, Z, ZETA, U, V, W));                       // :345-346
        if (with_stream) LPM_TRY(stream());                              // :352
        return sync();
    }

    // src/SphereBVE.f90:445-485
    int stream()
    {
        if (!has_absvort) return set_error(LPM_ERR_INVALID, "stream functions need absvort (pass it to lpm_bve_solver_new)");
        const int in[6] = {X, Y, Z, ZETA, ABSV, AREA};
        const int out[2] = {RELS, ABSS};
        const double sc[3] = {R, 0, 0};
        return eval<OpBveStream>(in, out, sc);
    }

    int diagnostics(double* ke, double* ens)
    {
        Replica& r = reps[0];
        LPM_CUDA(cudaSetDevice(r.dev->id));
        const int nb = 1024;
        LPM_TRY(r.dev->ws.reduce.reserve((2 * nb + 2) * sizeof(double)));
        double* part = r.dev->ws.reduce.as<double>();
        diag_partial<<<nb, 256, 0, r.dev->stream>>>(n, r.A(U), r.A(V), r.A(W), r.A(ZETA), r.A(AREA), r.mask.as<int32_t>(), part);
        diag_final<<<1, 1, 0, r.dev->stream>>>(nb, part, part + 2 * nb);
        count_launch(2);
        double h[2];
        LPM_CUDA(cudaMemcpyAsync(h, part + 2 * nb, sizeof(h), cudaMemcpyDeviceToHost, r.dev->stream));
        LPM_CUDA(cudaStreamSynchronize(r.dev->stream));
        if (ke) *ke = h[0];
        if (ens) *ens = h[1];
        return LPM_OK;
    }
};

// ---------------------------------------------------------------- plane
struct PlaneSolverDev : SolverBase {
    enum { X, Y, VORT, U, V, AREA, XIN, YIN, XS1, YS1, XS2, YS2, XS3, YS3, XS4, YS4, PSI, NARR };

    int velocity(int xi, int yi, int ox, int oy)
    {
        const int in[4] = {xi, yi, VORT, AREA};
        const int out[2] = {ox, oy};
        return eval<OpPlaneVel>(in, out, nullptr);
    }
    // src/PlaneIncompressibleSolver.f90:171-259
    int timestep(double dt, bool with_stream)
    {
        LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
            rk_scale_to<<<g, 256, 0, s>>>(n, dt, r.A(U), r.A(XS1));
            rk_scale_to<<<g, 256, 0, s>>>(n, dt, r.A(V), r.A(YS1));
            return 2;
        }));
        for (int st = 1; st < 4; ++st) {
            const double c = st < 3 ? 0.5 : 1.0;
            const int prev = XS1 + 2 * (st - 1), cur = XS1 + 2 * st;
            LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
                rk_input<<<g, 256, 0, s>>>(n, c, r.A(X), r.A(prev + 0), r.A(XIN));
                rk_input<<<g, 256, 0, s>>>(n, c, r.A(Y), r.A(prev + 1), r.A(YIN));
                return 2;
            }));
            LPM_TRY(velocity(XIN, YIN, cur + 0, cur + 1));
            LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
                rk_scale<<<g, 256, 0, s>>>(n, dt, r.A(cur + 0));
                rk_scale<<<g, 256, 0, s>>>(n, dt, r.A(cur + 1));
                return 3;
            }));
        }
        LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
            for (int c = 0; c < 2; ++c)
                rk_update<<<g, 256, 0, s>>>(n, r.A(X + c), r.A(XS1 + c), r.A(XS2 + c), r.A(XS3 + c), r.A(XS4 + c));
            return 2;
        }));
        LPM_TRY(velocity(X, Y, U, V));
        if (with_stream) {                          // :258 -> src/PlanarIncompressible.f90:470-505
            const int in[4] = {X, Y, VORT, AREA};
            const int out[1] = {PSI};
            LPM_TRY(eval<OpPlaneStream>(in, out, nullptr));
        }
        return sync();
    }
};

// ---------------------------------------------------------------- beta plane
struct BetaSolverDev : SolverBase {
    enum { X, Y, ZETA, U, V, AREA, ABSV, XIN, YIN, VIN, XS1, YS1, VS1, XS2, YS2, VS2, XS3, YS3, VS3, XS4, YS4, VS4,
           RELS, ABSS, NARR };
    double beta = 0.0;
    bool has_absvort = false;

    int velocity(int xi, int yi, int vi, int ox, int oy)
    {
        const int in[4] = {xi, yi, vi, AREA};
        const int out[2] = {ox, oy};
        return eval<OpBetaVel>(in, out, nullptr);
    }
    // src/BetaPlaneSolver.f90:142-219
    int timestep(double dt, bool with_stream)
    {
        LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
            rk_scale_to<<<g, 256, 0, s>>>(n, dt, r.A(U), r.A(XS1));
            rk_scale_to<<<g, 256, 0, s>>>(n, dt, r.A(V), r.A(YS1));
            rk_beta_vort<<<g, 256, 0, s>>>(n, dt, beta, r.A(V), r.A(VS1));
            return 3;
        }));
        for (int st = 1; st < 4; ++st) {
            const double c = st < 3 ? 0.5 : 1.0;
            const int prev = XS1 + 3 * (st - 1), cur = XS1 + 3 * st;
            LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
                rk_input<<<g, 256, 0, s>>>(n, c, r.A(X), r.A(prev + 0), r.A(XIN));
                rk_input<<<g, 256, 0, s>>>(n, c, r.A(Y), r.A(prev + 1), r.A(YIN));
                rk_input<<<g, 256, 0, s>>>(n, c, r.A(ZETA), r.A(prev + 2), r.A(VIN));
                return 3;
            }));
            LPM_TRY(velocity(XIN, YIN, VIN, cur + 0, cur + 1));
            LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
                rk_beta_vort<<<g, 256, 0, s>>>(n, dt, beta, r.A(cur + 1), r.A(cur + 2));
                rk_scale<<<g, 256, 0, s>>>(n, dt, r.A(cur + 0));
                rk_scale<<<g, 256, 0, s>>>(n, dt, r.A(cur + 1));
                return 3;
            }));
        }
        LPM_TRY(each([&](Replica& r, unsigned g, cudaStream_t s) {
            for (int c = 0; c < 3; ++c)
                rk_update<<<g, 256, 0, s>>>(n, r.A(X + c), r.A(XS1 + c), r.A(XS2 + c), r.A(XS3 + c), r.A(XS4 + c));
            return 3;
        }));
        LPM_TRY(velocity(X, Y, ZETA, U, V));        // :217 SetVelocityOnMesh
        if (with_stream) {                          // :218 -> src/BetaPlane.f90:399-442
            if (!has_absvort) return set_error(LPM_ERR_INVALID, "stream functions need absvort");
            const int in[5] = {X, Y, ZETA, ABSV, AREA};
            const int out[2] = {RELS, ABSS};
            LPM_TRY(eval<OpBetaStream>(in, out, nullptr));
        }
        return sync();
    }
};

// ---------------------------------------------------------------- planar shallow water
// type SWESolver + Timestep, src/SWEPlaneSolver.f90:47-110, 298-429.  Six prognostic arrays per particle
// (x, y, relVort, div, area, h), one fused right-hand-side sum per stage (SWEPlaneRHSIntegrals :457-565 ->
// OpSweRhsPlane: velocity, double dot product, PSE Laplacian of the surface), classical RK4.
//
// Stage tendencies with round-to-nearest intrinsics in the reference's operation order (:337-346):
//   relVort' = dt * ( -(zeta + f0 + beta y) * delta - beta * v )
//   div'     = dt * ( -doubleDot + (f0 + beta y) * zeta - g * lapSurf )
//   h'       = dt * ( -h * delta ),   area' = dt * ( area * delta ),   x' = dt u,  y' = dt v
__global__ void swe_stage(int64_t n, double dt, double f0, double beta, double g, const double* __restrict__ y,
                          const double* __restrict__ rv, const double* __restrict__ dv, const double* __restrict__ area,
                          const double* __restrict__ h, const double* __restrict__ u, const double* __restrict__ v,
                          const double* __restrict__ dd, const double* __restrict__ lap, double* __restrict__ xs,
                          double* __restrict__ ys, double* __restrict__ rvs, double* __restrict__ dvs,
                          double* __restrict__ as, double* __restrict__ hs, int whole_array_stage1)
{
    LPM_GRID_STRIDE(i, n)
    {
        xs[i] = __dmul_rn(dt, u[i]);
        ys[i] = __dmul_rn(dt, v[i]);
        hs[i] = __dmul_rn(dt, __dmul_rn(-h[i], dv[i]));
        as[i] = __dmul_rn(dt, __dmul_rn(area[i], dv[i]));
        // Stage 1 of the reference assigns these two to the WHOLE arrays inside its particle loop (:312-315, no
        // `(i)`), so every entry ends up with the LAST particle's value: k = n - 1 reproduces that.
        const int64_t k = whole_array_stage1 ? n - 1 : i;
        const double fy = __dadd_rn(f0, __dmul_rn(beta, y[k]));
        const double s = __dadd_rn(__dadd_rn(rv[k], f0), __dmul_rn(beta, y[k]));
        rvs[i] = __dmul_rn(dt, __dsub_rn(__dmul_rn(-s, dv[k]), __dmul_rn(beta, v[k])));
        dvs[i] = __dmul_rn(dt, __dsub_rn(__dadd_rn(-dd[k], __dmul_rn(fy, rv[k])), __dmul_rn(g, lap[k])));
    }
}
// surface height = h + topography (topo == nullptr: flat bottom, h + 0)
__global__ void swe_surface(int64_t n, const double* __restrict__ h, const double* __restrict__ topo, double* __restrict__ surf)
{
    LPM_GRID_STRIDE(i, n) surf[i] = __dadd_rn(h[i], topo ? topo[i] : 0.0);
}

typedef double (*TopoFn)(double x, double y, void* user);

struct SwePlaneSolverDev : SolverBase {
    // start state, right-hand side, stage inputs, then the four stages of the six prognostic arrays
    enum { X, Y, RV, DIV, AREA, H, U, V, DD, LAP, SURF, TOPO, XIN, YIN, RVIN, DIVIN, AREAIN, HIN, S1, NARR = S1 + 24 };
    double f0 = 0.0, beta = 0.0, g = 0.0, eps = 0.0;
    TopoFn topo = nullptr;
    void* topo_user = nullptr;
    std::vector<double> hx, hy, ht;

    // SWEPlaneRHSIntegrals at the state held in arrays (xi ... hi) -> U, V, DD, LAP
    int rhs(int xi, int yi, int rvi, int divi, int areai, int hi)
    {
        if (topo) {
            // the reference calls topoFn(x, y) inside its pair loop (:483, :490); here once per particle and stage,
            // on the host (a user-supplied Fortran / C function): two O(N) downloads and one upload per evaluation
            hx.resize((size_t)n); hy.resize((size_t)n); ht.resize((size_t)n);
            LPM_TRY(download(xi, hx.data()));
            LPM_TRY(download(yi, hy.data()));
            LPM_TRY(sync());
            for (int64_t i = 0; i < n; ++i) ht[(size_t)i] = topo(hx[(size_t)i], hy[(size_t)i], topo_user);
            LPM_TRY(upload(TOPO, ht.data()));
            LPM_TRY(sync());        // ht is reused by the next evaluation
        }
        LPM_TRY(each([&](Replica& r, unsigned gr, cudaStream_t s) {
            swe_surface<<<gr, 256, 0, s>>>(n, r.A(hi), topo ? r.A(TOPO) : nullptr, r.A(SURF));
            return 1;
        }));
        const int in[6] = {xi, yi, rvi, divi, SURF, areai};
        const int out[4] = {U, V, DD, LAP};
        const double sc[3] = {eps, 0, 0};
        return eval<OpSweRhsPlane>(in, out, sc);
    }

    int stage(int s, double dt, int yi, int rvi, int divi, int areai, int hi, bool whole_array)
    {
        const int b = S1 + 6 * s;
        return each([&](Replica& r, unsigned gr, cudaStream_t st) {
            swe_stage<<<gr, 256, 0, st>>>(n, dt, f0, beta, g, r.A(yi), r.A(rvi), r.A(divi), r.A(areai), r.A(hi), r.A(U), r.A(V),
                                         r.A(DD), r.A(LAP), r.A(b + 0), r.A(b + 1), r.A(b + 2), r.A(b + 3), r.A(b + 4), r.A(b + 5),
                                         whole_array ? 1 : 0);
            return 1;
        });
    }

    // src/SWEPlaneSolver.f90:298-429
    int timestep(double dt)
    {
        LPM_TRY(stage(0, dt, Y, RV, DIV, AREA, H, true));                     // :309-320
        for (int s = 1; s < 4; ++s) {
            const double c = s < 3 ? 0.5 : 1.0;
            const int prev = S1 + 6 * (s - 1);
            LPM_TRY(each([&](Replica& r, unsigned gr, cudaStream_t st) {       // :325-332, :350-357, :375-382
                for (int k = 0; k < 6; ++k) rk_input<<<gr, 256, 0, st>>>(n, c, r.A(X + k), r.A(prev + k), r.A(XIN + k));
                return 6;
            }));
            LPM_TRY(rhs(XIN, YIN, RVIN, DIVIN, AREAIN, HIN));
            LPM_TRY(stage(s, dt, YIN, RVIN, DIVIN, AREAIN, HIN, false));       // :337-346
        }
        LPM_TRY(each([&](Replica& r, unsigned gr, cudaStream_t st) {           // :400-413
            for (int k = 0; k < 6; ++k)
                rk_update<<<gr, 256, 0, st>>>(n, r.A(X + k), r.A(S1 + k), r.A(S1 + 6 + k), r.A(S1 + 12 + k), r.A(S1 + 18 + k));
            return 6;
        }));
        LPM_TRY(rhs(X, Y, RV, DIV, AREA, H));                                  // :416-417
        return sync();
    }
};

}  // namespace lpm
