// symmetric.cuh -- pair-symmetric evaluation of the BVE velocity and stream-function sums: the default for
// whole evaluations of large particle sets (sym_applicable below), with order-independent fixed-point
// accumulation.  Measured at icosTri 8 on one B200 (profiles/r02_ab_sym.log, the round-2 A/B): velocity
// 1410 -> 1325 ms, stream functions 2227 -> 1903 ms with the first fixed-point accumulation; the FP64-atomic
// builds (1271 / 1771 ms) were deleted because their results depended on the order in which atomics land.
// End of round 2 (profiles/r02o_bench_n1.json, r02m_ab_builds.log): 1201 ms and 1677 ms.
//
// The reference (src/SphereBVESolver.f90:396-420) visits every ordered pair (i, j): for the
// factored form used here, a_i = sum_j P_j / d_ij with d_ij = R^2 - x_i.x_j = d_ji.  Active
// particles are both targets and sources, so for two active particles c < c' the denominator --
// 3 DFMA -- and its reciprocal -- 3 more -- serve both a_c += P_c' / d and a_c' += P_c / d:
// 12 FP64 instructions for two interactions instead of 18 (stream functions: the denominator and the
// logarithm, 13 instead of 22).  At icosTri 8 two thirds of all interactions are active-active
// (1 310 720^2 of 2.577e12).
//
//   * passive targets (vertices) x all active sources: the one-sided engine (directsum.cuh) on the
//     gathered passive particles;
//   * active x active, on COMPACT indices (the packed source records are also the targets): a CTA
//     owns a block of BLOCK*T compact targets (registers) and a chunk of source tiles at or
//     above its own ("upper triangle").  Diagonal tiles (the block against itself) are
//     evaluated one-sided with the self pair excluded; for every tile above the diagonal each
//     pair is evaluated once: the thread adds P_j / d to its targets' sums and, per source, sums
//     P_t / d over its T targets.  Those per-source sums are reduced across the warp by
//     recursive halving over batches of SB sources (SB (2 SHFL + DADD) per level instead of a full
//     butterfly per source; the lanes take the batch in rotated order so that no level needs a select,
//     sym_reduce_red), the warps' sums are joined in shared memory in warp order (all three kernels; the
//     ones that carry the 64 KB log table use source tiles of 128 / 64 so that the buffers fit), and the
//     result is added to the source's accumulator in global memory -- exactly, in fixed point
//     (sym_red_add), so the total does not depend on the order of the adds nor on how target blocks
//     were dealt to ranks.  A CTA's own sums join the same accumulators when it ends.
//   * u_i = x_i cross a_i in a finalize kernel.
//
// Results are bit-identical from run to run and for any rank count (tests/test_sym_gpu.py,
// tests/test_multigpu.py); they differ from the one-sided path's by summation order only.
#pragma once
#include "ops.cuh"
#include "sym_kernels.cuh"

namespace lpm {

// Below this many active particles the triangular grid is too few CTAs to fill a B200 and the one-sided
// engine (whose targets-per-thread adapts) is faster: icosTri 7 (327 680) gains 5 %, icosTri 6 (81 920) loses.
constexpr int32_t kSymMinSources = 200000;
constexpr int32_t kSymChunkTiles = 16;      // source tiles (of 256) per CTA of the triangle kernel
constexpr int32_t kSymPanelBlocks = 256;    // target blocks per panel of its launch order (sym_kernel)

template <class K, int T, int BLOCK, int SB, int MINB, int ORDER = 0, bool COMBINE = false, int TS = kTile>
inline int launch_sym(cudaStream_t st, const SymParams& prm, SymGeom g, const double* src, double* acc)
{
    constexpr int TB = BLOCK * T;
    g.nblocks = (g.nsrc_pad + TB - 1) / TB;
    g.ntiles *= kTile / TS;             // the caller counts tiles of kTile; the kernel tiles of TS
    g.chunk_tiles *= kTile / TS;
    g.half_bin = 1 << (19 - kLogBits);
    g.panel_blocks = std::min<int32_t>(g.nblocks, rt().sym_panel_blocks);
    const int64_t npanels = (g.nblocks + g.panel_blocks - 1) / g.panel_blocks;
    constexpr size_t smem = sym_smem_bytes<K, T, BLOCK, COMBINE, TS>();
    static_assert(smem <= 227 * 1024, "shared memory of one CTA");
    const int64_t grid = npanels * g.panel_blocks * g.nchunks;
    if (grid <= 0 || grid > 0x7fffffffLL) return set_error(LPM_ERR_INVALID, "symmetric kernel grid %lld", (long long)grid);
    if (smem > 48 * 1024) {     // per device, as in launch_ds
        static bool configured[64] = {};
        int devid = 0;
        cudaGetDevice(&devid);
        if (devid < 0 || devid >= 64 || !configured[devid]) {
            LPM_CUDA(cudaFuncSetAttribute(sym_kernel<K, T, BLOCK, SB, MINB, ORDER, COMBINE, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (devid >= 0 && devid < 64) configured[devid] = true;
        }
    }
    sym_kernel<K, T, BLOCK, SB, MINB, ORDER, COMBINE, TS><<<(unsigned)grid, BLOCK, smem, st>>>(prm, g, src, acc);
    return LPM_OK;
}

// What differs between the two symmetric sums.
struct SymVel {
    using Op = OpBveVel;
    using SK = SymBveVel;
    static constexpr int NCOORD = 3;
    static constexpr int FX_MODE = 0;       // fixed-point window from F max|P| / (R^2 2^-110)
    static constexpr int FX_MODE2 = -1;     // no second window
    static void sym_params(SymParams& p, const Args& a) { p.R2 = a.sc[0] * a.sc[0]; }
    static void passive_params(BveVel::Params& p, const double* const* xyz, const Args& a)
    {
        p.x = xyz[0]; p.y = xyz[1]; p.z = xyz[2];
        p.R2 = a.sc[0] * a.sc[0];
    }
    // 8 targets per thread, batches of 8 sources taken in the lane-rotated order of sym_reduce_red (two groups of 4, phase by
    // phase: statement order 11), the warps' source sums combined in shared memory: one fixed-point add per (CTA, source,
    // component).  246 registers, two CTAs of 128 threads per SM.  Measured at icosTri 8, the triangle kernel alone:
    // 729-732 ms (profiles/r02m_ab_builds.log; with a scheduling fence after each group, order 43: 737; orders 27 / 59: 733 /
    // 738).  Before the rotated order (selects in the warp reduction; profiles/r02b_ab_paths.log, r02c_ab_paths.log,
    // r02e_order_sweep.log): order 43 765-769 ms, unfenced 790; batches of 4: 824; one source at a time, fenced: 848; warps
    // not combined: 869; 6 targets per thread at 2 / 3 CTAs per SM: 837 / 859; 4 targets, 3 CTAs: 881.
    static int launch(cudaStream_t st, const SymParams& prm, const SymGeom& g, const double* src, double* acc)
    {
#ifdef LPM_SYM_ORDER_SWEEP      // tools/order_sweep.py: one instantiation per statement order behind lpm_tune("sym_vel_order")
        switch (rt().sym_vel_order) {
#define LPM_O(n) case n: return launch_sym<SK, 8, 128, 8, 1, n, true>(st, prm, g, src, acc);
            LPM_O(0) LPM_O(1) LPM_O(2) LPM_O(3) LPM_O(4) LPM_O(5) LPM_O(6) LPM_O(7) LPM_O(8) LPM_O(9) LPM_O(10) LPM_O(11)
            LPM_O(16) LPM_O(17) LPM_O(18) LPM_O(19) LPM_O(20) LPM_O(21) LPM_O(22) LPM_O(23) LPM_O(24) LPM_O(25) LPM_O(26)
            LPM_O(32 + 8) LPM_O(32 + 9) LPM_O(32 + 10) LPM_O(32 + 11) LPM_O(32 + 24) LPM_O(32 + 25) LPM_O(32 + 26) LPM_O(32 + 27)
#undef LPM_O
            default: break;
        }
#endif
        return launch_sym<SK, 8, 128, 8, 1, 11, true>(st, prm, g, src, acc);
    }
    static void finalize(cudaStream_t st, const MaskPlan& mp, const double* src, const double* acc, const Outs<3>& out)
    {
        sym_bve_finalize<<<(unsigned)((mp.nsrc + 255) / 256), 256, 0, st>>>(mp.nsrc, mp.active.as<int32_t>(), src, acc, out);
    }
};
struct SymStream {
    using Op = OpBveStream;
    using SK = SymBveStream;
    static constexpr int NCOORD = 3;
    static constexpr int FX_MODE = 1;       // fixed-point window from F max|w| 2^12
    static constexpr int FX_MODE2 = -1;
    static void sym_params(SymParams& p, const Args& a) { p.R2 = a.sc[0] * a.sc[0]; }
    static void passive_params(BveStream::Params& p, const double* const* xyz, const Args& a)
    {
        p.x = xyz[0]; p.y = xyz[1]; p.z = xyz[2];
        p.R2 = a.sc[0] * a.sc[0];
    }
    // 64 KB table per CTA, two CTAs per SM: 8 targets per thread, 128 threads, batches of 8 sources in the rotated order, a
    // retry branch per source, and source tiles of 128 so that the combine buffers (16 KB) fit beside the table: 933 ms for
    // the triangle at icosTri 8 (profiles/r02m_ab_builds.log).  The same with selects in the reduction: 945; batches of 4:
    // 963; 4 targets per thread at 256 threads (128-register cap), combined: 1010 (rotated: 986; batches of 8, which spill:
    // 1035; tiles of 64: 1008; one CTA of 512 threads: 1015); not combined (round 2's first default): 1134 (rotated: 1125).
    static int launch(cudaStream_t st, const SymParams& prm, const SymGeom& g, const double* src, double* acc)
    {
        return launch_sym<SK, 8, 128, 8, 2, 1, true, 128>(st, prm, g, src, acc);
    }
    static void finalize(cudaStream_t st, const MaskPlan& mp, const double*, const double* acc, const Outs<2>& out)
    {
        sym_stream_finalize<<<(unsigned)((mp.nsrc + 255) / 256), 256, 0, st>>>(mp.nsrc, mp.active.as<int32_t>(), acc, out);
    }
};

// The end of an RK4 step (src/SphereBVESolver.f90:345-352): velocity and stream functions of the new state in one
// pass (SymBveVelStream / BveVelStream).  4 targets per thread, 128 threads, batches of 4 sources in the rotated order, a
// retry branch per source; 64 KB log table + source tiles of 64 (8 KB) + 20 KB of combine buffers per CTA: two CTAs per
// SM.  Triangle kernel at icosTri 8 (profiles/r02m_ab_builds.log): 1690 ms; with selects in the reduction 1803; not
// combined, tiles of 256 (round 2's first version) 1947 (rotated: 1836); one CTA of 256 threads per SM, tiles of 128,
// combined 1831 (rotated: 1705), not combined 2090.
struct SymVelStream {
    using Op = OpBveVelStream;
    using SK = SymBveVelStream;
    static constexpr int NCOORD = 3;
    static constexpr int FX_MODE = 0;       // components 0-2: the velocity window
    static constexpr int FX_MODE2 = 1;      // components 3-4: the stream-function window
    static void sym_params(SymParams& p, const Args& a) { p.R2 = a.sc[0] * a.sc[0]; }
    static void passive_params(BveVelStream::Params& p, const double* const* xyz, const Args& a)
    {
        p.x = xyz[0]; p.y = xyz[1]; p.z = xyz[2];
        p.R2 = a.sc[0] * a.sc[0];
    }
    static int launch(cudaStream_t st, const SymParams& prm, const SymGeom& g, const double* src, double* acc)
    {
        return launch_sym<SK, 4, 128, 4, 2, 0, true, 64>(st, prm, g, src, acc);
    }
    static void finalize(cudaStream_t st, const MaskPlan& mp, const double* src, const double* acc, const Outs<5>& out)
    {
        sym_bve_velstream_finalize<<<(unsigned)((mp.nsrc + 255) / 256), 256, 0, st>>>(mp.nsrc, mp.active.as<int32_t>(), src, acc, out);
    }
};

// Whole evaluation: one device, or -- rank mode -- collectively on every rank (each must call with its
// LoadBalance slice, sym_applicable() checks that): target blocks of the active x active part are dealt
// round-robin to the ranks and the accumulators summed with one ncclAllReduce (NC F doubles); the passive
// targets are sliced by LoadBalance and exchanged with the grouped broadcast of allgather_slices().
// Every rank then writes ALL n results into its own `out` (no peer stores), so the caller has nothing
// left to exchange.  a: the Op's arguments; out: where the results go (replica 0 only is used).
template <class S>
inline int sym_evaluate(Device& dev, cudaStream_t st, MaskPlan& mp, const Args& a,
                        const Outs<S::Op::NOUT>& out)
{
    using Op = typename S::Op;
    using K = typename Op::K;
    using SK = typename S::SK;
    constexpr int NOUT = Op::NOUT;
    static_assert(SK::NC == NOUT && (int)K::NS == (int)SK::NS, "the symmetric functor reads the one-sided kernel's records");
    Runtime& R = rt();
    Workspace& ws = dev.ws;
    const bool prof = R.profiling;
    auto body = [&]() -> int {
        LPM_TRY(Op::pack(dev, st, mp, a));          // records, and the log window for the stream functions
        const double* src = ws.sources.as<double>();
        SymGeom g{};
        int32_t chunk = 0;
        ds_chunks(mp.nsrc, &g.nsrc_pad, &chunk, &g.nchunks, mp.n);
        g.nsrc = mp.nsrc;
        g.ntiles = g.nsrc_pad / kTile;
        // The triangle kernel keeps no partial sums, so its chunks can be much shorter than the one-sided
        // engine's (whose chunk count is capped by its scratch): kSymChunkTiles tiles per CTA keep the last wave of
        // CTAs -- the tail -- below 1 % of a rank's share even on 8 GPUs (icosTri 8, 8 ranks: 101.3 ms with chunks
        // of 80 tiles, 17 waves of 4 ms CTAs).  Depends on the active count only, like ds_chunks.
        g.chunk_tiles = std::min<int32_t>(chunk / kTile, rt().sym_chunk_tiles);
        g.nchunks = (g.ntiles + g.chunk_tiles - 1) / g.chunk_tiles;
        g.world = R.rank_mode ? R.world : 1;
        g.rank = R.rank_mode ? R.rank : 0;
        SymParams prm{};
        S::sym_params(prm, a);
        if constexpr (SK::KS > 0) {
            prm.logtab = dev.logtab;
            prm.win = ws.logwin.as<int32_t>() + 1;
        }
        // ---- active x active
        if (mp.nsrc > 0) {
            const size_t nacc = (size_t)g.nsrc_pad * SK::NC;
            const size_t acc_bytes = nacc * kFxWords * sizeof(long long);
            LPM_TRY(ws.sym_acc.reserve(acc_bytes));
            LPM_CUDA(cudaMemsetAsync(ws.sym_acc.p, 0, acc_bytes, st));
            // the window of the fixed-point accumulators, from the largest entry of the records
            LPM_TRY(ws.sym_fx.reserve(2 * sizeof(FxWindow) + sizeof(int32_t)));
            LPM_TRY(ws.sym_acc2.reserve(nacc * sizeof(double)));
            FxWindow* fxw = ws.sym_fx.as<FxWindow>();
            int32_t* maxhi = reinterpret_cast<int32_t*>(fxw + 2);
            LPM_CUDA(cudaMemsetAsync(maxhi, 0, sizeof(int32_t), st));
            const int64_t nrec = (int64_t)g.nsrc_pad * SK::NS;
            const unsigned nb = (unsigned)std::min<int64_t>((nrec + 255) / 256, 4 * (int64_t)dev.sm_count);
            absmax_hi_kernel<<<nb, 256, 0, st>>>(nrec, src, nullptr, maxhi);
            sym_fx_scale_kernel<<<1, 1, 0, st>>>(S::FX_MODE, prm.R2, mp.nsrc, maxhi, fxw);
            if constexpr (S::FX_MODE2 >= 0)     // second component class (the fused sums' stream functions)
                sym_fx_scale_kernel<<<1, 1, 0, st>>>(S::FX_MODE2, prm.R2, mp.nsrc, maxhi, fxw + 1);
            prm.fx = fxw;
            cudaEvent_t pb = nullptr, pe = nullptr;
            if (prof) {     // the triangle kernel alone; the passive part's ds_kernel records its own pair
                if (dev.next_prof(&pb, &pe, 2 * prof_sum_of<K>::value + 1) != LPM_OK) return set_error(LPM_ERR_CUDA, "cannot create profiling events");
                LPM_CUDA(cudaEventRecord(pb, st));
            }
            LPM_TRY(S::launch(st, prm, g, src, ws.sym_acc.as<double>()));
            if (prof) LPM_CUDA(cudaEventRecord(pe, st));
            if (g.world > 1) {
                // The integer all-reduce of the accumulators (7 x 8 B per sum: 220 MB at icosTri 8) runs on the device's
                // communication stream while the passive part below computes on `st`; the conversion and the finalize
                // kernel wait for it at the end.  NCCL sees the same call order on every rank: this all-reduce, then
                // the grouped broadcast of the passive slices.
                if (!R.comm) return set_error(LPM_ERR_COMM, "world size %d but no communicator (lpm_comm_init_rank)", R.world);
                LPM_TRY(dev.ensure_comm_stream());
                LPM_CUDA(cudaEventRecord(dev.ev_comm[0], st));
                LPM_CUDA(cudaStreamWaitEvent(dev.comm_stream, dev.ev_comm[0], 0));
                LPM_NCCL(nccl().AllReduce(ws.sym_acc.p, ws.sym_acc.p, nacc * kFxWords, /*ncclInt64*/ 4, /*ncclSum*/ 0, R.comm, dev.comm_stream));
                LPM_CUDA(cudaEventRecord(dev.ev_comm[1], dev.comm_stream));
            }
        }
        // ---- passive targets x all active sources: the one-sided engine on the gathered passive particles
        const int64_t nv = mp.n - mp.nsrc;
        if (nv > 0) {
            LPM_TRY(ws.sorted_targets.reserve((size_t)nv * sizeof(int32_t)));
            int32_t* perm = ws.sorted_targets.as<int32_t>();
            passive_list_kernel<<<(unsigned)((mp.n + 255) / 256), 256, 0, st>>>(mp.n, mp.scan.as<int32_t>(), perm);
            LPM_TRY(ws.sort_vals[0].reserve((size_t)(nv + 1) * sizeof(int32_t)));       // "scan" of a list with no source in it
            LPM_CUDA(cudaMemsetAsync(ws.sort_vals[0].p, 0, (size_t)(nv + 1) * sizeof(int32_t), st));
            const unsigned gb = (unsigned)((nv + 255) / 256);
            const double* coords[3] = {nullptr, nullptr, nullptr};
            for (int k = 0; k < S::NCOORD; ++k) {       // the coordinates are the first NCOORD input arrays of every Op here
                LPM_TRY(ws.gathered[k].reserve((size_t)nv * sizeof(double)));
                gather_kernel<<<gb, 256, 0, st>>>(nv, perm, a.in[k], ws.gathered[k].as<double>());
                coords[k] = ws.gathered[k].as<double>();
            }
            count_launch(1 + S::NCOORD);
            typename K::Params prm1{};
            S::passive_params(prm1, coords, a);
            prm1.out.nrep = 1;
            double* bufs[NOUT];
            for (int k = 0; k < NOUT; ++k) {
                LPM_TRY(ws.sorted_out[k].reserve((size_t)nv * sizeof(double)));
                bufs[k] = ws.sorted_out[k].as<double>();
                prm1.out.p[0][k] = bufs[k];
            }
            MaskPlan view;                          // not owned: the sources of mp, no self pairs
            view.n = nv; view.nsrc = mp.nsrc;
            view.scan.p = ws.sort_vals[0].p; view.scan.cap = ws.sort_vals[0].cap;
            view.active.p = mp.active.p; view.active.cap = mp.active.cap;
            int64_t vb = 0, ve = nv;
            if (g.world > 1) load_balance0(nv, g.world, g.rank, &vb, &ve);
            const int rc = direct_sum<K>(dev, st, view, vb, ve, prm1, nv);
            view.scan = DevBuf{}; view.active = DevBuf{};
            LPM_TRY(rc);
            if (g.world > 1) LPM_TRY(allgather_slices(NOUT, bufs, nv, st));
            for (int k = 0; k < NOUT; ++k) {
                ScatterDst dst{};
                dst.nrep = 1;
                dst.p[0] = out.p[0][k];
                scatter_kernel<<<gb, 256, 0, st>>>(nv, perm, bufs[k], dst);
            }
            count_launch(NOUT);
        }
        // ---- active x active, second half: limbs -> doubles, cross product / copy into `out`
        if (mp.nsrc > 0) {
            const size_t nacc = (size_t)g.nsrc_pad * SK::NC;
            if (g.world > 1) LPM_CUDA(cudaStreamWaitEvent(st, dev.ev_comm[1], 0));
            constexpr int kSplit = S::FX_MODE2 >= 0 ? 3 : SK::NC;      // SK::window(): components >= 3 of the fused sums
            static_assert(SK::window(SK::NC - 1) == (S::FX_MODE2 >= 0 ? 1 : 0), "window classes");
            sym_fx_to_double_kernel<<<(unsigned)((nacc + 255) / 256), 256, 0, st>>>((int64_t)nacc, ws.sym_acc.as<long long>(),
                                                                                 ws.sym_fx.as<FxWindow>(), ws.sym_acc2.as<double>(),
                                                                                 SK::NC, kSplit);
            S::finalize(st, mp, src, ws.sym_acc2.as<double>(), out);
            count_launch(5);
        }
        LPM_CUDA(cudaGetLastError());
        return LPM_OK;
    };
    return body();
}

// May this evaluation take the symmetric path?  The path is on (lpm_set_symmetric), the particle set is large
// enough, and the call covers all targets: one device driving every target, or rank mode with this rank's
// LoadBalance slice (then every rank reaches the same answer and the collectives inside match up).
inline bool sym_applicable(int64_t tbeg, int64_t tend, int64_t nt, const MaskPlan& mp, int nrep)
{
    const Runtime& R = rt();
    if (!R.symmetric || mp.nsrc < R.sym_min_sources || nt != mp.n || nrep != 1 || R.devs.size() != 1) return false;
    int64_t b = 0, e = nt;
    if (R.rank_mode && R.world > 1) load_balance0(nt, R.world, R.rank, &b, &e);
    return tbeg == b && tend == e;
}

// Which sums have a symmetric form.
template <class Op> struct SymFor { using type = void; };
template <> struct SymFor<OpBveVel> { using type = SymVel; };
template <> struct SymFor<OpBveStream> { using type = SymStream; };
template <> struct SymFor<OpBveVelStream> { using type = SymVelStream; };

template <class Op>
inline int sym_try(Device& dev, cudaStream_t st, MaskPlan& mp, const Args& a, double* const* out, int64_t tbeg,
                   int64_t tend, int64_t nt, int nrep, bool* taken)
{
    using S = typename SymFor<Op>::type;
    *taken = false;
    if constexpr (!std::is_void<S>::value) {
        if (!sym_applicable(tbeg, tend, nt, mp, nrep)) return LPM_OK;
        Outs<Op::NOUT> o{};
        set_outs(o, out);
        LPM_TRY(sym_evaluate<S>(dev, st, mp, a, o));
        *taken = true;
    }
    return LPM_OK;
}

}  // namespace lpm
