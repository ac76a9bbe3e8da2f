// emu_lib.cpp -- TEST INFRASTRUCTURE ONLY (see emu_core.h): the experimental pair-symmetric kernels of
// lpm_v2_b200/csrc/sym_kernels.cuh, and the one-sided engine they lean on for the passive targets,
// compiled with g++ and run on CPU threads.  The host-side steps mirror csrc/symmetric.cuh
// (sym_evaluate): pack, zero the accumulators, symmetric kernel (once per emulated rank, into the same
// accumulators = the all-reduce), finalize, passive list, gather, one-sided kernel on a "view" whose
// scan is all zero, scatter.
#ifndef LPM_CUDA_EMU
#define LPM_CUDA_EMU 1
#endif
#include "cuda_runtime.h"

#include "sym_kernels.cuh"

using namespace lpm;

namespace {

struct Plan {
    std::vector<int32_t> scan, active, passive;
};
Plan make_plan(int64_t n, const int32_t* mask)
{
    Plan p;
    p.scan.resize(n + 1);
    int32_t c = 0;
    for (int64_t i = 0; i < n; ++i) {
        p.scan[i] = c;
        if (mask[i] != 0) { p.active.push_back((int32_t)i); ++c; }
    }
    p.scan[n] = c;
    // passive_list_kernel
    p.passive.assign(n - c, -1);
    const unsigned nb = (unsigned)((n + 255) / 256);
    emu_launch_seq(nb, 256, [&]() { passive_list_kernel(n, p.scan.data(), p.passive.data()); });
    return p;
}

void build_log_table()
{
    static bool done = false;
    if (done) return;
    for (int idx = 0; idx < kLogFull; ++idx) {
        const double q = log_bin_rcp(idx << (20 - kLogBits));
        g_log_full[idx] = (q > 0.0 && std::isfinite(q)) ? (double)(-logl((long double)q)) : 0.0;
    }
    done = true;
}

template <class SK, int T, int BLOCK, int SB, int ORDER>
void run_sym(const SymParams& prm, SymGeom g, const double* src, double* acc)
{
    constexpr int TB = BLOCK * T;
    g.nblocks = (g.nsrc_pad + TB - 1) / TB;
    g.half_bin = 1 << (19 - kLogBits);
    const int world = g.world;
    for (int r = 0; r < world; ++r) {            // every emulated rank adds into the same accumulators
        g.rank = r;
        emu_launch((unsigned)(g.nblocks * g.nchunks), BLOCK, [&]() { sym_kernel<SK, T, BLOCK, SB, 1, ORDER>(prm, g, src, acc); });
    }
}

// the one-sided engine on gathered targets with an all-zero scan (symmetric.cuh, passive part)
template <class K, int T, int BLOCK, int U>
void run_one_sided(typename K::Params prm, int64_t nv, int32_t nsrc, int32_t nsrc_pad, const double* src)
{
    DsGeom g{};
    g.tbeg = 0; g.tend = nv; g.ntgt = nv; g.nall = nv;
    g.nsrc = nsrc; g.nsrc_pad = nsrc_pad; g.chunk = nsrc_pad; g.nchunks = 1;
    g.tblk0 = 0;
    g.ntblocks = (int32_t)((nv + BLOCK * T - 1) / (BLOCK * T));
    g.bounds = nullptr;
    g.half_bin = 1 << (19 - kLogBits);
    std::vector<int32_t> zero_scan(nv + 1, 0);
    emu_launch((unsigned)g.ntblocks, BLOCK, [&]() { ds_kernel<K, T, BLOCK, U, 1>(prm, g, src, zero_scan.data(), nullptr); });
}

}  // namespace

// shape = lpm_set_bve_variant value - 200 (csrc/symmetric.cuh)
extern "C" __attribute__((visibility("default"))) int emu_sym_bve_velocity(int64_t n, const double* x, const double* y, const double* z, const double* zeta,
                                    const double* area, const int32_t* mask, double R, int shape, int chunk_tiles,
                                    int world, double* u, double* v, double* w)
{
    Plan pl = make_plan(n, mask);
    const int32_t nsrc = (int32_t)pl.active.size();
    int32_t pad = (nsrc + kTile - 1) / kTile * kTile;
    if (pad == 0) pad = kTile;
    std::vector<double> src((size_t)pad * 6);
    emu_launch_seq((unsigned)((pad + 255) / 256), 256,
                   [&]() { pack_bve_vel(nsrc, pad, pl.active.data(), x, y, z, zeta, area, R, src.data()); });
    SymGeom g{};
    g.nsrc = nsrc; g.nsrc_pad = pad; g.ntiles = pad / kTile;
    g.chunk_tiles = chunk_tiles;
    g.nchunks = (g.ntiles + chunk_tiles - 1) / chunk_tiles;
    g.world = world;
    SymParams prm{};
    prm.R2 = R * R;
    Outs<3> out{};
    out.nrep = 1;
    out.p[0][0] = u; out.p[0][1] = v; out.p[0][2] = w;
    if (nsrc > 0) {
        std::vector<double> acc((size_t)pad * 3, 0.0);
        switch (shape) {        // as SymVel::launch in csrc/symmetric.cuh
            case 1: run_sym<SymBveVel, 8, 128, 4, 35>(prm, g, src.data(), acc.data()); break;
            case 2: run_sym<SymBveVel, 4, 128, 8, 27>(prm, g, src.data(), acc.data()); break;
            case 3: run_sym<SymBveVel, 8, 128, 4, 27>(prm, g, src.data(), acc.data()); break;
            default: run_sym<SymBveVel, 4, 128, 8, 35>(prm, g, src.data(), acc.data()); break;
        }
        emu_launch_seq((unsigned)((nsrc + 255) / 256), 256,
                       [&]() { sym_bve_finalize(nsrc, pl.active.data(), src.data(), acc.data(), out); });
    }
    const int64_t nv = n - nsrc;
    if (nv > 0) {
        std::vector<double> gx(nv), gy(nv), gz(nv), ou(nv), ov(nv), ow(nv);
        for (int64_t c = 0; c < nv; ++c) { gx[c] = x[pl.passive[c]]; gy[c] = y[pl.passive[c]]; gz[c] = z[pl.passive[c]]; }
        BveVel::Params p1{};
        p1.x = gx.data(); p1.y = gy.data(); p1.z = gz.data();
        p1.R2 = R * R;
        p1.out.nrep = 1;
        p1.out.p[0][0] = ou.data(); p1.out.p[0][1] = ov.data(); p1.out.p[0][2] = ow.data();
        run_one_sided<BveVel, 4, 128, 2>(p1, nv, nsrc, pad, src.data());
        for (int64_t c = 0; c < nv; ++c) { u[pl.passive[c]] = ou[c]; v[pl.passive[c]] = ov[c]; w[pl.passive[c]] = ow[c]; }
    }
    return 0;
}

// shape = lpm_set_bve_variant value - 200
extern "C" __attribute__((visibility("default"))) int emu_sym_bve_stream(int64_t n, const double* x, const double* y, const double* z, const double* zeta,
                                  const double* omega, const double* area, const int32_t* mask, double R, int shape,
                                  int chunk_tiles, int world, double* rel, double* abs_)
{
    build_log_table();
    Plan pl = make_plan(n, mask);
    const int32_t nsrc = (int32_t)pl.active.size();
    int32_t pad = (nsrc + kTile - 1) / kTile * kTile;
    if (pad == 0) pad = kTile;
    std::vector<double> src((size_t)pad * 6);
    emu_launch_seq((unsigned)((pad + 255) / 256), 256,
                   [&]() { pack_bve_stream(nsrc, pad, pl.active.data(), x, y, z, zeta, omega, area, R, src.data()); });
    int32_t win[2] = {0, 0};
    log_window_kernel(0, 2.0 * R * R, win, SymBveStream::WINDOW_BINADES, win + 1);
    SymGeom g{};
    g.nsrc = nsrc; g.nsrc_pad = pad; g.ntiles = pad / kTile;
    g.chunk_tiles = chunk_tiles;
    g.nchunks = (g.ntiles + chunk_tiles - 1) / chunk_tiles;
    g.world = world;
    SymParams prm{};
    prm.R2 = R * R;
    prm.logtab = g_log_full;
    prm.win = win + 1;
    Outs<2> out{};
    out.nrep = 1;
    out.p[0][0] = rel; out.p[0][1] = abs_;
    if (nsrc > 0) {
        std::vector<double> acc((size_t)pad * 2, 0.0);
        switch (shape) {        // as SymStream::launch in csrc/symmetric.cuh
            case 1: run_sym<SymBveStream, 4, 128, 4, 0>(prm, g, src.data(), acc.data()); break;
            case 2: run_sym<SymBveStream, 4, 256, 4, 0>(prm, g, src.data(), acc.data()); break;
            case 3: run_sym<SymBveStream, 4, 128, 4, 1>(prm, g, src.data(), acc.data()); break;
            default: run_sym<SymBveStream, 4, 256, 4, 1>(prm, g, src.data(), acc.data()); break;
        }
        emu_launch_seq((unsigned)((nsrc + 255) / 256), 256,
                       [&]() { sym_stream_finalize(nsrc, pl.active.data(), acc.data(), out); });
    }
    const int64_t nv = n - nsrc;
    if (nv > 0) {
        std::vector<double> gx(nv), gy(nv), gz(nv), o0(nv), o1(nv);
        for (int64_t c = 0; c < nv; ++c) { gx[c] = x[pl.passive[c]]; gy[c] = y[pl.passive[c]]; gz[c] = z[pl.passive[c]]; }
        BveStream::Params p1{};
        p1.x = gx.data(); p1.y = gy.data(); p1.z = gz.data();
        p1.R2 = R * R;
        p1.logtab = g_log_full;
        p1.win = win + 1;
        p1.out.nrep = 1;
        p1.out.p[0][0] = o0.data(); p1.out.p[0][1] = o1.data();
        run_one_sided<BveStream, 4, 256, 2>(p1, nv, nsrc, pad, src.data());
        for (int64_t c = 0; c < nv; ++c) { rel[pl.passive[c]] = o0[c]; abs_[pl.passive[c]] = o1[c]; }
    }
    return 0;
}

// The default (one-sided) BVE velocity kernel over all targets, for comparison: ds_kernel<BveVel> with the real scan.
extern "C" __attribute__((visibility("default"))) int emu_default_bve_velocity(int64_t n, const double* x, const double* y, const double* z, const double* zeta,
                                        const double* area, const int32_t* mask, double R, double* u, double* v, double* w)
{
    Plan pl = make_plan(n, mask);
    const int32_t nsrc = (int32_t)pl.active.size();
    int32_t pad = (nsrc + kTile - 1) / kTile * kTile;
    if (pad == 0) pad = kTile;
    std::vector<double> src((size_t)pad * 6);
    emu_launch_seq((unsigned)((pad + 255) / 256), 256,
                   [&]() { pack_bve_vel(nsrc, pad, pl.active.data(), x, y, z, zeta, area, R, src.data()); });
    BveVel::Params p1{};
    p1.x = x; p1.y = y; p1.z = z;
    p1.R2 = R * R;
    p1.out.nrep = 1;
    p1.out.p[0][0] = u; p1.out.p[0][1] = v; p1.out.p[0][2] = w;
    constexpr int T = 4, BLOCK = 128;
    DsGeom g{};
    g.tbeg = 0; g.tend = n; g.ntgt = n; g.nall = n;
    g.nsrc = nsrc; g.nsrc_pad = pad; g.chunk = pad; g.nchunks = 1;
    g.tblk0 = 0;
    g.ntblocks = (int32_t)((n + BLOCK * T - 1) / (BLOCK * T));
    g.half_bin = 1 << (19 - kLogBits);
    emu_launch((unsigned)g.ntblocks, BLOCK, [&]() { ds_kernel<BveVel, T, BLOCK, 2, 1>(p1, g, src.data(), pl.scan.data(), nullptr); });
    return 0;
}

// shape = lpm_set_bve_variant value - 200 (SymPlane::launch in csrc/symmetric.cuh)
extern "C" __attribute__((visibility("default"))) int emu_sym_plane_velocity(int64_t n, const double* x, const double* y, const double* vort,
                                      const double* area, const int32_t* mask, int shape, int chunk_tiles, int world,
                                      double* u, double* v)
{
    Plan pl = make_plan(n, mask);
    const int32_t nsrc = (int32_t)pl.active.size();
    int32_t pad = (nsrc + kTile - 1) / kTile * kTile;
    if (pad == 0) pad = kTile;
    std::vector<double> src((size_t)pad * 4);
    const double inv_norm = 1.0 / (2.0 * LPM_PI);
    emu_launch_seq((unsigned)((pad + 255) / 256), 256,
                   [&]() { pack_plane(nsrc, pad, pl.active.data(), x, y, vort, area, inv_norm, src.data()); });
    SymGeom g{};
    g.nsrc = nsrc; g.nsrc_pad = pad; g.ntiles = pad / kTile;
    g.chunk_tiles = chunk_tiles;
    g.nchunks = (g.ntiles + chunk_tiles - 1) / chunk_tiles;
    g.world = world;
    SymParams prm{};
    Outs<2> out{};
    out.nrep = 1;
    out.p[0][0] = u; out.p[0][1] = v;
    if (nsrc > 0) {
        std::vector<double> acc((size_t)pad * 2, 0.0);
        switch (shape) {
            case 1: run_sym<SymPlaneVel, 8, 128, 4, 1>(prm, g, src.data(), acc.data()); break;
            case 2: run_sym<SymPlaneVel, 4, 128, 8, 0>(prm, g, src.data(), acc.data()); break;
            case 3: run_sym<SymPlaneVel, 8, 128, 4, 0>(prm, g, src.data(), acc.data()); break;
            default: run_sym<SymPlaneVel, 4, 128, 8, 1>(prm, g, src.data(), acc.data()); break;
        }
        emu_launch_seq((unsigned)((nsrc + 255) / 256), 256,
                       [&]() { sym_stream_finalize(nsrc, pl.active.data(), acc.data(), out); });
    }
    const int64_t nv = n - nsrc;
    if (nv > 0) {
        std::vector<double> gx(nv), gy(nv), o0(nv), o1(nv);
        for (int64_t c = 0; c < nv; ++c) { gx[c] = x[pl.passive[c]]; gy[c] = y[pl.passive[c]]; }
        PlaneVel::Params p1{};
        p1.x = gx.data(); p1.y = gy.data();
        p1.out.nrep = 1;
        p1.out.p[0][0] = o0.data(); p1.out.p[0][1] = o1.data();
        run_one_sided<PlaneVel, 4, 128, 1>(p1, nv, nsrc, pad, src.data());
        for (int64_t c = 0; c < nv; ++c) { u[pl.passive[c]] = o0[c]; v[pl.passive[c]] = o1[c]; }
    }
    return 0;
}
