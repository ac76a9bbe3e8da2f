// ops.cuh -- one description per direct sum: how to pack the active particles
// into source records and how to fill the kernel Params from device arrays.
// `in[]` are the input arrays in include/lpm_gpu.h order (after n); all device
// pointers.  Shared by the one-shot entry points and the resident solvers.
#pragma once
#include "runtime.cuh"
#include "pairs_pse.cuh"
#include "pairs_swe.cuh"

namespace lpm {

struct Args {
    int64_t n;
    const double* in[8];
    const int32_t* mask;
    double sc[3];           // scalar arguments (radius / eps / sphere radius)
    int64_t m = 0;          // separate target locations (interpolation): count and arrays
    const double* tgt[3] = {nullptr, nullptr, nullptr};
};

inline unsigned pack_grid(int32_t nsrc_pad) { return (unsigned)((nsrc_pad + 255) / 256); }

template <int NO>
inline void set_outs(Outs<NO>& o, double* const* out)
{
    o.nrep = 1;
    for (int k = 0; k < NO; ++k) o.p[0][k] = out[k];
}

// ---- BVE velocity: in = x y z relvort area; sc = radius; out = u v w
struct OpBveVel {
    using K = BveVel;
    static constexpr int NIN = 5, NOUT = 3, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_bve_vel<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                     a.in[3], a.in[4], a.sc[0], dev.ws.sources.as<double>());
        count_launch();
        return LPM_OK;
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.z = a.in[2];
        p.R2 = a.sc[0] * a.sc[0];
        return p;
    }
};

// ---- BVE stream: in = x y z relvort absvort area; sc = radius; out = relstream absstream
struct OpBveStream {
    using K = BveStream;
    static constexpr int NIN = 6, NOUT = 2, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_bve_stream<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                        a.in[3], a.in[4], a.in[5], a.sc[0], dev.ws.sources.as<double>());
        count_launch();
        return log_window<K>(dev, st, 0, 2.0 * a.sc[0] * a.sc[0], 0, nullptr, nullptr);     // d <= 2 R^2
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.z = a.in[2];
        p.R2 = a.sc[0] * a.sc[0];
        return p;
    }
};

// ---- BVE velocity + stream functions fused (the end of an RK4 step): in = x y z relvort absvort area; sc = radius;
//      out = u v w relstream absstream
struct OpBveVelStream {
    using K = BveVelStream;
    static constexpr int NIN = 6, NOUT = 5, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_bve_velstream<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                           a.in[3], a.in[4], a.in[5], a.sc[0], dev.ws.sources.as<double>());
        count_launch();
        return log_window<K>(dev, st, 0, 2.0 * a.sc[0] * a.sc[0], 0, nullptr, nullptr);     // d <= 2 R^2
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.z = a.in[2];
        p.R2 = a.sc[0] * a.sc[0];
        return p;
    }
};

// ---- planar velocity / stream: in = x y vort area
template <class KK, bool STREAM>
struct OpPlane {
    using K = KK;
    static constexpr int NIN = 4, NOUT = STREAM ? 1 : 2, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        const double inv_norm = STREAM ? 1.0 / (4.0 * LPM_PI) : 1.0 / (2.0 * LPM_PI);
        pack_plane<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                   a.in[3], inv_norm, dev.ws.sources.as<double>());
        count_launch();
        if constexpr (STREAM) return log_window<K>(dev, st, 1, 0.0, a.n, a.in[0], a.in[1]);
        return LPM_OK;
    }
    static typename K::Params params(const Args& a)
    {
        typename K::Params p{};
        p.x = a.in[0]; p.y = a.in[1];
        return p;
    }
};
using OpPlaneVel = OpPlane<PlaneVel, false>;
using OpPlaneStream = OpPlane<PlaneStream, true>;

// ---- beta-plane velocity: in = x y relvort area;  stream: in = x y relvort absvort area
struct OpBetaVel {
    using K = BetaVel;
    static constexpr int NIN = 4, NOUT = 2, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_beta<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                  nullptr, a.in[3], 0, dev.ws.sources.as<double>());
        count_launch();
        return LPM_OK;
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1];
        return p;
    }
};
struct OpBetaStream {
    using K = BetaStream;
    static constexpr int NIN = 5, NOUT = 2, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_beta<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                  a.in[3], a.in[4], 1, dev.ws.sources.as<double>());
        count_launch();
        return log_window<K>(dev, st, 2, 0.0, a.n, a.in[1], nullptr);
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1];
        return p;
    }
};

// ---- PSE sphere: in = x y z f area; sc = eps, sphere_radius
struct OpPseSphere {
    using K = PseSphere;
    static constexpr int NIN = 5, NOUT = 1, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_pse_sphere<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                        a.in[3], a.in[4], a.sc[0], dev.ws.sources.as<double>());
        count_launch();
        return LPM_OK;
    }
    static K::Params params(const Args& a);
};

// ---- PSE plane: in = x y f area; sc = eps
struct OpPsePlane {
    using K = PsePlane;
    static constexpr int NIN = 4, NOUT = 1, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_pse_plane<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                       a.in[3], a.sc[0], dev.ws.sources.as<double>());
        count_launch();
        return LPM_OK;
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.f = a.in[2];
        p.inv_eps2 = 1.0 / (a.sc[0] * a.sc[0]);
        return p;
    }
};

// ============================================================== remaining PSE operators
inline PseSphereConsts pse_sphere_consts(double eps, double sr, double scale)
{
    PseSphereConsts c;
    c.rad_over_eps = sr / eps;
    const double theta_cut = kPseCut * eps / sr;
    c.cos_cut = (theta_cut < LPM_PI) ? cos(theta_cut) : -2.0;
    c.chord_cut = sphere_chord_cut(eps, sr);
    c.scale = scale;
    c.k2_scale = 4.0 * c.rad_over_eps * c.rad_over_eps;
    c.nterms = rt().pse_series ? atan_sq_terms(theta_cut) : 0;
    return c;
}
inline OpPseSphere::K::Params OpPseSphere::params(const Args& a)
{
    K::Params p{};
    p.x = a.in[0]; p.y = a.in[1]; p.z = a.in[2]; p.f = a.in[3];
    p.c = pse_sphere_consts(a.sc[0], a.sc[1], 1.0 / (a.sc[0] * a.sc[0]));
    return p;
}

// ---- interpolation (sphere): in = x y z f area; sc = eps, sphere_radius; targets tgt[0..2] (m of them)
struct OpPseInterpSphere {
    using K = PseInterpSphere;
    static constexpr int NIN = 5, NOUT = 1, NTGT = 3;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_pse_interp<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                        a.in[3], a.in[4], a.sc[0], 1, dev.ws.sources.as<double>());
        count_launch();
        return LPM_OK;
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.tgt[0]; p.y = a.tgt[1]; p.z = a.tgt[2];
        p.c = pse_sphere_consts(a.sc[0], a.sc[1], 1.0);
        return p;
    }
};
// ---- interpolation (plane): in = x y f area; sc = eps; targets tgt[0..1]
struct OpPseInterpPlane {
    using K = PseInterpPlane;
    static constexpr int NIN = 4, NOUT = 1, NTGT = 2;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_pse_interp<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], nullptr,
                                                        a.in[2], a.in[3], a.sc[0], 0, dev.ws.sources.as<double>());
        count_launch();
        return LPM_OK;
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.tgt[0]; p.y = a.tgt[1];
        p.inv_eps2 = 1.0 / (a.sc[0] * a.sc[0]);
        return p;
    }
};

template <class KK>
inline int pack_generic(Device& dev, cudaStream_t st, const MaskPlan& mp, int layout, const double* x, const double* y,
                        const double* z, const double* q0, const double* q1, const double* q2, const double* area,
                        double wscale)
{
    int32_t pad;
    LPM_TRY(reserve_sources<KK>(dev, mp, &pad));
    pack_pse_generic<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), layout, x, y, z, q0, q1, q2,
                                                     area, wscale, dev.ws.sources.as<double>());
    count_launch();
    return LPM_OK;
}

// ---- gradient (plane): in = x y f area; sc = eps; out = gx gy
struct OpPseGradPlane {
    using K = PseGradPlane;
    static constexpr int NIN = 4, NOUT = 2, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        const double e = a.sc[0];
        return pack_generic<K>(dev, st, mp, 0, a.in[0], a.in[1], nullptr, a.in[2], nullptr, nullptr, a.in[3],
                               1.0 / (LPM_PI * e * e * e));
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.f = a.in[2];
        p.inv_eps2 = 1.0 / (a.sc[0] * a.sc[0]); p.scale = 1.0 / a.sc[0];
        return p;
    }
};
// ---- gradient (sphere): in = x y z f area; sc = eps, sphere_radius; out = gx gy gz
struct OpPseGradSphere {
    using K = PseGradSphere;
    static constexpr int NIN = 5, NOUT = 3, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        const double e = a.sc[0];
        return pack_generic<K>(dev, st, mp, 1, a.in[0], a.in[1], a.in[2], a.in[3], nullptr, nullptr, a.in[4],
                               1.0 / (LPM_PI * e * e));
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.z = a.in[2]; p.f = a.in[3];
        p.c = pse_sphere_consts(a.sc[0], a.sc[1], 1.0 / (a.sc[0] * a.sc[0]));
        return p;
    }
};
// ---- plane second partials (MODE 0, 3 outputs) / double dot (MODE 1, 1 output): in = x y gx gy area; sc = eps
template <int MODE>
struct OpPseTensorPlane {
    using K = PseTensorPlane<MODE>;
    static constexpr int NIN = 5, NOUT = MODE == 0 ? 3 : 1, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        const double e = a.sc[0];
        return pack_generic<K>(dev, st, mp, 2, a.in[0], a.in[1], nullptr, a.in[2], a.in[3], nullptr, a.in[4],
                               1.0 / (LPM_PI * e * e * e));
    }
    static typename K::Params params(const Args& a)
    {
        typename K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.gx = a.in[2]; p.gy = a.in[3];
        p.inv_eps2 = 1.0 / (a.sc[0] * a.sc[0]); p.inv_eps = 1.0 / a.sc[0];
        return p;
    }
};
// ---- sphere double dot: in = x y z u v w area; sc = eps, sphere_radius
struct OpPseDoubleDotSphere {
    using K = PseDoubleDotSphere;
    static constexpr int NIN = 7, NOUT = 1, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        const double e = a.sc[0];
        return pack_generic<K>(dev, st, mp, 3, a.in[0], a.in[1], a.in[2], a.in[3], a.in[4], a.in[5], a.in[6],
                               1.0 / (LPM_PI * e * e * e));
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.z = a.in[2]; p.u = a.in[3]; p.v = a.in[4];
        p.c = pse_sphere_consts(a.sc[0], a.sc[1], 1.0 / (a.sc[0] * a.sc[0]));
        return p;
    }
};
// ---- sphere divergence: in = x y z u v w area; sc = eps, sphere_radius
struct OpPseDivSphere {
    using K = PseDivSphere;
    static constexpr int NIN = 7, NOUT = 1, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        const double e = a.sc[0];
        return pack_generic<K>(dev, st, mp, 3, a.in[0], a.in[1], a.in[2], a.in[3], a.in[4], a.in[5], a.in[6],
                               1.0 / (LPM_PI * e * e * e));
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.z = a.in[2]; p.u = a.in[3]; p.v = a.in[4]; p.w = a.in[5];
        p.c = pse_sphere_consts(a.sc[0], a.sc[1], 1.0 / a.sc[0]);
        return p;
    }
};

// ---- planar SWE RHS integrals: in = x y vort div surf area; sc = eps; out = u v doubleDot lapSurf
struct OpSweRhsPlane {
    using K = SweRhsPlane;
    static constexpr int NIN = 6, NOUT = 4, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_swe_plane<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                       a.in[3], a.in[4], a.in[5], a.sc[0], dev.ws.sources.as<double>());
        count_launch();
        return LPM_OK;
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.surf = a.in[4];
        p.inv_eps2 = 1.0 / (a.sc[0] * a.sc[0]);
        return p;
    }
};

// ---- planar SWE velocity (PlanarSWE.f90:469-494): in = x y vort div area; out = u v
struct OpSwePlaneVel {
    using K = SwePlaneVel;
    static constexpr int NIN = 5, NOUT = 2, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_swe_plane_vel<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                           a.in[3], a.in[4], dev.ws.sources.as<double>());
        count_launch();
        return LPM_OK;
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1];
        return p;
    }
};

// ---- spherical SWE RHS integrals (SphereSWESolver.f90:296-375): in = x y z vort div surf area;
//      sc = radius, eps; out = u v w lapSurf
struct OpSweRhsSphere {
    using K = SweRhsSphere;
    static constexpr int NIN = 7, NOUT = 4, NTGT = 0;
    static int pack(Device& dev, cudaStream_t st, const MaskPlan& mp, const Args& a)
    {
        int32_t pad;
        LPM_TRY(reserve_sources<K>(dev, mp, &pad));
        pack_swe_sphere<<<pack_grid(pad), 256, 0, st>>>(mp.nsrc, pad, mp.active.as<int32_t>(), a.in[0], a.in[1], a.in[2],
                                                        a.in[3], a.in[4], a.in[5], a.in[6], a.sc[0], a.sc[1],
                                                        dev.ws.sources.as<double>());
        count_launch();
        return LPM_OK;
    }
    static K::Params params(const Args& a)
    {
        K::Params p{};
        p.x = a.in[0]; p.y = a.in[1]; p.z = a.in[2]; p.surf = a.in[5];
        p.R2 = a.sc[0] * a.sc[0];
        p.c = pse_sphere_consts(a.sc[1], a.sc[0], 1.0);      // SphereDistance uses the sphere's radius
        return p;
    }
};

}  // namespace lpm
