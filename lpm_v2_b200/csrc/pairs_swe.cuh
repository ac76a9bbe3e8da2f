// pairs_swe.cuh -- the fused right-hand-side integrals of the planar shallow-water
// solver, src/SWEPlaneSolver.f90:457-560 (SWEPlaneRHSIntegrals): in one pair loop
//   u, v         Biot-Savart + potential-flow velocity from vorticity and divergence (:498-499)
//   ux,uy,vx,vy  the velocity-gradient sums whose double dot product feeds the divergence equation (:501-508)
//   lapSurf      PSE Laplacian of the fluid surface h + topography (:492-494)
// SURVEY.md 8(f) rank 3.  The self pair contributes nothing (the reference cycles at
// i == j after a Laplacian term that is exactly zero there).
#pragma once
#include "pairs.cuh"

namespace lpm {

// Source record: x, y, zeta A/(2 pi), delta A/(2 pi), surface height, A/(pi eps^2).
struct SweRhsPlane : NoSharedTable {
    static constexpr int NS = 6, NA = 7;
    static constexpr bool SKIP_SELF = true;
    struct Params {
        const double *x, *y, *surf;
        double inv_eps2;
        Outs<4> out;        // u, v, doubleDot, lapSurf
    };
    struct Tgt { double x, y, s; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i], p.surf[i]}; }

    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params& p, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], const SharedCtx&)
    {
        double dx[T], dy[T], r2[T], r[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            dx[k] = t[k].x - s[0]; dy[k] = t[k].y - s[1];
            r2[k] = fma(dx[k], dx[k], dy[k] * dy[k]);
            const double k2 = r2[k] * p.inv_eps2;
            if (k2 <= kPseCut * kPseCut)       // at the self pair s[4] - t.s == 0 exactly
                acc[k][6] = fma(pse_eta_pi(k2) * (s[4] - t[k].s), s[5], acc[k][6]);
            if (CHECK) r2[k] = (j == self[k]) ? 1.0 : r2[k];
        }
        rcp_batch<T>(r2, r);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double w = r[k];
            if (CHECK) w = (j == self[k]) ? 0.0 : w;
            const double rot = s[2] * w, pot = s[3] * w;
            acc[k][0] = fma(dx[k], pot, fma(-dy[k], rot, acc[k][0]));
            acc[k][1] = fma(dy[k], pot, fma(dx[k], rot, acc[k][1]));
            const double w2 = 2.0 * w * w;
            const double a = fma(dx[k], s[3], -(dy[k] * s[2])) * w2;
            const double b = fma(dy[k], s[3], dx[k] * s[2]) * w2;
            acc[k][2] = fma(-dx[k], a, acc[k][2] + pot);     // ux
            acc[k][3] = fma(-dy[k], a, acc[k][3] - rot);     // uy
            acc[k][4] = fma(-dx[k], b, acc[k][4] + rot);     // vx
            acc[k][5] = fma(-dy[k], b, acc[k][5] + pot);     // vy
        }
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0]);
        p.out.store(1, i, a[1]);
        p.out.store(2, i, a[2] * a[2] + 2.0 * a[3] * a[4] + a[5] * a[5]);
        p.out.store(3, i, a[6] * p.inv_eps2);
    }
};

__global__ void pack_swe_plane(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                               const double* __restrict__ x, const double* __restrict__ y,
                               const double* __restrict__ vort, const double* __restrict__ div,
                               const double* __restrict__ surf, const double* __restrict__ area, double eps,
                               double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[6] = {LPM_PLANE_FAR, LPM_PLANE_FAR, 0.0, 0.0, 0.0, 0.0};      // null source: far away, zero strength / area
    if (c < nsrc) {
        int32_t j = active[c];
        r[0] = x[j]; r[1] = y[j];
        r[2] = vort[j] * area[j] / (2.0 * LPM_PI);
        r[3] = div[j] * area[j] / (2.0 * LPM_PI);
        r[4] = surf[j];
        r[5] = area[j] / (LPM_PI * eps * eps);
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 6);
    o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]); o[2] = make_double2(r[4], r[5]);
}

}  // namespace lpm
