// pairs_swe.cuh -- the fused right-hand-side integrals of the planar shallow-water
// solver, src/SWEPlaneSolver.f90:457-560 (SWEPlaneRHSIntegrals): in one pair loop
//   u, v         Biot-Savart + potential-flow velocity from vorticity and divergence (:498-499)
//   ux,uy,vx,vy  the velocity-gradient sums whose double dot product feeds the divergence equation (:501-508)
//   lapSurf      PSE Laplacian of the fluid surface h + topography (:492-494)
// SURVEY.md 8(f) rank 3.  The self pair contributes nothing (the reference cycles at
// i == j after a Laplacian term that is exactly zero there).
#pragma once
#include "pairs.cuh"
#include "pairs_pse.cuh"

namespace lpm {

// Source record: x, y, zeta A/(2 pi), delta A/(2 pi), surface height, A/(pi eps^2).
struct SweRhsPlane : NoSharedTable {
    static constexpr int NS = 6, NA = 7;
    static constexpr bool SKIP_SELF = true;
    static constexpr bool BATCHED_RCP = true;
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
        rcp_group<T, CHECK>(r2, r);
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

// =============================================================================
// Planar SWE velocity from vorticity and divergence, src/PlanarSWE.f90:469-494
// (SetVelocityFromFieldData; SWEComputeVelocity :261-290 is the same sum):
//   rot = zeta_j A_j / (2 pi r^2), pot = delta_j A_j / (2 pi r^2)
//   u_i += -(y_i - y_j) rot + (x_i - x_j) pot;  v_i += (x_i - x_j) rot + (y_i - y_j) pot
// Source record: x, y, zeta A/(2 pi), delta A/(2 pi).
struct SwePlaneVel : NoSharedTable {
    static constexpr int NS = 4, NA = 2;
    static constexpr bool SKIP_SELF = true;
    static constexpr bool BATCHED_RCP = true;
    struct Params {
        const double *x, *y;
        Outs<2> out;
    };
    struct Tgt { double x, y; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i]}; }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params&, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], const SharedCtx&)
    {
        double dx[T], dy[T], r2[T], r[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            dx[k] = t[k].x - s[0]; dy[k] = t[k].y - s[1];
            r2[k] = fma(dx[k], dx[k], dy[k] * dy[k]);
            if (CHECK) r2[k] = (j == self[k]) ? 1.0 : r2[k];
        }
        rcp_group<T, CHECK>(r2, r);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double w = r[k];
            if (CHECK) w = (j == self[k]) ? 0.0 : w;
            const double rot = s[2] * w, pot = s[3] * w;
            acc[k][0] = fma(dx[k], pot, fma(-dy[k], rot, acc[k][0]));
            acc[k][1] = fma(dy[k], pot, fma(dx[k], rot, acc[k][1]));
        }
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0]);
        p.out.store(1, i, a[1]);
    }
};

__global__ void pack_swe_plane_vel(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                                   const double* __restrict__ x, const double* __restrict__ y,
                                   const double* __restrict__ vort, const double* __restrict__ div,
                                   const double* __restrict__ area, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[4] = {LPM_PLANE_FAR, LPM_PLANE_FAR, 0.0, 0.0};
    if (c < nsrc) {
        int32_t j = active[c];
        r[0] = x[j]; r[1] = y[j];
        r[2] = vort[j] * area[j] / (2.0 * LPM_PI);
        r[3] = div[j] * area[j] / (2.0 * LPM_PI);
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 4);
    o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]);
}

// =============================================================================
// Spherical SWE right-hand-side integrals, src/SphereSWESolver.f90:296-375 (SWESphereRHSIntegrals),
// as far as the reference computes them: with d = R^2 - x_i . x_j,
//   rot = zeta_j A_j / (4 pi R^2 d), pot = R delta_j A_j / (4 pi R^2 d)
//   (u, v, w)_i -= (x_i cross x_j) rot + x_j pot                         (j /= i)
//   lapSurf_i  += eta(d_ij / eps) / eps^2 (s_j - s_i) A_j                  (j == i included: exactly 0)
// i.e. (u, v, w)_i = x_i cross A_i + B_i with A_i = sum_j cr_j x_j / d, B_i = sum_j cp_j x_j / d,
// cr = -zeta A/(4 pi R^2), cp = -delta A/(4 pi R).  doubleDot is returned as zero: the reference zeroes it and
// never accumulates it (:328), and applies no trailing 1/eps^2 to lapSurf (the plane routine does).
// Source record: x, y, z, cr, cp, surface height, A/(pi eps^2), |x|.
struct SweRhsSphere : NoSharedTable {
    static constexpr int NS = 8, NA = 7;
    static constexpr bool SKIP_SELF = true;
    static constexpr bool BATCHED_RCP = true;
    struct Params {
        const double *x, *y, *z, *surf;
        double R2;
        PseSphereConsts c;
        Outs<4> out;        // u, v, w, lapSurf (doubleDot is identically zero: filled by the entry point)
    };
    struct Tgt { double x, y, z, s, nrm, thr; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        Tgt t{p.x[i], p.y[i], p.z[i], p.surf[i], 0.0, 0.0};
        t.nrm = sqrt(t.x * t.x + t.y * t.y + t.z * t.z);
        t.thr = p.c.cos_cut * t.nrm;
        return t;
    }
    template <int T, bool CHECK>
    __device__ static __forceinline__ void group(const Params& p, const Tgt (&t)[T], const double (&s)[NS],
                                                 double (&acc)[T][NA], int32_t j, const int32_t (&self)[T], const SharedCtx&)
    {
        double d[T], r[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            const double dot = fma(t[k].x, s[0], fma(t[k].y, s[1], t[k].z * s[2]));
            if (dot >= t[k].thr * s[7]) {      // inside the PSE cut-off; at the self pair s[5] - t.s == 0 exactly
                const double k2 = sphere_k2(t[k].x, t[k].y, t[k].z, s[0], s[1], s[2], dot, t[k].nrm * s[7], p.c);
                acc[k][6] = fma((s[5] - t[k].s) * pse_eta_pi(k2), s[6], acc[k][6]);
            }
            d[k] = fma(-t[k].x, s[0], p.R2);
            d[k] = fma(-t[k].y, s[1], d[k]);
            d[k] = fma(-t[k].z, s[2], d[k]);
            if (CHECK) d[k] = (j == self[k]) ? 1.0 : d[k];
        }
        rcp_group<T, CHECK>(d, r);
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double w = r[k];
            if (CHECK) w = (j == self[k]) ? 0.0 : w;
            const double rot = s[3] * w, pot = s[4] * w;
            acc[k][0] = fma(rot, s[0], acc[k][0]);
            acc[k][1] = fma(rot, s[1], acc[k][1]);
            acc[k][2] = fma(rot, s[2], acc[k][2]);
            acc[k][3] = fma(pot, s[0], acc[k][3]);
            acc[k][4] = fma(pot, s[1], acc[k][4]);
            acc[k][5] = fma(pot, s[2], acc[k][5]);
        }
    }
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt& t, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, fma(t.y, a[2], -(t.z * a[1])) + a[3]);
        p.out.store(1, i, fma(t.z, a[0], -(t.x * a[2])) + a[4]);
        p.out.store(2, i, fma(t.x, a[1], -(t.y * a[0])) + a[5]);
        p.out.store(3, i, a[6]);
    }
};

__global__ void pack_swe_sphere(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                                const double* __restrict__ x, const double* __restrict__ y,
                                const double* __restrict__ z, const double* __restrict__ vort,
                                const double* __restrict__ div, const double* __restrict__ surf,
                                const double* __restrict__ area, double R, double eps, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, kNullNorm};     // null source: d = R^2, zero strengths, zero area, outside the PSE cut-off
    if (c < nsrc) {
        int32_t j = active[c];
        r[0] = x[j]; r[1] = y[j]; r[2] = z[j];
        r[3] = -vort[j] * area[j] / (4.0 * LPM_PI * R * R);
        r[4] = -div[j] * area[j] / (4.0 * LPM_PI * R);
        r[5] = surf[j];
        r[6] = area[j] / (LPM_PI * eps * eps);
        r[7] = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * 8);
    for (int q = 0; q < 4; ++q) o[q] = make_double2(r[2 * q], r[2 * q + 1]);
}

}  // namespace lpm
