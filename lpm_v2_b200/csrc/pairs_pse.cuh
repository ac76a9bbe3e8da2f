// pairs_pse.cuh -- the remaining particle-strength-exchange operators of
// src/PSEDirectSum.f90, as functors for ds_kernel:
//   interpolation        :128-168   (delta kernel, arbitrary target locations)
//   gradient             :180-267   (first-derivative kernel; sphere: tangent projection)
//   second partials      :269-320   (plane)
//   double dot product   :322-420   (plane and sphere)
//   divergence           :537-579   (sphere)
// All share the e^{-k^2} decay of the Laplacian kernel, so the same k > kPseCut
// cut-off applies.  j == i is included everywhere, as in the reference.
#pragma once
#include "pairs.cuh"

namespace lpm {

// pi * bivariateDeltaKernel8 (PSEDirectSum.f90:611-616) and pi * bivariateFirstDerivativeKernel8 (:629-634)
__device__ __forceinline__ double pse_delta_pi(double k2)
{
    return fma(fma(fma(-1.0 / 6.0, k2, 2.0), k2, -6.0), k2, 4.0) * pse_exp_neg(k2);
}
__device__ __forceinline__ double pse_dphi_pi(double k2)
{
    return fma(fma(fma(1.0 / 3.0, k2, -5.0), k2, 20.0), k2, -20.0) * pse_exp_neg(k2);
}

// ---------------------------------------------------------------- interpolation
// Source record (sphere): x, y, z, f A/(pi eps^2), |x|, 0;  (plane): x, y, f A/(pi eps^2), 0.
struct PseInterpSphere : CullSphere {
    static constexpr int NS = 6, NA = 1;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *z;          // TARGET locations (length = number of targets)
        PseSphereConsts c;
        Outs<1> out;
    };
    __device__ static __forceinline__ double cull_dist(const Params& p) { return p.c.chord_cut; }
    struct Tgt { double x, y, z, nrm, thr; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        Tgt t{p.x[i], p.y[i], p.z[i], 0.0, 0.0};
        t.nrm = sqrt(t.x * t.x + t.y * t.y + t.z * t.z);
        t.thr = p.c.cos_cut * t.nrm;
        return t;
    }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const SharedCtx&)
    {
        double dot = fma(t.x, s[0], fma(t.y, s[1], t.z * s[2]));
        if (dot < t.thr * s[4]) return;
        double k2 = sphere_k2(t.x, t.y, t.z, s[0], s[1], s[2], dot, t.nrm * s[4], p.c);
        acc[0] = fma(pse_delta_pi(k2), s[3], acc[0]);
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0]);
    }
};

struct PseInterpPlane : CullPlane {
    static constexpr int NS = 4, NA = 1;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y;
        double inv_eps2;
        Outs<1> out;
    };
    __device__ static __forceinline__ double cull_dist(const Params& p) { return kPseCut * rsqrt(p.inv_eps2); }
    struct Tgt { double x, y; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i]}; }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const SharedCtx&)
    {
        double dx = s[0] - t.x, dy = s[1] - t.y;
        double k2 = fma(dx, dx, dy * dy) * p.inv_eps2;
        if (k2 > kPseCut * kPseCut) return;
        acc[0] = fma(pse_delta_pi(k2), s[2], acc[0]);
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0]);
    }
};

// mode 0: interpolation record (f A/(pi eps^2)); geometry picks the layout
__global__ void pack_pse_interp(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active,
                                const double* __restrict__ x, const double* __restrict__ y,
                                const double* __restrict__ z, const double* __restrict__ f,
                                const double* __restrict__ area, double eps, int sphere, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[6] = {0.0, 0.0, 0.0, 0.0, sphere ? kNullNorm : 0.0, 0.0};      // null source (see pack_pse_sphere)
    if (c < nsrc) {
        int32_t j = active[c];
        double w = f[j] * area[j] / (LPM_PI * eps * eps);
        if (sphere) {
            r[0] = x[j]; r[1] = y[j]; r[2] = z[j]; r[3] = w;
            r[4] = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        } else {
            r[0] = x[j]; r[1] = y[j]; r[2] = w;
        }
    }
    if (sphere) {
        double2* o = reinterpret_cast<double2*>(src + (size_t)c * 6);
        o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]); o[2] = make_double2(r[4], r[5]);
    } else {
        double2* o = reinterpret_cast<double2*>(src + (size_t)c * 4);
        o[0] = make_double2(r[0], r[1]); o[1] = make_double2(r[2], r[3]);
    }
}

// ---------------------------------------------------------------- gradient
// plane (:180-218): grad_i = eps^-1 sum (f_j + f_i)(x_i - x_j) phi'(k)/eps^3 A_j.  Record: x, y, f, A/(pi eps^3).
struct PseGradPlane : CullPlane {
    static constexpr int NS = 4, NA = 2;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *f;
        double inv_eps2, scale;
        Outs<2> out;
    };
    __device__ static __forceinline__ double cull_dist(const Params& p) { return kPseCut * rsqrt(p.inv_eps2); }
    struct Tgt { double x, y, f; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i) { return Tgt{p.x[i], p.y[i], p.f[i]}; }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const SharedCtx&)
    {
        double dx = t.x - s[0], dy = t.y - s[1];
        double k2 = fma(dx, dx, dy * dy) * p.inv_eps2;
        if (k2 > kPseCut * kPseCut) return;
        double c = (s[2] + t.f) * pse_dphi_pi(k2) * s[3];
        acc[0] = fma(dx, c, acc[0]);
        acc[1] = fma(dy, c, acc[1]);
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0] * p.scale);
        p.out.store(1, i, a[1] * p.scale);
    }
};

// sphere (:221-267): grad_i = eps^-2 P_i sum (f_j + f_i)(x_i - x_j) phi'(k)/eps^2 A_j, P = I - x x^T.
// Record: x, y, z, f, A/(pi eps^2), |x|.
struct PseGradSphere : CullSphere {
    static constexpr int NS = 6, NA = 3;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *z, *f;
        PseSphereConsts c;
        Outs<3> out;
    };
    __device__ static __forceinline__ double cull_dist(const Params& p) { return p.c.chord_cut; }
    struct Tgt { double x, y, z, f, nrm, thr; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        Tgt t{p.x[i], p.y[i], p.z[i], p.f[i], 0.0, 0.0};
        t.nrm = sqrt(t.x * t.x + t.y * t.y + t.z * t.z);
        t.thr = p.c.cos_cut * t.nrm;
        return t;
    }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const SharedCtx&)
    {
        double dot = fma(t.x, s[0], fma(t.y, s[1], t.z * s[2]));
        if (dot < t.thr * s[5]) return;
        double k2 = sphere_k2(t.x, t.y, t.z, s[0], s[1], s[2], dot, t.nrm * s[5], p.c);
        double c = (s[3] + t.f) * pse_dphi_pi(k2) * s[4];
        acc[0] = fma(t.x - s[0], c, acc[0]);
        acc[1] = fma(t.y - s[1], c, acc[1]);
        acc[2] = fma(t.z - s[2], c, acc[2]);
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt& t, const double (&a)[NA], int64_t i)
    {
        double xg = fma(t.x, a[0], fma(t.y, a[1], t.z * a[2]));       // P g = g - x (x . g)
        p.out.store(0, i, fma(-t.x, xg, a[0]) * p.c.scale);
        p.out.store(1, i, fma(-t.y, xg, a[1]) * p.c.scale);
        p.out.store(2, i, fma(-t.z, xg, a[2]) * p.c.scale);
    }
};

// ---------------------------------------------------------------- plane: second partials / double dot
// Both accumulate the four sums  q_ab = sum (g_a,j + g_a,i)(x_i - x_j)_b phi'(k)/eps^3 A_j  (:287-303, :343-357).
// MODE 0: second partials -> (q_xx, (q_xy + q_yx)/2, q_yy)/eps;  MODE 1: double dot -> (q_xx^2 + 2 q_xy q_yx + q_yy^2)/eps^2.
// Record: x, y, g_x, g_y, A/(pi eps^3), 0.
template <int MODE>
struct PseTensorPlane : CullPlane {
    static constexpr int NS = 6, NA = 4;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *gx, *gy;
        double inv_eps2, inv_eps;
        Outs<3> out;
    };
    __device__ static __forceinline__ double cull_dist(const Params& p) { return kPseCut * rsqrt(p.inv_eps2); }
    struct Tgt { double x, y, gx, gy; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        return Tgt{p.x[i], p.y[i], p.gx[i], p.gy[i]};
    }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const SharedCtx&)
    {
        double dx = t.x - s[0], dy = t.y - s[1];
        double k2 = fma(dx, dx, dy * dy) * p.inv_eps2;
        if (k2 > kPseCut * kPseCut) return;
        double c = pse_dphi_pi(k2) * s[4];
        double a = (s[2] + t.gx) * c, b = (s[3] + t.gy) * c;
        acc[0] = fma(a, dx, acc[0]);
        acc[1] = fma(a, dy, acc[1]);
        acc[2] = fma(b, dx, acc[2]);
        acc[3] = fma(b, dy, acc[3]);
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        if (MODE == 0) {
            p.out.store(0, i, a[0] * p.inv_eps);
            p.out.store(1, i, 0.5 * (a[1] + a[2]) * p.inv_eps);
            p.out.store(2, i, a[3] * p.inv_eps);
        } else {
            p.out.store(0, i, (a[0] * a[0] + 2.0 * a[1] * a[2] + a[3] * a[3]) * p.inv_eps2);
        }
    }
};

// ---------------------------------------------------------------- sphere: double dot / divergence
// Record: x, y, z, u, v, w, A/(pi eps^3), |x|.
// Double dot (:367-420): nine sums; the w rows add yComp(i) (the reference's quirk at :408-413, kept).
struct PseDoubleDotSphere : CullSphere {
    static constexpr int NS = 8, NA = 9;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *z, *u, *v;
        PseSphereConsts c;
        Outs<1> out;
    };
    __device__ static __forceinline__ double cull_dist(const Params& p) { return p.c.chord_cut; }
    struct Tgt { double x, y, z, u, v, nrm, thr; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        Tgt t{p.x[i], p.y[i], p.z[i], p.u[i], p.v[i], 0.0, 0.0};
        t.nrm = sqrt(t.x * t.x + t.y * t.y + t.z * t.z);
        t.thr = p.c.cos_cut * t.nrm;
        return t;
    }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const SharedCtx&)
    {
        double dot = fma(t.x, s[0], fma(t.y, s[1], t.z * s[2]));
        if (dot < t.thr * s[7]) return;
        double k2 = sphere_k2(t.x, t.y, t.z, s[0], s[1], s[2], dot, t.nrm * s[7], p.c);
        double c = pse_dphi_pi(k2) * s[6];
        double dx = t.x - s[0], dy = t.y - s[1], dz = t.z - s[2];
        double a = (s[3] + t.u) * c, b = (s[4] + t.v) * c, e = (s[5] + t.v) * c;
        acc[0] = fma(a, dx, acc[0]); acc[1] = fma(a, dy, acc[1]); acc[2] = fma(a, dz, acc[2]);
        acc[3] = fma(b, dx, acc[3]); acc[4] = fma(b, dy, acc[4]); acc[5] = fma(b, dz, acc[5]);
        acc[6] = fma(e, dx, acc[6]); acc[7] = fma(e, dy, acc[7]); acc[8] = fma(e, dz, acc[8]);
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        // (ux ux + vy vy + wz wz + 2 (uy vx + uz wx + vz wy)) / eps^2
        double v = a[0] * a[0] + a[4] * a[4] + a[8] * a[8] + 2.0 * (a[1] * a[3] + a[2] * a[6] + a[5] * a[7]);
        p.out.store(0, i, v * p.c.scale);
    }
};

// Divergence (:537-579): div_i = eps^-1 sum [P_i (x_i - x_j)] . (v_j + v_i) phi'(k)/eps^3 A_j.
struct PseDivSphere : CullSphere {
    static constexpr int NS = 8, NA = 1;
    static constexpr bool SKIP_SELF = false;
    struct Params {
        const double *x, *y, *z, *u, *v, *w;
        PseSphereConsts c;
        Outs<1> out;
    };
    __device__ static __forceinline__ double cull_dist(const Params& p) { return p.c.chord_cut; }
    struct Tgt { double x, y, z, u, v, w, nrm, thr; };
    __device__ static __forceinline__ Tgt load_target(const Params& p, int64_t i)
    {
        Tgt t{p.x[i], p.y[i], p.z[i], p.u[i], p.v[i], p.w[i], 0.0, 0.0};
        t.nrm = sqrt(t.x * t.x + t.y * t.y + t.z * t.z);
        t.thr = p.c.cos_cut * t.nrm;
        return t;
    }
    template <bool CHECK>
    __device__ static __forceinline__ void pair(const Params& p, const Tgt& t, const double (&s)[NS],
                                                double (&acc)[NA], bool, const SharedCtx&)
    {
        double dot = fma(t.x, s[0], fma(t.y, s[1], t.z * s[2]));
        if (dot < t.thr * s[7]) return;
        double k2 = sphere_k2(t.x, t.y, t.z, s[0], s[1], s[2], dot, t.nrm * s[7], p.c);
        double c = pse_dphi_pi(k2) * s[6];
        double dx = t.x - s[0], dy = t.y - s[1], dz = t.z - s[2];
        double wu = s[3] + t.u, wv = s[4] + t.v, ww = s[5] + t.w;
        double dw = fma(dx, wu, fma(dy, wv, dz * ww));               // d . w
        double dxi = fma(dx, t.x, fma(dy, t.y, dz * t.z));           // d . x_i
        double xw = fma(t.x, wu, fma(t.y, wv, t.z * ww));            // x_i . w
        acc[0] = fma(c, fma(-dxi, xw, dw), acc[0]);                  // (P d) . w = d.w - (d.x)(x.w)
    }
    LPM_DEFAULT_GROUP()
    __device__ static __forceinline__ void finalize(const Params& p, const Tgt&, const double (&a)[NA], int64_t i)
    {
        p.out.store(0, i, a[0] * p.c.scale);
    }
};

// generic packers for the records above
// layout 0: x, y, q0, w                      (plane gradient;        w = A/(pi eps^p))
// layout 1: x, y, z, q0, w, |x|              (sphere gradient)
// layout 2: x, y, q0, q1, w, 0               (plane tensor)
// layout 3: x, y, z, q0, q1, q2, w, |x|      (sphere double dot / divergence)
__global__ void pack_pse_generic(int32_t nsrc, int32_t nsrc_pad, const int32_t* __restrict__ active, int layout,
                                 const double* __restrict__ x, const double* __restrict__ y,
                                 const double* __restrict__ z, const double* __restrict__ q0,
                                 const double* __restrict__ q1, const double* __restrict__ q2,
                                 const double* __restrict__ area, double wscale, double* __restrict__ src)
{
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nsrc_pad) return;
    double r[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int ns = (layout == 0) ? 4 : (layout == 3) ? 8 : 6;
    if (layout == 1) r[5] = kNullNorm;           // null source of the sphere layouts (see pack_pse_sphere)
    if (layout == 3) r[7] = kNullNorm;
    if (c < nsrc) {
        int32_t j = active[c];
        double w = area[j] * wscale;
        if (layout == 0) {
            r[0] = x[j]; r[1] = y[j]; r[2] = q0[j]; r[3] = w;
        } else if (layout == 1) {
            r[0] = x[j]; r[1] = y[j]; r[2] = z[j]; r[3] = q0[j]; r[4] = w;
            r[5] = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        } else if (layout == 2) {
            r[0] = x[j]; r[1] = y[j]; r[2] = q0[j]; r[3] = q1[j]; r[4] = w;
        } else {
            r[0] = x[j]; r[1] = y[j]; r[2] = z[j]; r[3] = q0[j]; r[4] = q1[j]; r[5] = q2[j]; r[6] = w;
            r[7] = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        }
    }
    double2* o = reinterpret_cast<double2*>(src + (size_t)c * ns);
    for (int q = 0; q < ns / 2; ++q) o[q] = make_double2(r[2 * q], r[2 * q + 1]);
}

}  // namespace lpm
