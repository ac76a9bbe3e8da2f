/* caller.c -- TEST INFRASTRUCTURE: a compiled C caller of the C ABI (include/lpm_gpu.h, include/lpm_mesh.h),
 * passing arrays by reference and scalars by value exactly as the bind(C) interfaces of
 * lpm_v2_b200/fortran/lpm_gpu.f90 declare them (integer(c_int64_t), value :: n; real(c_double) :: x(*);
 * real(c_double), value :: radius; type(c_ptr) handles).  It does what the patched Fortran would do for
 * src/SphereBVESolver.f90:80-90, 219-353: New -> Timestep -> read the state back -> Delete, plus one
 * BVESphereVelocity evaluation (:377-430), and checks every result against the CPU oracle
 * (oracle/liblpm_oracle.so, linked here as the checker only).
 *
 *   gcc -std=c99 -O1 -I include tests/c_caller/caller.c -L lpm_v2_b200 -llpmgpu -llpmmesh -L oracle -llpm_oracle -lm
 * Prints "C_CALLER_OK <max rel err>" and exits 0 on success. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lpm_gpu.h"
#include "lpm_mesh.h"

/* the checker (oracle/lpm_oracle.c) */
void oracle_bve_velocity(int64_t n, const double *x, const double *y, const double *z, const double *relVort,
                         const double *area, const int32_t *mask, double R, int64_t ibeg, int64_t iend,
                         double *u, double *v, double *w);
void oracle_bve_rk4_step(int64_t n, double *x, double *y, double *z, double *relVort, double *u, double *v, double *w,
                         const double *area, const int32_t *mask, double R, double Omega, double dt);

static double relerr(const double *a, const double *b, int64_t n)
{
    double num = 0.0, den = 1e-300;
    for (int64_t i = 0; i < n; ++i) {
        if (fabs(a[i] - b[i]) > num) num = fabs(a[i] - b[i]);
        if (fabs(b[i]) > den) den = fabs(b[i]);
    }
    return num / den;
}

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != LPM_OK) {                                                             \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, lpm_gpu_last_error());         \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

int main(void)
{
    /* mesh: icosTri level 3 through the host-only mesh library */
    lpm_mesh *mesh = NULL;
    if (lpm_mesh_create(LPM_ICOS_TRI_SPHERE_SEED, 3, 1.0, &mesh) != LPM_OK) return 3;
    const int64_t n = lpm_mesh_num_particles(mesh);
    double *buf = (double *)malloc(sizeof(double) * (size_t)n * 24);
    int32_t *mask = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    double *x = buf, *y = buf + n, *z = buf + 2 * n, *area = buf + 3 * n, *zeta = buf + 4 * n, *absv = buf + 5 * n;
    double *u = buf + 6 * n, *v = buf + 7 * n, *w = buf + 8 * n;
    double *ou = buf + 9 * n, *ov = buf + 10 * n, *ow = buf + 11 * n;
    double *rx = buf + 12 * n, *ry = buf + 13 * n, *rz = buf + 14 * n, *rq = buf + 15 * n;
    double *gx = buf + 16 * n, *gy = buf + 17 * n, *gz = buf + 18 * n, *gq = buf + 19 * n;
    double *gu = buf + 20 * n, *gv = buf + 21 * n, *gw = buf + 22 * n, *scratch = buf + 23 * n;
    if (lpm_mesh_get_particles(mesh, x, y, z, area, mask) != LPM_OK) return 3;
    lpm_mesh_destroy(mesh);
    const double R = 1.0, Omega = 6.283185307179586, dt = 0.01;
    for (int64_t i = 0; i < n; ++i) {          /* a smooth vorticity with zero mean by symmetry */
        zeta[i] = 3.0 * z[i] + 2.0 * x[i] * y[i];
        absv[i] = zeta[i] + 2.0 * Omega * z[i] / R;
    }

    int used = 0;
    CHECK(lpm_gpu_init(1, &used));
    if (used != 1) return 4;

    /* one BVESphereVelocity evaluation: arrays by reference, n and the radius by value */
    CHECK(lpm_bve_velocity(n, x, y, z, zeta, area, mask, R, u, v, w));
    oracle_bve_velocity(n, x, y, z, zeta, area, mask, R, 0, n, ou, ov, ow);
    double worst = relerr(u, ou, n);
    if (relerr(v, ov, n) > worst) worst = relerr(v, ov, n);
    if (relerr(w, ow, n) > worst) worst = relerr(w, ow, n);

    /* New -> Timestep -> get_state -> Delete */
    lpm_bve_solver *solver = NULL;
    CHECK(lpm_bve_solver_new(n, x, y, z, zeta, absv, u, v, w, area, mask, R, Omega, &solver));
    CHECK(lpm_bve_solver_timestep(solver, dt, 1));
    CHECK(lpm_bve_solver_get_state(solver, gx, gy, gz, gq, gu, gv, gw, scratch, NULL));
    CHECK(lpm_bve_solver_delete(solver));
    memcpy(rx, x, sizeof(double) * n); memcpy(ry, y, sizeof(double) * n); memcpy(rz, z, sizeof(double) * n);
    memcpy(rq, zeta, sizeof(double) * n);
    oracle_bve_rk4_step(n, rx, ry, rz, rq, ou, ov, ow, area, mask, R, Omega, dt);
    const double *got[7] = {gx, gy, gz, gq, gu, gv, gw};
    const double *ref[7] = {rx, ry, rz, rq, ou, ov, ow};
    for (int k = 0; k < 7; ++k) {
        const double e = relerr(got[k], ref[k], n);
        if (e > worst) worst = e;
    }
    CHECK(lpm_gpu_finalize());
    free(buf); free(mask);
    if (!(worst <= 1e-12)) {
        fprintf(stderr, "parity %.3e > 1e-12\n", worst);
        return 1;
    }
    printf("C_CALLER_OK %.3e\n", worst);
    return 0;
}
