"""Regenerates tests/golden/oracle_*.npz.

These fixtures are outputs of THIS repo's CPU oracle (oracle/lpm_oracle.c, parity
build), not of the reference: lpm-v2 is Fortran + MPI and cannot be compiled or
run in the build image (no Fortran compiler), and it ships no golden vectors for
these sums.  They pin the oracle against drift (compiler, libm, edits) and give
the GPU parity tests committed expected values.  The only reference-derived
known answers are in reference_thresholds.json.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpm_v2_b200 import mesh, problems  # noqa: E402
from oracle import binding as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # config 1 at a CPU-friendly level: icosTri L2, Gaussian vortex
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, 2)
    zeta = problems.gaussian_vortex(m)
    absv = problems.abs_vorticity(m, zeta, 2 * np.pi)
    u, v, w = O.bve_velocity(m.x, m.y, m.z, zeta, m.area, m.is_active, 1.0)
    rs, as_ = O.bve_stream(m.x, m.y, m.z, zeta, absv, m.area, m.is_active, 1.0)
    st = O.bve_rk4_step(m.x, m.y, m.z, zeta, u, v, w, m.area, m.is_active, 1.0, 2 * np.pi, 0.01)
    np.savez_compressed(os.path.join(HERE, "oracle_bve_icos2.npz"), zeta=zeta, absvort=absv, u=u, v=v, w=w,
                        relstream=rs, absstream=as_, **{"rk4_" + k: a for k, a in zip("x y z zeta u v w".split(), st)})
    # config 3: icosTri L3, PSE Laplacian of Y_5^4, eps = h^0.6
    m = mesh.PolyMesh2d(mesh.ICOS_TRI_SPHERE_SEED, 3)
    f = problems.spherical_harmonic54(m)
    eps = m.max_edge_length ** 0.6
    lap = O.pse_laplacian_sphere(m.x, m.y, m.z, f, m.area, m.is_active, eps, 1.0)
    np.savez_compressed(os.path.join(HERE, "oracle_pse_icos3.npz"), f=f, eps=eps, lap=lap)
    # config 2 at quadRect L3: colliding Lamb dipoles
    m = mesh.PolyMesh2d(mesh.QUAD_RECT_SEED, 3, 7.0)
    vort = problems.colliding_dipoles(m)
    u, v = O.plane_velocity(m.x, m.y, vort, m.area, m.is_active)
    psi = O.plane_stream(m.x, m.y, vort, m.area, m.is_active)
    st = O.plane_rk4_step(m.x, m.y, vort, u, v, m.area, m.is_active, 0.01)
    np.savez_compressed(os.path.join(HERE, "oracle_plane_quad3.npz"), vort=vort, u=u, v=v, psi=psi,
                        **{"rk4_" + k: a for k, a in zip("x y u v".split(), st)})
    # beta plane L3 Gaussian vortex
    m = mesh.PolyMesh2d(mesh.BETA_PLANE_SEED, 3)
    zeta = problems.betaplane_gaussian(m)
    u, v = O.betaplane_velocity(m.x, m.y, zeta, m.area, m.is_active)
    lu, lv = O.betaplane_velocity(m.x, m.y, zeta, m.area, m.is_active, variant="_ld")
    np.savez_compressed(os.path.join(HERE, "oracle_beta3.npz"), zeta=zeta, u=u, v=v, u_ld=lu, v_ld=lv)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
