"""CPU emulations (numpy, np.longdouble as the judge) of the scalar algorithms the CUDA pair
kernels use, so that their accuracy claims do not rest on GPU runs alone:

  * table logarithm of the stream-function kernels (csrc/pairs.cuh, log_tab / log_group_fast)
  * cubic-Newton reciprocal and the 4-way product tree (rcp_fast / rcp_batch)
  * theta^2 from tan^2(theta/2) and the atan^2 series of the sphere PSE kernels (sphere_k2)

MUFU.RCP64H is emulated as the high word of 1/x with a relative perturbation of up to 2^-19 --
worse than the hardware's ~2^-20."""
from fractions import Fraction

import numpy as np
import pytest


def _hi(d):
    return (d.view(np.uint64) >> np.uint64(32)).astype(np.int64)


def _from_hi(h):
    return (h.astype(np.uint64) << np.uint64(32)).view(np.float64)


def _rcp64h(x, rel_err=0.0):
    return _from_hi(_hi((1.0 / x) * (1.0 + rel_err)))


def _fma(a, b, c):
    """One rounding of a*b + c (long double carries 64 bits: enough for these magnitudes)."""
    return (a.astype(np.longdouble) * np.asarray(b, np.longdouble) + np.asarray(c, np.longdouble)).astype(np.float64)


@pytest.mark.parametrize("seed_err", [0.0, 2.0 ** -20, -2.0 ** -20, 2.0 ** -19])
def test_table_log_accuracy(seed_err):
    """Absolute error <= 1.5 ulp of max(|ln d|, 1/2) -- what matters for a sum of weighted logarithms."""
    rng = np.random.default_rng(3)
    for lo, hi in ((1e-12, 1e-9), (1e-6, 1e-3), (0.3, 0.7), (0.999, 1.001), (1.0, 4.0), (1e3, 1e7)):
        d = rng.uniform(lo, hi, 100000)
        c = _from_hi((_hi(d) & 0xfffff000) | 0x800)            # bin centre: top 8 mantissa bits + half a bin
        q = _rcp64h(c, seed_err)
        table = (-np.log(q.astype(np.longdouble))).astype(np.float64)
        r = _fma(d, q, -1.0)
        assert np.abs(r).max() <= 2.0 ** -9 + 2.0 ** -18
        p = _fma(r, 0.2, -0.25)
        p = _fma(r, p, 1.0 / 3.0)
        p = _fma(r, p, -0.5)
        p = _fma(r, p, 1.0)
        got = _fma(r, p, table)
        ref = np.log(d.astype(np.longdouble))
        err = np.abs(got.astype(np.longdouble) - ref)
        ulp = np.spacing(np.maximum(np.abs(ref.astype(np.float64)), 0.5))
        assert float((err / ulp).max()) <= 1.5


def test_reciprocal_newton_and_product_tree():
    rng = np.random.default_rng(4)
    d = np.exp(rng.uniform(np.log(1e-8), np.log(4.0), (4, 200000)))

    def rcp_fast(x, seed_err):
        r0 = _rcp64h(x, seed_err)
        e = _fma(-x, r0, 1.0)
        t = _fma(e, e, e)
        return _fma(r0, t, r0)

    for seed_err in (0.0, 2.0 ** -20, -2.0 ** -19):
        one = rcp_fast(d[0], seed_err)
        ref = 1.0 / d[0].astype(np.longdouble)
        assert float(np.abs((one - ref) / ref).max()) <= 2.3e-16            # <= 1 ulp
        p01, p23 = d[0] * d[1], d[2] * d[3]
        q = rcp_fast(p01 * p23, seed_err)
        q01, q23 = q * p23, q * p01
        got = [q01 * d[1], q01 * d[0], q23 * d[3], q23 * d[2]]
        for g, x in zip(got, d):
            ref = 1.0 / x.astype(np.longdouble)
            assert float(np.abs((g - ref) / ref).max()) <= 1.0e-15          # a few ulp, far inside 1e-12


_ATAN_SQ = [float(Fraction((-1) ** (k - 1), k) * sum(Fraction(1, 2 * j - 1) for j in range(1, k + 1))) for k in range(1, 29)]


def _atan_sq_terms(theta_cut):
    if not theta_cut < 0.9:
        return 0
    t = np.tan(0.5 * theta_cut) ** 2 * 1.02
    n, tn = 2, t
    while n < 28 and tn * t > 1e-18:
        tn *= t
        n += 1
    return 0 if tn * t > 1e-18 else n


def test_atan_sq_coefficients_match_the_kernel_source():
    import os
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "lpm_v2_b200", "csrc", "pairs.cuh")).read()
    m = re.search(r"kAtanSq\[kAtanSqMaxTerms\] = \{([^}]*)\}", src)
    coef = [float(v) for v in m.group(1).split(",")]
    assert coef == _ATAN_SQ


@pytest.mark.parametrize("theta_cut", [0.01, 0.05, 0.26, 0.5, 0.8])
def test_sphere_distance_series(theta_cut):
    nt = _atan_sq_terms(theta_cut)
    assert 2 <= nt <= 28
    rng = np.random.default_rng(5)
    n = 200000
    ang = rng.uniform(0, theta_cut, n) * rng.choice([1e-3, 1.0], n)
    ra, rb = rng.uniform(0.99, 1.01, n), rng.uniform(0.99, 1.01, n)      # particles slightly off the sphere
    a = np.stack([ra, 0 * ra, 0 * ra], 1)
    b = np.stack([rb * np.cos(ang), rb * np.sin(ang), 0 * rb], 1)
    al, bl = a.astype(np.longdouble), b.astype(np.longdouble)
    ref = np.arctan2(al[:, 0] * bl[:, 1] - al[:, 1] * bl[:, 0], (al * bl).sum(1)) ** 2     # true angle of the rounded vectors
    cr = np.cross(a, b)
    s2 = (cr * cr).sum(1)
    dot = (a * b).sum(1)
    nn = np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1)
    t = s2 / (nn + dot) ** 2
    acc = np.full(n, _ATAN_SQ[nt - 1])
    for k in range(nt - 2, -1, -1):
        acc = acc * t + _ATAN_SQ[k]
    th2 = 4.0 * t * acc
    ok = ref > 0
    series = float(np.max(np.abs(th2[ok] - ref[ok]) / ref[ok]))
    atan2_form = float(np.max(np.abs(np.arctan2(np.sqrt(s2), dot)[ok] ** 2 - ref[ok]) / ref[ok]))
    assert series <= 2e-15 and series <= 3.0 * atan2_form + 1e-16


def test_pse_exp_neg_accuracy():
    """csrc/pairs.cuh pse_exp_neg: exp(-t) from m = rint(64 t / ln 2), a two-step reduction, a degree-5 polynomial
    and a 64-entry table of 2^(-j/64): <= 2.5 ulp on [0, 700] (the PSE kernels use k^2 <= 64)."""
    rng = np.random.default_rng(11)
    tab = np.exp2(-(np.arange(64, dtype=np.longdouble)) / 64).astype(np.float64)
    magic = 6755399441055744.0
    hi, lo = float.fromhex("0x1.62e42fefa0000p-7"), float.fromhex("0x1.cf79abc9e3b3ap-46")
    for a, b in ((0.0, 1e-3), (0.0, 1.0), (1.0, 8.0), (8.0, 64.0), (64.0, 700.0)):
        t = np.concatenate([rng.uniform(a, b, 200000), [a, b]])
        u = _fma(t, 92.33248261689366, magic)
        m = (u.view(np.uint64) & np.uint64(0xffffffff)).astype(np.int64)
        mf = u - magic
        assert np.array_equal(mf, np.rint(mf)) and np.array_equal(m, mf.astype(np.int64))
        r = _fma(mf, -hi, t)
        r = _fma(mf, -lo, r)
        assert np.abs(r).max() <= np.log(2.0) / 128 * (1 + 1e-3)      # (the long-double emulation of the fma double-rounds ties)
        p = _fma(r, -1.0 / 120.0, 1.0 / 24.0)
        p = _fma(r, p, -1.0 / 6.0)
        p = _fma(r, p, 0.5)
        p = _fma(r, p, -1.0)
        p = _fma(r, p, 1.0)
        v = tab[m & 63] * p
        got = _from_hi(_hi(v) - ((m >> 6) << 20)) + 0.0
        got = (got.view(np.uint64) | (v.view(np.uint64) & np.uint64(0xffffffff))).view(np.float64)
        ref = np.exp(-t.astype(np.longdouble))
        err = np.abs(got.astype(np.longdouble) - ref) / np.spacing(ref.astype(np.float64))
        assert float(err.max()) <= 2.5, (a, b, float(err.max()))
