"""Headline size (icosTri level 8, BASELINE.json configs[3]): properties that do
not need an O(N^2) CPU pass, plus oracle spot checks on sampled targets, plus
the slice / device-pointer path the multi-GPU runs use."""
import numpy as np
import pytest

from lpm_v2_b200 import mesh as M, problems
from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-12
PI = problems.PI


@pytest.fixture(scope="module")
def l8(get_mesh):
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, 8)
    assert m.n == 1966082 and m.n_active == 1310720
    return m


@pytest.fixture(scope="module")
def l8_rh54(gpu, l8):
    zeta = problems.rossby_haurwitz54(l8)
    return zeta, gpu.bve_velocity(l8.x, l8.y, l8.z, zeta, l8.area, l8.is_active, 1.0)


def test_l8_spot_check_against_oracle(gpu, oracle, l8, l8_rh54):
    zeta, (u, v, w) = l8_rh54
    rng = np.random.default_rng(8)
    idx = np.concatenate([[0, 1, 11, 12, 31, 32, l8.n - 1], rng.integers(0, l8.n, 57)])
    scale = max(np.abs(u).max(), np.abs(v).max(), np.abs(w).max())
    worst = 0.0
    for i in idx:
        ou, ov, ow = oracle.bve_velocity(l8.x, l8.y, l8.z, zeta, l8.area, l8.is_active, 1.0, rng=(int(i), int(i) + 1))
        worst = max(worst, abs(u[i] - ou[i]), abs(v[i] - ov[i]), abs(w[i] - ow[i]))
    assert worst <= TOL * scale


def test_l8_velocity_is_tangent_and_finite(l8, l8_rh54):
    _, (u, v, w) = l8_rh54
    assert np.all(np.isfinite(u)) and np.all(np.isfinite(v)) and np.all(np.isfinite(w))
    radial = l8.x * u + l8.y * v + l8.z * w           # u = x cross a  =>  x . u = 0
    assert np.abs(radial).max() <= 1e-13 * np.abs(u).max()


def test_l8_linearity(gpu, l8, l8_rh54):
    """The sum is linear in the vorticity: u(a z1 + b z2) = a u(z1) + b u(z2)."""
    z1, (u1, v1, w1) = l8_rh54
    z2 = problems.gaussian_vortex(l8)
    u2, v2, w2 = gpu.bve_velocity(l8.x, l8.y, l8.z, z2, l8.area, l8.is_active, 1.0)
    a, b = 0.75, -1.5
    u3, v3, w3 = gpu.bve_velocity(l8.x, l8.y, l8.z, a * z1 + b * z2, l8.area, l8.is_active, 1.0)
    for c3, c1, c2 in ((u3, u1, u2), (v3, v1, v2), (w3, w1, w2)):
        assert relerr(c3, a * c1 + b * c2) <= TOL


def test_l8_slices_are_bitwise_independent_of_partition(gpu, l8, l8_rh54):
    """One-sided engine: the device-pointer API on LoadBalance slices for 8 ranks == the 1-GPU one-sided result,
    bit for bit (a target's sum does not depend on the slice it is in).  The default whole evaluation
    (pair-symmetric at this size) differs from it by summation order only; its own independence of the rank
    count is tested in tests/test_multigpu.py and tests/test_emu_abi.py."""
    import torch
    from lpm_v2_b200 import torch_api, api
    zeta, sym_uvw = l8_rh54
    gpu.set_symmetric(False)
    try:
        u, v, w = gpu.bve_velocity(l8.x, l8.y, l8.z, zeta, l8.area, l8.is_active, 1.0)
    finally:
        gpu.set_symmetric(True)
    for a, b in zip(sym_uvw, (u, v, w)):
        assert relerr(a, b) <= 1e-12
        assert not np.array_equal(a, b)          # the default did take the other path
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    t = {k: torch.from_numpy(np.ascontiguousarray(a)).to(dev) for k, a in
         dict(x=l8.x, y=l8.y, z=l8.z, q=zeta, a=l8.area).items()}
    mask = torch.from_numpy(l8.is_active).to(dev)
    out = [torch.full((l8.n,), float("nan"), dtype=torch.float64, device=dev) for _ in range(3)]
    s, e, _ = api.load_balance(l8.n, 8)
    for r in (0, 3, 7):                      # first, middle and the ragged last slice
        torch_api.bve_velocity_dev(t["x"], t["y"], t["z"], t["q"], t["a"], mask, 1.0, int(s[r]) - 1, int(e[r]), *out)
    torch.cuda.synchronize()
    for r in (0, 3, 7):
        sl = slice(int(s[r]) - 1, int(e[r]))
        for o, full in zip(out, (u, v, w)):
            assert np.array_equal(o[sl].cpu().numpy(), full[sl])
    # untouched slices stay untouched
    assert torch.isnan(out[0][int(s[1]) - 1:int(e[1])]).all()


def test_l8_active_list(gpu, oracle, l8):
    got = gpu.active_list(l8.is_active)
    assert np.array_equal(got, oracle.active_list(l8.is_active))
