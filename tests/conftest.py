import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _ensure_built(gpu_library=False):
    """The .so files are git-ignored build products: build what is absent.  The CPU oracle and the host-only
    mesh library need gcc / g++ only and are built at collection; liblpmgpu.so needs nvcc and is built lazily
    by the fixtures that use it, so the oracle / mesh / gloo tests also run on a box without the CUDA toolkit."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "liblpm_oracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    if not os.path.exists(os.path.join(ROOT, "lpm_v2_b200", "liblpmmesh.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lpm_v2_b200"), "liblpmmesh.so"])
    if gpu_library and not os.environ.get("LPM_GPU_LIBRARY") and \
            not os.path.exists(os.path.join(ROOT, "lpm_v2_b200", "liblpmgpu.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lpm_v2_b200"), "liblpmgpu.so"])


_ensure_built()


def _have_b200():
    """A usable device for the `gpu` tests: the emulated library (tests/test_emu_abi.py sets LPM_GPU_LIBRARY), or a
    CUDA device of compute capability 10.x."""
    if os.environ.get("LPM_GPU_LIBRARY"):
        return True
    if not os.path.exists("/dev/nvidiactl"):
        return False
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
        except Exception:
            return False
    n = ctypes.c_int(0)
    if cudart.cudaGetDeviceCount(ctypes.byref(n)) != 0 or n.value < 1:
        return False
    major = ctypes.c_int(0)
    cudart.cudaDeviceGetAttribute(ctypes.byref(major), 75, 0)       # cudaDevAttrComputeCapabilityMajor
    return major.value == 10


def pytest_collection_modifyitems(config, items):
    """`gpu` tests skip (instead of erroring in lpm_gpu_init) when there is no B200."""
    if _have_b200():
        return
    skip = pytest.mark.skip(reason="no B200 (sm_100) device on this box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding


@pytest.fixture(scope="session")
def lpm():
    _ensure_built(gpu_library=True)
    import lpm_v2_b200
    return lpm_v2_b200


@pytest.fixture(scope="session")
def gpu(lpm):
    """Initialised library on the GPU box (single-process mode, all devices)."""
    from lpm_v2_b200 import api
    api.init(1)
    yield api


_mesh_cache = {}


@pytest.fixture(scope="session")
def get_mesh():
    from lpm_v2_b200 import mesh

    def make(seed, nest, amp=1.0):
        key = (seed, nest, amp)
        if key not in _mesh_cache:
            _mesh_cache[key] = mesh.PolyMesh2d(seed, nest, amp)
        return _mesh_cache[key]
    return make


def relerr(a, b):
    """max_i |a - b| / max_i |b|  (the parity metric)."""
    a, b = np.asarray(a), np.asarray(b)
    d = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (d if d > 0 else 1.0))
