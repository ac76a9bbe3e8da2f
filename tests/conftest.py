import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    """The .so files are git-ignored build products: build them if absent."""
    lib = os.path.join(ROOT, "lpm_v2_b200", "liblpmgpu.so")
    ora = os.path.join(ROOT, "oracle", "liblpm_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(ora)):
        import __graft_entry__ as g
        g.build()


_ensure_built()


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding


@pytest.fixture(scope="session")
def lpm():
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
