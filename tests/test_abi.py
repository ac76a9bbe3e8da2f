"""C-ABI surface: the library loads on a CPU-only box, exports every symbol that
include/lpm_gpu.h declares, fails loudly (no CPU fallback) without a device,
and the product never touches oracle/."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="lpm_gpu.h"):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lpm_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported(lpm):
    from lpm_v2_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in lpm_gpu.h but not exported"
    # and the ctypes table binds exactly the header's set
    assert sorted(_lib.PROTOTYPES) == syms
    # the tuning hook of csrc/lpm_gpu_tuning.h is deliberately NOT part of the ABI header
    assert "lpm_tune" not in syms and "lpm_set_bve_variant" not in syms


def test_mesh_header_symbols_all_exported():
    """include/lpm_mesh.h <-> liblpmmesh.so, which must not pull in the GPU library or CUDA."""
    import subprocess
    from lpm_v2_b200 import _meshlib
    syms = _declared_symbols("lpm_mesh.h")
    assert len(syms) == 11
    for s in syms:
        assert hasattr(_meshlib.lib, s), f"{s} declared in lpm_mesh.h but not exported"
    assert sorted(_meshlib.PROTOTYPES) == syms
    out = subprocess.run(["ldd", _meshlib.LIB_PATH], capture_output=True, text=True).stdout
    assert "cuda" not in out.lower() and "lpmgpu" not in out


def test_reference_arm_does_not_map_the_gpu_library():
    """bench.py --impl reference times the CPU port: the process must never load liblpmgpu.so."""
    import subprocess
    import sys
    code = ("import sys; sys.argv=['bench.py','--impl','reference','--level','3','--steps','1','--warmup','0'];"
            "import runpy; runpy.run_path('bench.py', run_name='__main__');"
            "maps=open('/proc/self/maps').read(); assert 'liblpmgpu' not in maps, 'GPU library mapped'; "
            "assert 'liblpmmesh' in maps and 'liblpm_oracle' in maps; print('MAPS_OK')")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "MAPS_OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


def test_load_balance_matches_reference_rule(lpm, oracle):
    """src/MPISetup.f90:132-146: chunk = n / p, remainder to the last rank."""
    from lpm_v2_b200 import api
    for n, p in [(1922, 4), (1922, 1), (1922, 8), (30722, 7), (10, 3), (7, 8), (1966082, 8)]:
        s, e, m = api.load_balance(n, p)
        os_, oe, om = oracle.load_balance(n, p)
        assert np.array_equal(s, os_) and np.array_equal(e, oe) and np.array_equal(m, om)
        chunk = n // p
        assert s[0] == 1 and e[-1] == n
        assert all(s[r] == r * chunk + 1 for r in range(p))
        assert m.sum() == n


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback(lpm):
    """Without a device every compute entry point must fail, not compute."""
    from lpm_v2_b200 import api, LpmError
    x = np.ones(8)
    m = np.ones(8, np.int32)
    with pytest.raises(LpmError) as ei:
        api.init()
    assert ei.value.code == 2
    with pytest.raises(LpmError) as ei:
        api.bve_velocity(x, x, x, x, x, m, 1.0)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(LpmError):
        api.pse_laplacian_sphere(x, x, x, x, x, m, 0.1)
    with pytest.raises(LpmError):
        api.active_list(m)


def test_product_does_not_use_oracle():
    """The oracle is test infrastructure: nothing under lpm_v2_b200/ or include/
    may import, include or link it."""
    bad = []
    for base in ("lpm_v2_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".so", ".pyc", ".o")):
                    continue
                txt = open(os.path.join(dp, fn), errors="replace").read()
                if re.search(r"(from|import)\s+oracle|lpm_oracle|oracle/|liblpm_oracle", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
    # the shared object does not link the oracle either
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "lpm_v2_b200", "liblpmgpu.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_argument_validation(lpm):
    from lpm_v2_b200 import _lib, _meshlib
    s = np.zeros(2, np.int64)
    p = C.POINTER(C.c_int64)
    assert _lib.lib.lpm_load_balance(10, 0, s.ctypes.data_as(p), s.ctypes.data_as(p), s.ctypes.data_as(p)) == 1
    assert "lpm_load_balance" in _lib.last_error()
    h = C.c_void_p()
    assert _meshlib.lib.lpm_mesh_create(999, 1, 1.0, C.byref(h)) == 1     # invalid meshSeed
    assert _meshlib.lib.lpm_mesh_create(205, -1, 1.0, C.byref(h)) == 1
