"""A COMPILED C caller of the C ABI (tests/c_caller/caller.c): built with gcc -std=c99 against include/lpm_gpu.h
and include/lpm_mesh.h, linked against liblpmgpu.so / liblpmmesh.so, arrays passed by reference and scalars by
value as the bind(C) interfaces of lpm_v2_b200/fortran/lpm_gpu.f90 declare them.  It is the closest stand-in for
the Fortran host this image allows (no Fortran compiler): the headers are valid C, the symbols link, the calling
convention is the one a compiler generates (not ctypes' guess), and New / Timestep / Delete + one
BVESphereVelocity evaluation (src/SphereBVESolver.f90:80-90, 219-353, 377-430) meet the parity bound."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def caller(lpm, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("c_caller") / "caller")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_caller", "caller.c"),
           "-L", os.path.join(ROOT, "lpm_v2_b200"), "-llpmgpu", "-llpmmesh",
           "-L", os.path.join(ROOT, "oracle"), "-llpm_oracle", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(ROOT, "lpm_v2_b200"), os.path.join(ROOT, "oracle"),
                                              env.get("LD_LIBRARY_PATH", "")])
    return exe, env


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_c_caller_builds_links_and_fails_loudly_without_a_gpu(caller):
    exe, env = caller
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_caller_velocity_and_solver_step(caller):
    exe, env = caller
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "C_CALLER_OK" in r.stdout, (r.returncode, r.stdout, r.stderr)
