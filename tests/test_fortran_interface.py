"""The reference-side binding (lpm_v2_b200/fortran/lpm_gpu.f90 + the generated lpm_gpu_interface.inc) against the
C header.  No Fortran compiler exists in this image, so these are structural checks: the interface file is the
generator's output for the CURRENT header; every entry point of include/lpm_gpu.h has exactly one bind(C) body
with the same name and argument count; scalars are passed by value; and the wrapper module includes the file."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "lpm_v2_b200", "fortran", "lpm_gpu_interface.inc")


def _gen():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_fortran_interface as g
    return g


def test_interface_file_is_in_sync_with_the_header():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_fortran_interface.py"), "--check"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_every_entry_point_is_bound_with_matching_arguments():
    g = _gen()
    text = open(INC).read()
    text = re.sub(r"&\n\s*", "", text)                  # join continuation lines
    bodies = dict((m.group(2), (m.group(1), m.group(3), m.group(4))) for m in re.finditer(
        r"\t(?:[a-z_()0-9]+ function|subroutine) ((lpm_[a-z0-9_]+))\(([^)]*)\) bind\(C, name=\"\2\"\)\n(.*?)\n\tend (?:function|subroutine)",
        text, flags=re.S))
    protos = g.prototypes()
    assert len(protos) >= 60 and sorted(bodies) == sorted(p[1] for p in protos)
    for ret, fn, params in protos:
        _, arglist, body = bodies[fn]
        args = [a.strip() for a in arglist.split(",") if a.strip()]
        assert args == [p["name"] for p in params], fn
        for p in params:
            decl = [ln for ln in body.splitlines() if re.search(r"::.*\b" + p["name"] + r"\b", ln)]
            assert len(decl) == 1, (fn, p["name"])
            by_value = ", value" in decl[0]
            is_scalar_c = p["stars"] == 0 and p["dim"] is None
            assert by_value == (is_scalar_c or "type(c_ptr), value" in decl[0] or "type(c_funptr), value" in decl[0]), (fn, p["name"], decl[0])


def test_wrapper_module_includes_the_interface_and_binds_solvers():
    mod = open(os.path.join(ROOT, "lpm_v2_b200", "fortran", "lpm_gpu.f90")).read()
    assert 'include "lpm_gpu_interface.inc"' in mod
    inc = open(INC).read()
    # the entry points VERDICT r01 found missing from the hand-written module
    for fn in ("lpm_plane_solver_new", "lpm_plane_solver_timestep", "lpm_plane_solver_get_state", "lpm_plane_solver_delete",
               "lpm_betaplane_solver_new", "lpm_betaplane_solver_timestep", "lpm_betaplane_solver_get_state",
               "lpm_betaplane_solver_delete", "lpm_pse_double_dot_sphere", "lpm_pse_interpolate_plane", "lpm_set_symmetric"):
        assert f'name="{fn}"' in inc, fn
