"""lpm_v2_b200 -- B200-native direct-sum hot path of lpm-v2 behind the
reference's solver interface.  See DESIGN.md.

Submodules load lazily: `mesh` and `problems` need only the host-only liblpmmesh.so; `api`, `solvers`,
`torch_api` bind liblpmgpu.so (and raise at import if it has not been built -- there is no CPU fallback)."""
import importlib

_SUBMODULES = ("_lib", "_meshlib", "api", "mesh", "problems", "solvers", "torch_api", "dist")


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module("." + name, __name__)
    if name == "LpmError":
        from ._lib import LpmError
        return LpmError
    raise AttributeError(name)
