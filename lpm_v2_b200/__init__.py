"""lpm_v2_b200 -- B200-native direct-sum hot path of lpm-v2 behind the
reference's solver interface.  See DESIGN.md."""
from . import _lib, api, mesh  # noqa: F401
from ._lib import LpmError  # noqa: F401
