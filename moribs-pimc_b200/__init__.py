"""moribs-pimc_b200: B200-native (sm_100a) sampling hot path of MoRiBS-PIMC.

The directory name carries a hyphen (it mirrors the reference's repository
name), so import it through ``__graft_entry__.load_package()`` which registers
it as ``moribs_pimc_b200``.
"""
from . import configs  # noqa: F401
from . import gpu      # noqa: F401  (ctypes binding; loading the .so is deferred to first use and fails loudly)

__all__ = ["configs", "gpu"]
