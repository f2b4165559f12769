"""hp-vpinns_b200: B200-native engine for the hp-VPINN variational-residual hot path.

The directory name is not a Python identifier; import it through the loader module at the repository root:

    import hpv_b200                      # -> this package
    from hpv_b200 import Engine, poisson2d

Compute lives in libhpv.so (csrc/, C ABI in include/hpv.h), built in-tree by build.py.  There is no CPU
fallback: constructing an Engine without the library or without a B200 raises.
"""
from ._lib import HpvError, LIB_PATH, SIGNATURES, load  # noqa: F401
from .engine import Engine, SIN, TANH, POISSON1D, POISSON2D, ADVDIFF  # noqa: F401

__all__ = ["Engine", "HpvError", "load", "LIB_PATH", "SIGNATURES", "SIN", "TANH", "POISSON1D", "POISSON2D", "ADVDIFF"]
