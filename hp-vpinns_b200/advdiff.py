"""`from hpv_b200.advdiff import VPINN` -- the class a reference script binds in place of its own `class VPINN`."""
from .vpinn import VPINN_AdvDiff as VPINN  # noqa: F401
