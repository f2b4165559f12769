"""`from hpv_b200.poisson1d import VPINN` -- the class a reference script binds in place of its own `class VPINN`."""
from .vpinn import VPINN_Poisson1D as VPINN  # noqa: F401
