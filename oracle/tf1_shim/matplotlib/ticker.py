"""TEST INFRASTRUCTURE ONLY -- inert matplotlib sub-module stand-in (see matplotlib/__init__.py)."""
from . import _Anything


def __getattr__(name):
    return _Anything()
