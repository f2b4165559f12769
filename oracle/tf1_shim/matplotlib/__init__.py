"""TEST INFRASTRUCTURE ONLY -- import-time stand-in for matplotlib (absent from this image).  The reference
scripts import it at module scope (P1D:20-21, P2D:18, ADI:17-20); plotting itself is out of scope, so every
attribute/call is absorbed by an inert object."""


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))

    def __getitem__(self, k):
        return _Anything()

    def __setitem__(self, k, v):
        pass


def __getattr__(name):
    return _Anything()
