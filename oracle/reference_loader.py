"""TEST INFRASTRUCTURE ONLY -- executes the reference's own scripts from /root/reference in THIS container.

/root/reference does not exist on the GPU box, so nothing here is used at GPU-test / smoke / bench time;
the only caller is the fixture generator ``tests/golden/make_golden.py`` (and CPU tests that skip when the
reference tree is absent).  TensorFlow / pyDOE / matplotlib are replaced by ``oracle/tf1_shim``.

``load(which)``            -> module object holding the reference ``VPINN`` class (driver block NOT run).
``run_driver_setup(which)``-> namespace after running the reference ``__main__`` block UP TO (not including)
                              the ``model = VPINN(...)`` line, i.e. the unmodified problem set-up and RHS
                              assembly (P1D:231-329, P2D:279-426, ADI:351-483).
"""
import os
import sys
import types

REF_ROOT = "/root/reference"
SCRIPTS = {
    "P1D": "main/Poisson-1D/hp-VPINN-Poisson-1D.py",
    "P2D": "main/Poisson-2D/hp-VPINN-Poisson-2D.py",
    "ADI": "main/AdvDiff-Identification/hp-VPINN-AdvDiff-Identification.py",
}
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf1_shim")


def available():
    return all(os.path.exists(os.path.join(REF_ROOT, p)) for p in SCRIPTS.values())


def _prepare(which):
    path = os.path.join(REF_ROOT, SCRIPTS[which])
    for p in (os.path.dirname(path), _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    with open(path) as f:
        return path, f.read()


def load(which, **overrides):
    """Import the reference script as a module WITHOUT running its driver; ``overrides`` become the module
    globals the class reads at call time (var_form, scheme, LR, lossb_weight, V ...)."""
    path, src = _prepare(which)
    mod = types.ModuleType("hpvpinn_reference_" + which)
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    mod.__dict__.update(overrides)
    return mod


def run_driver_setup(which, stop_at="model = VPINN(", **overrides):
    """Run the reference driver block up to the first line starting with ``stop_at`` (default: the model
    construction); returns the namespace dict."""
    path, src = _prepare(which)
    lines = src.split("\n")
    cut = next(i for i, l in enumerate(lines) if l.strip().startswith(stop_at))
    body = "\n".join(lines[:cut])
    ns = {"__name__": "__main__", "__file__": path}
    if overrides:
        # hyper-parameters are plain assignments inside the driver: patch the literal lines
        out = []
        for l in body.split("\n"):
            key = l.strip().split("=")[0].strip() if "=" in l else None
            if key in overrides and l.strip().startswith(key) and not l.strip().startswith("#"):
                indent = l[:len(l) - len(l.lstrip())]
                l = "%s%s = %r" % (indent, key, overrides[key])
            out.append(l)
        body = "\n".join(out)
    exec(compile(body, path, "exec"), ns)
    return ns


def tf_shim():
    """Return the TF-1 stand-in module (puts oracle/tf1_shim on sys.path)."""
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    import tensorflow
    return tensorflow
