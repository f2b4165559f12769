"""TEST INFRASTRUCTURE ONLY: ctypes access to tests/emu/hpv_emu.cpp (the CUDA kernel bodies run on host
threads).  Built on demand with g++ into build/emu/ (git-ignored)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "hpv_emu.cpp")
OUT = os.path.join(ROOT, "build", "emu", "libhpv_emu.so")
CSRC = os.path.join(ROOT, "hp-vpinns_b200", "csrc")

_lib = None


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return any(os.path.getmtime(d) > t for d in deps)


def lib():
    global _lib
    if _lib is None:
        if _stale():
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-Wno-array-bounds", SRC, "-o", OUT],
                           check=True)
        _lib = ctypes.CDLL(OUT)
    return _lib


def _d(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _i(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


PROBLEM = {"poisson1d": 0, "poisson2d": 1, "advdiff": 2}
ACT = {"sin": 0, "tanh": 1}


def varloss(problem, var_form, layers, act, xi, w, T, D1, D2, d1b, lo, hi, ntx, nty, F, theta, eps=0.0, V=1.0,
            ntest=None, n_ctas_fwd=3, n_ctas_bwd=2, bwd_block=128, backward=True):
    L = lib()
    dim = 1 if problem == "poisson1d" else 2
    lo = np.asarray(lo, dtype=np.float64).reshape(-1, dim)
    n_el = lo.shape[0]
    if dim == 1:
        nty = 1
    keep = []
    def D(a):
        arr, p = _d(a); keep.append(arr); return p
    def I(a):
        arr, p = _i(a); keep.append(arr); return p
    loss = ctypes.c_double(0)
    geps = ctypes.c_double(0)
    res = np.zeros((n_el, nty, ntx), dtype=np.float32)
    el = np.zeros(n_el, dtype=np.float64)
    P = sum(layers[i] * layers[i + 1] + layers[i + 1] for i in range(len(layers) - 1))
    g = np.zeros(P, dtype=np.float64)
    rc = L.hpv_emu_varloss(
        ctypes.c_int(dim), I(layers), ctypes.c_int(len(layers)), ctypes.c_int(ACT[act]), ctypes.c_int(len(xi)), D(xi), D(w),
        ctypes.c_int(T.shape[0]), D(T), D(D1), D(D2), D(d1b), ctypes.c_int(PROBLEM[problem]), ctypes.c_int(var_form),
        ctypes.c_double(V), ctypes.c_int(n_el), D(lo), D(hi), I(ntest), ctypes.c_int(ntx), ctypes.c_int(nty), D(F), D(theta),
        ctypes.c_double(eps), ctypes.c_int(n_ctas_fwd), ctypes.c_int(n_ctas_bwd), ctypes.c_int(bwd_block),
        ctypes.byref(loss), res.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), el.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        g.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if backward else None, ctypes.byref(geps))
    if rc != 0:
        raise RuntimeError("hpv_emu_varloss failed: %d" % rc)
    return loss.value, res, el, g, geps.value


def points(layers, act, theta, pts, eps=0.0, target=None, a0=None, a1=None, weight=1.0, mode=2, backward=False, bwd_block=64):
    L = lib()
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n, dim = pts.shape
    keep = []
    def D(a):
        arr, p = _d(a); keep.append(arr); return p
    def I(a):
        arr, p = _i(a); keep.append(arr); return p
    u = np.zeros(n); d1 = np.zeros((n, dim)); d2 = np.zeros((n, dim))
    P = sum(layers[i] * layers[i + 1] + layers[i + 1] for i in range(len(layers) - 1))
    g = np.zeros(P)
    loss = ctypes.c_double(0); geps = ctypes.c_double(0)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    rc = L.hpv_emu_points(ctypes.c_int(dim), I(layers), ctypes.c_int(len(layers)), ctypes.c_int(ACT[act]), D(theta),
                          ctypes.c_double(eps), ctypes.c_int(n), D(pts), D(target), D(a0), D(a1), ctypes.c_double(weight),
                          ctypes.c_int(mode), dp(u), dp(d1), dp(d2), ctypes.byref(loss), dp(g) if backward else None,
                          ctypes.byref(geps), ctypes.c_int(bwd_block))
    if rc != 0:
        raise RuntimeError("hpv_emu_points failed: %d" % rc)
    return u, d1, d2, loss.value, g, geps.value
