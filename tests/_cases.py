"""Shared helpers for the parity tests: load a golden fixture and evaluate it with the float64 oracle."""
import glob
import os

import numpy as np

from oracle import hpvpinn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ACT = {"poisson1d": "sin", "poisson2d": "tanh", "advdiff": "tanh"}


def load(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    d["name"] = name
    d["kind"] = str(d["kind"])
    d["layers"] = [int(v) for v in d["layers"]]
    return d


def case_names(prefix=""):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz"))):
        n = os.path.basename(p)[:-4]
        if n.startswith(("p1d_", "p2d_", "adi_")):
            out.append(n)
    return out


def nodes_from_flat(XY):
    """Recover the 1-D nodes from the reference's flattened tensor grid (P2D:362-365)."""
    Q = int(round(np.sqrt(XY.shape[0])))
    return XY[:Q, 0].copy()


def oracle_lossv(c, theta=None, eps=None):
    """(lossv, residuals, grad_theta[, grad_eps]) of the factorised float64 oracle for golden case ``c``."""
    theta = c["theta"] if theta is None else theta
    Ws, bs = O.unpack_theta(theta, c["layers"])
    vf = int(c["var_form"])
    if c["kind"] == "poisson2d":
        X = nodes_from_flat(c["X_quad"]); Q = len(X)
        WX = c["W_quad"][:Q, 0]
        fn = lambda W, b: O.varloss_2d_factorised(W, b, X, WX, c["F_ext"], c["gridx"], c["gridy"], int(c["Ntx"]), int(c["Nty"]), vf)
    elif c["kind"] == "advdiff":
        X = c["T_quad"]; WX = c["WT_quad"]
        e = float(c["eps0"]) if eps is None else eps
        fn3 = lambda W, b, ee: O.varloss_2d_factorised(W, b, X, WX, None, c["grid_x"], c["grid_t"], int(c["Ntx"]), int(c["Ntt"]), vf,
                                                      problem="advdiff", eps=ee, V=float(c["V"]))
        res = fn3(Ws, bs, e)[1].detach().numpy()
        l, g, ge = O.loss_and_grad(lambda W, b, ee: fn3(W, b, ee)[0], Ws, bs, extra=np.array([e]))
        return l, res, g, ge
    else:
        fn = lambda W, b: O.varloss_1d_factorised(W, b, c["X_quad"], c["W_quad"], c["F_ext"], c["grid"], vf)
    res = fn(Ws, bs)[1].detach().numpy()
    l, g = O.loss_and_grad(lambda W, b: fn(W, b)[0], Ws, bs)
    return l, res, g


def engine_inputs(c, N=None):
    """Inputs of the engine (C-ABI / emulation harness) for golden case ``c``: element corners in the
    reference's loop order (ex outer, ey inner: P2D:69-70), 1-D nodes/weights, test tables on those nodes."""
    kind = c["kind"]
    out = dict(problem=kind, var_form=int(c["var_form"]), layers=c["layers"], act=ACT[kind], theta=c["theta"], eps=0.0, V=1.0)
    if kind == "poisson1d":
        xi = c["X_quad"].ravel(); w = c["W_quad"].ravel()
        g = c["grid"]
        lo, hi = g[:-1, None], g[1:, None]
        ntx, nty = int(c["N"]), 1
        F = np.asarray(c["F_ext"]).reshape(len(g) - 1, 1, ntx)
    else:
        if kind == "poisson2d":
            xi = nodes_from_flat(c["X_quad"]); w = c["W_quad"][:len(xi), 0]
            gx, gy = c["gridx"], c["gridy"]
            ntx, nty = int(c["Ntx"]), int(c["Nty"])
            F = np.asarray(c["F_ext"]).reshape(-1, nty, ntx)
        else:
            xi = c["T_quad"]; w = c["WT_quad"]
            gx, gy = c["grid_x"], c["grid_t"]
            ntx, nty = int(c["Ntx"]), int(c["Ntt"])
            F = None
            out["eps"] = float(c["eps0"]); out["V"] = float(c["V"])
        lo = np.array([[gx[ex], gy[ey]] for ex in range(len(gx) - 1) for ey in range(len(gy) - 1)])
        hi = np.array([[gx[ex + 1], gy[ey + 1]] for ex in range(len(gx) - 1) for ey in range(len(gy) - 1)])
    Nmax = max(ntx, nty) if N is None else N
    T = O.Test_fcn(Nmax, xi)
    D1, D2 = O.dTest_fcn(Nmax, xi)
    d1b, _ = O.dTest_fcn(Nmax, np.array([-1.0, 1.0]))
    out.update(xi=xi, w=w, T=T, D1=D1, D2=D2, d1b=d1b, lo=lo, hi=hi, ntx=ntx, nty=nty, F=F)
    return out
