"""TEST INFRASTRUCTURE ONLY -- float64 CPU oracle for the hp-VPINN variational-residual hot path.

This file restates, on the CPU in float64, what the reference computes for the path
``net_u -> input derivatives -> projection on Jacobi test functions -> element residual -> lossv``.
It is the CHECKER for the CUDA engine; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``hp-vpinns_b200/``) never does.

Pinning status: the reference's arithmetic lives in TensorFlow 1.x (third-party, un-vendored, version not
pinned anywhere in the reference; TF cannot be installed here), and the reference ships no tests or golden
vectors.  The oracle is therefore pinned (tests/test_oracle_vs_golden.py) against fixtures produced by
running the reference's OWN, UNMODIFIED ``VPINN`` classes and ``GaussJacobiQuadRule_V3`` in this container
on a TF-1 API stand-in (``oracle/tf1_shim``; generator ``tests/golden/make_golden.py``).  Against real
TensorFlow arithmetic it is "parity unpinned": float64 matmul/tanh/sin/reduce_sum in TF and torch agree
only up to summation order (~1e-15 rel).

Abbreviations (all under /root/reference): GJQ = Utilities/GaussJacobiQuadRule_V3.py,
P1D = main/Poisson-1D/hp-VPINN-Poisson-1D.py, P2D = main/Poisson-2D/hp-VPINN-Poisson-2D.py,
ADI = main/AdvDiff-Identification/hp-VPINN-AdvDiff-Identification.py.

Two restatements are kept:
* ``*_literal``: torch-float64 autograd with the reference's op granularity (MLP re-evaluated per derivative
  helper, one reduction per (k, r) test-function pair) -- the CPU baseline the bench times;
* ``*_factorised``: the sum-factorised form ``U = c * (T*w) @ G @ (D*w).T`` with analytic forward-mode MLP
  derivatives -- the form the CUDA kernels implement; differentiable (torch) for d(theta), d(eps).
"""
import numpy as np
import torch
from scipy.special import gamma as _gamma
from scipy.special import jacobi as _sp_jacobi
from scipy.special import roots_jacobi as _roots_jacobi

F64 = torch.float64


# --------------------------------------------------------------------------------------------------
# Quadrature and test functions (setup side)
# --------------------------------------------------------------------------------------------------
def Jacobi(n, a, b, x):
    """GJQ:24-26 -- value of the Jacobi polynomial P_n^{(a,b)} at x."""
    x = np.array(x)
    return _sp_jacobi(n, a, b)(x)


def DJacobi(n, a, b, x, k):
    """GJQ:30-33 -- k-th derivative of P_n^{(a,b)}."""
    x = np.array(x)
    ctemp = _gamma(a + b + n + 1 + k) / (2 ** k) / _gamma(a + b + n + 1)
    return ctemp * Jacobi(n - k, a + k, b + k, x)


def GaussLobattoJacobiWeights(Q, a, b):
    """GJQ:46-61 -- Gauss-Lobatto-Jacobi nodes and weights (end points included)."""
    X = _roots_jacobi(Q - 2, a + 1, b + 1)[0]
    if a == 0 and b == 0:
        W = 2 / ((Q - 1) * Q * (Jacobi(Q - 1, 0, 0, X) ** 2))
        Wl = 2 / ((Q - 1) * Q * (Jacobi(Q - 1, 0, 0, -1) ** 2))
        Wr = 2 / ((Q - 1) * Q * (Jacobi(Q - 1, 0, 0, 1) ** 2))
    else:
        c = 2 ** (a + b + 1) * _gamma(a + Q) * _gamma(b + Q) / ((Q - 1) * _gamma(Q) * _gamma(a + b + Q + 1))
        W = c / (Jacobi(Q - 1, a, b, X) ** 2)
        Wl = (b + 1) * c / (Jacobi(Q - 1, a, b, -1) ** 2)
        Wr = (a + 1) * c / (Jacobi(Q - 1, a, b, 1) ** 2)
    W = np.append(W, Wr)
    W = np.append(Wl, W)
    X = np.append(X, 1)
    X = np.append(-1, X)
    return [X, W]


def Test_fcn(N_test, x):
    """P1D:157-162, P2D:196-208, ADI:257-262 -- phi_n = P_{n+1} - P_{n-1}, n = 1..N."""
    return np.asarray([Jacobi(n + 1, 0, 0, x) - Jacobi(n - 1, 0, 0, x) for n in range(1, N_test + 1)])


def dTest_fcn(N_test, x):
    """P1D:164-183, P2D:210-229, ADI:265-284 -- first and second derivative of phi_n (reference coords)."""
    d1, d2 = [], []
    for n in range(1, N_test + 1):
        if n == 1:
            d1.append(((n + 2) / 2) * Jacobi(n, 1, 1, x))
            d2.append(((n + 2) * (n + 3) / (2 * 2)) * Jacobi(n - 1, 2, 2, x))
        elif n == 2:
            d1.append(((n + 2) / 2) * Jacobi(n, 1, 1, x) - (n / 2) * Jacobi(n - 2, 1, 1, x))
            d2.append(((n + 2) * (n + 3) / (2 * 2)) * Jacobi(n - 1, 2, 2, x))
        else:
            d1.append(((n + 2) / 2) * Jacobi(n, 1, 1, x) - (n / 2) * Jacobi(n - 2, 1, 1, x))
            d2.append(((n + 2) * (n + 3) / (2 * 2)) * Jacobi(n - 1, 2, 2, x)
                      - (n * (n + 1) / (2 * 2)) * Jacobi(n - 3, 2, 2, x))
    return np.asarray(d1), np.asarray(d2)


def tensor_quadrature(Q):
    """P2D:360-365 / ADI:395-400 -- flattened tensor grid: point p = j*Q + i -> (X[i], Y[j])."""
    X, WX = GaussLobattoJacobiWeights(Q, 0, 0)
    xx, yy = np.meshgrid(X, X)
    wxx, wyy = np.meshgrid(WX, WX)
    XY = np.hstack((xx.flatten()[:, None], yy.flatten()[:, None]))
    WXY = np.hstack((wxx.flatten()[:, None], wyy.flatten()[:, None]))
    return X, WX, XY, WXY


# --------------------------------------------------------------------------------------------------
# Manufactured solutions and RHS assembly (driver side; inputs of the hot path)
# --------------------------------------------------------------------------------------------------
def u_ext_2d(x, y, omegax=2 * np.pi, omegay=2 * np.pi, r1=10):
    """P2D:303-305."""
    return (0.1 * np.sin(omegax * x) + np.tanh(r1 * x)) * np.sin(omegay * y)


def f_ext_2d(x, y, omegax=2 * np.pi, omegay=2 * np.pi, r1=10):
    """P2D:307-310."""
    return ((-0.1 * (omegax ** 2) * np.sin(omegax * x) - (2 * r1 ** 2) * (np.tanh(r1 * x)) / ((np.cosh(r1 * x)) ** 2))
            * np.sin(omegay * y)
            + (0.1 * np.sin(omegax * x) + np.tanh(r1 * x)) * (-omegay ** 2 * np.sin(omegay * y)))


def u_ext_1d(x, omega=8 * np.pi, amp=1, r1=80):
    """P1D:251-253."""
    return amp * (0.1 * np.sin(omega * x) + np.tanh(r1 * x))


def f_ext_1d(x, omega=8 * np.pi, amp=1, r1=80):
    """P1D:255-257."""
    g = -0.1 * (omega ** 2) * np.sin(omega * x) - (2 * r1 ** 2) * (np.tanh(r1 * x)) / ((np.cosh(r1 * x)) ** 2)
    return -amp * g


def rhs_2d_literal(grid_x, grid_y, N_test_x, N_test_y, XY_quad, WXY_quad, f=f_ext_2d):
    """P2D:384-414, loops kept literal: F[ex,ey,k,r] = J*sum(wx*phi_r(x)*wy*phi_k(y)*f)."""
    x_quad, y_quad, w_quad = XY_quad[:, 0:1], XY_quad[:, 1:2], WXY_quad
    NE_x, NE_y = len(grid_x) - 1, len(grid_y) - 1
    F_total = []
    for ex in range(NE_x):
        for ey in range(NE_y):
            Ntx, Nty = N_test_x[ex], N_test_y[ey]
            xe = grid_x[ex] + (grid_x[ex + 1] - grid_x[ex]) / 2 * (x_quad + 1)
            ye = grid_y[ey] + (grid_y[ey + 1] - grid_y[ey]) / 2 * (y_quad + 1)
            jac = ((grid_x[ex + 1] - grid_x[ex]) / 2) * ((grid_y[ey + 1] - grid_y[ey]) / 2)
            tx = np.asarray([Jacobi(n + 1, 0, 0, x_quad) - Jacobi(n - 1, 0, 0, x_quad) for n in range(1, Ntx + 1)])
            ty = np.asarray([Jacobi(n + 1, 0, 0, y_quad) - Jacobi(n - 1, 0, 0, y_quad) for n in range(1, Nty + 1)])
            fq = f(xe, ye)
            F_total.append(np.asarray([[jac * np.sum(w_quad[:, 0:1] * tx[r] * w_quad[:, 1:2] * ty[k] * fq)
                                        for r in range(Ntx)] for k in range(Nty)]))
    return np.reshape(F_total, [NE_x, NE_y, N_test_y[0], N_test_x[0]])


def rhs_2d_factorised(grid_x, grid_y, Ntx, Nty, X, WX, f=f_ext_2d):
    """Same numbers as :func:`rhs_2d_literal` (P2D:384-414) via F = J * (Ty*w) @ f(x,y) @ (Tx*w).T."""
    A = Test_fcn(max(Ntx, Nty), X) * WX[None, :]
    NE_x, NE_y = len(grid_x) - 1, len(grid_y) - 1
    out = np.zeros((NE_x, NE_y, Nty, Ntx))
    for ex in range(NE_x):
        for ey in range(NE_y):
            xe = grid_x[ex] + (grid_x[ex + 1] - grid_x[ex]) / 2 * (X + 1)
            ye = grid_y[ey] + (grid_y[ey + 1] - grid_y[ey]) / 2 * (X + 1)
            jac = ((grid_x[ex + 1] - grid_x[ex]) / 2) * ((grid_y[ey + 1] - grid_y[ey]) / 2)
            G = f(xe[None, :], ye[:, None])
            out[ex, ey] = jac * A[:Nty] @ G @ A[:Ntx].T
    return out


def rhs_1d(grid, N_test_total, x_quad, w_quad, f=f_ext_1d):
    """P1D:275-294: F[e][i] = J*sum(w*f(x_e)*phi_i)."""
    F_total = []
    for e in range(len(grid) - 1):
        xe = grid[e] + (grid[e + 1] - grid[e]) / 2 * (x_quad + 1)
        jac = (grid[e + 1] - grid[e]) / 2
        Nt = N_test_total[e]
        tf_e = np.asarray([Jacobi(n + 1, 0, 0, x_quad) - Jacobi(n - 1, 0, 0, x_quad) for n in range(1, Nt + 1)])
        fq = f(xe)
        Fe = jac * np.asarray([sum(w_quad * fq * tf_e[i]) for i in range(Nt)])
        F_total.append(Fe[:, None])
    return np.asarray(F_total)


# --------------------------------------------------------------------------------------------------
# Parameters
# --------------------------------------------------------------------------------------------------
def xavier_params(layers, seed=1234):
    """Shape/scale of P1D:110-126 (truncated normal, std sqrt(2/(in+out)), zero biases) from numpy's
    ``default_rng`` -- TF's RNG stream is not reproducible, so the parity tests inject these."""
    rng = np.random.default_rng(seed)
    Ws, bs = [], []
    for l in range(len(layers) - 1):
        std = np.sqrt(2.0 / (layers[l] + layers[l + 1]))
        w = rng.standard_normal((layers[l], layers[l + 1]))
        bad = np.abs(w) > 2
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2
        Ws.append(std * w)
        bs.append(np.zeros((1, layers[l + 1])))
    return Ws, bs


def pack_theta(Ws, bs):
    """Flat parameter vector used at the C-ABI: for each layer W (in x out, row-major) then b (out)."""
    return np.concatenate([np.concatenate([np.asarray(W).ravel(), np.asarray(b).ravel()]) for W, b in zip(Ws, bs)])


def unpack_theta(theta, layers):
    Ws, bs, o = [], [], 0
    for l in range(len(layers) - 1):
        n = layers[l] * layers[l + 1]
        Ws.append(np.asarray(theta[o:o + n]).reshape(layers[l], layers[l + 1])); o += n
        bs.append(np.asarray(theta[o:o + layers[l + 1]]).reshape(1, layers[l + 1])); o += layers[l + 1]
    return Ws, bs


def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a), dtype=F64)


# --------------------------------------------------------------------------------------------------
# MLP: value and input derivatives
# --------------------------------------------------------------------------------------------------
def neural_net(X, Ws, bs, act):
    """P1D:128-138 (act='sin'), P2D:158-169 / ADI:219-231 (act='tanh')."""
    f = torch.sin if act == "sin" else torch.tanh
    H = X
    for W, b in zip(Ws[:-1], bs[:-1]):
        H = f(torch.add(torch.matmul(H, W), b))
    return torch.add(torch.matmul(H, Ws[-1]), bs[-1])


def _grad(y, x):
    """tf.gradients(y, x)[0] == d(sum y)/dx, differentiable again (P2D:177-178)."""
    return torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=True)[0]


def net_d_autograd(cols, Ws, bs, act, wrt, second=True):
    """P2D:175-185 / P1D:144-148 / ADI:236-245: rebuild net_u, differentiate twice w.r.t. column ``wrt``."""
    u = neural_net(torch.cat(cols, 1), Ws, bs, act)
    d1 = _grad(u, cols[wrt])
    d2 = _grad(d1, cols[wrt]) if second else None
    return d1, d2


def mlp_forward_mode(pts, Ws, bs, act, need_d2=True):
    """Analytic forward-mode restatement of net_u + tf.gradients chains: returns
    u (n,), d1 (n, dim), d2 (n, dim) [pure second derivatives].  torch float64, differentiable in Ws/bs."""
    pts = _t(pts)
    n, dim = pts.shape
    Ws = [_t(W) for W in Ws]
    bs = [_t(b) for b in bs]
    h = pts
    dh = [torch.zeros(n, dim, dtype=F64) for _ in range(dim)]
    for d in range(dim):
        dh[d][:, d] = 1.0
    ddh = [torch.zeros(n, dim, dtype=F64) for _ in range(dim)]
    for l, (W, b) in enumerate(zip(Ws, bs)):
        z = h @ W + b
        dz = [t @ W for t in dh]
        ddz = [t @ W for t in ddh]
        if l == len(Ws) - 1:
            return z[:, 0], torch.stack([t[:, 0] for t in dz], 1), torch.stack([t[:, 0] for t in ddz], 1)
        if act == "sin":
            a, s1 = torch.sin(z), torch.cos(z)
            s2 = -a
        else:
            a = torch.tanh(z)
            s1 = 1 - a * a
            s2 = -2 * a * s1
        h = a
        ddh = [s2 * dzd * dzd + s1 * ddzd for dzd, ddzd in zip(dz, ddz)]
        dh = [s1 * dzd for dzd in dz]


# --------------------------------------------------------------------------------------------------
# 2-D Poisson variational loss (P2D:68-120)
# --------------------------------------------------------------------------------------------------
def varloss_2d_literal(Ws, bs, XY_quad, WXY_quad, F_ext_total, gridx, gridy, N_testfcn, var_form=1,
                       elements=None):
    """Line-by-line torch-float64 restatement of P2D:68-120 (one reduce_sum per (k, r) pair, MLP rebuilt in
    every derivative helper).  ``elements`` optionally restricts the (ex, ey) pairs (bounded CPU-baseline
    samples).  Returns (lossv, list of residual arrays in loop order)."""
    Ws = [_t(W) for W in Ws]
    bs = [_t(b) for b in bs]
    xquad, yquad, wquad = XY_quad[:, 0:1], XY_quad[:, 1:2], WXY_quad
    NEx, NEy = np.size(N_testfcn[0]), np.size(N_testfcn[1])
    total, residuals = 0, []
    pairs = elements if elements is not None else [(ex, ey) for ex in range(NEx) for ey in range(NEy)]
    for ex, ey in pairs:
        F_el = _t(F_ext_total[ex, ey])
        Ntx, Nty = N_testfcn[0][ex], N_testfcn[1][ey]
        x = _t(gridx[ex] + (gridx[ex + 1] - gridx[ex]) / 2 * (xquad + 1)).requires_grad_(True)
        y = _t(gridy[ey] + (gridy[ey + 1] - gridy[ey]) / 2 * (yquad + 1)).requires_grad_(True)
        jx = (gridx[ex + 1] - gridx[ex]) / 2
        jy = (gridy[ey + 1] - gridy[ey]) / 2
        jac = jx * jy
        u = neural_net(torch.cat([x, y], 1), Ws, bs, "tanh")
        d1x, d2x = net_d_autograd([x, y], Ws, bs, "tanh", 0)
        d1y, d2y = net_d_autograd([x, y], Ws, bs, "tanh", 1)
        tx = Test_fcn(Ntx, xquad); d1tx, d2tx = dTest_fcn(Ntx, xquad)
        ty = Test_fcn(Nty, yquad); d1ty, d2ty = dTest_fcn(Nty, yquad)
        w0, w1 = wquad[:, 0:1], wquad[:, 1:2]
        if var_form == 0:
            integrand = d2x + d2y
            U = torch.stack([torch.stack([jac * torch.sum(_t(w0 * tx[r] * w1 * ty[k]) * integrand)
                                          for r in range(Ntx)]) for k in range(Nty)])
        elif var_form == 1:
            U1 = torch.stack([torch.stack([jac / jx * torch.sum(_t(w0 * d1tx[r] * w1 * ty[k]) * d1x)
                                           for r in range(Ntx)]) for k in range(Nty)])
            U2 = torch.stack([torch.stack([jac / jy * torch.sum(_t(w0 * tx[r] * w1 * d1ty[k]) * d1y)
                                           for r in range(Ntx)]) for k in range(Nty)])
            U = -U1 - U2
        else:
            U1 = torch.stack([torch.stack([jac * torch.sum(_t(w0 * d2tx[r] * w1 * ty[k]) * u)
                                           for r in range(Ntx)]) for k in range(Nty)])
            U2 = torch.stack([torch.stack([jac * torch.sum(_t(w0 * tx[r] * w1 * d2ty[k]) * u)
                                           for r in range(Ntx)]) for k in range(Nty)])
            U = U1 + U2
        Res = (U - F_el).reshape(1, -1)
        total = total + torch.mean(torch.square(Res))
        residuals.append(Res.detach().numpy().reshape(Nty, Ntx))
    return total, residuals


def _tables(N, X, W):
    T = Test_fcn(N, X)
    D1, D2 = dTest_fcn(N, X)
    return _t(T * W[None, :]), _t(D1 * W[None, :]), _t(D2 * W[None, :])


def varloss_2d_factorised(Ws, bs, X, WX, F_ext_total, gridx, gridy, Ntx, Nty, var_form=1, problem="poisson2d",
                          eps=None, V=1.0):
    """Sum-factorised form of P2D:68-120 (problem='poisson2d') and ADI:108-182 (problem='advdiff'; second
    coordinate = t, Res = U, no RHS).  Returns (lossv [torch scalar], residuals [NEx*NEy, Nty, Ntx] torch)."""
    Ws = [_t(W) for W in Ws]
    bs = [_t(b) for b in bs]
    Q = len(X)
    eps = _t(eps) if eps is not None else None
    A, B, C2 = _tables(max(Ntx, Nty), np.asarray(X), np.asarray(WX))
    Ax, Bx, Cx = A[:Ntx], B[:Ntx], C2[:Ntx]
    Ay, By, Cy = A[:Nty], B[:Nty], C2[:Nty]
    NEx, NEy = len(gridx) - 1, len(gridy) - 1
    res = []
    for ex in range(NEx):
        for ey in range(NEy):
            xe = gridx[ex] + (gridx[ex + 1] - gridx[ex]) / 2 * (np.asarray(X) + 1)
            ye = gridy[ey] + (gridy[ey + 1] - gridy[ey]) / 2 * (np.asarray(X) + 1)
            jx = (gridx[ex + 1] - gridx[ex]) / 2
            jy = (gridy[ey + 1] - gridy[ey]) / 2
            jac = jx * jy
            xx, yy = np.meshgrid(xe, ye)
            pts = np.hstack((xx.flatten()[:, None], yy.flatten()[:, None]))
            u, d1, d2 = mlp_forward_mode(pts, Ws, bs, "tanh")
            g = lambda f: f.reshape(Q, Q)
            if problem == "poisson2d":
                if var_form == 0:
                    U = jac * Ay @ g(d2[:, 0] + d2[:, 1]) @ Ax.T
                elif var_form == 1:
                    U = -(jac / jx) * Ay @ g(d1[:, 0]) @ Bx.T - (jac / jy) * By @ g(d1[:, 1]) @ Ax.T
                else:
                    U = jac * Ay @ g(u) @ Cx.T + jac * Cy @ g(u) @ Ax.T
                Res = U - _t(F_ext_total[ex, ey])
            else:
                if var_form == 0:
                    U = jac * Ay @ g(d1[:, 1] + V * d1[:, 0] - eps * d2[:, 0]) @ Ax.T
                else:
                    U = jac * Ay @ g(d1[:, 1] + V * d1[:, 0]) @ Ax.T + eps * (jac / jx) * Ay @ g(d1[:, 0]) @ Bx.T
                Res = U
            res.append(Res)
    res = torch.stack(res)
    return torch.sum(torch.mean(res.reshape(res.shape[0], -1) ** 2, dim=1)), res


# --------------------------------------------------------------------------------------------------
# AdvDiff-identification variational loss (ADI:108-182)
# --------------------------------------------------------------------------------------------------
def varloss_adi_literal(Ws, bs, eps, XT_quad, W_quad, grid_x, grid_t, N_testfcn, var_form=0, V=1.0,
                        elements=None):
    """Literal restatement of ADI:108-182 (the dead 1-D boundary evaluations ADI:134-155 are omitted: they
    never enter either var_form).  ``eps`` may be a torch tensor requiring grad."""
    Ws = [_t(W) for W in Ws]
    bs = [_t(b) for b in bs]
    eps = _t(eps)
    xquad, tquad, wquad = XT_quad[:, 0:1], XT_quad[:, 1:2], W_quad
    NEx, NEt = np.size(N_testfcn[0]), np.size(N_testfcn[1])
    total, residuals = 0, []
    pairs = elements if elements is not None else [(ex, et) for ex in range(NEx) for et in range(NEt)]
    for ex, et in pairs:
        Ntx, Ntt = N_testfcn[0][ex], N_testfcn[1][et]
        jac = (grid_t[et + 1] - grid_t[et]) / 2 * (grid_x[ex + 1] - grid_x[ex]) / 2
        jx = (grid_x[ex + 1] - grid_x[ex]) / 2
        x = _t(grid_x[ex] + (grid_x[ex + 1] - grid_x[ex]) / 2 * (xquad + 1)).requires_grad_(True)
        t = _t(grid_t[et] + (grid_t[et + 1] - grid_t[et]) / 2 * (tquad + 1)).requires_grad_(True)
        d1x, d2x = net_d_autograd([x, t], Ws, bs, "tanh", 0)
        d1t, _ = net_d_autograd([x, t], Ws, bs, "tanh", 1, second=False)
        tx = Test_fcn(Ntx, xquad); d1tx, _ = dTest_fcn(Ntx, xquad)
        tt = Test_fcn(Ntt, tquad)
        w0, w1 = wquad[:, 0:1], wquad[:, 1:2]
        if var_form == 0:
            U = torch.stack([torch.stack([
                jac * torch.sum(_t(w0 * tx[r] * w1 * tt[k]) * (d1t + V * d1x - eps * d2x))
                for r in range(Ntx)]) for k in range(Ntt)])
        else:
            U = torch.stack([torch.stack([
                jac * torch.sum(_t(w0 * tx[r] * w1 * tt[k]) * (d1t + V * d1x))
                + eps * jac / jx * torch.sum(_t(w0 * d1tx[r] * w1 * tt[k]) * d1x)
                for r in range(Ntx)]) for k in range(Ntt)])
        Res = U.reshape(1, -1)
        total = total + torch.mean(torch.square(Res))
        residuals.append(Res.detach().numpy().reshape(Ntt, Ntx))
    return total, residuals


# --------------------------------------------------------------------------------------------------
# 1-D Poisson variational loss (P1D:64-96)
# --------------------------------------------------------------------------------------------------
def varloss_1d_literal(Ws, bs, x_quad, w_quad, F_ext_total, grid, var_form=1):
    """Literal restatement of P1D:64-96.  x_quad, w_quad are (Q,1) columns; F_ext_total is (NE, N, 1)."""
    Ws = [_t(W) for W in Ws]
    bs = [_t(b) for b in bs]
    total, residuals = 0, []
    for e in range(np.shape(F_ext_total)[0]):
        F_el = _t(F_ext_total[e])
        Nt = np.shape(F_ext_total[e])[0]
        x = _t(grid[e] + (grid[e + 1] - grid[e]) / 2 * (x_quad + 1)).requires_grad_(True)
        xb = _t(np.array([[grid[e]], [grid[e + 1]]]))
        jac = (grid[e + 1] - grid[e]) / 2
        tq = Test_fcn(Nt, x_quad); d1q, d2q = dTest_fcn(Nt, x_quad)
        u = neural_net(x, Ws, bs, "sin")
        d1u, d2u = net_d_autograd([x], Ws, bs, "sin", 0)
        ub = neural_net(xb, Ws, bs, "sin")
        d1b, _ = dTest_fcn(Nt, np.array([[-1], [1]]))
        w = _t(w_quad)
        if var_form == 1:
            U = torch.stack([-jac * torch.sum(w * d2u * _t(tq[i])) for i in range(Nt)]).reshape(-1, 1)
        elif var_form == 2:
            U = torch.stack([torch.sum(w * d1u * _t(d1q[i])) for i in range(Nt)]).reshape(-1, 1)
        else:
            U = torch.stack([-1 / jac * torch.sum(w * u * _t(d2q[i]))
                             + 1 / jac * torch.sum(ub * _t(np.array([-d1b[i][0], d1b[i][-1]])))
                             for i in range(Nt)]).reshape(-1, 1)
        Res = U - F_el
        total = total + torch.mean(torch.square(Res))
        residuals.append(Res.detach().numpy().reshape(Nt))
    return total, residuals


def varloss_1d_factorised(Ws, bs, X, WX, F_ext_total, grid, var_form=1):
    """Factorised form of P1D:64-96: U = c * table @ field (+ boundary term for var_form 3).
    Returns (lossv, residuals [NE, N])."""
    Ws = [_t(W) for W in Ws]
    bs = [_t(b) for b in bs]
    X = np.asarray(X).ravel(); WX = np.asarray(WX).ravel()
    NE = np.shape(F_ext_total)[0]
    N = np.shape(F_ext_total)[1]
    A, B, C2 = _tables(N, X, WX)
    d1b, _ = dTest_fcn(N, np.array([-1.0, 1.0]))
    d1b = _t(d1b)
    res = []
    for e in range(NE):
        jac = (grid[e + 1] - grid[e]) / 2
        xe = grid[e] + (grid[e + 1] - grid[e]) / 2 * (X + 1)
        u, d1, d2 = mlp_forward_mode(xe[:, None], Ws, bs, "sin")
        if var_form == 1:
            U = -jac * A @ d2[:, 0]
        elif var_form == 2:
            U = B @ d1[:, 0]
        else:
            ub, _, _ = mlp_forward_mode(np.array([[grid[e]], [grid[e + 1]]]), Ws, bs, "sin")
            U = -1 / jac * C2 @ u + 1 / jac * (ub[1] * d1b[:, 1] - ub[0] * d1b[:, 0])
        res.append(U - _t(np.asarray(F_ext_total[e]).reshape(N)))
    res = torch.stack(res)
    return torch.sum(torch.mean(res ** 2, dim=1)), res


# --------------------------------------------------------------------------------------------------
# Boundary / data loss, PINN residual and the optimiser ("next" rows of SURVEY 8f)
# --------------------------------------------------------------------------------------------------
def lossb(Ws, bs, pts, u_train, act):
    """P1D:98, P2D:122, ADI:184 (without ADI's factor 10): mean((u_b - net_u(x_b))^2)."""
    u = neural_net(_t(pts), [_t(W) for W in Ws], [_t(b) for b in bs], act)
    return torch.mean(torch.square(_t(u_train) - u))


def net_f(Ws, bs, pts, act, problem, eps=None, V=1.0):
    """Strong-form residual field: P1D:150-155 (-u_xx), P2D:187-194 (u_xx+u_yy), ADI:247-253
    (u_t + V u_x - eps u_xx)."""
    u, d1, d2 = mlp_forward_mode(pts, Ws, bs, act)
    if problem == "poisson1d":
        return -d2[:, 0]
    if problem == "poisson2d":
        return d2[:, 0] + d2[:, 1]
    return d1[:, 1] + V * d1[:, 0] - _t(eps) * d2[:, 0]


def adam_tf1_step(theta, grad, m, v, t, lr=0.001, b1=0.9, b2=0.999, eps_hat=1e-8):
    """One ``tf.train.AdamOptimizer`` update (P2D:131-132): returns (theta, m, v) after step number t>=1."""
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return theta - lr_t * m / (np.sqrt(v) + eps_hat), m, v


def loss_and_grad(fn, Ws, bs, extra=None):
    """Evaluate ``fn(Ws, bs[, extra])`` -> torch scalar with autograd; return (loss, flat d(theta)[, d(extra)])."""
    Wt = [_t(W).clone().requires_grad_(True) for W in Ws]
    bt = [_t(b).clone().requires_grad_(True) for b in bs]
    params = [p for pair in zip(Wt, bt) for p in pair]
    if extra is not None:
        ex = _t(extra).clone().requires_grad_(True)
        loss = fn(Wt, bt, ex)
        g = torch.autograd.grad(loss, params + [ex], allow_unused=True)
        flat = np.concatenate([(gi if gi is not None else torch.zeros_like(p)).detach().numpy().ravel()
                               for gi, p in zip(g[:-1], params)])
        return float(loss.detach()), flat, (g[-1].detach().numpy() if g[-1] is not None else None)
    loss = fn(Wt, bt)
    g = torch.autograd.grad(loss, params, allow_unused=True)
    flat = np.concatenate([(gi if gi is not None else torch.zeros_like(p)).detach().numpy().ravel()
                           for gi, p in zip(g, params)])
    return float(loss.detach()), flat
