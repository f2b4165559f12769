"""CPU: why the gradient tolerance of the 1-D forms that integrate by parts twice is loose (tests/_gpu.py: GRAD_RTOL).
The round-1 review asked whether a compensated / float64 accumulation of the projection (the var_form-3 boundary terms
in particular) would tighten it.  This evaluates the same formulas three ways on the golden cases -- everything in
float64, everything in float32, and a float32 network with the whole projection, residual and loss in float64 -- and
shows that the error belongs to the float32 NETWORK VALUES, which the form multiplies by table entries of size ~N^4
before they cancel: float64 accumulation after a float32 network does not remove it."""
import numpy as np
import pytest
import torch

from oracle import hpvpinn_oracle as O
from tests import _cases as C


def _grad_1d(name, mlp_dt, proj_dt):
    c = C.load(name)
    layers = [int(v) for v in c["layers"]]
    Ws0, bs0 = O.unpack_theta(c["theta"], layers)
    Ws = [torch.tensor(W, dtype=mlp_dt, requires_grad=True) for W in Ws0]
    bs = [torch.tensor(b, dtype=mlp_dt, requires_grad=True) for b in bs0]
    X, WX, grid = np.asarray(c["X_quad"]).ravel(), np.asarray(c["W_quad"]).ravel(), np.asarray(c["gridx"] if "gridx" in c else c["grid"]).ravel()
    F = np.asarray(c["F_ext"])
    NE, N = F.shape[0], F.shape[1]
    A, B, C2 = O._tables(N, X, WX)
    d1b, _ = O.dTest_fcn(N, np.array([-1.0, 1.0]))
    A, B, C2, d1b = [torch.tensor(np.asarray(t), dtype=proj_dt) for t in (A, B, C2, d1b)]
    vf = int(c["var_form"])

    def mlp(x):                                         # forward mode, sin (P1D:128-148), in the network's dtype
        h = torch.tensor(x, dtype=mlp_dt)
        dh, ddh = torch.ones_like(h), torch.zeros_like(h)
        for l, (W, b) in enumerate(zip(Ws, bs)):
            z, dz, ddz = h @ W + b, dh @ W, ddh @ W
            if l == len(Ws) - 1:
                return z[:, 0], dz[:, 0], ddz[:, 0]
            a, s1 = torch.sin(z), torch.cos(z)
            h, dh, ddh = a, s1 * dz, -a * dz * dz + s1 * ddz

    tot = 0
    for e in range(NE):
        jac = (grid[e + 1] - grid[e]) / 2
        u, d1, d2 = [t.to(proj_dt) for t in mlp((grid[e] + jac * (X + 1))[:, None])]
        if vf == 1:
            U = -jac * A @ d2
        elif vf == 2:
            U = B @ d1
        else:
            ub = mlp(np.array([[grid[e]], [grid[e + 1]]]))[0].to(proj_dt)
            U = -1 / jac * C2 @ u + 1 / jac * (ub[1] * d1b[:, 1] - ub[0] * d1b[:, 0])
        tot = tot + torch.mean((U - torch.tensor(F[e].reshape(N), dtype=proj_dt)) ** 2)
    g = torch.autograd.grad(tot, Ws + bs, allow_unused=True)
    g = [gi if gi is not None else torch.zeros_like(p) for gi, p in zip(g, Ws + bs)]
    flat = np.concatenate([np.concatenate([g[l].double().numpy().ravel(), g[len(Ws) + l].double().numpy().ravel()]) for l in range(len(Ws))])
    gref = np.asarray(c["grad_lossv"]).ravel()
    return np.abs(flat - gref).max() / np.abs(gref).max()


@pytest.mark.parametrize("name", ["p1d_vf1", "p1d_vf2", "p1d_vf3"])
def test_float64_projection_after_a_float32_network_keeps_the_gradient_error(name):
    e64 = _grad_1d(name, torch.float64, torch.float64)
    e32 = _grad_1d(name, torch.float32, torch.float32)
    emix = _grad_1d(name, torch.float32, torch.float64)
    assert e64 < 1e-9                                             # the restatement itself matches the fixture
    assert emix > 0.3 * e32                                       # float64 accumulation does not buy an order of magnitude
    # measured: vf1 1.5e-7 / 2.2e-7, vf2 1.4e-5 / 2.3e-5, vf3 5.6e-3 / 4.3e-3 (float32 / mixed): the tolerances of
    # tests/_gpu.py (1e-4, 5e-4, 3e-2) sit above what float32 network values allow for each form
    assert e32 < {"p1d_vf1": 1e-5, "p1d_vf2": 5e-4, "p1d_vf3": 3e-2}[name]
