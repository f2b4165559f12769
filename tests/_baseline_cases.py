"""The BASELINE.json configurations that are parity cases rather than bench lines, at their full sizes, built from the
oracle (float64): C1 (1-D Poisson, 4 elements, Q=50, 5 test functions, MLP [1,5,5,1] -- the reference's own CPU run),
C2 (1-D Poisson, 16 elements, Q=80, 60 test functions, [1,20,20,20,1]) and C5 (AdvDiff identification, 20 elements in
x over one slab in t, Q=80x80, 60x60 test functions, [2,20,20,20,1] + eps), each with its default var_form."""
import numpy as np

from oracle import hpvpinn_oracle as O


def build(name, seed=3):
    rng = np.random.default_rng(seed)
    if name in ("c1", "c2"):
        n_el, Q, N, layers = (4, 50, 5, [1, 5, 5, 1]) if name == "c1" else (16, 80, 60, [1, 20, 20, 20, 1])
        X, W = O.GaussLobattoJacobiWeights(Q, 0, 0)
        D1, D2 = O.dTest_fcn(N, X)
        Ws, bs = O.xavier_params(layers, 1234)
        bs = [0.05 * rng.standard_normal(b.shape) for b in bs]
        g = np.linspace(-1, 1, n_el + 1)
        F = np.asarray(O.rhs_1d(g, n_el * [N], X, W)).reshape(n_el, N)
        fn = lambda Wt, bt: O.varloss_1d_factorised(Wt, bt, X, W, F, g, 1)
        l_ref, g_ref = O.loss_and_grad(lambda Wt, bt: fn(Wt, bt)[0], Ws, bs)
        inp = dict(problem="poisson1d", var_form=1, layers=layers, act="sin", xi=X, w=W, T=O.Test_fcn(N, X), D1=D1, D2=D2,
                   d1b=O.dTest_fcn(N, np.array([-1.0, 1.0]))[0], lo=g[:-1, None], hi=g[1:, None], ntx=N, nty=1,
                   F=F.reshape(n_el, 1, N), theta=O.pack_theta(Ws, bs))
        return inp, float(l_ref), g_ref, fn(Ws, bs)[1].detach().numpy(), None
    if name == "c5":
        Q, N, layers, n_el = 80, 60, [2, 20, 20, 20, 1], 20
        X, W = O.GaussLobattoJacobiWeights(Q, 0, 0)
        D1, D2 = O.dTest_fcn(N, X)
        Ws, bs = O.xavier_params(layers, 1234)
        bs = [0.05 * rng.standard_normal(b.shape) for b in bs]
        gx, gt = np.linspace(-1, 1, n_el + 1), np.array([0.0, 1.0])
        eps0, V = 1.0, 1.0
        fn = lambda Wt, bt, e: O.varloss_2d_factorised(Wt, bt, X, W, None, gx, gt, N, N, 0, problem="advdiff", eps=e, V=V)
        l_ref, g_ref, ge_ref = O.loss_and_grad(lambda Wt, bt, e: fn(Wt, bt, e)[0], Ws, bs, extra=np.array([eps0]))
        lo = np.array([[gx[i], gt[0]] for i in range(n_el)])
        hi = np.array([[gx[i + 1], gt[1]] for i in range(n_el)])
        inp = dict(problem="advdiff", var_form=0, layers=layers, act="tanh", xi=X, w=W, T=O.Test_fcn(N, X), D1=D1, D2=D2, d1b=None,
                   lo=lo, hi=hi, ntx=N, nty=N, F=None, theta=O.pack_theta(Ws, bs), eps=eps0, V=V)
        return inp, float(l_ref), g_ref, fn(Ws, bs, np.array([eps0]))[1].detach().numpy(), float(ge_ref[0])
    raise KeyError(name)
