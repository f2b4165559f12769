"""Generator of tests/golden/c3_converged.npz -- the NEAR-CONVERGED parameter set of SURVEY 8(d) / BASELINE.md:
a theta at which the C3 problem's variational residual has largely cancelled (U ~ F), the regime where fp32
(and any split-precision tensor-core product) is most exposed.

The network [2,20,20,20,1] is trained in float64 on the CPU (oracle/, factorised form + autograd) on the C3
problem itself -- 8x8 elements on [-1,1]^2, manufactured solution of P2D:300-310, loss = lossv + 10*lossb
(P2D:125-128) -- with a reduced rule (Q=24, N=12 per direction) and L-BFGS so that it converges in minutes,
then evaluated ONCE at the full C3 size (Q=80, N=60) with the same oracle: lossv, element losses, a slab of the
residuals and d lossv / d theta at that size are stored as the fixture the GPU test compares with.

    python tests/golden/make_converged_theta.py          (about 45 minutes on 8 cores; the output is committed)

Result of the committed run: training loss 5.49 -> 1.84e-2 in 6431 evaluations; C3 lossv 9.05e-2 -> 7.08e-4 (ratio 7.8e-3),
max|Res| = 0.30 against max|F| = 3.34.  L-BFGS stalls there on the steep tanh(10 x) component of the manufactured
solution; the deeper cancellation regime is covered synthetically (tests/test_gpu_parity_regimes.py,
test_c3_deep_cancellation_reaches_the_fp32_floor).
"""
import os
import sys
import time

import numpy as np
import torch
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hpvpinn_oracle as O  # noqa: E402

LAYERS = [2, 20, 20, 20, 1]
NE = 8


def main():
    torch.set_num_threads(8)
    g = np.linspace(-1, 1, NE + 1)
    Qt, Nt = 24, 12
    Xt, Wt = O.GaussLobattoJacobiWeights(Qt, 0, 0)
    Ft = O.rhs_2d_factorised(g, g, Nt, Nt, Xt, Wt)
    s = np.linspace(-1, 1, 81)[:, None]
    one = np.ones_like(s)
    Xb = np.vstack([np.hstack([s, -one]), np.hstack([s, one]), np.hstack([-one, s]), np.hstack([one, s])])
    ub = O.u_ext_2d(Xb[:, 0:1], Xb[:, 1:2])
    Ws, bs = O.xavier_params(LAYERS, 1234)
    theta0 = O.pack_theta(Ws, bs)

    def total(Wl, bl):
        return O.varloss_2d_factorised(Wl, bl, Xt, Wt, Ft, g, g, Nt, Nt, 1)[0] + 10 * O.lossb(Wl, bl, Xb, ub, "tanh")

    n_eval = [0]

    def fun(th):
        W_, b_ = O.unpack_theta(th, LAYERS)
        l, gr = O.loss_and_grad(total, W_, b_)
        n_eval[0] += 1
        if n_eval[0] % 200 == 0:
            print("  eval %5d  loss %.4e" % (n_eval[0], l), flush=True)
        return l, gr

    t0 = time.time()
    r = minimize(fun, theta0, jac=True, method="L-BFGS-B", options=dict(maxiter=6000, maxfun=8000, ftol=1e-16, gtol=1e-12, maxcor=50))
    theta = r.x
    print("trained: loss %.4e -> %.4e in %d evaluations, %.0f s" % (fun(theta0)[0], r.fun, n_eval[0], time.time() - t0), flush=True)

    # full C3 size, once
    Q, N = 80, 60
    X, W = O.GaussLobattoJacobiWeights(Q, 0, 0)
    F = O.rhs_2d_factorised(g, g, N, N, X, W)

    def lossv_full(th):
        W_, b_ = O.unpack_theta(th, LAYERS)
        l, gr = O.loss_and_grad(lambda a, b: O.varloss_2d_factorised(a, b, X, W, F, g, g, N, N, 1)[0], W_, b_)
        res = O.varloss_2d_factorised(W_, b_, X, W, F, g, g, N, N, 1)[1].numpy().reshape(NE * NE, N, N)
        return l, gr, res

    l0, _, _ = lossv_full(theta0)
    l1, g1, res1 = lossv_full(theta)
    el = (res1 ** 2).mean(axis=(1, 2))
    print("C3 lossv: initial %.6e, near-converged %.6e (ratio %.3e)" % (l0, l1, l1 / l0))
    Fnorm = float(np.abs(F).max())
    print("max|Res| %.3e vs max|F| %.3e" % (np.abs(res1).max(), Fnorm))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c3_converged.npz")
    np.savez_compressed(out, theta=theta, layers=np.array(LAYERS), lossv=l1, lossv_initial=l0, grad=g1, el_loss=el,
                        res_elements=np.array([0, 27, 63]), res=res1[[0, 27, 63]].astype(np.float64), F_max=Fnorm,
                        train_rule=np.array([Qt, Nt]), ne=NE, Q=Q, N=N)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
