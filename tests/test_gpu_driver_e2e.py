"""GPU (-m gpu): end to end with the reference DRIVERS' own data.  The fixtures tests/golden/driver_*.npz were
produced by running the unmodified `__main__` set-up blocks of the three reference scripts (grids, quadrature
layout, boundary points, RHS assembly F_ext_total: P1D:231-329, P2D:279-426, ADI:351-483) in the build container;
here the drop-in VPINN classes are constructed exactly as those scripts construct theirs (P2D:430-431, P1D:333-334,
ADI:488-489) and trained: the loss must fall and the prediction must approach the manufactured solution."""
import os

import numpy as np
import pytest

from oracle import hpvpinn_oracle as O
from tests import _cases as C

pytestmark = pytest.mark.gpu


def _fx(name):
    return dict(np.load(os.path.join(C.GOLDEN, name + ".npz")))


def test_poisson2d_driver_defaults_train(capsys):
    from hpv_b200.poisson2d import VPINN
    d = _fx("driver_p2d")
    rng = np.random.default_rng(0)
    X_f = 2 * rng.random((100, 2)) - 1
    f = O.f_ext_2d(X_f[:, 0:1], X_f[:, 1:2])
    xs = np.linspace(-1, 1, 41)
    X_test = np.array([[a, b] for a in xs for b in xs])
    u_test = O.u_ext_2d(X_test[:, 0:1], X_test[:, 1:2])
    N_testfcn = [list(d["N_test_x"]), list(d["N_test_y"])]
    # the reference class reads var_form / scheme / loss_his from the script's globals: same here
    global var_form, scheme, loss_his
    var_form, scheme, loss_his = 1, "VPINNs", []
    model = VPINN(d["X_u_train"], d["u_train"], X_f, f, d["XY_quad"], d["WXY_quad"], None, d["F_ext_total"],
                  d["grid_x"], d["grid_y"], N_testfcn, X_test, u_test, [int(v) for v in d["layers"]])
    assert model.var_form == 1 and model.scheme == "VPINNs" and model.loss_his is loss_his
    # RHS produced by the reference driver == the engine-side assembly of the same quantity (sanity of layout)
    X, W = O.GaussLobattoJacobiWeights(10, 0, 0)
    assert np.allclose(O.rhs_2d_factorised(d["grid_x"], d["grid_y"], 5, 5, X, W), d["F_ext_total"], rtol=1e-10, atol=1e-12)
    # the same optimisation in float64 on the CPU (factorised oracle + TF1-Adam restatement), same initial weights
    import torch
    theta = np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in zip(model.weights, model.biases)])
    layers = [int(v) for v in d["layers"]]
    Xb, ub = d["X_u_train"], d["u_train"]

    def total(Wt, bt):
        lv = O.varloss_2d_factorised(Wt, bt, X, W, d["F_ext_total"], d["grid_x"], d["grid_y"], 5, 5, 1)[0]
        return lv + 10 * O.lossb(Wt, bt, Xb, ub, "tanh")

    n_it = 150
    th, m_, v_ = theta.copy(), np.zeros_like(theta), np.zeros_like(theta)
    ref_after = []
    for t in range(1, n_it + 1):
        Ws, bs = O.unpack_theta(th, layers)
        _, g = O.loss_and_grad(total, Ws, bs)
        th, m_, v_ = O.adam_tf1_step(th, g, m_, v_, t)
        Ws, bs = O.unpack_theta(th, layers)
        ref_after.append(float(total([torch.as_tensor(w_) for w_ in Ws], [torch.as_tensor(b_) for b_ in bs])))
    model.train(n_it)
    out = capsys.readouterr().out
    assert "It: 0, Loss:" in out and "It: 100, Loss:" in out
    assert len(loss_his) == n_it
    assert np.allclose(loss_his, ref_after, rtol=2e-3)                 # 150 fp32 GPU steps track the float64 path
    assert loss_his[-1] < loss_his[0]
    theta_gpu = np.concatenate([np.concatenate([W_.ravel(), b_.ravel()]) for W_, b_ in zip(model.weights, model.biases)])
    assert np.abs(theta_gpu - th).max() < 2e-3
    u_ref = O.neural_net(torch.as_tensor(X_test), *[[torch.as_tensor(t_) for t_ in part] for part in O.unpack_theta(th, layers)], "tanh")
    assert np.abs(model.predict() - u_ref.numpy()).max() < 5e-3
    model.sess.close()


@pytest.mark.parametrize("fx", ["driver_p1d", "driver_p1d_3el"])
def test_poisson1d_driver_defaults_train(fx):
    from hpv_b200.poisson1d import VPINN
    d = _fx(fx)
    Xb = np.array([[-1.0], [1.0]])
    ub = O.u_ext_1d(Xb)
    Xt = np.linspace(-1, 1, 2001)[:, None]
    ut = O.u_ext_1d(Xt)
    rng = np.random.default_rng(0)
    Xf = 2 * rng.random((500, 1)) - 1
    rec = []
    model = VPINN(Xb, ub, d["x_quad"], d["w_quad"], d["F_ext_total"], d["grid"], Xt, ut, [int(v) for v in d["layers"]],
                  Xf, O.f_ext_1d(Xf), var_form=1, lossb_weight=1, LR=0.001, total_record=rec)
    model.train(300, 2e-32)
    losses = np.array(rec)[:, 1]
    assert len(rec) == 30 and np.all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert model.predict(Xt).shape == ut.shape
    model.sess.close()


def test_advdiff_driver_defaults_train():
    from hpv_b200.advdiff import VPINN
    d = _fx("driver_adi")
    rng = np.random.default_rng(0)
    XT_f = np.hstack((2 * rng.random((500, 1)) - 1, rng.random((500, 1))))
    XT_test = np.hstack((2 * rng.random((256, 1)) - 1, rng.random((256, 1))))
    u_test = np.zeros((256, 1))
    N_testfcn = [[5], [5]]
    model = VPINN(d["XT_u_train"], d["u_train"], XT_f, d["XT_quad"], d["W_quad"], d["T_quad"], d["WT_quad"], d["grid_x"],
                  d["grid_t"], N_testfcn, XT_test, u_test, [int(v) for v in d["layers"]], XT_test.min(0), XT_test.max(0),
                  var_form=0, V=float(d["V"]), LR=0.001)
    err, total, u_rec, u_his, t_train = model.train(400, 2e-32)
    total = np.array([[r[0], r[1], r[2]] for r in total])
    assert total.shape[0] == 40 and total[-1, 1] < total[0, 1]
    assert total[-1, 2] < 1.0                       # the diffusivity moves away from its initial value 1 (ADI:63)
    assert u_rec is not None and u_rec.shape == u_test.shape
    model.sess.close()
