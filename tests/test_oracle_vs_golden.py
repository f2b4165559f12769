"""CPU: pins the float64 oracle against fixtures produced by the reference's own VPINN classes
(tests/golden/make_golden.py) and against the reference quadrature module's tables."""
import numpy as np
import pytest
import torch

from oracle import hpvpinn_oracle as O
from tests import _cases as C

RTOL = 1e-11


@pytest.mark.parametrize("name", C.case_names())
def test_factorised_lossv_and_grad_match_reference_class(name):
    c = C.load(name)
    out = C.oracle_lossv(c)
    assert out[0] == pytest.approx(float(c["lossv"]), rel=RTOL)
    g = c["grad_lossv"]
    assert np.abs(out[2] - g).max() <= 1e-9 * max(1.0, np.abs(g).max())
    if c["kind"] == "advdiff":
        assert out[3][0] == pytest.approx(float(c["grad_lossv_eps"][0]), rel=1e-9)


@pytest.mark.parametrize("name", C.case_names())
def test_literal_restatement_matches_reference_class(name):
    c = C.load(name)
    Ws, bs = O.unpack_theta(c["theta"], c["layers"])
    vf = int(c["var_form"])
    if c["kind"] == "poisson2d":
        N = [(len(c["gridx"]) - 1) * [int(c["Ntx"])], (len(c["gridy"]) - 1) * [int(c["Nty"])]]
        l, res = O.varloss_2d_literal(Ws, bs, c["X_quad"], c["W_quad"], c["F_ext"], c["gridx"], c["gridy"], N, vf)
    elif c["kind"] == "advdiff":
        N = [(len(c["grid_x"]) - 1) * [int(c["Ntx"])], (len(c["grid_t"]) - 1) * [int(c["Ntt"])]]
        l, res = O.varloss_adi_literal(Ws, bs, float(c["eps0"]), c["XT_quad"], c["W_quad"], c["grid_x"], c["grid_t"], N, vf, float(c["V"]))
    else:
        l, res = O.varloss_1d_literal(Ws, bs, c["X_quad"], c["W_quad"], c["F_ext"], c["grid"], vf)
    assert float(l.detach()) == pytest.approx(float(c["lossv"]), rel=RTOL)
    # literal and factorised residual vectors agree entry by entry
    fac = C.oracle_lossv(c)[1]
    lit = np.stack(res).reshape(fac.shape)
    assert np.abs(fac - lit).max() <= 1e-10 * max(1.0, np.abs(lit).max())


@pytest.mark.parametrize("name", C.case_names())
def test_net_u_derivatives_lossb_and_adam(name):
    c = C.load(name)
    Ws, bs = O.unpack_theta(c["theta"], c["layers"])
    act = C.ACT[c["kind"]]
    xt = c["XT_test"] if c["kind"] == "advdiff" else c["X_test"]
    u, d1, d2 = O.mlp_forward_mode(xt, Ws, bs, act)
    assert np.allclose(u.numpy(), c["u_test_pred"][:, 0], rtol=1e-12, atol=1e-13)
    if c["kind"] == "poisson2d":
        assert np.allclose(d1[:, 0].numpy(), c["d1x"][:, 0], rtol=1e-11, atol=1e-12)
        assert np.allclose(d2[:, 0].numpy(), c["d2x"][:, 0], rtol=1e-11, atol=1e-12)
        assert np.allclose(d1[:, 1].numpy(), c["d1y"][:, 0], rtol=1e-11, atol=1e-12)
        assert np.allclose(d2[:, 1].numpy(), c["d2y"][:, 0], rtol=1e-11, atol=1e-12)
        f = O.net_f(Ws, bs, c["X_f_train"], act, "poisson2d")
        assert np.allclose(f.numpy(), c["f_pred"][:, 0], rtol=1e-11, atol=1e-11)
        lb = O.lossb(Ws, bs, c["X_u_train"], c["u_train"], act)
        assert float(lb) == pytest.approx(float(c["lossb"]), rel=1e-12)
    elif c["kind"] == "poisson1d":
        assert np.allclose(d1[:, 0].numpy(), c["d1"][:, 0], rtol=1e-11, atol=1e-12)
        assert np.allclose(d2[:, 0].numpy(), c["d2"][:, 0], rtol=1e-11, atol=1e-11)
        lb = O.lossb(Ws, bs, c["X_u_train"], c["u_train"], act)
        assert float(lb) == pytest.approx(float(c["lossb"]), rel=1e-12)
    else:
        f = O.net_f(Ws, bs, c["XT_f_train"], act, "advdiff", eps=float(c["eps0"]), V=float(c["V"]))
        assert np.allclose(f.detach().numpy(), c["f_pred"][:, 0], rtol=1e-11, atol=1e-11)
        lb = 10 * O.lossb(Ws, bs, c["XT_u_train"], c["u_train"], act)     # ADI:184 folds the 10 in
        assert float(lb) == pytest.approx(float(c["lossb"]), rel=1e-12)


def _total_loss_fn(c):
    """Differentiable float64 total loss (lossv + weighted lossb | lossp) as the reference assembles it
    (P1D:98-100, P2D:122-129, ADI:184-187)."""
    act = C.ACT[c["kind"]]
    vf = int(c["var_form"])
    if c["kind"] == "poisson2d":
        X = C.nodes_from_flat(c["X_quad"]); WX = c["W_quad"][:len(X), 0]

        def fn(W, b):
            lb = O.lossb(W, b, c["X_u_train"], c["u_train"], act)
            if str(c["scheme"]) == "PINNs":
                f = O.net_f(W, b, c["X_f_train"], act, "poisson2d")
                return 10 * lb + torch.mean((f[:, None] - torch.as_tensor(c["f_train"])) ** 2)
            lv = O.varloss_2d_factorised(W, b, X, WX, c["F_ext"], c["gridx"], c["gridy"], int(c["Ntx"]), int(c["Nty"]), vf)[0]
            return 10 * lb + lv
        return fn
    if c["kind"] == "poisson1d":
        def fn(W, b):
            lv = O.varloss_1d_factorised(W, b, c["X_quad"], c["W_quad"], c["F_ext"], c["grid"], vf)[0]
            return float(c["lossb_weight"]) * O.lossb(W, b, c["X_u_train"], c["u_train"], act) + lv
        return fn

    def fn(W, b, eps):
        lv = O.varloss_2d_factorised(W, b, c["T_quad"], c["WT_quad"], None, c["grid_x"], c["grid_t"], int(c["Ntx"]), int(c["Ntt"]),
                                     vf, problem="advdiff", eps=eps, V=float(c["V"]))[0]
        return 10 * O.lossb(W, b, c["XT_u_train"], c["u_train"], act) + lv
    return fn


@pytest.mark.parametrize("name", C.case_names())
def test_total_loss_grad_and_tf1_adam_trajectory(name):
    c = C.load(name)
    fn = _total_loss_fn(c)
    theta = c["theta"].copy()
    adi = c["kind"] == "advdiff"
    eps = np.array([float(c["eps0"])]) if adi else None
    Ws, bs = O.unpack_theta(theta, c["layers"])
    if adi:
        l, g, ge = O.loss_and_grad(fn, Ws, bs, extra=eps)
        assert ge[0] == pytest.approx(float(c["grad_loss_eps"][0]), rel=1e-9)
    else:
        l, g = O.loss_and_grad(fn, Ws, bs)
    assert l == pytest.approx(float(c["loss"]), rel=RTOL)
    assert np.abs(g - c["grad_loss"]).max() <= 1e-9 * max(1.0, np.abs(c["grad_loss"]).max())
    # replay the Adam steps the reference train() took
    if c["kind"] == "poisson2d":
        nsteps = len(c["adam_loss_his"])
    else:
        nsteps = 21
    lr = 0.001 if c["kind"] == "poisson2d" else float(c["LR"])
    m = np.zeros_like(theta); v = np.zeros_like(theta)
    me = np.zeros(1); ve = np.zeros(1)
    his = []
    for t in range(1, nsteps + 1):
        Ws, bs = O.unpack_theta(theta, c["layers"])
        if adi:
            _, g, ge = O.loss_and_grad(fn, Ws, bs, extra=eps)
            eps, me, ve = O.adam_tf1_step(eps, ge, me, ve, t, lr)
        else:
            _, g = O.loss_and_grad(fn, Ws, bs)
        theta, m, v = O.adam_tf1_step(theta, g, m, v, t, lr)
        Ws, bs = O.unpack_theta(theta, c["layers"])
        his.append(float(fn(Ws, bs, eps)) if adi else float(fn(Ws, bs)))
    assert np.abs(theta - c["adam_theta"]).max() <= 1e-10
    if c["kind"] == "poisson2d":
        assert np.allclose(his, c["adam_loss_his"], rtol=1e-9)
    elif c["kind"] == "poisson1d":
        assert np.allclose(his[0::10], c["adam_total_record"][:, 1], rtol=1e-9)
    else:
        assert np.allclose(his[0::10], c["adam_total_records"][:, 1], rtol=1e-9)
        assert eps[0] == pytest.approx(float(c["adam_eps"][0]), rel=1e-10)


def test_quadrature_and_tables_match_reference_module():
    t = dict(np.load(C.GOLDEN + "/tables.npz"))
    for Q in (5, 10, 50, 80):
        x, w = O.GaussLobattoJacobiWeights(Q, 0, 0)
        assert np.array_equal(x, t["gll_x_%d" % Q]) and np.array_equal(w, t["gll_w_%d" % Q])
        assert abs(w.sum() - 2.0) < 1e-13 and x[0] == -1 and x[-1] == 1
    x80 = t["gll_x_80"]
    assert np.array_equal(O.Test_fcn(60, x80), t["T_60_80"])
    d1, d2 = O.dTest_fcn(60, x80)
    assert np.array_equal(d1, t["D1_60_80"]) and np.array_equal(d2, t["D2_60_80"])
    # known answers: test functions vanish at the end points; dTest_fcn == DJacobi differences
    assert np.abs(t["T_60_80"][:, [0, -1]]).max() < 1e-12
    for n in (2, 3, 17, 60):
        ref1 = O.DJacobi(n + 1, 0, 0, x80, 1) - O.DJacobi(n - 1, 0, 0, x80, 1)
        assert np.allclose(d1[n - 1], ref1, rtol=1e-11, atol=1e-9)


def test_rhs_assembly_matches_reference_driver():
    d = dict(np.load(C.GOLDEN + "/driver_p2d.npz"))
    F = O.rhs_2d_literal(d["grid_x"], d["grid_y"], list(d["N_test_x"]), list(d["N_test_y"]), d["XY_quad"], d["WXY_quad"])
    assert np.array_equal(F, d["F_ext_total"])
    X = C.nodes_from_flat(d["XY_quad"]); WX = d["WXY_quad"][:len(X), 0]
    F2 = O.rhs_2d_factorised(d["grid_x"], d["grid_y"], 5, 5, X, WX)
    assert np.abs(F2 - F).max() < 1e-13 * np.abs(F).max() * 10
    # quadrature layout: p = j*Q + i -> (X[i], Y[j])
    Q = len(X)
    assert np.array_equal(d["XY_quad"][:, 0].reshape(Q, Q)[3], X) and np.array_equal(d["XY_quad"][:, 1].reshape(Q, Q)[:, 3], X)
    for tag in ("driver_p1d", "driver_p1d_3el"):
        d = dict(np.load(C.GOLDEN + "/%s.npz" % tag))
        NE = len(d["grid"]) - 1
        F = O.rhs_1d(d["grid"], NE * [int(d["N_testfcn"])], d["x_quad"][:, 0], d["w_quad"][:, 0])
        assert np.array_equal(F, d["F_ext_total"])


def test_exact_solution_substitution_known_answer():
    """with the exact u_xx+u_yy in place of the network, var_form 0 gives U == F_ext to rounding."""
    d = dict(np.load(C.GOLDEN + "/driver_p2d.npz"))
    X = C.nodes_from_flat(d["XY_quad"]); WX = d["WXY_quad"][:len(X), 0]
    A = O.Test_fcn(5, X) * WX
    ex, ey = 1, 2
    gx, gy = d["grid_x"], d["grid_y"]
    xe = gx[ex] + (gx[ex + 1] - gx[ex]) / 2 * (X + 1)
    ye = gy[ey] + (gy[ey + 1] - gy[ey]) / 2 * (X + 1)
    J = (gx[ex + 1] - gx[ex]) / 2 * (gy[ey + 1] - gy[ey]) / 2
    U = J * A @ O.f_ext_2d(xe[None, :], ye[:, None]) @ A.T
    assert np.abs(U - d["F_ext_total"][ex, ey]).max() < 1e-12
