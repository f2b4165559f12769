"""Generates tests/golden/*.npz by running the reference's OWN code in this container.

Run:  python tests/golden/make_golden.py        (needs /root/reference; not available on the GPU box)

What is executed, unmodified, from /root/reference:
  * the three ``VPINN`` classes (graph construction loops P1D:64-96, P2D:68-120, ADI:108-182, ``net_u``,
    ``net_d*``, ``Test_fcn*``, ``dTest_fcn``, ``train``) -- on the TF-1 API stand-in ``oracle/tf1_shim``;
  * ``GaussJacobiQuadRule_V3`` (real numpy/scipy code);
  * the driver blocks up to the model construction (problem set-up and RHS assembly).
What is injected: explicit network parameters (numpy ``default_rng`` Xavier draw -- TensorFlow's RNG stream is
not reproducible), small grids / test-function counts so the literal graphs evaluate in seconds, and the module
globals the classes read (var_form, scheme, LR, lossb_weight, V).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_loader as RL          # noqa: E402
from oracle import hpvpinn_oracle as O             # noqa: E402


def flat(arrs_w, arrs_b):
    return np.concatenate([np.concatenate([np.asarray(w).ravel(), np.asarray(b).ravel()]) for w, b in zip(arrs_w, arrs_b)])


def set_params(model, Ws, bs):
    for v, w in zip(model.weights, Ws):
        v.load(w)
    for v, b in zip(model.biases, bs):
        v.load(b)


def get_params(model):
    return flat([v.value.detach().numpy() for v in model.weights], [v.value.detach().numpy() for v in model.biases])


def grads(tf, model, loss, feed, extra=()):
    vs = [p for pair in zip(model.weights, model.biases) for p in pair] + list(extra)
    g = model.sess.run(tf.gradients(loss, vs), feed)
    n = len(vs) - len(extra)
    gt = np.concatenate([np.zeros(v.value.numel()) if gi is None else np.asarray(gi).ravel() for gi, v in zip(g[:n], vs[:n])])
    return gt, [np.asarray(gi).ravel() for gi in g[n:]]


def perturbed(layers, seed):
    """Xavier weights plus small non-zero biases (so that bias gradients/paths are exercised)."""
    Ws, bs = O.xavier_params(layers, seed)
    rng = np.random.default_rng(seed + 7)
    bs = [0.1 * rng.standard_normal(b.shape) for b in bs]
    return Ws, bs


# ---------------------------------------------------------------------------------------------------------
def case_p2d(name, layers, gridx, gridy, Q, Ntx, Nty, var_form, seed, scheme="VPINNs", adam_steps=4):
    tf = RL.tf_shim()
    tf.reset_default_graph()
    np.random.seed(seed)
    mod = RL.load("P2D", var_form=var_form, scheme=scheme, loss_his=[])
    X, WX, XY, WXY = O.tensor_quadrature(Q)
    gx, wgx = mod.GaussLobattoJacobiWeights(Q, 0, 0)
    assert np.array_equal(gx, X) and np.array_equal(wgx, WX)
    NEx, NEy = len(gridx) - 1, len(gridy) - 1
    N_testfcn = [NEx * [Ntx], NEy * [Nty]]
    F = O.rhs_2d_literal(gridx, gridy, N_testfcn[0], N_testfcn[1], XY, WXY)
    rng = np.random.default_rng(seed)
    Xb = 2 * rng.random((24, 2)) - 1
    Xb[:6, 1] = 1; Xb[6:12, 1] = -1; Xb[12:18, 0] = 1; Xb[18:, 0] = -1
    ub = O.u_ext_2d(Xb[:, 0:1], Xb[:, 1:2])
    Xf = 2 * rng.random((16, 2)) - 1
    ff = O.f_ext_2d(Xf[:, 0:1], Xf[:, 1:2])
    Xt = 2 * rng.random((32, 2)) - 1
    ut = O.u_ext_2d(Xt[:, 0:1], Xt[:, 1:2])
    model = mod.VPINN(Xb, ub, Xf, ff, XY, WXY, None, F, gridx, gridy, N_testfcn, Xt, ut, layers)
    Ws, bs = perturbed(layers, seed)
    set_params(model, Ws, bs)
    feed = {model.x_tf: model.x, model.y_tf: model.y, model.u_tf: model.utrain,
            model.x_test: model.xtest, model.y_test: model.ytest, model.x_f_tf: model.xf, model.y_f_tf: model.yf}
    out = dict(kind="poisson2d", layers=np.array(layers), gridx=gridx, gridy=gridy, Q=Q, Ntx=Ntx, Nty=Nty,
               var_form=var_form, scheme=scheme, theta=flat(Ws, bs), F_ext=F, X_quad=XY, W_quad=WXY,
               X_u_train=Xb, u_train=ub, X_f_train=Xf, f_train=ff, X_test=Xt, u_test=ut)
    run = model.sess.run
    out["lossv"], out["lossb"], out["lossp"], out["loss"] = run([model.lossv, model.lossb, model.lossp, model.loss], feed)
    out["grad_lossv"], _ = grads(tf, model, model.lossv, feed)
    out["grad_loss"], _ = grads(tf, model, model.loss, feed)
    out["u_test_pred"] = run(model.u_test, feed)
    out["f_pred"] = run(model.f_pred, feed)
    d1x, d2x = model.net_dxu(model.x_test, model.y_test)
    d1y, d2y = model.net_dyu(model.x_test, model.y_test)
    out["d1x"], out["d2x"], out["d1y"], out["d2y"] = run([d1x, d2x, d1y, d2y], feed)
    out["testx"] = model.Test_fcnx(Ntx, XY[:, 0:1])
    out["d1testx"], out["d2testx"] = model.dTest_fcn(Ntx, XY[:, 0:1])
    model.train(adam_steps)
    out["adam_loss_his"] = np.array(mod.loss_his, dtype=np.float64)
    out["adam_theta"] = get_params(model)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "lossv=%.10e loss=%.10e" % (out["lossv"], out["loss"]))


def case_p1d(name, layers, grid, Q, N, var_form, seed, lossb_weight=1.0, LR=0.001, adam_steps=21):
    tf = RL.tf_shim()
    tf.reset_default_graph()
    np.random.seed(seed)
    mod = RL.load("P1D", var_form=var_form, lossb_weight=lossb_weight, LR=LR, total_record=[])
    xq, wq = mod.GaussLobattoJacobiWeights(Q, 0, 0)
    NE = len(grid) - 1
    F = O.rhs_1d(grid, NE * [N], xq, wq)
    Xb = np.array([[-1.0], [1.0]])
    ub = O.u_ext_1d(Xb)
    rng = np.random.default_rng(seed)
    Xf = 2 * rng.random((16, 1)) - 1
    ff = O.f_ext_1d(Xf)
    Xt = np.linspace(-1, 1, 41)[:, None]
    ut = O.u_ext_1d(Xt)
    model = mod.VPINN(Xb, ub, xq[:, None], wq[:, None], F, grid, Xt, ut, layers, Xf, ff)
    Ws, bs = perturbed(layers, seed)
    set_params(model, Ws, bs)
    feed = {model.x_tf: model.x, model.u_tf: model.u, model.x_quad: model.xquad, model.x_test: model.xtest,
            model.xf_tf: model.xf, model.f_tf: model.f}
    out = dict(kind="poisson1d", layers=np.array(layers), grid=grid, Q=Q, N=N, var_form=var_form,
               lossb_weight=lossb_weight, LR=LR, theta=flat(Ws, bs), F_ext=F, X_quad=xq[:, None], W_quad=wq[:, None],
               X_u_train=Xb, u_train=ub, X_f_train=Xf, f_train=ff, X_test=Xt, u_test=ut)
    run = model.sess.run
    out["lossv"], out["lossb"], out["loss"] = run([model.lossv, model.lossb, model.loss], feed)
    out["grad_lossv"], _ = grads(tf, model, model.lossv, feed)
    out["grad_loss"], _ = grads(tf, model, model.loss, feed)
    out["u_test_pred"] = run(model.u_NN_test, feed)
    out["f_pred"] = run(model.f_pred, feed)
    d1, d2 = model.net_du(model.x_test)
    out["d1"], out["d2"] = run([d1, d2], feed)
    out["test"] = model.Test_fcn(N, xq[:, None])
    out["d1test"], out["d2test"] = model.dTest_fcn(N, xq[:, None])
    model.train(adam_steps, 1e-30)
    out["adam_total_record"] = np.array(mod.total_record, dtype=np.float64)
    out["adam_theta"] = get_params(model)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "lossv=%.10e loss=%.10e" % (out["lossv"], out["loss"]))


def case_adi(name, layers, grid_x, grid_t, Q, Ntx, Ntt, var_form, seed, eps0=0.7, V=1.0, LR=0.001, adam_steps=21):
    tf = RL.tf_shim()
    tf.reset_default_graph()
    np.random.seed(seed)
    mod = RL.load("ADI", var_form=var_form, V=V, LR=LR)
    X, WX, XT, WXT = O.tensor_quadrature(Q)
    NEx, NEt = len(grid_x) - 1, len(grid_t) - 1
    N_testfcn = [NEx * [Ntx], NEt * [Ntt]]
    rng = np.random.default_rng(seed)
    Xb = np.hstack((2 * rng.random((24, 1)) - 1, rng.random((24, 1))))
    Xb[:8, 0] = 1; Xb[8:16, 0] = -1; Xb[16:, 1] = 0
    ub = np.where(Xb[:, 1:2] == 0, -np.sin(np.pi * Xb[:, 0:1]), 0.0)
    Xf = np.hstack((2 * rng.random((16, 1)) - 1, rng.random((16, 1))))
    Xt = np.hstack((2 * rng.random((32, 1)) - 1, rng.random((32, 1))))
    ut = rng.standard_normal((32, 1))
    model = mod.VPINN(Xb, ub, Xf, XT, WXT, X, WX, grid_x, grid_t, N_testfcn, Xt, ut, layers, Xt.min(0), Xt.max(0))
    Ws, bs = perturbed(layers, seed)
    set_params(model, Ws, bs)
    model.epsilon.load(np.array([eps0]))
    feed = {model.x_tf: model.x, model.t_tf: model.t, model.u_tf: model.u, model.x_f_tf: model.x_f,
            model.t_f_tf: model.t_f, model.x_quad: model.xquad, model.t_quad: model.tquad,
            model.x_test: model.xtest, model.t_test: model.ttest}
    out = dict(kind="advdiff", layers=np.array(layers), grid_x=grid_x, grid_t=grid_t, Q=Q, Ntx=Ntx, Ntt=Ntt,
               var_form=var_form, V=V, LR=LR, eps0=eps0, theta=flat(Ws, bs), XT_quad=XT, W_quad=WXT, T_quad=X, WT_quad=WX,
               XT_u_train=Xb, u_train=ub, XT_f_train=Xf, XT_test=Xt, u_test=ut)
    run = model.sess.run
    out["lossv"], out["lossb"], out["lossp"], out["loss"] = run([model.lossv, model.lossb, model.lossp, model.loss], feed)
    out["grad_lossv"], ge = grads(tf, model, model.lossv, feed, extra=[model.epsilon])
    out["grad_lossv_eps"] = ge[0]
    out["grad_loss"], ge = grads(tf, model, model.loss, feed, extra=[model.epsilon])
    out["grad_loss_eps"] = ge[0]
    out["u_test_pred"] = run(model.u_NN_test, feed)
    out["f_pred"] = run(model.f_pred, feed)
    # ADI:321 builds np.array([it, loss, epsilon(shape (1,)), a]) -- ragged, an error on numpy >= 1.24 -- so the
    # reference train() loop (ADI:305-321) is driven by hand: same train op, same read-back every 10 iterations.
    rec = []
    for it in range(adam_steps):
        run(model.train_op_Adam, feed)
        if it % 10 == 0:
            rec.append([it, float(run(model.loss, feed)), float(run(model.epsilon, feed)[0])])
    out["adam_total_records"] = np.array(rec)
    out["adam_theta"] = get_params(model)
    out["adam_eps"] = model.epsilon.value.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "lossv=%.10e loss=%.10e" % (out["lossv"], out["loss"]))


def driver_fixtures():
    """Unmodified driver set-up blocks: pins RHS assembly, quadrature layout and grids."""
    ns = RL.run_driver_setup("P2D")
    np.savez_compressed(os.path.join(HERE, "driver_p2d.npz"), grid_x=ns["grid_x"], grid_y=ns["grid_y"],
                        F_ext_total=ns["F_ext_total"], XY_quad=ns["XY_quad_train"], WXY_quad=ns["WXY_quad_train"],
                        N_test_x=np.array(ns["N_test_x"]), N_test_y=np.array(ns["N_test_y"]), layers=np.array(ns["Net_layer"]),
                        X_u_train=ns["X_u_train"], u_train=ns["u_train"])
    for tag, ov in (("driver_p1d", {}), ("driver_p1d_3el", {"N_Element": 3})):
        ns = RL.run_driver_setup("P1D", **ov)
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), grid=ns["grid"], F_ext_total=ns["F_ext_total"],
                            x_quad=ns["X_quad_train"], w_quad=ns["W_quad_train"], N_testfcn=ns["N_testfcn"],
                            layers=np.array(ns["Net_layer"]))
    # ADI:451 builds a ragged array from u_ext()'s (1,1) results -- an error on numpy >= 1.24 -- so the ADI driver is
    # run up to the exact-solution block only (grids, quadrature layout, boundary/initial training points).
    ns = RL.run_driver_setup("ADI", stop_at="def u_ext(")
    np.savez_compressed(os.path.join(HERE, "driver_adi.npz"), grid_x=ns["grid_x"], grid_t=ns["grid_t"],
                        XT_quad=ns["XT_quad_train"], W_quad=ns["WXT_quad_train"], T_quad=ns["T_quad"], WT_quad=ns["WT_quad"],
                        XT_u_train=ns["XT_u_train"], u_train=ns["u_train"], layers=np.array(ns["Net_layer"]),
                        epsilon=ns["epsilon"], V=ns["V"])
    print("driver fixtures written")


def table_fixtures():
    """Quadrature nodes/weights and test-function tables straight from the reference module at the sizes
    BASELINE.json uses (Q=80, N=60) and a few small ones."""
    mod = RL.load("P2D")
    out = {}
    for Q in (5, 10, 50, 80):
        x, w = mod.GaussLobattoJacobiWeights(Q, 0, 0)
        out["gll_x_%d" % Q], out["gll_w_%d" % Q] = x, w
    x80 = out["gll_x_80"]
    m = mod.VPINN.__new__(mod.VPINN)
    out["T_60_80"] = m.Test_fcnx(60, x80)
    out["D1_60_80"], out["D2_60_80"] = m.dTest_fcn(60, x80)
    out["D1_bound_60"], out["D2_bound_60"] = m.dTest_fcn(60, np.array([-1.0, 1.0]))
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **out)
    print("tables written")


if __name__ == "__main__":
    assert RL.available(), "needs /root/reference"
    gx = np.array([-1.0, -0.2, 1.0])
    gy = np.array([-1.0, -0.5, 0.3, 1.0])
    for vf in (0, 1, 2):
        case_p2d("p2d_vf%d" % vf, [2, 5, 5, 5, 1], gx, gy, 10, 5, 4, vf, seed=100 + vf)
    case_p2d("p2d_vf1_w20", [2, 20, 20, 20, 1], np.array([-1.0, 0.0, 1.0]), np.array([-1.0, 0.0, 1.0]), 12, 6, 6, 1, seed=7)
    case_p2d("p2d_pinns", [2, 5, 5, 1], gx, gy, 6, 3, 3, 1, seed=11, scheme="PINNs")
    g3 = np.array([-1.0, -0.1, 0.1, 1.0])
    for vf in (1, 2, 3):
        case_p1d("p1d_vf%d" % vf, [1, 5, 5, 1], g3, 20, 8, vf, seed=200 + vf, lossb_weight=1.0)
    case_p1d("p1d_vf1_w20", [1, 20, 20, 20, 20, 1], np.linspace(-1, 1, 5), 30, 12, 1, seed=9, lossb_weight=2.5)
    for vf in (0, 1):
        case_adi("adi_vf%d" % vf, [2, 5, 5, 5, 1], np.array([-1.0, 0.1, 1.0]), np.array([0.0, 0.4, 1.0]), 10, 5, 4, vf,
                 seed=300 + vf, V=1.0 if vf == 0 else 0.6)
    driver_fixtures()
    table_fixtures()
