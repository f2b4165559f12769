"""GPU (-m gpu): the drop-in VPINN classes (reference constructor signatures) against fixtures recorded from the
reference's own classes: losses at injected weights, then short Adam trajectories (TF1 Adam semantics)."""
import numpy as np
import pytest

from oracle import hpvpinn_oracle as O
from tests import _cases as C

pytestmark = pytest.mark.gpu


def _n_testfcn(c, kx, ky, nx, ny):
    return [(len(c[kx]) - 1) * [int(c[nx])], (len(c[ky]) - 1) * [int(c[ny])]]


@pytest.mark.parametrize("name", ["p2d_vf0", "p2d_vf1", "p2d_vf2", "p2d_vf1_w20", "p2d_pinns"])
def test_poisson2d_class(name, capsys):
    from hpv_b200.poisson2d import VPINN
    c = C.load(name)
    N = _n_testfcn(c, "gridx", "gridy", "Ntx", "Nty")
    his = []
    m = VPINN(c["X_u_train"], c["u_train"], c["X_f_train"], c["f_train"], c["X_quad"], c["W_quad"], None, c["F_ext"],
              c["gridx"], c["gridy"], N, c["X_test"], c["u_test"], c["layers"], var_form=int(c["var_form"]),
              scheme=str(c["scheme"]), loss_his=his)
    Ws, bs = O.unpack_theta(c["theta"], c["layers"])
    m.set_weights(Ws, bs)
    assert m.sess.run(m.loss) == pytest.approx(float(c["loss"]), rel=2e-5)
    assert m.sess.run(m.lossb) == pytest.approx(float(c["lossb"]), rel=2e-5)
    assert m.sess.run(m.lossv) == pytest.approx(float(c["lossv"]), rel=2e-5)
    assert m.sess.run(m.lossp) == pytest.approx(float(c["lossp"]), rel=2e-5)
    assert np.allclose(m.predict(), c["u_test_pred"], rtol=1e-5, atol=2e-6)
    assert np.allclose(m.net_f(m.xf, m.yf), c["f_pred"], rtol=1e-4, atol=1e-4 * np.abs(c["f_pred"]).max())
    assert np.array_equal(m.Test_fcnx(3, m.xquad[:, 0]), m.Test_fcny(3, m.xquad[:, 0]))
    g, _ = (m.engine.loss_and_grad(), m.engine.read_grad())[1]
    assert np.abs(g - c["grad_loss"]).max() <= 1e-4 * np.abs(c["grad_loss"]).max()
    n = len(c["adam_loss_his"])
    m.train(n)
    assert "It: 0, Loss:" in capsys.readouterr().out
    assert np.allclose(his, c["adam_loss_his"], rtol=5e-5)
    theta = np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in zip(m.weights, m.biases)])
    assert np.abs(theta - c["adam_theta"]).max() <= 2e-5
    m.sess.close()


@pytest.mark.parametrize("name", ["p1d_vf1", "p1d_vf2", "p1d_vf3", "p1d_vf1_w20"])
def test_poisson1d_class(name, capsys):
    from hpv_b200.poisson1d import VPINN
    c = C.load(name)
    rec = []
    m = VPINN(c["X_u_train"], c["u_train"], c["X_quad"], c["W_quad"], c["F_ext"], c["grid"], c["X_test"], c["u_test"],
              c["layers"], c["X_f_train"], c["f_train"], var_form=int(c["var_form"]), lossb_weight=float(c["lossb_weight"]),
              LR=float(c["LR"]), total_record=rec)
    Ws, bs = O.unpack_theta(c["theta"], c["layers"])
    m.set_weights(Ws, bs)
    assert m.sess.run(m.loss) == pytest.approx(float(c["loss"]), rel=2e-5)
    assert m.sess.run(m.lossb) == pytest.approx(float(c["lossb"]), rel=2e-5)
    assert m.sess.run(m.lossv) == pytest.approx(float(c["lossv"]), rel=2e-5)
    assert np.allclose(m.predict(c["X_test"]), c["u_test_pred"], rtol=1e-5, atol=2e-6)
    d1, d2 = m.net_du(c["X_test"])
    assert np.allclose(d1, c["d1"], rtol=1e-4, atol=1e-5)
    assert np.allclose(m.net_f(c["X_test"]), c["f_pred"], rtol=1e-4, atol=1e-5 * max(1.0, np.abs(c["f_pred"]).max()))
    m.train(21, 1e-30)
    out = capsys.readouterr().out
    assert "It: 0, Lossb:" in out
    want = c["adam_total_record"]
    got = np.array(rec)
    assert got.shape == want.shape and np.array_equal(got[:, 0], want[:, 0])
    tol = 2e-3 if name == "p1d_vf3" else 1e-4            # var_form 3: fp32-ill-conditioned gradient
    assert np.allclose(got[:, 1], want[:, 1], rtol=tol)
    with pytest.raises(AttributeError):
        m.predict_subdomain(c["grid"])
    m.sess.close()


@pytest.mark.parametrize("name", ["adi_vf0", "adi_vf1"])
def test_advdiff_class(name, capsys):
    from hpv_b200.advdiff import VPINN
    c = C.load(name)
    N = _n_testfcn(c, "grid_x", "grid_t", "Ntx", "Ntt")
    xt = c["XT_test"]
    m = VPINN(c["XT_u_train"], c["u_train"], c["XT_f_train"], c["XT_quad"], c["W_quad"], c["T_quad"], c["WT_quad"],
              c["grid_x"], c["grid_t"], N, xt, c["u_test"], c["layers"], xt.min(0), xt.max(0), var_form=int(c["var_form"]),
              V=float(c["V"]), LR=float(c["LR"]))
    assert m.sess.run(m.epsilon)[0] == 1.0                           # ADI:63
    Ws, bs = O.unpack_theta(c["theta"], c["layers"])
    m.set_weights(Ws, bs, epsilon=float(c["eps0"]))
    assert m.sess.run(m.loss) == pytest.approx(float(c["loss"]), rel=2e-5)
    assert m.sess.run(m.lossb) == pytest.approx(float(c["lossb"]), rel=2e-5)
    assert m.sess.run(m.lossv) == pytest.approx(float(c["lossv"]), rel=2e-5)
    assert m.sess.run(m.lossp) == pytest.approx(float(c["lossp"]), rel=5e-5)
    assert np.allclose(m.net_f(m.x_f, m.t_f), c["f_pred"], rtol=1e-4, atol=1e-5 * max(1.0, np.abs(c["f_pred"]).max()))
    m.engine.loss_and_grad()
    g, ge = m.engine.read_grad()
    assert np.abs(g - c["grad_loss"]).max() <= 1e-4 * np.abs(c["grad_loss"]).max()
    assert ge == pytest.approx(float(c["grad_loss_eps"][0]), rel=1e-4)
    err, total, u_rec, u_his, t_train = m.train(21, 1e-30)
    want = c["adam_total_records"]
    got = np.array([[r[0], r[1], r[2]] for r in total])
    assert got.shape == want.shape
    assert np.allclose(got[:, 1], want[:, 1], rtol=1e-4) and np.allclose(got[:, 2], want[:, 2], rtol=1e-5)
    assert m.epsilon_value == pytest.approx(float(c["adam_eps"][0]), rel=1e-5)
    assert t_train > 0 and "epsilon:" in capsys.readouterr().out
    m.sess.close()
