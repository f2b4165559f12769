"""GPU parity (-m gpu): the CUDA path, called through the C ABI (ctypes), against the float64 oracle and the
golden fixtures produced by the reference's own classes.  Tolerances: the north star asks for the loss within
1e-5 relative of the reference's float64 TF-CPU path; residual entries are checked to 2e-5 of the largest
residual entry, gradients to 1e-4 of the largest gradient entry (looser for the two ill-conditioned 1-D
forms, see tests/_gpu.py)."""
import numpy as np
import pytest

from tests import _cases as C
from tests import _gpu as G

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
RES_RTOL = 2e-5
GRAD_RTOL = 1e-4


@pytest.mark.parametrize("name", C.case_names())
def test_varloss_forward_backward_vs_oracle_and_golden(name):
    c = C.load(name)
    inp = C.engine_inputs(c)
    eng = G.make_engine(inp)
    loss, res, el = eng.varloss_forward(want_residual=True, want_el_loss=True)
    o = C.oracle_lossv(c)
    assert loss == pytest.approx(o[0], rel=LOSS_RTOL)
    assert loss == pytest.approx(float(c["lossv"]), rel=LOSS_RTOL)          # the reference class's own number
    ores = o[1].reshape(res.shape)
    assert np.abs(res - ores).max() <= RES_RTOL * np.abs(ores).max()
    assert np.allclose(el.sum(), loss, rtol=1e-6)
    g, ge = eng.varloss_backward()
    gref = c["grad_lossv"]
    tol = G.GRAD_RTOL.get(name, GRAD_RTOL)
    assert np.abs(g - gref).max() <= tol * np.abs(gref).max()
    if c["kind"] == "advdiff":
        assert ge == pytest.approx(float(c["grad_lossv_eps"][0]), rel=1e-4)
    # determinism: a second evaluation is bit-identical
    loss2, res2 = eng.varloss_forward(want_residual=True)
    g2, _ = eng.varloss_backward()
    assert loss2 == loss and np.array_equal(res, res2) and np.array_equal(g, g2)
    eng.close()


@pytest.mark.parametrize("name", C.case_names())
def test_net_u_and_derivatives(name):
    c = C.load(name)
    inp = C.engine_inputs(c)
    eng = G.make_engine(inp)
    xt = c["XT_test"] if c["kind"] == "advdiff" else c["X_test"]
    u, d1, d2 = eng.net_u(xt, d1=True, d2=True)
    assert np.allclose(u, c["u_test_pred"][:, 0], rtol=1e-5, atol=2e-6)
    if c["kind"] == "poisson2d":
        s = max(1.0, np.abs(c["d2x"]).max())
        assert np.allclose(d1[:, 0], c["d1x"][:, 0], rtol=1e-4, atol=1e-5)
        assert np.allclose(d1[:, 1], c["d1y"][:, 0], rtol=1e-4, atol=1e-5)
        assert np.abs(d2[:, 0] - c["d2x"][:, 0]).max() <= 1e-5 * s
        assert np.abs(d2[:, 1] - c["d2y"][:, 0]).max() <= 1e-5 * s
    elif c["kind"] == "poisson1d":
        assert np.allclose(d1[:, 0], c["d1"][:, 0], rtol=1e-4, atol=1e-5)
        assert np.abs(d2[:, 0] - c["d2"][:, 0]).max() <= 1e-5 * max(1.0, np.abs(c["d2"]).max())
    eng.close()
