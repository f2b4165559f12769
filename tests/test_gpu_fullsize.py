"""GPU (-m gpu): the BASELINE.json-size configurations.  C3 (8x8 elements, Q=80, 60x60 test functions,
[2,20,20,20,1]) is compared directly with the factorised float64 oracle (a few seconds of CPU); on top of that,
size-independent properties of the path: linearity of U = Res + F in the output-layer weights, sum of element
losses = lossv, bit-identical re-evaluation, and the known-answer RHS (projection of an analytic field)."""
import numpy as np
import pytest

from oracle import hpvpinn_oracle as O
from tests import _gpu as G

pytestmark = pytest.mark.gpu


def _c3(ne=8, Q=80, N=60, layers=(2, 20, 20, 20, 1), seed=1234, vf=1):
    X, W = O.GaussLobattoJacobiWeights(Q, 0, 0)
    g = np.linspace(-1, 1, ne + 1)
    lo = np.array([[g[i], g[j]] for i in range(ne) for j in range(ne)])
    hi = np.array([[g[i + 1], g[j + 1]] for i in range(ne) for j in range(ne)])
    F = O.rhs_2d_factorised(g, g, N, N, X, W)
    Ws, bs = O.xavier_params(list(layers), seed)
    rng = np.random.default_rng(seed + 1)
    bs = [0.05 * rng.standard_normal(b.shape) for b in bs]
    D1, D2 = O.dTest_fcn(N, X)
    inp = dict(problem="poisson2d", var_form=vf, layers=list(layers), act="tanh", theta=O.pack_theta(Ws, bs), xi=X, w=W,
               T=O.Test_fcn(N, X), D1=D1, D2=D2, d1b=None, lo=lo, hi=hi, ntx=N, nty=N, F=F.reshape(ne * ne, N, N))
    return inp, (Ws, bs, X, W, F, g)


def test_c3_direct_parity_with_oracle():
    inp, (Ws, bs, X, W, F, g) = _c3()
    eng = G.make_engine(inp)
    loss, res, el = eng.varloss_forward(want_residual=True, want_el_loss=True)
    grad, _ = eng.varloss_backward()
    l_ref, g_ref = O.loss_and_grad(lambda Wt, bt: O.varloss_2d_factorised(Wt, bt, X, W, F, g, g, 60, 60, 1)[0], Ws, bs)
    r_ref = O.varloss_2d_factorised(Ws, bs, X, W, F, g, g, 60, 60, 1)[1].numpy()
    assert loss == pytest.approx(l_ref, rel=1e-5)
    assert np.abs(res - r_ref).max() <= 2e-5 * np.abs(r_ref).max()
    assert np.abs(grad - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
    assert el.sum() == pytest.approx(loss, rel=1e-6)
    loss2, res2 = eng.varloss_forward()
    grad2, _ = eng.varloss_backward()
    assert loss2 == loss and np.array_equal(res, res2) and np.array_equal(grad, grad2)
    eng.close()


def test_c3_linearity_in_output_layer():
    inp, _ = _c3()
    eng = G.make_engine(inp)
    F = inp["F"]
    theta = inp["theta"].copy()
    n_out = 21                                           # W_out (20) + b_out (1) are the last 21 parameters
    _, r1 = eng.varloss_forward()
    t2 = theta.copy(); t2[-n_out:] *= 2.5
    eng.set_params(t2)
    _, r2 = eng.varloss_forward()
    U1, U2 = r1 + F, r2 + F
    assert np.abs(U2 - 2.5 * U1).max() <= 2e-5 * np.abs(U2).max()
    t0 = theta.copy(); t0[-n_out:] = 0
    eng.set_params(t0)
    l0, r0 = eng.varloss_forward()
    assert np.abs(r0 + F).max() <= 1e-6 * np.abs(F).max()           # zero network -> Res = -F
    assert l0 == pytest.approx(sum(np.mean(F[e].astype(np.float32) ** 2) for e in range(F.shape[0])), rel=1e-5)
    eng.close()


def test_c4_shard_sum_equals_whole():
    """Element sharding: lossv and d theta of two half batches add up to those of the whole batch."""
    inp, _ = _c3(ne=6)
    whole = G.make_engine(inp)
    l, _ = whole.varloss_forward()
    g, _ = whole.varloss_backward()
    parts = []
    for sl in (slice(0, 18), slice(18, 36)):
        sub = dict(inp, lo=inp["lo"][sl], hi=inp["hi"][sl], F=inp["F"][sl])
        e = G.make_engine(sub)
        ls, _ = e.varloss_forward()
        gs, _ = e.varloss_backward()
        parts.append((ls, gs))
        e.close()
    assert parts[0][0] + parts[1][0] == pytest.approx(l, rel=1e-6)
    assert np.abs(parts[0][1] + parts[1][1] - g).max() <= 1e-5 * np.abs(g).max()
    whole.close()


def test_limits_are_reported_not_silently_wrong():
    import hpv_b200
    eng = hpv_b200.Engine(0)
    with pytest.raises(hpv_b200.HpvError):
        eng.set_network([2, 64, 64, 1], "tanh")                      # hidden width > 32
    eng.set_network([2, 5, 1], "tanh")
    with pytest.raises(hpv_b200.HpvError):
        eng.set_quadrature(np.linspace(-1, 1, 200), np.ones(200))    # Q > 128
    with pytest.raises(hpv_b200.HpvError):
        eng.varloss_forward()                                        # nothing set up
    with pytest.raises(hpv_b200.HpvError):
        eng.set_elements(np.zeros((0, 2)), np.zeros((0, 2)), 5, 5)    # empty element batch
    eng.close()


def test_rhs_assembly_on_gpu_matches_reference_driver():
    """F_ext_total assembled by the fused projection kernel (hpv_project_field) against the array the reference
    driver itself builds (fixture driver_p2d.npz / driver_p1d_3el.npz, produced by the unmodified P2D:384-414 and
    P1D:275-294 loops) and against the float64 oracle at the C3 size."""
    import os
    from tests import _cases as C
    d = dict(np.load(os.path.join(C.GOLDEN, "driver_p2d.npz")))
    X, W = O.GaussLobattoJacobiWeights(10, 0, 0)
    gx, gy = d["grid_x"], d["grid_y"]
    lo = np.array([[gx[i], gy[j]] for i in range(4) for j in range(4)])
    hi = np.array([[gx[i + 1], gy[j + 1]] for i in range(4) for j in range(4)])
    D1, D2 = O.dTest_fcn(5, X)
    inp = dict(problem="poisson2d", var_form=1, layers=[2, 5, 5, 5, 1], act="tanh", theta=np.zeros(81), xi=X, w=W,
               T=O.Test_fcn(5, X), D1=D1, D2=D2, d1b=None, lo=lo, hi=hi, ntx=5, nty=5, F=None)
    eng = G.make_engine(inp)
    F = eng.assemble_rhs(O.f_ext_2d, lo, hi, X)
    ref = d["F_ext_total"].reshape(16, 5, 5)
    assert np.abs(F - ref).max() <= 2e-6 * np.abs(ref).max()
    eng.close()
    # 1-D, non-uniform 3-element grid of the reference driver
    d1 = dict(np.load(os.path.join(C.GOLDEN, "driver_p1d_3el.npz")))
    xq, wq, g = d1["x_quad"].ravel(), d1["w_quad"].ravel(), d1["grid"]
    Dq1, Dq2 = O.dTest_fcn(60, xq)
    inp1 = dict(problem="poisson1d", var_form=1, layers=[1, 5, 1], act="sin", theta=np.zeros(16), xi=xq, w=wq, T=O.Test_fcn(60, xq),
                D1=Dq1, D2=Dq2, d1b=O.dTest_fcn(60, np.array([-1.0, 1.0]))[0], lo=g[:-1, None], hi=g[1:, None], ntx=60, nty=1, F=None)
    e1 = G.make_engine(inp1)
    F1 = e1.assemble_rhs(O.f_ext_1d, g[:-1, None], g[1:, None], xq)
    ref1 = d1["F_ext_total"].reshape(3, 1, 60)
    assert np.abs(F1 - ref1).max() <= 1e-5 * np.abs(ref1).max()
    e1.close()
    # C3 size against the factorised float64 assembly
    inp3, (Ws, bs, X3, W3, F3, g3) = _c3()
    e3 = G.make_engine(inp3)
    Fg = e3.assemble_rhs(O.f_ext_2d, inp3["lo"], inp3["hi"], X3)
    assert np.abs(Fg - inp3["F"]).max() <= 2e-6 * np.abs(inp3["F"]).max()
    e3.close()
