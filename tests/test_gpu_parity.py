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


def _random_case(problem, layers, act, var_form, Q, N, seed, grid=(np.array([-1.0, -0.3, 1.0]), np.array([-1.0, 0.4, 1.0]))):
    from oracle import hpvpinn_oracle as O
    rng = np.random.default_rng(seed)
    X, W = O.GaussLobattoJacobiWeights(Q, 0, 0)
    Ws, bs = O.xavier_params(list(layers), seed)
    bs = [0.1 * rng.standard_normal(b.shape) for b in bs]
    D1, D2 = O.dTest_fcn(N, X)
    inp = dict(problem=problem, var_form=var_form, layers=list(layers), act=act, theta=O.pack_theta(Ws, bs), xi=X, w=W,
               T=O.Test_fcn(N, X), D1=D1, D2=D2, d1b=O.dTest_fcn(N, np.array([-1.0, 1.0]))[0])
    if problem == "poisson1d":
        g = np.array([-1.0, -0.2, 0.5, 1.0])
        F = O.rhs_1d(g, 3 * [N], X, W)
        inp.update(lo=g[:-1, None], hi=g[1:, None], ntx=N, nty=1, F=np.asarray(F).reshape(3, 1, N))
        fn = lambda Wt, bt: O.varloss_1d_factorised(Wt, bt, X, W, F, g, var_form)
    else:
        gx, gy = grid
        F = O.rhs_2d_factorised(gx, gy, N, N, X, W)
        lo = np.array([[gx[i], gy[j]] for i in range(len(gx) - 1) for j in range(len(gy) - 1)])
        hi = np.array([[gx[i + 1], gy[j + 1]] for i in range(len(gx) - 1) for j in range(len(gy) - 1)])
        inp.update(lo=lo, hi=hi, ntx=N, nty=N, F=F.reshape(-1, N, N))
        fn = lambda Wt, bt: O.varloss_2d_factorised(Wt, bt, X, W, F, gx, gy, N, N, var_form)
    l_ref, g_ref = O.loss_and_grad(lambda Wt, bt: fn(Wt, bt)[0], Ws, bs)
    r_ref = fn(Ws, bs)[1].detach().numpy()
    return inp, l_ref, g_ref, r_ref


@pytest.mark.parametrize("problem,layers,act,vf,Q,N", [
    ("poisson2d", [2, 32, 32, 1], "tanh", 1, 14, 7),            # padded width 32
    ("poisson2d", [2, 27, 31, 29, 1], "tanh", 0, 12, 6),        # unequal widths, padded to 32, second derivatives
    ("poisson2d", [2, 12, 12, 1], "tanh", 2, 16, 9),            # width padded to 20, var_form 2
    ("poisson1d", [1, 24, 24, 24, 1], "sin", 1, 40, 20),        # padded width 32, 1-D
    ("poisson1d", [1, 3, 1], "sin", 2, 9, 4),                   # one hidden layer, width padded to 8
    ("poisson2d", [2, 6, 6, 6, 6, 6, 6, 6, 6, 1], "tanh", 1, 10, 5),   # eight hidden layers
])
def test_other_network_shapes(problem, layers, act, vf, Q, N):
    inp, l_ref, g_ref, r_ref = _random_case(problem, layers, act, vf, Q, N, seed=Q + N)
    eng = G.make_engine(inp)
    loss, res = eng.varloss_forward()
    g, _ = eng.varloss_backward()
    assert loss == pytest.approx(l_ref, rel=LOSS_RTOL)
    assert np.abs(res - r_ref.reshape(res.shape)).max() <= RES_RTOL * np.abs(r_ref).max()
    tol = 5e-4 if (problem == "poisson1d" and vf == 2) else GRAD_RTOL
    assert np.abs(g - g_ref).max() <= tol * np.abs(g_ref).max()
    eng.close()


@pytest.mark.parametrize("name", ["p2d_vf1", "p2d_vf1_w20", "adi_vf1"])
def test_directional_and_two_tangent_reverse_sweeps(name, monkeypatch):
    """The reverse sweep of first-derivative forms without an eps-dependent coefficient (Poisson-2D var_form 1)
    runs in the directional mode by default; HPV_BWD_DIR=0 (read by hpv_create) keeps the two-tangent sweep.  Both
    must give the reference gradient.  AdvDiff var_form 1 has an eps-dependent term and never takes the
    directional kernel: its gradients (incl. d/d eps) must not depend on the switch."""
    c = C.load(name)
    inp = C.engine_inputs(c)
    gref = c["grad_lossv"]
    got = {}
    for d in ("0", "1"):
        monkeypatch.setenv("HPV_BWD_DIR", d)
        eng = G.make_engine(inp)
        eng.forward_async()
        got[d] = eng.varloss_backward()
        info = eng.kernel_info()
        eng.close()
        assert np.abs(got[d][0] - gref).max() <= GRAD_RTOL * np.abs(gref).max()
        assert info["bwd_directional"] == (1 if (d == "1" and c["kind"] == "poisson2d") else 0)
    if c["kind"] == "advdiff":
        assert np.array_equal(got["0"][0], got["1"][0]) and got["0"][1] == got["1"][1]
        assert got["1"][1] == pytest.approx(float(c["grad_lossv_eps"][0]), rel=1e-4)
    else:
        assert not np.array_equal(got["0"][0], got["1"][0])
        assert np.abs(got["0"][0] - got["1"][0]).max() <= 2e-5 * np.abs(gref).max()


def test_parameter_uploads_are_stream_ordered_and_combined_read():
    """hpv_set_params stages the parameters in pinned memory and returns without synchronising; back-to-back
    uploads followed by hpv_loss_and_grad + hpv_read_losses_and_grad (one synchronisation) must see each upload."""
    c = C.load("p2d_vf1_w20")
    inp = C.engine_inputs(c)
    eng = G.make_engine(inp)
    ref = G.make_engine(inp)
    eng.configure_training(wv=1.0, point_slots=())
    rng = np.random.default_rng(7)
    for k in range(4):
        theta = inp["theta"] * (1.0 + 0.02 * rng.standard_normal(inp["theta"].shape))
        eng.set_params(inp["theta"], 0.0)           # immediately superseded by the next upload
        eng.set_params(theta, 0.0)
        eng.loss_and_grad()
        losses, g, _ = eng.read_losses_and_grad()
        ref.set_params(theta, 0.0)
        lref = ref.varloss_forward(want_residual=False)
        gref, _ = ref.varloss_backward()
        assert losses[1] == pytest.approx(lref, rel=1e-6) and losses[0] == pytest.approx(lref, rel=1e-6)
        assert np.array_equal(g, gref)
    eng.close(); ref.close()


def test_ragged_test_function_counts_on_gpu():
    c = C.load("p2d_vf1")
    inp = C.engine_inputs(c)
    n_el, ntx, nty = inp["lo"].shape[0], inp["ntx"], inp["nty"]
    ntest = np.array([[ntx - (e % 2), nty - (e % 3 == 0)] for e in range(n_el)], dtype=np.int32)
    eng = G.make_engine(dict(inp, ntest=ntest))
    loss, res = eng.varloss_forward()
    full = C.oracle_lossv(c)[1].reshape(n_el, nty, ntx)
    want = sum(np.mean(full[e, :ntest[e, 1], :ntest[e, 0]] ** 2) for e in range(n_el))
    assert loss == pytest.approx(want, rel=LOSS_RTOL)
    eng.close()


def test_many_contexts_alternating_on_one_device():
    """Contexts share the constant-memory copy of the parameters per kernel translation unit; the library re-stages
    it when they alternate."""
    cases = [C.load(n) for n in ("p2d_vf1", "p2d_vf0", "adi_vf0", "p2d_vf2")]      # all padded width 8 -> same copies
    engs = [G.make_engine(C.engine_inputs(c)) for c in cases]
    for _ in range(3):
        for c, e in zip(cases, engs):
            assert e.varloss_forward(want_residual=False) == pytest.approx(float(c["lossv"]), rel=LOSS_RTOL)
        for c, e in zip(reversed(cases), reversed(engs)):
            e.forward_async()
            g, _ = e.varloss_backward()
            assert np.abs(g - c["grad_lossv"]).max() <= GRAD_RTOL * np.abs(c["grad_lossv"]).max()
    for e in engs:
        e.close()


def test_sin_network_in_two_dimensions_and_tanh_in_one():
    """The oracle's loss restatements fix the activation per problem as the reference does (sin in 1-D, tanh in 2-D);
    the engine takes it as a parameter, so the other combinations are checked on net_u and its derivatives."""
    import hpv_b200
    from oracle import hpvpinn_oracle as O
    rng = np.random.default_rng(3)
    for layers, act in (([2, 12, 12, 1], "sin"), ([1, 9, 9, 9, 1], "tanh")):
        Ws, bs = O.xavier_params(layers, 5)
        bs = [0.2 * rng.standard_normal(b.shape) for b in bs]
        pts = 2 * rng.random((333, layers[0])) - 1
        eng = hpv_b200.Engine(0)
        eng.set_network(layers, act)
        eng.set_params(O.pack_theta(Ws, bs))
        u, d1, d2 = eng.net_u(pts, d1=True, d2=True)
        uo, d1o, d2o = O.mlp_forward_mode(pts, Ws, bs, act)
        assert np.allclose(u, uo.numpy(), rtol=1e-5, atol=2e-6)
        assert np.allclose(d1, d1o.numpy(), rtol=1e-4, atol=1e-5)
        assert np.abs(d2 - d2o.numpy()).max() <= 1e-5 * max(1.0, np.abs(d2o.numpy()).max())
        eng.close()


@pytest.mark.parametrize("name", ["p2d_vf1", "p1d_vf1", "adi_vf1"])
def test_loss_total_of_a_step_equals_the_forward_kernels_own_total(name):
    """In a step (hpv_loss_and_grad) the forward kernel leaves the sum of the element losses to the loss assembly of the
    gradient reduction (defer_total); the stand-alone forward call forms it in its very last CTA.  Same element losses,
    both summed in float64: the two totals agree to float rounding, and both match the fixture."""
    c = C.load(name)
    eng = G.make_engine(C.engine_inputs(c))
    lossv, el = eng.varloss_forward(want_residual=False, want_el_loss=True)
    eng.configure_training(wv=1.0, point_slots=(), train_eps=(c["kind"] == "advdiff"))
    eng.loss_and_grad()
    losses = eng.read_losses()
    assert losses[1] == pytest.approx(np.float32(el.sum()), rel=2e-7)
    assert losses[1] == pytest.approx(lossv, rel=2e-7)
    assert lossv == pytest.approx(float(c["lossv"]), rel=1e-5)
    eng.close()


def test_forward_partition_weighting_does_not_change_the_result(monkeypatch):
    """The tensor-core forward kernel distributes its tiles over the CTAs by a cost model (HPV_FWD_BALANCE, read when
    the launch plan is made); with equal tile counts the partial sums of an element are formed over different point
    ranges, so the residuals may differ in the last bits -- and by no more."""
    c = C.load("p2d_vf1_w20")
    inp = C.engine_inputs(c)
    got = {}
    for b in ("0", "1"):
        monkeypatch.setenv("HPV_FWD_BALANCE", b)
        eng = G.make_engine(inp)
        got[b] = eng.varloss_forward(want_residual=True)
        eng.close()
    assert got["0"][0] == pytest.approx(got["1"][0], rel=1e-6)
    assert np.abs(got["0"][1] - got["1"][1]).max() <= 2e-6 * np.abs(got["1"][1]).max()
    assert got["1"][0] == pytest.approx(float(c["lossv"]), rel=1e-5)
