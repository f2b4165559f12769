"""CPU: the CUDA kernel BODIES (hp-vpinns_b200/csrc/*.cuh) executed on host threads by tests/emu, against the
float64 oracle.  This exercises, without a GPU, the arithmetic (forward-mode MLP, hand-written reverse sweep),
the sum-factorised projection and its adjoint, the work partition with elements split over several CTAs, the
last-arriver reductions and the self-resetting counters.  fp32 tolerances as in the GPU parity tests."""
import numpy as np
import pytest

from oracle import hpvpinn_oracle as O
from tests import _cases as C
from tests import _emu as E
from tests._gpu import GRAD_RTOL

CASES = ["p2d_vf0", "p2d_vf1", "p2d_vf2", "p1d_vf1", "p1d_vf2", "p1d_vf3", "adi_vf0", "adi_vf1"]


@pytest.mark.parametrize("name", CASES)
def test_emulated_forward_backward(name):
    c = C.load(name)
    inp = C.engine_inputs(c)
    loss, res, el, g, ge = E.varloss(**inp)
    o = C.oracle_lossv(c)
    assert loss == pytest.approx(o[0], rel=1e-5)
    assert loss == pytest.approx(float(c["lossv"]), rel=1e-5)
    ores = o[1].reshape(res.shape)
    assert np.abs(res - ores).max() <= 2e-5 * np.abs(ores).max()
    assert el.sum() == pytest.approx(loss, rel=1e-6)
    assert np.abs(g - o[2]).max() <= GRAD_RTOL.get(name, 1e-4) * np.abs(o[2]).max()
    if c["kind"] == "advdiff":
        assert ge == pytest.approx(o[3][0], rel=1e-4)


@pytest.mark.parametrize("name", ["p2d_vf1", "p2d_vf1_w20"])
def test_directional_reverse_sweep_matches_two_tangent_sweep(name):
    """Poisson-2D var_form 1 projects first derivatives only, so the reverse sweep carries ONE tangent along the
    per-point direction (gbar_x, gbar_y) (HpvMode<2,1,0>::DIR) instead of the x and y tangents.  Both sweeps give
    the gradient of the same loss: each against the oracle, and against each other."""
    c = C.load(name)
    inp = C.engine_inputs(c)
    gref = C.oracle_lossv(c)[2]
    out = {}
    try:
        for d in (0, 1):
            E.lib().hpv_emu_set_bwd_dir(d)
            out[d] = E.varloss(**inp)[3]
            assert np.abs(out[d] - gref).max() <= 1e-4 * np.abs(gref).max()
    finally:
        E.lib().hpv_emu_set_bwd_dir(1)
    assert not np.array_equal(out[0], out[1])           # two different kernels ...
    assert np.abs(out[0] - out[1]).max() <= 2e-5 * np.abs(gref).max()      # ... one gradient


@pytest.mark.parametrize("layers,var_form", [([2, 27, 31, 1], 1), ([2, 32, 32, 1], 0)])
def test_padded_width_32_on_emulated_kernels(layers, var_form):
    """Hidden width padded to 32: the weight-gradient GEMM has more register tiles (72) than a warp has lanes and
    walks them in batches; unequal widths exercise the zero padding.  var_form 1 takes the directional sweep,
    var_form 0 the five-channel one."""
    rng = np.random.default_rng(11)
    Q, N = 8, 4
    X, W = O.GaussLobattoJacobiWeights(Q, 0, 0)
    gx, gy = np.array([-1.0, 0.1, 1.0]), np.array([-1.0, 1.0])
    F = O.rhs_2d_factorised(gx, gy, N, N, X, W)
    Ws, bs = O.xavier_params(layers, 4)
    bs = [0.1 * rng.standard_normal(b.shape) for b in bs]
    theta = O.pack_theta(Ws, bs)
    lo = np.array([[gx[i], gy[0]] for i in range(2)])
    hi = np.array([[gx[i + 1], gy[1]] for i in range(2)])
    D1, D2 = O.dTest_fcn(N, X)
    kw = dict(problem="poisson2d", var_form=var_form, layers=layers, act="tanh", xi=X, w=W, T=O.Test_fcn(N, X), D1=D1, D2=D2,
              d1b=None, lo=lo, hi=hi, ntx=N, nty=N, F=F.reshape(2, N, N), theta=theta)
    l_ref, g_ref = O.loss_and_grad(lambda Wt, bt: O.varloss_2d_factorised(Wt, bt, X, W, F, gx, gy, N, N, var_form)[0], Ws, bs)
    loss, res, el, g, _ = E.varloss(n_ctas_fwd=2, n_ctas_bwd=2, bwd_block=64, **kw)
    assert loss == pytest.approx(float(l_ref), rel=1e-5)
    assert np.abs(g - g_ref).max() <= 1e-4 * np.abs(g_ref).max()


@pytest.mark.parametrize("bwd_block", [32, 96, 160])
def test_reverse_sweep_block_shapes(bwd_block):
    """The reverse sweep is launched as one CTA of W warps per SM, W chosen per launch (1 warp for a handful of
    boundary points, 11-15 for the element batch): any number of warps, including non-powers of two, must give the
    same gradient (the CTA-wide sums and the trip count of the tile groups depend on it)."""
    c = C.load("p2d_vf1_w20")
    inp = C.engine_inputs(c)
    gref = C.oracle_lossv(c)[2]
    g = E.varloss(n_ctas_fwd=2, n_ctas_bwd=2, bwd_block=bwd_block, **inp)[3]
    assert np.abs(g - gref).max() <= 1e-4 * np.abs(gref).max()
    xb, ub = c["X_u_train"], c["u_train"]
    Ws, bs = O.unpack_theta(c["theta"], c["layers"])
    lref, gref = O.loss_and_grad(lambda W, b: 10 * O.lossb(W, b, xb, ub, "tanh"), Ws, bs)
    _, _, _, lb, g, _ = E.points(c["layers"], "tanh", c["theta"], xb, target=ub.ravel(), a0=[1, 0, 0, 0, 0], weight=10.0, mode=0,
                                 backward=True, bwd_block=bwd_block)
    assert lb == pytest.approx(lref, rel=1e-5)
    assert np.abs(g - gref).max() <= 1e-4 * np.abs(gref).max()


@pytest.mark.parametrize("name", ["c1", "c2", "c5"])
def test_baseline_configs_at_full_size_on_emulated_kernels(name):
    """C1, C2 and C5 of BASELINE.json at their full sizes (Q = 50 / 80, up to 60 x 60 test functions) through the kernel
    bodies on host threads; the GPU twin is tests/test_gpu_zz_baseline_configs.py."""
    from tests import _baseline_cases as B
    inp, l_ref, g_ref, r_ref, ge_ref = B.build(name)
    loss, res, el, g, ge = E.varloss(n_ctas_fwd=8, n_ctas_bwd=8, bwd_block=128, **inp)
    assert loss == pytest.approx(l_ref, rel=1e-5)
    assert np.abs(res - r_ref.reshape(res.shape)).max() <= 2e-5 * np.abs(r_ref).max()
    assert np.abs(g - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
    if ge_ref is not None:
        assert ge == pytest.approx(ge_ref, rel=1e-4)


def test_partition_independence_and_determinism():
    """Same numbers whatever the number of CTAs an element is split over (fixed-order reductions)."""
    c = C.load("p2d_vf1_w20")          # Q = 12 -> 144 points per element -> one tile per element
    inp = C.engine_inputs(c)
    base = E.varloss(n_ctas_fwd=1, n_ctas_bwd=1, **inp)
    for nf, nb in ((2, 3), (4, 5)):
        out = E.varloss(n_ctas_fwd=nf, n_ctas_bwd=nb, **inp)
        assert out[0] == pytest.approx(base[0], rel=1e-6)
        assert np.allclose(out[1], base[1], rtol=1e-5, atol=1e-6)
        assert np.abs(out[3] - base[3]).max() <= 1e-5 * np.abs(base[3]).max()
    again = E.varloss(n_ctas_fwd=4, n_ctas_bwd=5, **inp)
    assert again[0] == out[0] and np.array_equal(again[1], out[1]) and np.array_equal(again[3], out[3])


def test_element_split_over_several_ctas():
    """Q = 20 -> 400 points per element = 2 tiles: with 5 CTAs over 3 elements (6 tiles) elements are shared
    between CTAs and the partial-U reduction path is taken."""
    rng = np.random.default_rng(5)
    layers = [2, 5, 5, 1]
    Q, N = 20, 6
    X, W = O.GaussLobattoJacobiWeights(Q, 0, 0)
    gx, gy = np.array([-1.0, 0.2, 1.0]), np.array([-1.0, 0.5, 1.0])
    F = O.rhs_2d_factorised(gx, gy, N, N, X, W)
    Ws, bs = O.xavier_params(layers, 3)
    bs = [0.1 * rng.standard_normal(b.shape) for b in bs]
    theta = O.pack_theta(Ws, bs)
    lo = np.array([[gx[i], gy[j]] for i in range(2) for j in range(2)])
    hi = np.array([[gx[i + 1], gy[j + 1]] for i in range(2) for j in range(2)])
    D1, D2 = O.dTest_fcn(N, X)
    kw = dict(problem="poisson2d", var_form=1, layers=layers, act="tanh", xi=X, w=W, T=O.Test_fcn(N, X), D1=D1, D2=D2,
              d1b=None, lo=lo, hi=hi, ntx=N, nty=N, F=F.reshape(4, N, N), theta=theta)
    ref_l, ref_res = O.varloss_2d_factorised(Ws, bs, X, W, F, gx, gy, N, N, 1)
    l_ref, g_ref = O.loss_and_grad(lambda Wt, bt: O.varloss_2d_factorised(Wt, bt, X, W, F, gx, gy, N, N, 1)[0], Ws, bs)
    for nf in (1, 3, 5, 8):
        loss, res, el, g, _ = E.varloss(n_ctas_fwd=nf, n_ctas_bwd=3, bwd_block=64, **kw)
        assert loss == pytest.approx(float(ref_l), rel=1e-5)
        assert np.abs(res - ref_res.numpy()).max() <= 2e-5 * np.abs(ref_res.numpy()).max()
        assert np.abs(g - g_ref).max() <= 1e-4 * np.abs(g_ref).max()


def test_ragged_test_function_counts():
    """Per-element numbers of test functions (the reference's nominal p-refinement, P2D:72-73): entries beyond an
    element's (ntx, nty) do not enter its mean."""
    c = C.load("p2d_vf1")
    inp = C.engine_inputs(c)
    n_el = inp["lo"].shape[0]
    ntx, nty = inp["ntx"], inp["nty"]
    ntest = np.array([[ntx - (e % 2), nty - (e % 3 == 0)] for e in range(n_el)], dtype=np.int32)
    loss, res, el, g, _ = E.varloss(ntest=ntest, **inp)
    full = C.oracle_lossv(c)[1].reshape(n_el, nty, ntx)
    want = sum(np.mean(full[e, :ntest[e, 1], :ntest[e, 0]] ** 2) for e in range(n_el))
    assert loss == pytest.approx(want, rel=1e-5)
    for e in range(n_el):
        assert np.all(res[e, ntest[e, 1]:, :] == 0) and np.all(res[e, :, ntest[e, 0]:] == 0)


@pytest.mark.parametrize("name", ["p2d_vf0", "p1d_vf1", "adi_vf0"])
def test_emulated_net_u_and_point_losses(name):
    c = C.load(name)
    act = C.ACT[c["kind"]]
    layers = c["layers"]
    xt = c["XT_test"] if c["kind"] == "advdiff" else c["X_test"]
    u, d1, d2, _, _, _ = E.points(layers, act, c["theta"], xt)
    uo, d1o, d2o = O.mlp_forward_mode(xt, *O.unpack_theta(c["theta"], layers), act)
    assert np.allclose(u, uo.numpy(), rtol=1e-5, atol=2e-6)
    assert np.allclose(d1, d1o.numpy(), rtol=1e-4, atol=1e-5)
    assert np.abs(d2 - d2o.numpy()).max() <= 1e-5 * max(1.0, np.abs(d2o.numpy()).max())
    # boundary loss (value channel only) and its gradient
    xb = c["XT_u_train"] if c["kind"] == "advdiff" else c["X_u_train"]
    ub = c["u_train"]
    Ws, bs = O.unpack_theta(c["theta"], layers)
    lref, gref = O.loss_and_grad(lambda W, b: 10 * O.lossb(W, b, xb, ub, act), Ws, bs)
    _, _, _, lb, g, _ = E.points(layers, act, c["theta"], xb, target=ub.ravel(), a0=[1, 0, 0, 0, 0], weight=10.0, mode=0, backward=True)
    assert lb == pytest.approx(lref, rel=1e-5)
    assert np.abs(g - gref).max() <= 1e-4 * np.abs(gref).max()
    # strong-form residual loss (second derivatives, and d/d eps for AdvDiff) and its gradient
    import torch
    if c["kind"] == "advdiff":
        xf = c["XT_f_train"]; eps0 = float(c["eps0"]); V = float(c["V"])
        fn = lambda W, b, e: torch.mean(O.net_f(W, b, xf, act, "advdiff", eps=e, V=V) ** 2)
        lref, gref, geref = O.loss_and_grad(fn, Ws, bs, extra=np.array([eps0]))
        _, _, _, lp, g, ge = E.points(layers, act, c["theta"], xf, eps=eps0, target=np.zeros(len(xf)), a0=[0, V, 1, 0, 0],
                                      a1=[0, 0, 0, -1, 0], backward=True)
        assert ge == pytest.approx(geref[0], rel=1e-4)
    elif c["kind"] == "poisson2d":
        xf, ff = c["X_f_train"], c["f_train"].ravel()
        fn = lambda W, b: torch.mean((O.net_f(W, b, xf, act, "poisson2d") - torch.as_tensor(ff)) ** 2)
        lref, gref = O.loss_and_grad(fn, Ws, bs)
        _, _, _, lp, g, _ = E.points(layers, act, c["theta"], xf, target=ff, a0=[0, 0, 0, 1, 1], backward=True)
    else:
        xf, ff = c["X_f_train"], c["f_train"].ravel()
        fn = lambda W, b: torch.mean((O.net_f(W, b, xf, act, "poisson1d") - torch.as_tensor(ff)) ** 2)
        lref, gref = O.loss_and_grad(fn, Ws, bs)
        _, _, _, lp, g, _ = E.points(layers, act, c["theta"], xf, target=ff, a0=[0, 0, 0, -1, 0], backward=True)
    assert lp == pytest.approx(lref, rel=2e-5)
    assert np.abs(g - gref).max() <= 2e-4 * np.abs(gref).max()
