"""GPU (-m gpu), needs >= 2 GPUs of one node (skipped otherwise): the multi-GPU training step whose gradient sum
runs inside the step's last kernel over NVLink peer memory (include/hpv.h: hpv_peer_export / hpv_peer_connect).
Two processes, one per GPU, each with its block of elements; after a few Adam steps both ranks must hold bitwise
identical parameters, equal (to fp32 summation order) to those of one engine that trained on all elements."""
import os
import socket

import numpy as np
import pytest

from tests import _cases as C
from tests import _gpu as G

pytestmark = pytest.mark.gpu

NSTEPS = 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, collective, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hpv_b200 import distributed as D
    c = C.load(name)
    inp = C.engine_inputs(c)
    sl = D.shard_slice(inp["lo"].shape[0], rank, world)
    sub = dict(inp, lo=inp["lo"][sl], hi=inp["hi"][sl], F=None if inp["F"] is None else inp["F"][sl])
    eng = G.make_engine(sub, device=rank)
    eng.configure_training(wv=1.0, point_slots=(), lr=1e-3, train_eps=(c["kind"] == "advdiff"))
    assert D.connect_peers(eng)
    hist = eng.train_steps(NSTEPS)
    theta, eps = eng.get_params()
    out[rank] = (hist[:, :2].copy(), theta, eps)
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["p2d_vf1", "adi_vf0"])
def test_peer_exchange_step_matches_single_engine(name):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), name, "peer", out), nprocs=world, join=True)
    c = C.load(name)
    eng = G.make_engine(C.engine_inputs(c))
    eng.configure_training(wv=1.0, point_slots=(), lr=1e-3, train_eps=(c["kind"] == "advdiff"))
    hist = eng.train_steps(NSTEPS)
    theta, eps = eng.get_params()
    eng.close()
    h0, t0, e0 = out[0]
    h1, t1, e1 = out[1]
    assert np.array_equal(t0, t1) and e0 == e1 and np.array_equal(h0, h1)          # identical on every rank
    assert np.allclose(h0, hist[:, :2], rtol=2e-6)                                 # global loss history
    assert np.abs(t0 - theta).max() <= 2e-6 * max(1.0, np.abs(theta).max())
    assert e0 == pytest.approx(eps, rel=1e-5, abs=1e-7)


def _p2d_class_args():
    """Constructor arguments of the Poisson-2D class from the reference driver's own arrays (driver_p2d.npz)."""
    from oracle import hpvpinn_oracle as O
    d = dict(np.load(os.path.join(C.GOLDEN, "driver_p2d.npz")))
    rng = np.random.default_rng(0)
    X_f = 2 * rng.random((100, 2)) - 1
    f = O.f_ext_2d(X_f[:, 0:1], X_f[:, 1:2])
    xs = np.linspace(-1, 1, 11)
    X_test = np.array([[a, b] for a in xs for b in xs])
    u_test = O.u_ext_2d(X_test[:, 0:1], X_test[:, 1:2])
    N_testfcn = [list(d["N_test_x"]), list(d["N_test_y"])]
    args = (d["X_u_train"], d["u_train"], X_f, f, d["XY_quad"], d["WXY_quad"], None, d["F_ext_total"], d["grid_x"], d["grid_y"],
            N_testfcn, X_test, u_test, [int(v) for v in d["layers"]])
    return args, len(N_testfcn[0]) * len(N_testfcn[1])


def _class_worker(rank, world, port, out):
    """The INTEGRATION.md multi-GPU recipe: class on this rank's element shard, attach_distributed(), train()."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hpv_b200 import distributed as D
    from hpv_b200.poisson2d import VPINN
    args, n_el = _p2d_class_args()
    his = []
    model = VPINN(*args, var_form=1, scheme="VPINNs", loss_his=his, device=rank, seed=7,
                  elements=D.shard_slice(n_el, rank, world))           # owns_point_losses: inferred, rank 0 only
    assert model.owns_point_losses == (rank == 0)
    assert model.attach_distributed()
    first = model.sess.run(model.loss)                                 # global value on every rank
    model.train(NSTEPS)
    theta = np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in zip(model.weights, model.biases)])
    out[rank] = (first, list(his), theta)
    model.sess.close()
    dist.destroy_process_group()


def test_sharded_class_with_boundary_loss_matches_single_engine():
    """with the boundary slot enabled on every rank the exchanged objective was
    world_size*10*lossb + lossv.  The sharded classes must train the SAME objective as one engine."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from hpv_b200.poisson2d import VPINN
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_class_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    args, _ = _p2d_class_args()
    his = []
    model = VPINN(*args, var_form=1, scheme="VPINNs", loss_his=his, seed=7)
    first = model.sess.run(model.loss)
    model.train(NSTEPS)
    theta = np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in zip(model.weights, model.biases)])
    model.sess.close()
    f0, h0, t0 = out[0]
    f1, h1, t1 = out[1]
    assert np.array_equal(t0, t1) and h0 == h1 and f0 == f1                        # identical on every rank
    assert f0 == pytest.approx(first, rel=2e-6)
    assert np.allclose(h0, his, rtol=5e-6)
    assert np.abs(t0 - theta).max() <= 5e-6 * max(1.0, np.abs(theta).max())
