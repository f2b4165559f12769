"""CPU, world_size 2, gloo: the element-sharding logic of the N>1 path (hp-vpinns_b200/distributed.py).  Each rank
evaluates ITS block of elements (kernel bodies on host threads, tests/emu), one all-reduce of [loss | gradient]
follows, and every rank must hold the whole-batch values of the float64 oracle."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import hpv_b200  # noqa: F401  (registers the package alias)
from hpv_b200 import distributed as D
from tests import _cases as C


def test_shard_bounds_cover_every_element_once():
    for n in (1, 2, 7, 64, 1024):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                b, e = D.shard_bounds(n, r, w)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
            sizes = [D.shard_bounds(n, r, w)[1] - D.shard_bounds(n, r, w)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import _emu as E
    c = C.load(name)
    inp = C.engine_inputs(c)
    sl = D.shard_slice(inp["lo"].shape[0], rank, world)
    sub = dict(inp, lo=inp["lo"][sl], hi=inp["hi"][sl], F=None if inp["F"] is None else inp["F"][sl])
    loss, res, el, g, ge = E.varloss(**sub)
    gl, gg = D.allreduce_loss_grad_cpu(loss, np.concatenate([g, [ge]]))
    out[rank] = (gl, gg)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["p2d_vf1", "adi_vf0"])
def test_two_rank_sharding_reproduces_the_whole_batch(name):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    c = C.load(name)
    o = C.oracle_lossv(c)
    for r in range(world):
        gl, gg = out[r]
        assert gl == pytest.approx(o[0], rel=1e-5)
        assert np.abs(gg[:-1] - o[2]).max() <= 1e-4 * np.abs(o[2]).max()
        if c["kind"] == "advdiff":
            assert gg[-1] == pytest.approx(o[3][0], rel=1e-4)
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])       # identical on every rank
