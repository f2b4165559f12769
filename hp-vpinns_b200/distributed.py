"""Multi-GPU data parallelism over elements: the variational loss is a sum of independent element
losses (`varloss_total += loss_element`, P2D:120), so each rank owns a contiguous block of the (ex, ey) element
order, evaluates loss and gradient of its block, and ONE all-reduce of [gradient | d eps | losses] (<= 4 KB, fp32)
makes every rank hold the global values; every rank then applies the identical Adam update to its replica of the
parameters.  No activations or residuals are ever exchanged.

One process per GPU (torchrun).  Default: the sum runs inside the step's last kernel over NVLink peer memory
(`connect_peers`); `collective="nccl"` keeps the NCCL all-reduce through torch.distributed on the engine's stream.
The point-wise losses (boundary data, PINN residual) are replicated data: only rank 0 adds them before the reduce.
"""
import numpy as np


def shard_bounds(n_items, rank, world_size):
    """Contiguous block [begin, end) of `n_items` owned by `rank`; blocks differ by at most one item."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return (n_items * rank) // world_size, (n_items * (rank + 1)) // world_size


def shard_slice(n_items, rank, world_size):
    b, e = shard_bounds(n_items, rank, world_size)
    return slice(b, e)


class _DeviceBuffer:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


def reduce_tensor(engine):
    """torch fp32 CUDA tensor aliasing the engine's reduce buffer ([grad_pad | d eps | pad | losses])."""
    import torch
    ptr, n = engine.reduce_buffer()
    return torch.as_tensor(_DeviceBuffer(ptr, n), device="cuda")


def connect_peers(engine, group=None):
    """Connect the engines of the ranks of ONE node for the peer-memory gradient exchange (include/hpv.h:
    hpv_peer_export / hpv_peer_connect): the CUDA IPC handles of every rank's inbox are gathered through
    torch.distributed; afterwards `engine.train_steps` sums the gradient over the ranks inside its last kernel
    (NVLink peer stores + flags) and applies Adam in the same launch -- no NCCL call per step."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world < 2:
        return False
    mine = engine.peer_export(world)
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    engine.peer_connect(rank, world, b"".join(handles))
    dist.barrier(group=group)           # nobody pushes before every rank has mapped every inbox
    return True


class ShardedStep:
    """loss_and_grad -> all-reduce(SUM) -> adam_step, on the current torch CUDA stream.

    `engine` must have been built on this rank's element shard; `point_slots` are enabled on rank 0 only (their
    data is replicated, so summing them over ranks would count them world_size times)."""

    def __init__(self, engine, group=None, collective="peer"):
        import torch
        import torch.distributed as dist
        self.engine, self.dist, self.group = engine, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.peer = False
        stream = torch.cuda.current_stream()
        if stream.cuda_stream == 0:
            raise RuntimeError("run inside `with torch.cuda.stream(torch.cuda.Stream())`: the engine needs a "
                               "non-default stream handle to share with NCCL")
        engine.set_stream(stream.cuda_stream)
        self.red = reduce_tensor(engine) if self.world > 1 else None
        if self.world > 1 and collective == "peer":
            self.peer = connect_peers(engine, group)

    def loss_and_grad(self):
        self.engine.loss_and_grad()
        if self.world > 1:
            self.dist.all_reduce(self.red, group=self.group)

    def step(self):
        if self.world == 1 or self.peer:
            self.engine.train_steps(1, want_history=False)      # reduction, exchange and Adam in the last kernel
        else:
            self.loss_and_grad()
            self.engine.adam_step()


def allreduce_loss_grad_cpu(loss, grad, group=None):
    """Host-side restatement of the reduce (gloo, CPU tensors) used by the world_size-2 CPU tests of the sharding
    logic: returns the global (loss, grad) from this rank's shard values."""
    import torch
    import torch.distributed as dist
    buf = torch.from_numpy(np.concatenate([np.atleast_1d(np.asarray(loss, dtype=np.float64)),
                                           np.asarray(grad, dtype=np.float64).ravel()]))
    dist.all_reduce(buf, group=group)
    out = buf.numpy()
    return float(out[0]), out[1:].copy()
