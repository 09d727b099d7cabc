"""Host-side logic of the data-parallel path on CPU: two gloo ranks average a network's flat gradient buffer with
one all-reduce and fold the 1/world factor into the optimizer (dwc_gan_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(7, 5)
        self.b = nn.Conv2d(3, 4, 3)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dwc_gan_b200 import parallel
    from dwc_gan_b200.flat import FlatParams
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(10 + rank)                     # different weights per rank before the broadcast
    net = _Toy()
    flat = FlatParams(net)
    net.ensure_flat = lambda: flat                   # the two methods GradSync / broadcast_parameters rely on
    parallel.broadcast_parameters(net)
    ref = flat.data.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, flat.data)               # every rank now holds rank 0's weights
    flat.grad.copy_(torch.arange(flat.total, dtype=torch.float32) * (rank + 1))
    sync = parallel.GradSync(async_stream=False)
    sync(net)
    expect = torch.arange(flat.total, dtype=torch.float32) * sum(range(1, world + 1))
    assert torch.equal(flat.grad, expect)            # SUM over ranks; the mean is applied inside the fused Adam
    assert sync.bytes_reduced == flat.total * 4
    # parameters still alias the flat buffers after the collective (conv weights as channels_last views)
    assert net.b.weight.grad.data_ptr() == flat.grad.data_ptr() + flat.offsets["b.weight"] * 4
    out.put((rank, float(flat.grad.sum())))
    dist.destroy_process_group()


def test_gradient_average_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got[0][1] == got[1][1]


class _Toy2(nn.Module):
    """Three 'layer groups' so that bucket ranges and their complement are both non-trivial."""

    def __init__(self):
        super().__init__()
        self.early = nn.Conv2d(3, 4, 3)
        self.mid = nn.Linear(11, 7)
        self.late = nn.Conv2d(4, 5, 3)
        self.tail = nn.Linear(5, 3)


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dwc_gan_b200 import ops, parallel
    from dwc_gan_b200.flat import FlatParams
    parallel.init_from_env(backend="gloo")
    torch.manual_seed(3)
    net = _Toy2()
    flat = FlatParams(net)
    net.ensure_flat = lambda: flat
    sync = parallel.GradSync(async_stream=False)
    sync.set_buckets(net, [("late.",), ("mid.",)])
    total = flat.total
    expect = torch.arange(total, dtype=torch.float32) * sum(range(1, world + 1))
    for it in range(3):
        flat.grad.copy_(torch.arange(total, dtype=torch.float32) * (rank + 1))
        c0 = sync.collectives
        sync.begin(net, key="k")
        assert ops.RT.wgrad_hook is not None
        # backward order: late layer first (two contributions), then mid, then early
        mid_state = None
        for name in ("late.weight", "late.weight", "mid.weight", "early.weight"):
            ops.RT.wgrad_hook(name)
            if name == "late.weight" and mid_state is None:
                mid_state = "after first late"
                if it > 0:          # one of two expected contributions: the bucket must NOT have fired yet
                    o = flat.offsets["late.weight"]
                    assert flat.grad[o + 1] == (o + 1) * (rank + 1)
        sync(net)
        assert ops.RT.wgrad_hook is None
        pad = torch.ones(total, dtype=torch.bool)          # alignment padding between parameters is never reduced
        for nme in flat.names:
            pad[flat.offsets[nme]:flat.offsets[nme] + flat.numels[nme]] = False
        assert torch.equal(flat.grad[~pad], expect[~pad]), it          # every parameter element reduced exactly once
        n = sync.collectives - c0
        # first call: learn (one whole-buffer reduce); afterwards two buckets + ONE coalesced collective for the pieces
        # in front of and behind them
        assert n == (1 if it == 0 else 3), (it, n)
    out.put((rank, sync.bytes_reduced))
    dist.destroy_process_group()


def test_bucketed_gradient_average_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got[0][1] == got[1][1] > 0
