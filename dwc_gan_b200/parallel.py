"""Data parallelism: one process per GPU, gradients averaged with ONE NCCL all-reduce per phase.

The batch shards naturally (no BatchNorm anywhere; every loss is a batch mean), so the
only exchange step is the gradient average: the discriminator's flat gradient buffer after
``loss_dis_all.backward()`` and the generator's after ``loss_gen_total.backward()``.  Because
all gradients of a network live in one flat fp32 buffer (flat.py) that is a single collective
of 56 MB (D) / 81 MB (G) over NVLink 5 / NVSwitch; it is issued on a side stream so the next
phase's independent work (weight packing, zeroing) can overlap, and the optimizer waits on it.

The reference has no distributed code at all (train.py:42); parity for N > 1 is defined as "N
reference replicas each stepping its own shard with averaged gradients" (the text encoder's batch
row mixing makes the forward depend on the local batch).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


class GradSync:
    """Averages a network's flat gradient buffer across ranks.  Parameters that received no gradient this step
    (attention head while attention is off) are skipped consistently: every rank runs the same schedule, and the
    optimizer's 'touched' set, not the buffer content, decides what is updated."""

    def __init__(self, group=None, async_stream=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stream = None
        self.async_stream = async_stream
        self.bytes_reduced = 0

    def __call__(self, net):
        if self.world == 1:
            return
        flat = net.ensure_flat()
        g = flat.grad
        self.bytes_reduced += g.numel() * 4
        if g.is_cuda and self.async_stream:
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=g.device)
            cur = torch.cuda.current_stream(g.device)
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            cur.wait_stream(self.stream)
        else:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        # the 1/world factor is folded into the fused Adam kernel (FusedAdam.grad_scale)


def broadcast_parameters(net, src=0, group=None):
    """Make every rank start from rank `src`'s weights (one broadcast of the flat buffer)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = net.ensure_flat()
    dist.broadcast(flat.data, src=src, group=group)
    flat.bump()


def attach(solver, group=None):
    """Turn a Solver into a data-parallel replica: broadcast weights once, average gradients every phase."""
    broadcast_parameters(solver.gen, group=group)
    broadcast_parameters(solver.dis, group=group)
    solver._dp_sync = GradSync(group)
    solver.gen_opt.grad_scale = solver.dis_opt.grad_scale = 1.0 / solver._dp_sync.world
    return solver
