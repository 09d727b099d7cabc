"""Data parallelism: one process per GPU, gradients averaged with bucketed NCCL all-reduces overlapped with backward.

The batch shards naturally (no BatchNorm anywhere; every loss is a batch mean), so the
only exchange step is the gradient average: the discriminator's flat gradient buffer during / after
``loss_dis_all.backward()`` and the generator's during / after ``loss_gen_total.backward()``.  All
gradients of a network live in one flat fp32 buffer (flat.py); contiguous ranges of it whose gradients
are final early in backward are reduced on a communication stream while backward continues (GradSync),
the rest when backward ends, and the optimizer waits on the communication stream (56 MB for D, 81 MB
for G per step over NVLink 5 / NVSwitch).

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
    """Averages a network's flat gradient buffer across ranks, in BUCKETS that overlap the rest of backward.

    A bucket is a contiguous range of the flat buffer (a group of layers named by prefixes) whose gradients are final
    long before backward ends: the discriminator's two deepest layers per scale (90 % of its parameters; backward visits
    them first) and the whole decoder (final once the batched decode's backward is through, while the encoders' backward
    still runs).  ops.side_launch reports every weight-gradient launch; when all launches a bucket expects have been
    enqueued, the bucket's all-reduce is issued on the communication stream behind the weight-gradient stream and the
    calling stream, and runs under the remaining backward kernels.  What is left (encoders, text encoder, MLP) is
    reduced at the end of the phase.  The expected launch counts are learned from the first call of a phase (which
    reduces the whole buffer at once); every rank runs the same schedule, so the collectives match.

    Parameters that received no gradient this step (attention head while attention is off) are reduced as zeros: the
    optimizer's 'touched' set, not the buffer content, decides what is updated."""

    def __init__(self, group=None, async_stream=True, buckets=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stream = None
        self.async_stream = async_stream
        self.bytes_reduced = 0
        self.bucket_prefixes = buckets or {}        # {id(net): [(prefix, ...), ...]}
        self.use_buckets = os.environ.get("DWC_DP_BUCKETS", "1") != "0"
        # optional bf16 wire format in bf16 mode (DWC_DP_BF16=1; fp32 mode always exchanges fp32): halves the bytes, but
        # measured without effect at 2 and 8 GPUs (20.71 vs 20.73 ms at 8) - the exchange is bound by the latency of the
        # collectives left at the end of a phase and by rank skew, not by bytes (profiles/r02c_scaling.md) - so off
        self.bf16_wire = os.environ.get("DWC_DP_BF16", "0") == "1"
        self._wire = {}
        self.coalesce = os.environ.get("DWC_DP_COALESCE", "1") != "0" and hasattr(dist, "_coalescing_manager")
        self.plans = {}                             # (id(net), key) -> {"need": {name: count}, "buckets": [...]}
        self.cur = None
        self.collectives = 0

    # ---- phase protocol ------------------------------------------------------------------------------
    def set_buckets(self, net, prefix_groups):
        self.bucket_prefixes[id(net)] = [tuple(g) for g in prefix_groups]

    def begin(self, net, key=None):
        """Start of a phase (after zero_grad): arm the weight-gradient hook."""
        if self.world == 1:
            return
        from . import ops
        plan = self.plans.get((id(net), key))
        self.cur = dict(net=net, key=key, plan=plan, counts={}, fired=[])
        ops.RT.wgrad_hook = self._on_wgrad

    def _ranges(self, flat, prefixes):
        names = [n for n in flat.names if n.startswith(tuple(prefixes))]
        if not names:
            return None
        s = min(flat.offsets[n] for n in names)
        e = max(flat.offsets[n] + flat.numels[n] for n in names)
        inside = [n for n in flat.names if s <= flat.offsets[n] < e]
        return (s, e, names) if set(inside) == set(names) else None        # contiguous run only

    def _on_wgrad(self, name):
        cur = self.cur
        if cur is None:
            return
        cur["counts"][name] = cur["counts"].get(name, 0) + 1
        plan = cur["plan"]
        if plan is None:
            return
        for i, b in enumerate(plan["buckets"]):
            if i in cur["fired"] or name not in b["need"]:
                continue
            if all(cur["counts"].get(n, 0) >= c for n, c in b["need"].items()):
                cur["fired"].append(i)
                self._reduce(cur["net"].ensure_flat().grad[b["start"]:b["end"]], also_side=True)

    def _reduce(self, g, also_side=False):
        """g: one range of the flat gradient buffer, or a list of ranges reduced as ONE coalesced collective (the
        end-of-phase leftovers: one launch latency instead of one per piece)."""
        pieces = list(g) if isinstance(g, (list, tuple)) else [g]
        g = pieces[0]
        self.bytes_reduced += sum(t.numel() for t in pieces) * 4
        self.collectives += 1
        if g.is_cuda and self.async_stream:
            if self.stream is None:
                # high priority: the collective's CTAs are placed ahead of the next compute kernel's when SMs free up
                # at a kernel boundary (they cannot share an SM with a one-CTA-per-SM persistent kernel)
                prio = int(os.environ.get("DWC_DP_PRIO", "-1"))
                self.stream = torch.cuda.Stream(device=g.device, priority=prio)
            cur = torch.cuda.current_stream(g.device)
            self.stream.wait_stream(cur)
            if also_side:
                from . import ops
                side = ops.RT.side_streams.get(g.device.index)
                if side is not None:
                    self.stream.wait_stream(side)
            with torch.cuda.stream(self.stream):
                self._all_reduce_many(pieces)
        else:
            self._all_reduce_many(pieces)

    def _all_reduce_many(self, pieces):
        if len(pieces) == 1 or self.bf16_wire or not self.coalesce or dist.get_backend(self.group) != "nccl":
            for t in pieces:
                self._all_reduce(t)
            return
        with dist._coalescing_manager(group=self.group, device=pieces[0].device, async_ops=False):
            for t in pieces:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _all_reduce(self, g):
        from . import ops
        if self.bf16_wire and g.is_cuda and g.dtype == torch.float32 and ops.RT.dtype == torch.bfloat16:
            key = (g.data_ptr(), g.numel())
            buf = self._wire.get(key)
            if buf is None:
                buf = self._wire[key] = torch.empty(g.numel(), dtype=torch.bfloat16, device=g.device)
            buf.copy_(g)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            g.copy_(buf)
        else:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)

    def __call__(self, net):
        """End of a phase (after backward, weight-gradient stream joined): reduce what the buckets did not cover and
        make the calling stream wait for all of it.  The 1/world factor is folded into the fused Adam kernel."""
        if self.world == 1:
            return
        from . import ops
        ops.RT.wgrad_hook = None
        flat = net.ensure_flat()
        g = flat.grad
        cur, self.cur = self.cur, None
        if cur is None or cur["net"] is not net:
            cur = dict(plan=None, counts={}, fired=[], key=None)
        plan = cur["plan"]
        if plan is None:
            self._reduce(g)
            # learn the schedule for the next call of this phase
            buckets = []
            for prefixes in (self.bucket_prefixes.get(id(net), []) if self.use_buckets else []):
                r = self._ranges(flat, prefixes)
                need = {n: c for n, c in cur["counts"].items() if n.startswith(tuple(prefixes))}
                if r is not None and need:
                    buckets.append(dict(start=r[0], end=r[1], need=need))
            buckets.sort(key=lambda b: b["start"])
            self.plans[(id(net), cur["key"])] = dict(buckets=buckets)
        else:
            done = sorted((plan["buckets"][i]["start"], plan["buckets"][i]["end"]) for i in cur["fired"])
            pos, rest = 0, []
            for s, e in done + [(flat.total, flat.total)]:
                if s - pos >= 64:                       # (shorter gaps are alignment padding: nothing to reduce)
                    rest.append(g[pos:s])
                pos = max(pos, e)
            if rest:
                self._reduce(rest)
        if g.is_cuda and self.async_stream and self.stream is not None:
            torch.cuda.current_stream(g.device).wait_stream(self.stream)


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
    # buckets whose gradients are final early in backward (see GradSync): the deepest discriminator layers of each scale
    # (backward starts there) and the decoder (done before the first encode's backward)
    nl = solver.dis.n_layer
    solver._dp_sync.set_buckets(solver.dis, [tuple("cnns_feat.%d.%d." % (sc, l) for l in (nl - 2, nl - 1))
                                             for sc in range(solver.dis.num_scales)])
    solver._dp_sync.set_buckets(solver.gen, [("dec.",)])
    solver.gen_opt.grad_scale = solver.dis_opt.grad_scale = 1.0 / solver._dp_sync.world
    return solver
