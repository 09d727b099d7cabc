"""Flat fp32 parameter / gradient storage and the fused Adam + EMA on top of it.

All parameters of a network live in ONE fp32 buffer (conv weights physically [Cout, KH, KW, Cin],
i.e. the channels_last storage of the logical OIHW tensor the state_dict exposes), gradients in
a second buffer of the same layout.  That gives: one gradient all-reduce per phase, one fused
optimizer launch per contiguous run of active parameters, tight "fusion groups" (the 16 style
heads as one [128, K] matrix, the two decoder heads as one 4-channel conv), and bf16 GEMM
operands packed straight from the master copy.

Replaces torch.optim.Adam (solver.py:65-68: lr 1e-4, betas (0.5, 0.999), eps 1e-8, coupled L2
1e-4; torch>=2 semantics: parameters without a gradient this step are skipped) and
utils.moving_average (utils.py:52-54).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib as L

ALIGN = 64  # elements (256 B)
_VERSION = [0]
_DRYRUN = [False]   # tests/dryrun_cpu.py only: bookkeeping without kernels


def _next_version():
    _VERSION[0] += 1
    return _VERSION[0]


class FlatParams:
    def __init__(self, module: nn.Module, fuse_groups: Sequence[Sequence[str]] = ()):
        self.module = module
        self.fuse_groups = [list(g) for g in fuse_groups]
        self.version = _next_version()
        self.touched = set()
        self.names: List[str] = []
        self.offsets: Dict[str, int] = {}
        self.numels: Dict[str, int] = {}
        self.data: Optional[torch.Tensor] = None
        self.grad: Optional[torch.Tensor] = None
        self.rebuild()

    # ---- layout -------------------------------------------------------------------------
    def _order(self, named):
        in_group = {}
        for g in self.fuse_groups:
            for nme in g:
                in_group[nme] = g
        order, seen = [], set()
        for nme in named:
            if nme in seen:
                continue
            grp = in_group.get(nme)
            if grp is None:
                order.append([nme])
                seen.add(nme)
            else:
                order.append([x for x in grp if x in named])
                seen.update(grp)
        return order

    def rebuild(self):
        """(Re)create the flat buffers from the module's current parameter values / device."""
        named = dict(self.module.named_parameters())
        if not named:
            return
        device = next(iter(named.values())).device
        groups = self._order(named)
        off = 0
        self.names, self.offsets, self.numels = [], {}, {}
        for grp in groups:
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            for nme in grp:
                self.names.append(nme)
                self.offsets[nme] = off
                self.numels[nme] = named[nme].numel()
                off += named[nme].numel()
        total = (off + ALIGN - 1) // ALIGN * ALIGN
        data = torch.zeros(total, dtype=torch.float32, device=device)
        grad = torch.zeros(total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for nme in self.names:
                p = named[nme]
                dview = self._view(data, nme, p.shape)
                dview.copy_(p.detach().to(torch.float32))
                p.data = dview
                p.grad = self._view(grad, nme, p.shape) if p.requires_grad else None
        self.data, self.grad = data, grad
        self.total = total
        self.version = _next_version()
        self.touched = set()

    def _view(self, flat, nme, shape):
        o, n = self.offsets[nme], self.numels[nme]
        seg = flat[o:o + n]
        if len(shape) == 4:
            co, ci, kh, kw = shape
            return seg.view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return seg.view(shape)

    def ok(self) -> bool:
        """True while every parameter still aliases the flat buffer (deepcopy / .to() break that)."""
        if self.data is None:
            return False
        named = dict(self.module.named_parameters())
        if set(named) != set(self.names):
            return False
        base = self.data.data_ptr()
        for nme in (self.names[0], self.names[-1]):
            if named[nme].data_ptr() != base + self.offsets[nme] * 4:
                return False
        return named[self.names[0]].device == self.data.device

    def ensure(self):
        if not self.ok():
            self.rebuild()

    # ---- raw views used by the kernels ----------------------------------------------------
    def raw(self, nme, count=None):
        o = self.offsets[nme]
        return self.data[o:o + (count or self.numels[nme])]

    def raw_grad(self, nme, count=None, touch=True):
        o = self.offsets[nme]
        if touch:
            self.touched.add(nme)
        return self.grad[o:o + (count or self.numels[nme])]

    def touch(self, *names):
        self.touched.update(names)

    def zero_grad(self):
        if self.grad is not None:
            self.grad.zero_()
        self.touched = set()
        named = dict(self.module.named_parameters())
        for nme in self.names:
            p = named[nme]
            if p.requires_grad and p.grad is None:
                p.grad = self._view(self.grad, nme, p.shape)

    def bump(self):
        self.version = _next_version()


class FusedAdam(torch.optim.Optimizer):
    """Adam with coupled L2 over a FlatParams buffer: one kernel launch per contiguous run of
    parameters that (a) received a gradient this step and (b) share a step count."""

    def __init__(self, flat: FlatParams, lr, betas, weight_decay, eps=1e-8):
        self.flat = flat
        named = dict(flat.module.named_parameters())
        params = [named[n] for n in flat.names if named[n].requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._names = [n for n in flat.names if named[n].requires_grad]
        self.steps = {n: 0 for n in self._names}
        self.m = None
        self.v = None
        self._hyper_dev = None
        self.captured = None
        self.grad_scale = 1.0

    def _ensure_state(self):
        f = self.flat
        if self.m is None or self.m.numel() != f.total or self.m.device != f.data.device:
            self.m = torch.zeros_like(f.data)
            self.v = torch.zeros_like(f.data)

    def zero_grad(self, set_to_none: bool = False):  # noqa: D401  (flat buffers are never dropped)
        self.flat.zero_grad()

    def active_ranges(self):
        """[(start, end, step_count)] merged over adjacent touched parameters with equal step."""
        f = self.flat
        out = []
        trainable = set(self._names)
        prev_merged = False          # was the previous parameter (in buffer order) part of the last range?
        for n in f.names:
            if n not in trainable or n not in f.touched:
                prev_merged = False
                continue
            s, e, st = f.offsets[n], f.offsets[n] + f.numels[n], self.steps[n] + 1
            if out and prev_merged and out[-1][2] == st:
                out[-1] = (out[-1][0], e, st)      # the gap is alignment padding only (zeros stay zeros)
            else:
                out.append((s, e, st))
            prev_merged = True
        return out

    MAX_RANGES = 64

    def _hyper_row(self, st):
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        return [g["lr"], b1, b2, g["eps"], g["weight_decay"], 1 - b1 ** st, 1 - b2 ** st, self.grad_scale]

    def _launch(self, ranges, hd):
        f = self.flat
        lib = L.lib()
        for i, (s, e, _) in enumerate(ranges):
            L.check(lib.dwc_adam_step(L.ptr(f.data[s:]), L.ptr(f.grad[s:]), L.ptr(self.m[s:]), L.ptr(self.v[s:]),
                                      L.i64(e - s), None, L.ptr(hd[i]), L.stream()), "adam")

    def graph_buffers(self):
        """Persistent hyper-parameter buffers of the CUDA-graph path (allocate before capturing)."""
        if self._hyper_dev is None or self._hyper_dev.device != self.flat.data.device:
            self._hyper_dev = torch.zeros(self.MAX_RANGES, 8, dtype=torch.float32, device=self.flat.data.device)
            self._hyper_ring = [torch.zeros(self.MAX_RANGES, 8, dtype=torch.float32).pin_memory() for _ in range(4)]
            self._hyper_events = [None] * 4
            self._hyper_slot = 0
        self._ensure_state()
        return self._hyper_dev

    @torch.no_grad()
    def step(self, closure=None):
        f = self.flat
        if not f.data.is_cuda and not _DRYRUN[0]:
            raise RuntimeError("FusedAdam runs on CUDA only (no CPU fallback)")
        ranges = self.active_ranges()
        if not ranges:
            return
        if _DRYRUN[0]:
            print("adam ranges:", [(a, b - a, st) for a, b, st in ranges][:8])
            for n in self._names:
                if n in f.touched:
                    self.steps[n] += 1
            f.bump()
            return
        if torch.cuda.is_current_stream_capturing():
            # CUDA-graph capture: the kernels read their hyper-parameters from the persistent device rows that
            # graph_prepare() refreshes before every replay; step counters advance in graph_finish().
            if self._hyper_dev is None or len(ranges) > self.MAX_RANGES:
                raise RuntimeError("FusedAdam: call graph_buffers() before capturing")
            self.captured = dict(ranges=ranges, touched=[n for n in self._names if n in f.touched],
                                 rep=[next(n for n in self._names if n in f.touched and s <= f.offsets[n] < e)
                                      for s, e, _ in ranges])
            self._launch(ranges, self._hyper_dev)
            f.bump()
            return
        self._ensure_state()
        hyper = torch.empty(len(ranges), 8, dtype=torch.float32, pin_memory=True)
        for i, (_, _, st) in enumerate(ranges):
            hyper[i] = torch.tensor(self._hyper_row(st))
        hd = hyper.to(f.data.device, non_blocking=True)
        self._keep = (hyper, hd)
        self._launch(ranges, hd)
        for n in self._names:
            if n in f.touched:
                self.steps[n] += 1
        f.bump()

    def graph_prepare(self, rec):
        """Before a replay of a graph captured with `rec = self.captured`: upload this step's hyper-parameters."""
        slot = self._hyper_slot
        self._hyper_slot = (slot + 1) % len(self._hyper_ring)
        ev = self._hyper_events[slot]
        if ev is not None:
            ev.synchronize()
        host = self._hyper_ring[slot]
        for i, n in enumerate(rec["rep"]):
            host[i] = torch.tensor(self._hyper_row(self.steps[n] + 1))
        self._hyper_dev.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._hyper_events[slot] = ev

    def graph_finish(self, rec):
        for n in rec["touched"]:
            self.steps[n] += 1
        self.flat.bump()

    def _param_order(self):
        """Names of the trainable parameters in ``module.parameters()`` order - the order torch.optim.Adam (the
        reference's optimizer, solver.py:63-68) numbers its state entries in.  The flat buffer's own order differs
        (fusion groups are contiguous), so checkpoints are emitted / consumed through this map."""
        return [n for n, p in self.flat.module.named_parameters() if p.requires_grad]

    def state_dict(self):
        """torch.optim.Adam-shaped state (what Solver.save writes, solver.py:413): entry i belongs to the i-th
        trainable parameter of ``module.parameters()``."""
        named = dict(self.flat.module.named_parameters())
        order = self._param_order()
        state = {}
        if self.m is not None:
            for i, n in enumerate(order):
                if self.steps[n] == 0:
                    continue
                shape = named[n].shape
                state[i] = dict(step=torch.tensor(float(self.steps[n])),
                                exp_avg=self.flat._view(self.m, n, shape).clone().contiguous(),
                                exp_avg_sq=self.flat._view(self.v, n, shape).clone().contiguous())
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(order)))
        return dict(state=state, param_groups=[group])

    def load_state_dict(self, sd):
        self._ensure_state()
        named = dict(self.flat.module.named_parameters())
        order = self._param_order()
        for i, st in sd.get("state", {}).items():
            n = order[int(i)]
            self.steps[n] = int(st["step"])
            self.flat._view(self.m, n, named[n].shape).copy_(st["exp_avg"])
            self.flat._view(self.v, n, named[n].shape).copy_(st["exp_avg_sq"])
        if sd.get("param_groups"):
            for k in ("lr", "betas", "eps", "weight_decay", "initial_lr"):
                if k in sd["param_groups"][0]:
                    self.param_groups[0][k] = sd["param_groups"][0][k]


@torch.no_grad()
def ema_update(flat: FlatParams, flat_avg: FlatParams, beta: float = 0.999):
    """p_avg = lerp(p, p_avg, beta) over whole flat buffers (utils.py:52-54)."""
    flat.ensure()
    flat_avg.ensure()
    if flat.total != flat_avg.total:
        raise RuntimeError("EMA copy has a different parameter layout")
    L.check(L.lib().dwc_ema_step(L.ptr(flat.data), L.ptr(flat_avg.data), L.i64(flat.total), beta, L.stream()), "ema")
    flat_avg.bump()
