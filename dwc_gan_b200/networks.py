"""B200-native generator and discriminator behind the reference's nn.Module API.

Same constructor signatures, sub-module names and state_dict keys/shapes as the reference
(networks/networks_v2.py: AdaINGen_v2, StyleEncoder, Decoder, TxtEncoder; networks/networks.py:
MsImageDis, ContentEncoder, ResBlocks, ResBlock, MLP, Conv2dBlock, LinearBlock,
AdaptiveInstanceNorm2d, LayerNorm), so reference checkpoints load both ways - but every
forward/backward arithmetic step is one of the CUDA kernels behind include/dwc_b200.h
(ops.py).  Parameters are created by the same torch.nn constructors in the same order as the
reference so that a given seed yields bit-identical initial weights.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .flat import FlatParams
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, NORM_ADAIN, NORM_IN, NORM_LN, NORM_NONE, RT
from .plan import HB

_ACT = {"relu": ACT_RELU, "lrelu": ACT_LRELU, "none": ACT_NONE}


def _hb_from_tensor(x: torch.Tensor) -> HB:
    """Logical NCHW tensor -> halo-0 NHWC buffer in the compute dtype (zero-copy when it already is one)."""
    n, c, h, w = x.shape
    if x.dtype == RT.dtype and x.is_contiguous(memory_format=torch.channels_last):
        return HB(x.permute(0, 2, 3, 1), n, h, w, c, 0, 0)
    if c % 8 != 0:
        raise RuntimeError("feature maps must have a multiple of 8 channels")
    xc = ops._cast(x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1), RT.dtype) \
        if x.dtype != RT.dtype else x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)
    return HB(xc, n, h, w, c, 0, 0)


def _hb_to_tensor(hb: HB) -> torch.Tensor:
    assert hb.halo == 0 and hb.layout == 0
    return hb.t.permute(0, 3, 1, 2)


class _FlatOwner(nn.Module):
    """Mixin for top-level networks: owns the FlatParams buffer and keeps it valid."""

    _fuse_groups: List[List[str]] = []

    def _init_flat(self):
        object.__setattr__(self, "_flat", FlatParams(self, self._fuse_groups))
        for m in self.modules():
            if isinstance(m, (Conv2dBlock, _LinearHolder, LayerNorm, TxtEncoder)):
                object.__setattr__(m, "_owner", self)
            if isinstance(m, Conv2dBlock):
                m._packed = {}

    @property
    def flat(self) -> FlatParams:
        f = self.__dict__.get("_flat")
        if f is None:
            self._init_flat()
            f = self.__dict__["_flat"]
        return f

    def ensure_flat(self):
        f = self.flat
        if not f.ok():
            f.rebuild()
            for m in self.modules():
                if isinstance(m, (Conv2dBlock, _LinearHolder, LayerNorm, TxtEncoder)):
                    object.__setattr__(m, "_owner", self)
        return f

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        if "_flat" in self.__dict__:
            self.__dict__["_flat"].bump()
        return r

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_flat", "_name_cache"):
                continue
            object.__setattr__(new, k, copy.deepcopy(v, memo))
        new._init_flat()
        return new

    def param_name_of(self, mod: nn.Module, leaf: str) -> str:
        cache = self.__dict__.setdefault("_name_cache", {})
        key = (id(mod), leaf)
        if key not in cache:
            for nme, m in self.named_modules():
                if m is mod:
                    cache[key] = (nme + "." if nme else "") + leaf
                    break
            else:
                raise KeyError("module not found in owner")
        return cache[key]


# ---------------------------------------------------------------------------------------------
# basic blocks
# ---------------------------------------------------------------------------------------------

class AdaptiveInstanceNorm2d(nn.Module):
    """networks/networks.py:693-722.  weight/bias ([B*C]) are assigned by assign_adain_params; the two
    buffers are dummies that only exist for state_dict compatibility."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = None
        self.bias = None
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def __repr__(self):
        return self.__class__.__name__ + "(" + str(self.num_features) + ")"


class LayerNorm(nn.Module):
    """MUNIT LayerNorm (networks/networks.py:725-752): per-sample mean / unbiased std over C*H*W,
    (x-mean)/(std+eps), per-channel gamma (init U(0,1)) and beta."""

    def __init__(self, num_features, eps=1e-5, affine=True):
        super().__init__()
        self.num_features, self.affine, self.eps = num_features, affine, eps
        if self.affine:
            self.gamma = nn.Parameter(torch.Tensor(num_features).uniform_())
            self.beta = nn.Parameter(torch.zeros(num_features))

    def grad_buffers(self):
        owner = self.__dict__["_owner"]
        f = owner.flat
        return (f.raw_grad(owner.param_name_of(self, "gamma")), f.raw_grad(owner.param_name_of(self, "beta")))


class Conv2dBlock(nn.Module):
    """pad -> conv -> norm -> activation (networks/networks.py:524-585) on haloed NHWC buffers."""

    def __init__(self, input_dim, output_dim, kernel_size, stride, padding=0, norm="none", activation="relu",
                 pad_type="zero"):
        super().__init__()
        if pad_type != "reflect":
            raise NotImplementedError("only reflect padding is on the B200 hot path (configs/celeba_faces.yaml:53,68)")
        self.k, self.stride, self.padding = kernel_size, stride, padding
        self.cin, self.cout = input_dim, output_dim
        assert (stride == 1 and kernel_size == 2 * padding + 1) or (stride == 2 and kernel_size == 4 and padding == 1)
        self.norm_kind = {"none": NORM_NONE, "in": NORM_IN, "adain": NORM_ADAIN, "ln": NORM_LN}[norm]
        # same construction order as the reference: norm (LayerNorm draws gamma) before conv
        if norm == "in":
            self.norm = nn.InstanceNorm2d(output_dim)
        elif norm == "ln":
            self.norm = LayerNorm(output_dim)
        elif norm == "adain":
            self.norm = AdaptiveInstanceNorm2d(output_dim)
        elif norm == "none":
            self.norm = None
        else:
            raise NotImplementedError("norm %s is not on the hot path" % norm)
        self.act_name = activation
        if activation not in ("relu", "lrelu", "none", "tanh", "sigmoid"):
            raise NotImplementedError("activation %s is not on the hot path" % activation)
        self.activation = None     # kept for attribute compatibility; fused into the post pass
        self.conv = nn.Conv2d(input_dim, output_dim, kernel_size, stride, bias=True)
        self._packed = {}
        object.__setattr__(self, "extra_cols", None)      # fused sibling (decoder heads), set by Decoder

    # ---- parameter plumbing -------------------------------------------------------------
    def _names(self):
        owner = self.__dict__["_owner"]
        return owner, owner.param_name_of(self, "conv.weight"), owner.param_name_of(self, "conv.bias")

    @property
    def weight_param(self):
        return self.conv.weight

    def total_cout(self):
        return self.cout + (self.extra_cols.cout if self.extra_cols is not None else 0)

    def _raw_weight(self):
        owner, wn, bn = self._names()
        f = owner.flat
        tot = self.total_cout()
        return f, f.raw(wn, tot * self.k * self.k * self.cin), f.raw(bn, tot)

    def bias_f32(self):
        return self._raw_weight()[2]

    @staticmethod
    def _pack_buffer(hit, shape, dtype, device):
        """Packed operands are rewritten IN PLACE when the weights change: captured CUDA graphs of both phases keep
        pointing at the same buffers, so a network is packed once per optimizer step (by whichever phase first uses
        it after the step) and every later user - eager or replayed - reads current values."""
        if hit is not None and tuple(hit[1].shape) == tuple(shape) and hit[1].dtype == dtype and hit[1].device == device:
            return hit[1]
        return torch.empty(shape, dtype=dtype, device=device)

    def _pack(self, f, key, w, tot, mode, out, rows_p):
        """(Re)pack one operand.  The first request registers it with the network's flat buffer; once the set of
        operands is stable, the first stale operand of a step repacks ALL of them in one launch
        (dwc_pack_weights_batch) instead of one small kernel in front of every convolution."""
        reg = f.__dict__.get("_pack_reg")
        if reg is None or reg["base"] != f.data.data_ptr():          # new flat buffer: every pointer is stale
            reg = f.__dict__["_pack_reg"] = {"entries": {}, "table": None, "dirty": True, "base": f.data.data_ptr()}
        ek = (id(self), key)
        ent = reg["entries"].get(ek)
        if ent is None or ent["out"] is not out or ent["w_ptr"] != w.data_ptr():
            reg["entries"][ek] = dict(layer=self, key=key, out=out, w_ptr=w.data_ptr(), tot=tot, mode=mode,
                                      rows_p=rows_p)
            reg["dirty"] = True
        capturing = torch.cuda.is_current_stream_capturing() if out.is_cuda else False
        if ops.RT.batch_pack and not (reg["dirty"] and capturing):
            if reg["dirty"]:
                ents = list(reg["entries"].values())
                arr = (L.PackEntry * len(ents))()
                for i, e in enumerate(ents):
                    o, ly = e["out"], e["layer"]
                    arr[i] = L.PackEntry(e["w_ptr"], o.data_ptr(), e["tot"], ly.k, ly.k, ly.cin, e["mode"], e["rows_p"],
                                         L.dt(o), 0, o.numel())
                host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
                reg["table"] = host.to(out.device)
                reg["order"] = ents
                reg["dirty"] = False
                reg["calls"] = 0
            reg["calls"] += 1
            # the batch is only worth it (and only complete) once a whole step has registered its operands
            if reg["calls"] > 1 or len(reg["order"]) > 8:
                ops._call("dwc_pack_weights_batch", L.ptr(reg["table"]), len(reg["order"]), L.stream())
                for e in reg["order"]:
                    e["layer"]._packed[e["key"]] = (f.version, e["out"])
                return self._packed[key]
        ops._call("dwc_pack_weights", L.ptr(w), tot, self.k, self.k, self.cin, mode, L.ptr(out), L.dt(out), rows_p,
                  L.stream())
        hit = (f.version, out)
        self._packed[key] = hit
        return hit

    def packed_fwd(self, dtype):
        f, w, _ = self._raw_weight()
        tot = self.total_cout()
        rows_p = tot if tot % 64 == 0 else ((tot + 15) // 16 * 16 if tot <= 16 else (tot + 63) // 64 * 64)
        if dtype == torch.float32 and rows_p == tot:
            return w.view(tot, -1), rows_p
        key = ("f", dtype)
        hit = self._packed.get(key)
        if hit is None or hit[0] != f.version:
            out = self._pack_buffer(hit, (rows_p, self.k * self.k * self.cin), dtype, w.device)
            hit = self._pack(f, key, w, tot, 0, out, rows_p)
        return hit[1], rows_p

    def packed_dgrad(self, dtype, pad_rows=False):
        """(GEMM operand of the data gradient, its padded row count): rows = input channels (16 when a few-channel
        image gradient runs on the tensor cores)."""
        f, w, _ = self._raw_weight()
        tot = self.total_cout()
        rows_p = 16 if (pad_rows and self.cin <= 16) else self.cin
        key = ("d", dtype, rows_p)
        hit = self._packed.get(key)
        if hit is None or hit[0] != f.version:
            if self.stride == 1:
                out = self._pack_buffer(hit, (rows_p, self.k * self.k * tot), dtype, w.device)
                mode = 1
            else:
                out = self._pack_buffer(hit, (4, rows_p, 4 * tot), dtype, w.device)
                mode = 2
            hit = self._pack(f, key, w, tot, mode, out, rows_p)
        return hit[1], rows_p

    def packed_rows(self, dtype, mode):
        """Weights over a row-im2col operand: mode 3 forward [cout][k*64], mode 4 data gradient [cin][k*64]."""
        f, w, _ = self._raw_weight()
        tot = self.total_cout()
        key = ("r", dtype, mode)
        hit = self._packed.get(key)
        if hit is None or hit[0] != f.version:
            rows = tot if mode == 3 else self.cin
            out = self._pack_buffer(hit, (rows, self.k * 64), dtype, w.device)
            hit = self._pack(f, key, w, tot, mode, out, rows)
        return hit[1]

    def grad_buffers(self):
        owner, wn, bn = self._names()
        f = owner.flat
        tot = self.total_cout()
        gw = f.raw_grad(wn, tot * self.k * self.k * self.cin)
        gb = f.raw_grad(bn, tot)
        return gw, gb

    def bias_grad_needed(self):
        """A bias that feeds InstanceNorm / AdaIN cancels in the mean subtraction: its gradient is exactly zero (the
        reference computes round-off there); the parameter still counts as 'touched' for Adam's bookkeeping."""
        return self.norm_kind not in (NORM_IN, NORM_ADAIN)

    def touch_params(self):
        owner, wn, bn = self._names()
        owner.flat.touch(wn, bn)

    # ---- compute ---------------------------------------------------------------------------
    def in_layout(self):
        return 1 if self.stride == 2 else 0

    def run(self, xp: HB, out_halo=0, out_layout=0, res: Optional[HB] = None, raw=False, skip_box=None,
            res_box=None) -> HB:
        """xp: reflect-haloed input (halo == padding, parity planes if stride 2).  skip_box / res_box: see ResBlock."""
        assert xp.halo == self.padding and xp.layout == self.in_layout() and xp.c == self.cin
        epi = (_ACT[self.act_name], out_halo, out_layout) if not raw and res is None and self.norm_kind == NORM_NONE \
            else None
        y = _ConvProxy.conv(xp, self, skip_box, epi)
        if raw:
            return y
        return self.finish(y, out_halo, out_layout, res, res_box)

    def run_first(self, img, rows_t, pool, out_halo=0, out_layout=0, raw=False) -> HB:
        """First layer of a network: img NCHW fp32 (3 channels), rows_t = ops.image_rows(img, pool, self)."""
        epi = (_ACT[self.act_name], out_halo, out_layout) if not raw and self.norm_kind == NORM_NONE else None
        y = ops.first_conv(img, rows_t, self, pool, epi)
        return y if raw else self.finish(y, out_halo, out_layout, None)

    def finish(self, y: HB, out_halo=0, out_layout=0, res: Optional[HB] = None, res_box=None) -> HB:
        n = y.n
        nw = nb = None
        ln = None
        if self.norm_kind == NORM_ADAIN:
            assert self.norm.weight is not None and self.norm.bias is not None, \
                "Please assign weight and bias before calling AdaIN!"
            nw = self.norm.weight.view(n, self.cout)
            nb = self.norm.bias.view(n, self.cout)
        elif self.norm_kind == NORM_LN:
            owner = self.__dict__["_owner"]
            f = owner.flat
            nw = f.raw(owner.param_name_of(self.norm, "gamma"))
            nb = f.raw(owner.param_name_of(self.norm, "beta"))
            ln = self.norm
        eps = self.norm.eps if self.norm is not None else 1e-5
        return ops.post(y, self.norm_kind, _ACT[self.act_name], nw, nb, res, out_halo, out_layout, ln, eps,
                        anchor=ln.gamma if ln is not None else None, skip_box=res_box)

    def forward(self, x):
        """API-compatible entry: logical NCHW tensor in, logical NCHW tensor (compute dtype) out."""
        self.__dict__["_owner"].ensure_flat()
        if self.act_name in ("tanh", "sigmoid"):
            raise NotImplementedError("decoder heads run fused through Decoder.forward")
        if x.shape[1] % 8 != 0:
            return _hb_to_tensor(self.run_first(x, ops.image_rows(x, 1, self), 1))
        xp = ops.post(_hb_from_tensor(x), out_halo=self.padding, out_layout=self.in_layout())
        return _hb_to_tensor(self.run(xp))


class _ConvProxy:
    @staticmethod
    def conv(xp: HB, layer: Conv2dBlock, skip_box=None, epi=None) -> HB:
        return ops.conv(xp, layer, skip_box, epi)


class ResBlock(nn.Module):
    """networks/networks.py:509-522: [conv3+norm+act, conv3+norm] + residual."""

    def __init__(self, dim, norm="in", activation="relu", pad_type="zero"):
        super().__init__()
        self.model = nn.Sequential(
            Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation=activation, pad_type=pad_type),
            Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation="none", pad_type=pad_type))

    def run(self, p0: HB, out_halo) -> HB:
        # The block input receives two gradients: the skip connection's (produced by the backward of the second
        # conv's norm pass) and the first conv's data gradient.  The box carries the former to the latter's dgrad
        # kernel, which adds onto it in its epilogue (backward order guarantees it is there first).
        box = {} if ops.RT.fuse_skip_grad and p0.t.requires_grad else None
        p1 = self.model[0].run(p0, out_halo=1, skip_box=box)
        return self.model[1].run(p1, out_halo=out_halo, res=p0, res_box=box)


class ResBlocks(nn.Module):
    def __init__(self, num_blocks, dim, norm="in", activation="relu", pad_type="zero"):
        super().__init__()
        self.model = nn.Sequential(*[ResBlock(dim, norm=norm, activation=activation, pad_type=pad_type)
                                     for _ in range(num_blocks)])

    def run(self, p0: HB, out_halo) -> HB:
        nblk = len(self.model)
        for i, blk in enumerate(self.model):
            p0 = blk.run(p0, out_halo if i == nblk - 1 else 1)
        return p0


class _LinearHolder(nn.Module):
    """Marker base class for modules that run nn.Linear parameters through dwc_sgemm."""

    def _lin(self, x, lin: nn.Linear, act, rows=None, prefix=None):
        owner = self.__dict__["_owner"]
        f = owner.flat
        wn = owner.param_name_of(lin, "weight")
        bn = owner.param_name_of(lin, "bias") if lin.bias is not None else None
        n_out = rows or lin.out_features
        k = lin.in_features
        w = f.raw(wn, n_out * k).view(n_out, k)
        b = f.raw(bn, n_out) if bn is not None else None
        extra = prefix or []

        def grads():
            gw = f.raw_grad(wn, n_out * k).view(n_out, k)
            gb = f.raw_grad(bn, n_out) if bn is not None else None
            for nme in extra:
                f.touch(nme)
            return gw, gb

        return ops.linear(x, w, b, act, grads, lin.weight)


class LinearBlock(_LinearHolder):
    """networks/networks.py:587-634 (norm none; relu or none)."""

    def __init__(self, input_dim, output_dim, norm="none", activation="relu"):
        super().__init__()
        if norm != "none" or activation not in ("relu", "none"):
            raise NotImplementedError("LinearBlock(norm=%s, activation=%s) is not on the hot path" % (norm, activation))
        self.fc = nn.Linear(input_dim, output_dim, bias=True)
        self.norm = None
        self.activation = None
        self.act = 1 if activation == "relu" else 0

    def forward(self, x):
        return self._lin(x, self.fc, self.act)


class MLP(nn.Module):
    """networks/networks.py:491-503."""

    def __init__(self, input_dim, output_dim, dim, n_blk, norm="none", activ="relu"):
        super().__init__()
        model = [LinearBlock(input_dim, dim, norm=norm, activation=activ)]
        for _ in range(n_blk - 2):
            model += [LinearBlock(dim, dim, norm=norm, activation=activ)]
        model += [LinearBlock(dim, output_dim, norm="none", activation="none")]
        self.model = nn.Sequential(*model)

    def forward(self, x):
        return self.model(x.reshape(x.size(0), -1).float())


# ---------------------------------------------------------------------------------------------
# generator parts
# ---------------------------------------------------------------------------------------------

class StyleEncoder(_LinearHolder):
    """networks/networks_v2.py:98-141."""

    def __init__(self, n_downsample, input_dim, dim, norm, activ, pad_type, c_dim, num_class, use_map=False):
        super().__init__()
        self.num_class, self.use_map, self.c_dim = num_class, use_map, c_dim
        model = [Conv2dBlock(input_dim, dim, 7, 1, 3, norm=norm, activation=activ, pad_type=pad_type)]
        for _ in range(2):
            model += [Conv2dBlock(dim, 2 * dim, 4, 2, 1, norm=norm, activation=activ, pad_type=pad_type)]
            dim *= 2
        for _ in range(n_downsample - 2):
            model += [Conv2dBlock(dim, dim, 4, 2, 1, norm=norm, activation=activ, pad_type=pad_type)]
        model += [nn.AdaptiveAvgPool2d(1)]
        self.model = nn.Sequential(*model)
        if self.use_map:
            self.mapping = nn.Sequential(nn.Linear(dim, dim), nn.ReLU(inplace=True), nn.Dropout(p=0.1),
                                         nn.Linear(dim, dim), nn.ReLU(inplace=True))
        self.fcs = nn.ModuleList()
        self.fcvars = nn.ModuleList()
        for _ in range(self.num_class):
            self.fcs.append(nn.Linear(dim, c_dim))
            self.fcvars.append(nn.Linear(dim, c_dim))
        self.output_dim = dim
        assert activ == "relu"

    def run(self, img, rows_t):
        convs = [m for m in self.model if isinstance(m, Conv2dBlock)]
        h = convs[0].run_first(img, rows_t, 1, out_halo=1, out_layout=1)
        for i, cv in enumerate(convs[1:], start=1):
            last = i == len(convs) - 1
            if last:
                y = cv.run(h, raw=True)
                feats = ops.relu_gap(y)
            else:
                h = cv.run(h, out_halo=1, out_layout=1)
        if self.use_map:
            feats = self._lin(feats, self.mapping[0], 1)
            p = self.mapping[2].p
            if self.training and p > 0:
                mask = (torch.rand_like(feats) >= p).float() / (1.0 - p)
                feats = ops.MulFn.apply(feats, mask)
            feats = self._lin(feats, self.mapping[3], 1)
        # all 16 heads as one [2*num_class*c_dim, dim] GEMM (fusion group in the flat buffer)
        owner = self.__dict__["_owner"]
        names = []
        for grp in (self.fcs, self.fcvars):
            for lin in grp:
                names += [owner.param_name_of(lin, "weight"), owner.param_name_of(lin, "bias")]
        out = self._lin(feats, self.fcs[0], 0, rows=2 * self.num_class * self.c_dim, prefix=names)
        half = self.num_class * self.c_dim
        return out[:, :half], out[:, half:]

    def forward(self, x):
        self.__dict__["_owner"].ensure_flat()
        mu, lv = self.run(x, ops.image_rows(x, 1, self.model[0]))
        return list(mu.split(self.c_dim, 1)), list(lv.split(self.c_dim, 1))


class ContentEncoder(nn.Module):
    """networks/networks.py:428-446."""

    def __init__(self, n_downsample, n_res, input_dim, dim, norm, activ, pad_type):
        super().__init__()
        model = [Conv2dBlock(input_dim, dim, 7, 1, 3, norm=norm, activation=activ, pad_type=pad_type)]
        prev = dim
        for _ in range(n_downsample):
            dim = min(dim * 2, 256)
            model += [Conv2dBlock(prev, dim, 4, 2, 1, norm=norm, activation=activ, pad_type=pad_type)]
            prev = dim
        model += [ResBlocks(n_res, dim, norm=norm, activation=activ, pad_type=pad_type)]
        self.model = nn.Sequential(*model)
        self.output_dim = dim

    def run(self, img, rows_t) -> HB:
        mods = list(self.model)
        h = None
        for i, m in enumerate(mods[:-1]):
            nxt = mods[i + 1]
            layout = 1 if isinstance(nxt, Conv2dBlock) and nxt.stride == 2 else 0
            if i == 0:
                h = m.run_first(img, rows_t, 1, out_halo=1, out_layout=layout)
            else:
                h = m.run(h, out_halo=1, out_layout=layout)
        return mods[-1].run(h, out_halo=0)

    def forward(self, x):
        self.model[0].__dict__["_owner"].ensure_flat()
        return _hb_to_tensor(self.run(x, ops.image_rows(x, 1, self.model[0])))


class Decoder(nn.Module):
    """networks/networks_v2.py:144-169.  The two 7x7 heads run as one 4-channel convolution."""

    def __init__(self, n_upsample, n_res, dim, output_dim, res_norm="adain", activ="relu", pad_type="zero",
                 use_attention=False):
        super().__init__()
        self.use_attention = use_attention
        model = [ResBlocks(n_res, dim, res_norm, activ, pad_type=pad_type)]
        for _ in range(n_upsample):
            model += [nn.Upsample(scale_factor=2, mode="bilinear"),
                      Conv2dBlock(dim, dim // 2, 5, 1, 2, norm="ln", activation=activ, pad_type=pad_type)]
            dim //= 2
        self.model = nn.Sequential(*model)
        self.image_content = Conv2dBlock(dim, output_dim, 7, 1, 3, norm="none", activation="tanh", pad_type=pad_type)
        self.image_attention = Conv2dBlock(dim, 1, 7, 1, 3, norm="none", activation="sigmoid", pad_type=pad_type)
        object.__setattr__(self.image_content, "extra_cols", self.image_attention)   # fused sibling, not a child

    def run(self, content: HB):
        mods = list(self.model)
        h = ops.post(content, out_halo=1)                       # reflect pad for the first 3x3 conv
        h = mods[0].run(h, out_halo=0)
        convs = [m for m in mods[1:] if isinstance(m, Conv2dBlock)]
        for i, cv in enumerate(convs):
            h = ops.upsample_pad(h, 2)
            h = cv.run(h, out_halo=3 if i == len(convs) - 1 else 0)
        if not convs:
            h = ops.post(h, out_halo=3)
        # both heads as one 4-channel conv: [N, H, W, 4] = content(3) | attention(1)
        return ops.heads_conv(h, self.image_content, self.image_attention.touch_params)

    def forward(self, x):
        self.image_content.__dict__["_owner"].ensure_flat()
        img, att = self.run(_hb_from_tensor(x))
        return img, (att if self.use_attention else None)


class TxtEncoder(_LinearHolder):
    """networks/networks_v2.py:171-254: embedding + style concat -> 2-layer bi-LSTM -> 16 linear heads.
    The sort by length of the reference is a no-op on the result and is skipped;
    per-sample lengths are handled inside the LSTM step kernels."""

    def __init__(self, vocab, embed_dim=512, hidden_size=512, c_dim=8, num_class=8, num_layers=1, dropout_in=0.1,
                 dropout_out=0.1, bidirectional=True, pretrained_embed=None):
        super().__init__()
        if not bidirectional:
            raise NotImplementedError("only the bidirectional text encoder is on the hot path")
        self.vocab, self.embed_dim, self.hidden_size = vocab, embed_dim, hidden_size
        self.num_layers, self.dropout_in, self.dropout_out = num_layers, dropout_in, dropout_out
        self.bidirectional, self.num_class, self.c_dim = bidirectional, num_class, c_dim
        self.style_dim = c_dim * num_class
        self.embed_tokens = nn.Embedding(vocab.size, embed_dim, vocab.padding_idx)
        if pretrained_embed is not None:
            wm = np.zeros((vocab.size, embed_dim))
            for i, word in enumerate(vocab.itos):
                try:
                    wm[i] = pretrained_embed[word]
                except KeyError:
                    wm[i] = np.random.normal(scale=0.6, size=(embed_dim,))
            self.embed_tokens.load_state_dict({"weight": torch.from_numpy(wm)})
            self.embed_tokens.weight.requires_grad = False
        self.lstm = nn.LSTM(input_size=embed_dim + self.style_dim, hidden_size=hidden_size, num_layers=num_layers,
                            dropout=self.dropout_out if num_layers > 1 else 0., bidirectional=bidirectional)
        hidden_dim = hidden_size * num_layers * 4
        self.fcs = nn.ModuleList()
        self.fcvars = nn.ModuleList()
        for _ in range(self.num_class):
            self.fcs.append(nn.Linear(hidden_dim, c_dim))
            self.fcvars.append(nn.Linear(hidden_dim, c_dim))

    def forward(self, style_ord, src_tokens, src_lengths):
        from .text import txt_encode
        owner = self.__dict__["_owner"]
        owner.ensure_flat()
        out = txt_encode(self, style_ord, src_tokens, src_lengths)           # [B, 4*L*H] with the cat/view quirk
        names = []
        for grp in (self.fcs, self.fcvars):
            for lin in grp:
                names += [owner.param_name_of(lin, "weight"), owner.param_name_of(lin, "bias")]
        o = self._lin(out, self.fcs[0], 0, rows=2 * self.num_class * self.c_dim, prefix=names)
        half = self.num_class * self.c_dim
        mu, lv = o[:, :half], o[:, half:]
        return list(mu.split(self.c_dim, 1)), list(lv.split(self.c_dim, 1))


class AdaINGen_v2(_FlatOwner):
    """networks/networks_v2.py:9-95."""

    def __init__(self, input_dim, vocab, params, pretrained_embed=None):
        super().__init__()
        dim, n_res, activ, pad_type = params["dim"], params["n_res"], params["activ"], params["pad_type"]
        mlp_dim, use_attention = params["mlp_dim"], params["use_attention"]
        c_dim, num_cls = params["c_dim"], params["num_cls"]
        style_dim = c_dim * num_cls
        self.c_dim, self.num_cls = c_dim, num_cls
        self.enc_style = StyleEncoder(params["style_downsample"], input_dim, dim, norm="none", activ=activ,
                                      pad_type=pad_type, c_dim=c_dim, num_class=num_cls, use_map=params["use_map"])
        self.enc_content = ContentEncoder(params["content_downsample"], n_res, input_dim, dim, "in", activ,
                                          pad_type=pad_type)
        self.dec = Decoder(params["content_downsample"], n_res, self.enc_content.output_dim, input_dim,
                           res_norm="adain", activ=activ, pad_type=pad_type, use_attention=use_attention)
        self.enc_txt = TxtEncoder(vocab, params["embed_dim"], params["hidden_size"], c_dim, num_cls,
                                  params["num_layers"], params["dropout_in"], params["dropout_out"],
                                  pretrained_embed=pretrained_embed)
        self.mlp = MLP(style_dim, self.get_num_adain_params(self.dec), mlp_dim, 3, norm="none", activ=activ)
        heads = lambda pre: ([f"{pre}.fcs.{i}.weight" for i in range(num_cls)] +
                             [f"{pre}.fcvars.{i}.weight" for i in range(num_cls)],
                             [f"{pre}.fcs.{i}.bias" for i in range(num_cls)] +
                             [f"{pre}.fcvars.{i}.bias" for i in range(num_cls)])
        from .text import _lstm_groups
        self._fuse_groups = [*heads("enc_style"), *heads("enc_txt"), *_lstm_groups(params["num_layers"]),
                             ["dec.image_content.conv.weight", "dec.image_attention.conv.weight"],
                             ["dec.image_content.conv.bias", "dec.image_attention.conv.bias"]]
        # flat buffer is created lazily (first use) so that weight init sees plain contiguous tensors

    # ---- reference API --------------------------------------------------------------------
    def forward(self, images):
        content, mus, logvar = self.encode(images)
        return self.decode(content, torch.cat(mus, 1) if isinstance(mus, (list, tuple)) else mus)

    def encode_fused(self, images, after_style=None):
        """(content tensor, mu [B, S], logvar [B, S]) with one shared padded copy of the image.  `after_style(mu)` is
        called between the two encoders: the Solver forks the text encoder (which only needs mu) onto its own stream
        there, so that its latency-bound LSTM / small-GEMM kernels run under the content encoder's convolutions."""
        self.ensure_flat()
        images = images.contiguous().float()
        rows = ops.image_rows(images, 1, self.enc_style.model[0])     # shared by both 7x7 first layers
        mu, lv = self.enc_style.run(images, rows)
        if after_style is not None:
            after_style(mu)
        content = _hb_to_tensor(self.enc_content.run(images, rows))
        return content, mu, lv

    def encode_forked(self, images):
        """encode_fused with the style encoder on its own stream: returns (content, mu, lv, join) where join() makes the
        current stream wait for mu / lv.  The caller keeps working on the content code (e.g. the cycle decode) while the
        style encoder's small feature maps are processed; autograd replays both backward passes the same way."""
        self.ensure_flat()
        images = images.contiguous().float()
        rows = ops.image_rows(images, 1, self.enc_style.model[0])
        if not (ops.RT.use_sty_stream and images.is_cuda):
            mu, lv = self.enc_style.run(images, rows)
            return _hb_to_tensor(self.enc_content.run(images, rows)), mu, lv, (lambda: None)
        main = torch.cuda.current_stream()
        ss = ops.RT.aux_stream(images.device, 'sty')
        ss.wait_stream(main)
        with torch.cuda.stream(ss):
            mu, lv = self.enc_style.run(images, rows)
        content = _hb_to_tensor(self.enc_content.run(images, rows))

        def join():
            cur = torch.cuda.current_stream()
            cur.wait_stream(ss)
            mu.record_stream(cur)
            lv.record_stream(cur)
        return content, mu, lv, join

    def encode(self, images):
        content, mu, lv = self.encode_fused(images)
        return content, list(mu.split(self.c_dim, 1)), list(lv.split(self.c_dim, 1))

    def encode_txt(self, style_ord, txt_org2trg, txt_lens):
        self.ensure_flat()
        return self.enc_txt(style_ord, txt_org2trg, txt_lens)

    def decode(self, content, style):
        self.ensure_flat()
        adain_params = self.mlp(style)
        self.assign_adain_params(adain_params, self.dec)
        return self.dec(content)

    def assign_adain_params(self, adain_params, model):
        """networks_v2.py:89-95 / networks.py:463-472.  All AdaIN layers of the decoder have the same width, so the
        [N, L*2*F] MLP output is split by ONE transposing copy into L x (mean, std) contiguous [N*F] vectors (and one
        gather in backward) instead of 2L slice copies, 2L zero-fills and 2L gradient accumulations per decode."""
        mods = [m for m in model.modules() if m.__class__.__name__ == "AdaptiveInstanceNorm2d"]
        feats = {m.num_features for m in mods}
        if mods and len(feats) == 1 and adain_params.is_cuda and adain_params.dim() == 2 and \
                adain_params.size(1) == 2 * len(mods) * mods[0].num_features:
            parts = ops.AdainSplitFn.apply(adain_params, len(mods), mods[0].num_features)
            for i, m in enumerate(mods):
                m.bias = parts[2 * i]
                m.weight = parts[2 * i + 1]
            return
        for m in model.modules():
            if m.__class__.__name__ == "AdaptiveInstanceNorm2d":
                mean = adain_params[:, :m.num_features]
                std = adain_params[:, m.num_features:2 * m.num_features]
                m.bias = mean.contiguous().view(-1)
                m.weight = std.contiguous().view(-1)
                if adain_params.size(1) > 2 * m.num_features:
                    adain_params = adain_params[:, 2 * m.num_features:]

    def get_num_adain_params(self, model):
        return sum(2 * m.num_features for m in model.modules() if m.__class__.__name__ == "AdaptiveInstanceNorm2d")


# ---------------------------------------------------------------------------------------------
# discriminator
# ---------------------------------------------------------------------------------------------

class MsImageDis(_FlatOwner, _LinearHolder):
    """networks/networks.py:43-170 (lsgan; norm none; lrelu; reflect)."""

    def __init__(self, input_dim, params, device=None):
        super().__init__()
        self.n_layer, self.gan_type, self.dim = params["n_layer"], params["gan_type"], params["dim"]
        self.norm, self.activ, self.num_scales = params["norm"], params["activ"], params["num_scales"]
        self.pad_type, self.num_cls, self.input_dim = params["pad_type"], params["num_cls"], input_dim
        self.image_size, self.dataset = params["image_size"], params["dataset"]
        self.device = device if device is not None else torch.device("cpu")
        if self.gan_type != "lsgan" or self.norm != "none":
            raise NotImplementedError("only gan_type lsgan / norm none are on the hot path (celeba_faces.yaml:62-67)")
        self.cnns_feat, self.cnns_src, self.cnns_cls = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for i in range(self.num_scales):
            feat, src, cls = self._make_net(self.image_size // (2 ** i))
            self.cnns_feat.append(feat)
            self.cnns_src.append(src)
            self.cnns_cls.append(cls)
        # flat buffer is created lazily

    def _make_net(self, im_size):
        dim = self.dim
        cnn = [Conv2dBlock(self.input_dim, dim, 4, 2, 1, norm="none", activation=self.activ, pad_type=self.pad_type)]
        pre = dim
        for _ in range(self.n_layer - 1):
            dim = min(dim * 2, 512)
            cnn += [Conv2dBlock(pre, dim, 4, 2, 1, norm=self.norm, activation=self.activ, pad_type=self.pad_type)]
            pre = dim
        src = nn.Conv2d(dim, 1, 1, 1, 0)
        cls = nn.Conv2d(dim, self.num_cls, kernel_size=im_size // (2 ** self.n_layer), stride=1, padding=0, bias=False)
        return nn.Sequential(*cnn), src, cls

    def _head(self, h: HB, src: nn.Conv2d, cls: nn.Conv2d):
        f = self.flat
        x2 = h.t.reshape(h.n * h.h * h.w, h.c)

        def lin(x, conv, rows, k):
            wn = self.param_name_of(conv, "weight")
            bn = self.param_name_of(conv, "bias") if conv.bias is not None else None
            w = f.raw(wn).view(rows, k)
            b = f.raw(bn) if bn is not None else None

            def grads():
                return f.raw_grad(wn).view(rows, k), (f.raw_grad(bn) if bn is not None else None)
            return ops.linear(x, w, b, 0, grads, conv.weight)

        out_src = lin(x2, src, 1, h.c).view(h.n, h.h, h.w, 1).permute(0, 3, 1, 2)
        kk = cls.kernel_size[0]
        assert kk == h.h and kk == h.w, "classification head expects a full-extent kernel"
        out_cls = lin(h.t.reshape(h.n, h.h * h.w * h.c), cls, self.num_cls, h.h * h.w * h.c)
        return out_src, out_cls

    def forward(self, x, use_multiscales=True):
        self.ensure_flat()
        outputs = []
        x = x.contiguous().float()
        for i in range(self.num_scales):
            convs = list(self.cnns_feat[i])
            pool = 2 ** i
            h = convs[0].run_first(x, ops.image_rows(x, pool, convs[0]), pool, out_halo=1, out_layout=1)
            for j, cv in enumerate(convs[1:], start=1):
                last = j == len(convs) - 1
                h = cv.run(h, out_halo=0 if last else 1, out_layout=0 if last else 1)
            outputs.append(list(self._head(h, self.cnns_src[i], self.cnns_cls[i])))
            if not use_multiscales:
                break
        return outputs

    def _classification_loss(self, logit, target, dataset="CelebA"):
        return ops.bce_logits(logit, target)

    def calc_dis_loss(self, input_fake, input_real, fake_cls, real_cls, weight_gan=1.0, weight_cls=1.0):
        outs0 = self.forward(input_fake)
        outs1 = self.forward(input_real)
        loss = 0.0
        for out_fake, out_real in zip(outs0, outs1):
            loss = loss + (ops.mse_const(out_fake[0], 0.0) + ops.mse_const(out_real[0], 1.0)) * weight_gan
            loss = loss + self._classification_loss(out_real[1], real_cls, self.dataset) * weight_cls
        return loss

    def calc_gen_loss(self, input_fake, target_cls, weight_gan=1.0, weight_cls=1.0):
        loss = 0
        for out in self.forward(input_fake):
            loss = loss + ops.mse_const(out[0], 1.0) * weight_gan
            loss = loss + self._classification_loss(out[1], target_cls, self.dataset) * weight_cls
        return loss
