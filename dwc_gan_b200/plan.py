"""Geometry of the haloed NHWC buffers and the gconv / wgrad launch plans.

A convolution of the reference (reflect-pad + nn.Conv2d, networks/networks.py:531,577-580) becomes
  forward : gconv over the reflect-haloed input (4-D boxes of 128 output pixels),
  dgrad   : gconv over the zero-haloed output gradient in "flat" mode (tap = constant row offset);
            stride-2 4x4 convs split into the four parity planes of the padded input,
  wgrad   : pixel-reduction GEMM between the output gradient and the shifted padded input.
The plans are plain Python objects so that tests can execute them with a CPU emulator
(tests/emu.py) without a GPU; `launch()` turns them into the C-ABI structs.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

from . import _lib as L


def _pow2_floor(v: int) -> int:
    p = 1
    while p * 2 <= v:
        p *= 2
    return p


def choose_box(w: int, h: int, n: int, rows: int) -> Tuple[int, int, int]:
    """(bx, by, bn) with bx*by*bn == rows covering a (w, h, n) pixel grid with little waste."""
    bx = min(_pow2_floor(w), rows)
    by = min(_pow2_floor(h), rows // bx)
    bn = rows // (bx * by)
    return bx, by, bn


class HB:
    """Haloed NHWC activation buffer.  layout 0: [N, H+2h, W+2h, C]; layout 1: four parity planes
    [N, 4, (H+2h)/2, (W+2h)/2, C] (input of a stride-2 conv)."""

    __slots__ = ("t", "n", "h", "w", "c", "halo", "layout", "stats", "act_out")

    def __init__(self, t, n, h, w, c, halo, layout=0, stats=None, act_out=None):
        self.t, self.n, self.h, self.w, self.c, self.halo, self.layout = t, n, h, w, c, halo, layout
        self.stats = stats            # (partial sums tensor, splits) written by the producing convolution's epilogue
        self.act_out = act_out        # (tensor, act, halo, layout): activated + haloed copy written by the same epilogue

    @staticmethod
    def shape_of(n, h, w, c, halo, layout=0):
        hp, wp = h + 2 * halo, w + 2 * halo
        if layout == 0:
            return (n, hp, wp, c)
        assert hp % 2 == 0 and wp % 2 == 0
        return (n, 4, hp // 2, wp // 2, c)

    @classmethod
    def empty(cls, n, h, w, c, halo, layout, dtype, device, zero=False):
        shape = cls.shape_of(n, h, w, c, halo, layout)
        t = torch.zeros(shape, dtype=dtype, device=device) if zero else torch.empty(shape, dtype=dtype, device=device)
        return cls(t, n, h, w, c, halo, layout)

    def like(self, t):
        return HB(t, self.n, self.h, self.w, self.c, self.halo, self.layout)

    @property
    def hp(self):
        return self.h + 2 * self.halo

    @property
    def wp(self):
        return self.w + 2 * self.halo

    def struct(self):
        assert self.t.is_contiguous()
        return L.HBuf(self.t.data_ptr(), self.n, self.h, self.w, self.c, self.halo, self.layout, L.dt(self.t))

    def interior(self):
        """[N, H, W, C] strided view of the interior (plain layout only)."""
        assert self.layout == 0
        h = self.halo
        return self.t[:, h:h + self.h, h:h + self.w, :]

    def interior_offset(self):
        assert self.layout == 0
        return (self.halo * self.wp + self.halo) * self.c

    def padded_nhwc(self):
        """[N, Hp, Wp, C] tensor of the padded image (gathers the parity planes if needed)."""
        if self.layout == 0:
            return self.t
        n, _, hq, wq, c = self.t.shape
        out = self.t.new_empty(n, hq * 2, wq * 2, c)
        for py in range(2):
            for px in range(2):
                out[:, py::2, px::2, :] = self.t[:, py * 2 + px]
        return out


@dataclass
class GConvPlan:
    a: torch.Tensor                # storage tensor of the A operand
    a_off: int                     # element offset of the view origin inside `a`
    a_dim: Tuple[int, ...]         # C, X, Y, Z, N
    a_str: Tuple[int, ...]
    box: Tuple[int, int, int]
    tiles: Tuple[int, int, int]
    valid: Tuple[int, int, int]
    flat: Tuple[int, int, int, int, int]   # flag, img rows, pitch, h, w
    taps: List[Tuple[int, int, int]]
    w: torch.Tensor                # [ncols_padded, ntaps*C]
    w_off: int
    ncols: int
    ncols_padded: int
    bias: Optional[torch.Tensor]
    out: torch.Tensor
    out_off: int
    o_str: Tuple[int, int, int]
    accumulate: bool = False
    backend: int = L.SIMT
    stats: Optional[torch.Tensor] = None   # fused per-(n, tile, column) {sum, sumsq} partials (see dwc_b200.h)
    out2: Optional[Tuple] = None   # (tensor, act, halo, layout): activated, reflect-haloed second output (dwc_b200.h)
    nphase: int = 1                # independent problems sharing A / tiling / taps (stride-2 dgrad parity phases)
    phase_w_off: int = 0           # element offset of phase ph's weights: w_off + ph * phase_w_off
    phase_out_off: int = 0         # element offset of phase ph's output: out_off + ph * phase_out_off

    def launch(self):
        g = L.GConv()
        es = self.a.element_size()
        g.dtype = L.dt(self.a)
        g.backend = self.backend
        g.a = self.a.data_ptr() + self.a_off * es
        g.a_dim = (C.c_int64 * 5)(*self.a_dim)
        g.a_str = (C.c_int64 * 5)(*self.a_str)
        g.box = (C.c_int32 * 3)(*self.box)
        g.tiles = (C.c_int32 * 3)(*self.tiles)
        g.valid = (C.c_int32 * 3)(*self.valid)
        g.flat, g.flat_img, g.flat_pitch, g.flat_h, g.flat_w = self.flat
        g.ntaps = len(self.taps)
        flat_taps = [v for t in self.taps for v in t]
        taps_arr = (C.c_int32 * len(flat_taps))(*flat_taps)
        g.taps = C.cast(taps_arr, C.POINTER(C.c_int32))
        g.w = self.w.data_ptr() + self.w_off * self.w.element_size()
        g.ncols, g.ncols_padded = self.ncols, self.ncols_padded
        g.bias = self.bias.data_ptr() if self.bias is not None else 0
        g.out = self.out.data_ptr() + self.out_off * self.out.element_size()
        g.o_str = (C.c_int64 * 3)(*self.o_str)
        g.out_dtype = L.dt(self.out)
        g.accumulate = int(self.accumulate)
        g.nphase, g.phase_w_off, g.phase_out_off = self.nphase, self.phase_w_off, self.phase_out_off
        g.stats = self.stats.data_ptr() if self.stats is not None else None
        if self.out2 is not None:
            g.out2, g.out2_act, g.out2_halo, g.out2_layout = (self.out2[0].data_ptr(),) + tuple(self.out2[1:])
        L.check(L.lib().dwc_gconv(C.byref(g), L.stream()), "gconv")


@dataclass
class WGradPlan:
    a: torch.Tensor
    a_off: int
    a_dim: Tuple[int, ...]
    a_str: Tuple[int, ...]
    b: torch.Tensor
    b_off: int
    b_dim: Tuple[int, ...]
    b_str: Tuple[int, ...]
    box: Tuple[int, int, int]
    tiles: Tuple[int, int, int]
    taps: List[Tuple[int, int, int]]
    ca: int
    cb: int
    dw: torch.Tensor               # fp32 gradient buffer, element (ia, t, ib) at ia*s_a + t*s_t + ib*s_b
    s_a: int
    s_t: int
    s_b: int
    dbias: Optional[torch.Tensor]
    accumulate: bool = True
    backend: int = L.SIMT
    remap: Optional[Tuple[int, int, int, int, int, int]] = None   # axis, div, lo_limit, hi_limit, hi_stride, lo_stride

    def _struct(self, workspace=None):
        g = L.WGrad()
        es = self.a.element_size()
        g.dtype = L.dt(self.a)
        g.backend = self.backend
        g.a = self.a.data_ptr() + self.a_off * es
        g.a_dim = (C.c_int64 * 5)(*self.a_dim)
        g.a_str = (C.c_int64 * 5)(*self.a_str)
        g.b = self.b.data_ptr() + self.b_off * es
        g.b_dim = (C.c_int64 * 5)(*self.b_dim)
        g.b_str = (C.c_int64 * 5)(*self.b_str)
        g.box = (C.c_int32 * 3)(*self.box)
        g.tiles = (C.c_int32 * 3)(*self.tiles)
        g.ntaps = len(self.taps)
        flat_taps = [v for t in self.taps for v in t]
        self._taps_arr = (C.c_int32 * len(flat_taps))(*flat_taps)
        g.taps = C.cast(self._taps_arr, C.POINTER(C.c_int32))
        g.ca, g.cb = self.ca, self.cb
        g.dw = self.dw.data_ptr()
        g.s_a, g.s_t, g.s_b = self.s_a, self.s_t, self.s_b
        g.dbias = self.dbias.data_ptr() if self.dbias is not None else 0
        g.accumulate = int(self.accumulate)
        if self.remap is not None:
            (g.remap_axis, g.remap_div, g.remap_lo_limit, g.remap_hi_limit, g.remap_hi_stride,
             g.remap_lo_stride) = self.remap
        if workspace is not None:
            g.workspace = workspace.data_ptr()
            g.workspace_bytes = workspace.numel() * workspace.element_size()
        return g

    def workspace_bytes(self):
        return int(L.lib().dwc_wgrad_workspace_bytes(C.byref(self._struct())))

    def launch(self, workspace_fn):
        need = self.workspace_bytes()
        ws = workspace_fn(need)
        L.check(L.lib().dwc_wgrad(C.byref(self._struct(ws)), L.stream()), "wgrad")


# --------------------------------------------------------------------------------------------
# plan builders
# --------------------------------------------------------------------------------------------

def conv_taps(k: int, stride: int) -> List[Tuple[int, int, int]]:
    """(dx, dy, z) per filter tap in (kh, kw) order for the forward / wgrad B operand."""
    taps = []
    for kh in range(k):
        for kw in range(k):
            if stride == 1:
                taps.append((kw, kh, 0))
            else:
                taps.append((kw // 2, kh // 2, (kh % 2) * 2 + (kw % 2)))
    return taps


def input_view(xp: HB):
    """rank-5 (C, X, Y, Z, N) dims/strides of a padded input buffer."""
    c = xp.c
    if xp.layout == 0:
        return (c, xp.wp, xp.hp, 1, xp.n), (1, c, xp.wp * c, xp.hp * xp.wp * c, xp.hp * xp.wp * c)
    hq, wq = xp.hp // 2, xp.wp // 2
    return (c, wq, hq, 4, xp.n), (1, c, wq * c, hq * wq * c, 4 * hq * wq * c)


def stats_splits(plan: "GConvPlan") -> int:
    """Number of per-image partials the fused-statistics epilogue of `plan` writes, or 0 if it cannot (tiles that
    span images, non-tensor-core backends, few-channel outputs, experimental kernels)."""
    if plan.backend != L.TC or plan.box[2] != 1 or plan.flat[0] or plan.nphase != 1 or plan.accumulate:
        return 0
    if plan.ncols % 64 != 0 or plan.ncols != plan.ncols_padded or plan.out.dtype != torch.bfloat16:
        return 0
    if HALO_K or os.environ.get("DWC_CG2") or os.environ.get("DWC_GCONV_DEBUG"):
        return 0
    return plan.tiles[0] * plan.tiles[1]


def out_size(xp: HB, k: int, stride: int):
    return (xp.hp - k) // stride + 1, (xp.wp - k) // stride + 1


# Which stride-1 window sizes run on the halo-tile kernel (csrc/gconv_halo.cu) instead of the tap-by-tap kernel:
# none by default.  Measured on B200 (profiles/r01c_microbench_halo_vs_tap.md, r01d_*): fetching the input k times
# instead of k*k times, and halving the weight bytes per MMA cycle with two accumulators, does NOT make these layers
# faster - the tap-by-tap kernel is not bound by L2 operand traffic but by per-CTA prologue/epilogue time and by the
# SS-mode A-operand read (>= 64 cycles per M=128 MMA, which caps N<=64 layers at <= 50 %).  DWC_HALO_K="3,5,7"
# switches the halo kernel on for experiments; it passes the same parity tests.
HALO_K = tuple(int(v) for v in os.environ.get("DWC_HALO_K", "").split(",") if v)
HALO_MT = int(os.environ.get("DWC_HALO_MT", "1"))       # accumulators (128-row M tiles) per CTA


def halo_backend(backend, k, stride, c, ncols_padded=None):
    if backend == L.TC and stride == 1 and k in HALO_K and c % 64 == 0:
        return L.TC_HALO
    return backend


def halo_box():
    return (16, 16, 1) if HALO_MT == 2 else (8, 16, 1)


def plan_conv_fwd(xp: HB, w_packed, ncols, ncols_padded, bias, y: HB, k, stride, backend) -> GConvPlan:
    ho, wo = out_size(xp, k, stride)
    assert (y.h, y.w, y.n) == (ho, wo, xp.n) and y.layout == 0, ((y.h, y.w), (ho, wo))
    assert (stride == 1 and xp.layout == 0) or (stride == 2 and xp.layout == 1 and k == 4)
    dims, strs = input_view(xp)
    backend = halo_backend(backend, k, stride, xp.c, ncols_padded)
    box = halo_box() if backend == L.TC_HALO else choose_box(wo, ho, xp.n, 128)
    tiles = (-(-wo // box[0]), -(-ho // box[1]), -(-xp.n // box[2]))
    cy = y.c
    return GConvPlan(a=xp.t, a_off=0, a_dim=dims, a_str=strs, box=box, tiles=tiles, valid=(wo, ho, xp.n),
                     flat=(0, 0, 0, 0, 0), taps=conv_taps(k, stride), w=w_packed, w_off=0, ncols=ncols,
                     ncols_padded=ncols_padded, bias=bias, out=y.t, out_off=y.interior_offset(),
                     o_str=(cy, y.wp * cy, y.hp * y.wp * cy), backend=backend)


def plan_conv_dgrad(dy: HB, w_packed, dxp: HB, k, stride, backend, cin_padded=None, accumulate=False) -> List[GConvPlan]:
    """dy: zero-haloed output gradient (halo k-1 for stride 1, 1 for the stride-2 4x4 conv).
    dxp: gradient of the padded input (same geometry as the forward input buffer); fully overwritten, or added to
    when `accumulate` (stride 1 only: the skip-connection gradient of a ResBlock is already in it)."""
    assert not accumulate or stride == 1
    cout, cin = dy.c, dxp.c
    cin_padded = cin_padded or cin
    rows = dy.n * dy.hp * dy.wp
    tiles = (-(-rows // 128), 1, 1)
    a_dim = (cout, rows, 1, 1, 1)
    a_str = (1, cout, rows * cout, rows * cout, rows * cout)
    plans = []
    hb = halo_backend(backend, k, stride, cout, cin_padded)
    if stride == 1 and hb != backend:
        # full correlation of the zero-haloed gradient with the flipped filter = a valid k x k window over dy's buffer
        assert dy.halo == k - 1 and dxp.layout == 0 and (dxp.hp, dxp.wp) == (dy.hp - k + 1, dy.wp - k + 1)
        dims = (cout, dy.wp, dy.hp, 1, dy.n)
        strs = (1, cout, dy.wp * cout, dy.hp * dy.wp * cout, dy.hp * dy.wp * cout)
        hbx = halo_box()
        plans.append(GConvPlan(a=dy.t, a_off=0, a_dim=dims, a_str=strs, box=hbx,
                               tiles=(-(-dxp.wp // hbx[0]), -(-dxp.hp // 16), dy.n), valid=(dxp.wp, dxp.hp, dy.n),
                               flat=(0, 0, 0, 0, 0), taps=conv_taps(k, 1), w=w_packed, w_off=0, ncols=cin,
                               ncols_padded=cin_padded, bias=None, out=dxp.t, out_off=0,
                               o_str=(cin, dxp.wp * cin, dxp.hp * dxp.wp * cin), backend=hb, accumulate=accumulate))
    elif stride == 1:
        assert dy.halo == k - 1 and dxp.layout == 0
        taps = [(kh * dy.wp + kw, 0, 0) for kh in range(k) for kw in range(k)]
        plans.append(GConvPlan(a=dy.t, a_off=0, a_dim=a_dim, a_str=a_str, box=(128, 1, 1), tiles=tiles,
                               valid=(rows, 1, dy.n), flat=(1, dy.hp * dy.wp, dy.wp, dxp.hp, dxp.wp), taps=taps,
                               w=w_packed, w_off=0, ncols=cin, ncols_padded=cin_padded, bias=None, out=dxp.t, out_off=0,
                               o_str=(cin, dxp.wp * cin, dxp.hp * dxp.wp * cin), backend=backend,
                               accumulate=accumulate))
    else:
        assert k == 4 and dy.halo == 1 and dxp.layout == 1
        hq, wq = dxp.hp // 2, dxp.wp // 2
        taps = [(i * dy.wp + j, 0, 0) for i in range(2) for j in range(2)]
        kk = 4 * cout
        # the four parity phases of the padded input share dy, the tiling and the 2x2 taps: one launch
        plans.append(GConvPlan(a=dy.t, a_off=0, a_dim=a_dim, a_str=a_str, box=(128, 1, 1), tiles=tiles,
                               valid=(rows, 1, dy.n), flat=(1, dy.hp * dy.wp, dy.wp, hq, wq), taps=taps,
                               w=w_packed, w_off=0, ncols=cin, ncols_padded=cin_padded,
                               bias=None, out=dxp.t, out_off=0,
                               o_str=(cin, wq * cin, 4 * hq * wq * cin), backend=backend,
                               nphase=4, phase_w_off=cin_padded * kk, phase_out_off=hq * wq * cin))
    return plans


def plan_conv_wgrad(dy: HB, xp: HB, dw, dbias, k, stride, backend, accumulate=True) -> WGradPlan:
    """dw: fp32 [Cout, k, k, Cin] contiguous (channels_last storage of the OIHW parameter)."""
    cout, cin = dy.c, xp.c
    b_dim, b_str = input_view(xp)
    a_dim = (cout, dy.w, dy.h, 1, dy.n)
    a_str = (1, cout, dy.wp * cout, dy.hp * dy.wp * cout, dy.hp * dy.wp * cout)
    rows = 64
    box = choose_box(dy.w, dy.h, dy.n, rows)
    tiles = (-(-dy.w // box[0]), -(-dy.h // box[1]), -(-dy.n // box[2]))
    return WGradPlan(a=dy.t, a_off=dy.interior_offset(), a_dim=a_dim, a_str=a_str, b=xp.t, b_off=0, b_dim=b_dim,
                     b_str=b_str, box=box, tiles=tiles, taps=conv_taps(k, stride), ca=cout, cb=cin, dw=dw,
                     s_a=k * k * cin, s_t=cin, s_b=1, dbias=dbias, accumulate=accumulate, backend=backend)


# --------------------------------------------------------------------------------------------
# few-channel layers on the tensor-core kernels through row-im2col buffers
# --------------------------------------------------------------------------------------------

def rows_shape(n, hp, wo, ys):
    """[N, ys, Hp/ys, Wo, 64] row-im2col buffer of a padded image (8-pixel x 8-channel windows)."""
    return (n, ys, hp // ys, wo, 64)


def rows_view(n, hp, wo, ys):
    yr = hp // ys
    return (64, wo, yr, ys, n), (1, 64, wo * 64, yr * wo * 64, ys * yr * wo * 64)


def first_conv_taps(k, stride):
    return [(0, kh, 0) for kh in range(k)] if stride == 1 else [(0, kh // 2, kh % 2) for kh in range(k)]


def plan_first_conv_fwd(rows_t, n, hp, wo, ho, k, stride, w_packed, cout, bias, y: HB, backend) -> GConvPlan:
    dims, strs = rows_view(n, hp, wo, stride)
    box = choose_box(wo, ho, n, 128)
    tiles = (-(-wo // box[0]), -(-ho // box[1]), -(-n // box[2]))
    cy = y.c
    return GConvPlan(a=rows_t, a_off=0, a_dim=dims, a_str=strs, box=box, tiles=tiles, valid=(wo, ho, n),
                     flat=(0, 0, 0, 0, 0), taps=first_conv_taps(k, stride), w=w_packed, w_off=0, ncols=cout,
                     ncols_padded=cout, bias=bias, out=y.t, out_off=y.interior_offset(),
                     o_str=(cy, y.wp * cy, y.hp * y.wp * cy), backend=backend)


def plan_first_conv_wgrad(dy: HB, rows_t, n, hp, wo, k, stride, cin, dw, dbias, backend, accumulate=True) -> WGradPlan:
    """dw: fp32 master gradient [Cout, k, k, cin]; the 64-wide window index (j, ch) is scattered to (kw=j, ci=ch)."""
    cout = dy.c
    b_dim, b_str = rows_view(n, hp, wo, stride)
    a_dim = (cout, dy.w, dy.h, 1, dy.n)
    a_str = (1, cout, dy.wp * cout, dy.hp * dy.wp * cout, dy.hp * dy.wp * cout)
    box = choose_box(dy.w, dy.h, dy.n, 64)
    tiles = (-(-dy.w // box[0]), -(-dy.h // box[1]), -(-dy.n // box[2]))
    return WGradPlan(a=dy.t, a_off=dy.interior_offset(), a_dim=a_dim, a_str=a_str, b=rows_t, b_off=0, b_dim=b_dim,
                     b_str=b_str, box=box, tiles=tiles, taps=first_conv_taps(k, stride), ca=cout, cb=64, dw=dw,
                     s_a=k * k * cin, s_t=k * cin, s_b=1, dbias=dbias, accumulate=accumulate, backend=backend,
                     remap=(2, 8, cin, k, cin, 1))


def plan_heads_dgrad(rows_d, n, hh, wh, w_packed, dxp: HB, k, backend) -> GConvPlan:
    """rows_d [N, Hh, Wh, 64]: 8-pixel windows of the zero-haloed 4(+4)-channel head gradient."""
    rows = n * hh * wh
    cin = dxp.c
    taps = [(kh * wh, 0, 0) for kh in range(k)]
    return GConvPlan(a=rows_d, a_off=0, a_dim=(64, rows, 1, 1, 1), a_str=(1, 64, rows * 64, rows * 64, rows * 64),
                     box=(128, 1, 1), tiles=(-(-rows // 128), 1, 1), valid=(rows, 1, n),
                     flat=(1, hh * wh, wh, dxp.hp, dxp.wp), taps=taps, w=w_packed, w_off=0, ncols=cin, ncols_padded=cin,
                     bias=None, out=dxp.t, out_off=0, o_str=(cin, dxp.wp * cin, dxp.hp * dxp.wp * cin), backend=backend)


def plan_heads_wgrad(win, xp: HB, dw, k, cout, backend, accumulate=True) -> WGradPlan:
    """win [N, H, W+halo, 64] with win[.., u, j*8+co] = dy[.., u-j, co]; dw: fp32 master gradient [cout, k, k, cin]."""
    n, h, wu = xp.n, xp.h, xp.wp
    cin = xp.c
    b_dim, b_str = input_view(xp)
    a_dim = (64, wu, h, 1, n)
    a_str = (1, 64, wu * 64, h * wu * 64, h * wu * 64)
    box = choose_box(wu, h, n, 64)
    tiles = (-(-wu // box[0]), -(-h // box[1]), -(-n // box[2]))
    return WGradPlan(a=win, a_off=0, a_dim=a_dim, a_str=a_str, b=xp.t, b_off=0, b_dim=b_dim, b_str=b_str, box=box,
                     tiles=tiles, taps=[(0, kh, 0) for kh in range(k)], ca=64, cb=cin, dw=dw, s_a=0, s_t=k * cin, s_b=1,
                     dbias=None, accumulate=accumulate, backend=backend,
                     remap=(1, 8, cout, k, cin, k * k * cin))
