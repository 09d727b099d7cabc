"""torch.autograd.Function wrappers around the C ABI (include/dwc_b200.h).

PyTorch is plumbing here (device memory, streams, the autograd tape); every arithmetic step of
the hot path is one of our CUDA kernels.  Conventions:
  * activations travel as haloed NHWC buffers (plan.HB); gradient buffers handed to a
    convolution's backward always carry a ZERO halo (written by the producer kernels);
  * parameter gradients are accumulated by the kernels straight into ``param.grad`` (views of a
    flat fp32 buffer, see flat.py); the parameter is still passed to the Function as an anchor so
    that the tape is built when only parameters require grad;
  * there is no CPU fallback: every entry point raises if the CUDA library is not usable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib as L
from . import plan as P
from .plan import HB

# --------------------------------------------------------------------------------------------
# runtime configuration
# --------------------------------------------------------------------------------------------


class Runtime:
    """Process-wide knobs: compute dtype (bf16 product mode / fp32 validation mode), kernel backend."""

    def __init__(self):
        self.dtype = torch.bfloat16
        self.use_tc = True
        self._ws = {}
        self.launches = 0
        self.side_streams = {}
        self.side_keep = []
        self.side_main = None
        self.use_side_stream = os.environ.get("DWC_SIDE", "1") != "0"
        self.fuse_skip_grad = os.environ.get("DWC_FUSE_SKIP_GRAD", "1") != "0"
        self.batch_pack = os.environ.get("DWC_BATCH_PACK", "1") != "0"
        self.use_conv7 = os.environ.get("DWC_CONV7", "1") != "0"
        self.use_txt_stream = os.environ.get("DWC_TXT_STREAM", "1") != "0"
        self.use_dis_stream = os.environ.get("DWC_DIS_STREAM", "1") != "0"
        self.aux_streams = {}
        self.home_stream = None                            # stream that called backward() (set by the Solver)
        self.use_sty_stream = os.environ.get("DWC_STY_STREAM", "1") != "0"
        self.conv7_which = os.environ.get("DWC_CONV7", "1")          # diagnostics: "h" heads only, "d" dgrad only
        self.use_fused_norm = os.environ.get("DWC_FUSED_NORM", "0") != "0"   # opt-in: the row-streaming passes are faster
        self.epi_stats = os.environ.get("DWC_EPI_STATS", "1") != "0"        # norm statistics from the conv epilogue
        # activation + reflect halo of the norm-less blocks from the conv epilogue (dwc_gconv_t.out2): bit-identical to
        # the separate pass and 35 launches fewer per step, but measured 0.4% SLOWER on the training step (the extra
        # scattered epilogue stores cost what the streaming pass did, profiles/r02f) - opt-in
        self.epi_act = os.environ.get("DWC_EPI_ACT", "0") == "1"
        # parked experiment (csrc/experimental/normbwd_cluster.cu, DWC_EXPERIMENTAL=1 builds only): one-pass cluster backward
        self.norm_cluster = os.environ.get("DWC_NORM_CLUSTER", "0") != "0"
        self.wgrad_hook = None       # data parallel: callable(weight name) after a layer's weight gradient is enqueued

    def set_mode(self, mode: str):
        if mode == "bf16":
            self.dtype, self.use_tc = torch.bfloat16, True
        elif mode == "bf16_simt":
            self.dtype, self.use_tc = torch.bfloat16, False
        elif mode == "fp32":
            self.dtype, self.use_tc = torch.float32, False
        else:
            raise ValueError("mode must be bf16, bf16_simt or fp32")
        # product mode: the large fp32 GEMMs of the text encoder run as tf32 tensor-core GEMMs (csrc/dense_tc.cu);
        # the validation modes keep exact fp32
        if os.path.exists(L.LIB_PATH):
            L.lib().dwc_set_tf32(1 if (mode == "bf16" and os.environ.get("DWC_TF32", "1") != "0") else 0)

    def tc_ok(self, *channels):
        return self.use_tc and self.dtype == torch.bfloat16 and all(c % 64 == 0 for c in channels)

    def aux_stream(self, device, name):
        key = (device.index if device.index is not None else torch.cuda.current_device(), name)
        st = self.aux_streams.get(key)
        if st is None:
            st = self.aux_streams[key] = torch.cuda.Stream(device=key[0])
        return st

    def workspace(self, nbytes, device=None):
        device = device or torch.device("cuda", torch.cuda.current_device())
        key = (device.index, torch.cuda.current_stream().cuda_stream)
        t = self._ws.get(key)
        if t is None or t.numel() * 4 < nbytes:
            t = torch.empty(max(nbytes // 4 + 1024, 16 << 20), dtype=torch.float32, device=device)
            self._ws[key] = t
        return t


RT = Runtime()


def side_launch(fn, keep, name=None):
    """Run `fn` (weight-gradient kernels: off the critical path of backward) on the device's side stream, ordered
    after everything already enqueued on the current stream.  `keep` are the tensors the kernels read: they are held
    until side_join() so that their memory is not reused while the side stream still needs it.  `name`: the weight
    whose gradient `fn` produces - reported to RT.wgrad_hook (bucketed gradient all-reduce, parallel.GradSync)."""
    try:
        return _side_launch(fn, keep)
    finally:
        if name is not None and RT.wgrad_hook is not None:
            RT.wgrad_hook(name)


def _side_launch(fn, keep):
    if os.environ.get("DWC_ABLATE_WGRAD"):       # diagnostics only (tools/ablate_step.sh): what the weight gradients cost
        return None
    if not RT.use_side_stream:
        return fn()
    dev = torch.cuda.current_device()
    side = RT.side_streams.get(dev)
    if side is None:
        side = RT.side_streams[dev] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        fn()
    if not RT.side_keep:
        # first side launch of this backward pass: join the streams when the pass ends, so that whoever called
        # .backward() sees complete parameter gradients on its own stream.  `main` may be an auxiliary stream (a
        # forked encoder replays its backward there); the Solver names the stream that called backward() as home.
        RT.side_main = RT.home_stream if RT.home_stream is not None else main
        torch.autograd.Variable._execution_engine.queue_callback(side_join)
    RT.side_keep.append(keep)


def side_join():
    """Make the current stream wait for the side stream (before the gradients are consumed)."""
    if RT.side_keep:
        main = RT.side_main
        main.wait_stream(RT.side_streams[main.device.index])
    RT.side_keep.clear()


def _require_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("dwc_gan_b200: the hot path runs on CUDA only (got a %s tensor); "
                           "there is no CPU fallback" % t.device)


def _call(fn_name, *args):
    RT.launches += 1
    L.check(getattr(L.lib(), fn_name)(*args), fn_name)


def _stats_splits(hw, n=1):
    """Splits of the per-(n,c) reductions: (split, sample) blocks that fill the GPU as ONE wave of two resident
    CTAs per SM (the row-streaming kernels give each block a long contiguous range of rows, so a partial second wave
    would cost a full one), at least 64 pixels per block."""
    return max(1, min(64, hw // 64, 296 // max(1, n)))


# --------------------------------------------------------------------------------------------
# image boundary
# --------------------------------------------------------------------------------------------

class ImagePadFn(torch.autograd.Function):
    """NCHW fp32 image -> reflect-haloed NHWC buffer (optional 2x avg-pool): networks.py:113,531."""

    @staticmethod
    def forward(ctx, img, pool, pad, layout):
        _require_cuda(img)
        img = img.contiguous().float()
        n, c, h, w = img.shape
        out = HB.empty(n, h // pool, w // pool, c, pad, layout, RT.dtype, img.device)
        hs = out.struct()
        _call("dwc_image_pad_fwd", L.ptr(img), n, c, h, w, pool, C.byref(hs), L.stream())
        ctx.meta = (n, c, h, w, pool, pad, layout)
        return out.t

    @staticmethod
    def backward(ctx, dout):
        n, c, h, w, pool, pad, layout = ctx.meta
        d = HB(dout.contiguous(), n, h // pool, w // pool, c, pad, layout)
        dimg = torch.empty(n, c, h, w, dtype=torch.float32, device=dout.device)
        hs = d.struct()
        _call("dwc_image_pad_bwd", C.byref(hs), pool, L.ptr(dimg), n, c, h, w, 0, L.stream())
        return dimg, None, None, None


def image_pad(img, pool, pad, layout) -> HB:
    n, c, h, w = img.shape
    t = ImagePadFn.apply(img, pool, pad, layout)
    return HB(t, n, h // pool, w // pool, c, pad, layout)


# --------------------------------------------------------------------------------------------
# convolution
# --------------------------------------------------------------------------------------------

class ConvFn(torch.autograd.Function):
    """reflect-padded conv + bias as gconv; backward = gconv dgrad + pixel-reduction wgrad
    (networks.py:577-580 and aten::convolution_backward)."""

    @staticmethod
    def forward(ctx, xp_t, weight, layer, xp: HB, skip_box=None, epi=None):
        _require_cuda(xp_t)
        ctx.skip_box = skip_box
        k, s, cout = layer.k, layer.stride, layer.total_cout()
        ho, wo = P.out_size(xp, k, s)
        hy = k - 1 if s == 1 else 1
        y = HB.empty(xp.n, ho, wo, cout, hy, 0, xp_t.dtype, xp_t.device)
        wf, rows_p = layer.packed_fwd(xp_t.dtype)
        tc = RT.tc_ok(xp.c) and (rows_p % 64 == 0 or rows_p == 16)
        pl = P.plan_conv_fwd(HB(xp_t, xp.n, xp.h, xp.w, xp.c, xp.halo, xp.layout), wf, cout, rows_p, layer.bias_f32(),
                             y, k, s, L.TC if tc else L.SIMT)
        stats = _attach_stats(pl, layer, xp.n, cout, xp_t.device)
        if stats is None:
            stats = _attach_act_out(pl, layer, epi, xp.n, ho, wo, cout, xp_t.dtype, xp_t.device)
        RT.launches += 1
        pl.launch()
        ctx.layer, ctx.xp_meta, ctx.y_meta = layer, (xp.n, xp.h, xp.w, xp.c, xp.halo, xp.layout), (ho, wo, hy)
        ctx.save_for_backward(xp_t)
        if stats is None:
            return y.t
        ctx.mark_non_differentiable(stats)
        return y.t, stats

    @staticmethod
    def backward(ctx, dy_t, dstats=None):
        layer = ctx.layer
        (xp_t,) = ctx.saved_tensors
        n, h, w, c, halo, layout = ctx.xp_meta
        ho, wo, hy = ctx.y_meta
        k, s, cout = layer.k, layer.stride, layer.total_cout()
        dy = HB(dy_t.contiguous(), n, ho, wo, cout, hy, 0)
        xp = HB(xp_t, n, h, w, c, halo, layout)
        dxp_t = None
        # Output channels that are not a multiple of 64 (64 -> 32 stage of the 256x256 variant): the gradient is
        # zero-extended to 64 channels once, so that both gradients run on the tcgen05 kernels (zero weights / zero
        # gradient rows for the padding), instead of falling back to the CUDA-core kernels
        padc = RT.tc_ok(c) and not RT.tc_ok(cout) and cout >= 16 and s == 1 and layer.extra_cols is None
        if padc:
            cp = _pad64(cout)
            dy = _pad_channels(dy, cp)
        if ctx.needs_input_grad[0]:
            # ResBlock: the skip-connection gradient (PostFn.backward of the block's second conv left it in the box,
            # same geometry as this conv's input, zero halo) is the buffer the data gradient is added to, instead of
            # a separate element-wise add of two full tensors by the autograd engine
            skip = ctx.skip_box.pop("dres", None) if ctx.skip_box is not None else None
            dxp = skip if skip is not None else HB.empty(n, h, w, c, halo, layout, dy_t.dtype, dy_t.device)
            assert (dxp.n, dxp.h, dxp.w, dxp.c, dxp.halo, dxp.layout) == (n, h, w, c, halo, layout)
            tc = RT.tc_ok(c, cout) or padc
            if padc:
                _, wm, _ = layer._raw_weight()
                wpad = torch.zeros(cp, k * k * c, dtype=torch.float32, device=dy_t.device)
                wpad[:cout] = wm.view(cout, -1)
                wd, rows_p = _pack(wpad, cp, k, c, 1, c, (c, k * k * cp), dy_t.dtype), c
                RT.launches += 1
            else:
                wd, rows_p = layer.packed_dgrad(dy_t.dtype)
            for q in P.plan_conv_dgrad(dy, wd, dxp, k, s, L.TC if tc else L.SIMT, cin_padded=rows_p,
                                       accumulate=skip is not None):
                RT.launches += 1
                q.launch()
            dxp_t = dxp.t
        if ctx.needs_input_grad[1]:
            gw, gb = layer.grad_buffers()
            if not layer.bias_grad_needed():
                gb = None          # bias in front of InstanceNorm / AdaIN: its gradient is exactly zero
            tc = RT.tc_ok(c, cout) or padc
            RT.launches += 3
            if padc:
                def run(dy=dy):
                    gwp = torch.empty(cp, k * k * c, dtype=torch.float32, device=dy_t.device)
                    gbp = torch.empty(cp, dtype=torch.float32, device=dy_t.device) if gb is not None else None
                    P.plan_conv_wgrad(dy, xp, gwp, gbp, k, s, L.TC, accumulate=False).launch(
                        lambda nbytes: RT.workspace(nbytes, dy_t.device))
                    gw.view(cout, -1).add_(gwp[:cout])
                    if gb is not None:
                        gb.add_(gbp[:cout])
                side_launch(run, (dy.t, xp_t), name=layer._names()[1])
            else:
                wp = P.plan_conv_wgrad(dy, xp, gw, gb, k, s, L.TC if tc else L.SIMT, accumulate=True)
                side_launch(lambda: wp.launch(lambda nbytes: RT.workspace(nbytes, dy_t.device)), (dy.t, xp_t),
                            name=layer._names()[1])
        return dxp_t, None, None, None, None, None


def _pad_channels(hb: HB, c_new: int) -> HB:
    """Zero-extend (or cut back) the channel dimension of a haloed buffer: a plain copy that lets layers whose channel
    count is not a multiple of 64 (the 32-channel last decoder stage of the 256x256 variant) run on the
    tcgen05 kernels, which stage 64-channel (128-byte) slabs."""
    if c_new > hb.c:
        t = torch.nn.functional.pad(hb.t, (0, c_new - hb.c))
    else:
        t = hb.t[..., :c_new].contiguous()
    return HB(t, hb.n, hb.h, hb.w, c_new, hb.halo, hb.layout)


def _pad64(c):
    return (c + 63) // 64 * 64


def _pack(w, cout, k, cin, mode, rows, out_shape, dtype):
    out = torch.empty(out_shape, dtype=dtype, device=w.device)
    _call("dwc_pack_weights", L.ptr(w), cout, k, k, cin, mode, L.ptr(out), L.dt(out), rows, L.stream())
    return out


def _attach_stats(pl, layer, n, cout, device):
    """Ask the convolution's epilogue for the per-(n, tile, column) partial sums its norm layer needs (InstanceNorm /
    AdaIN / LayerNorm statistics without re-reading y); None when the layer has no norm or the launch cannot do it."""
    if not RT.epi_stats or getattr(layer, "norm_kind", NORM_NONE) == NORM_NONE:
        return None
    splits = P.stats_splits(pl)
    if splits == 0 or splits > 256:
        return None
    pl.stats = torch.empty(n * splits * cout * 2, dtype=torch.float32, device=device)
    return pl.stats


def _attach_act_out(pl, layer, epi, n, ho, wo, cout, dtype, device):
    """Ask the convolution's epilogue to also write act(y) into the next block's reflect-haloed input buffer (layers
    without a norm: style encoder, discriminator), so that the forward needs no separate activation / pad pass.
    epi = (act, out_halo, out_layout) of the consumer; returns the buffer or None when the launch cannot do it."""
    if epi is None or not RT.epi_act or pl.backend != L.TC or dtype != torch.bfloat16:
        return None
    act, halo, layout = epi
    if getattr(layer, "norm_kind", NORM_NONE) != NORM_NONE or act not in (1, 2) or cout % 64 or \
            pl.ncols != pl.ncols_padded or pl.flat[0] or pl.nphase != 1:
        return None
    if halo and (ho < 2 * halo + 2 or wo < 2 * halo + 2):
        return None
    if cout % 256 and cout > 128 and cout % 128:
        return None
    o = HB.empty(n, ho, wo, cout, halo, layout, dtype, device)
    pl.out2 = (o.t, act, halo, layout)
    return o.t


def _hb_with_stats(out, n, ho, wo, cout, halo, epi=None):
    if isinstance(out, tuple):
        t, extra = out
        if epi is not None:
            return HB(t, n, ho, wo, cout, halo, 0, act_out=(extra,) + tuple(epi))
        return HB(t, n, ho, wo, cout, halo, 0, stats=(extra, extra.numel() // (2 * n * cout)))
    return HB(out, n, ho, wo, cout, halo, 0)


def conv(xp: HB, layer, skip_box=None, epi=None) -> HB:
    """epi = (act, out_halo, out_layout) of the consumer when the layer has no norm and no residual: the epilogue then
    also writes the activated, haloed tensor (HB.act_out) that PostFn.forward would otherwise produce."""
    k, s = layer.k, layer.stride
    ho, wo = P.out_size(xp, k, s)
    out = ConvFn.apply(xp.t, layer.weight_param, layer, xp, skip_box, epi)
    return _hb_with_stats(out, xp.n, ho, wo, layer.total_cout(), k - 1 if s == 1 else 1,
                          epi if isinstance(out, tuple) and getattr(layer, "norm_kind", NORM_NONE) == NORM_NONE else None)


def image_rows(img, pool, layer):
    """Row-im2col buffer of the (pooled, reflect-padded) image for a few-channel first convolution; no autograd."""
    _require_cuda(img)
    img = img.contiguous().float()
    n, c, h, w = img.shape
    k, s, p = layer.k, layer.stride, layer.padding
    hp = h // pool + 2 * p
    wo = (w // pool + 2 * p - k) // s + 1
    rows = torch.empty(P.rows_shape(n, hp, wo, s), dtype=RT.dtype, device=img.device)
    _call("dwc_image_rows_fwd", L.ptr(img), n, c, h, w, pool, p, s, s, wo, L.ptr(rows), L.dt(rows), L.stream())
    return rows


class FirstConvFn(torch.autograd.Function):
    """First convolution of an encoder / discriminator scale (3 -> 64 channels, 7x7 s1 or 4x4 s2) on the tensor
    cores: the image is expanded once into 8-pixel x 8-channel windows (image_rows), which turns the layer into a
    k-tap, 64-channel gconv / wgrad.  Backward to the image (generated images only) is a regular dgrad."""

    @staticmethod
    def forward(ctx, img, weight, layer, rows_t, pool, epi=None):
        n, c, h, w = img.shape
        k, s, p, cout = layer.k, layer.stride, layer.padding, layer.cout
        hi, wi = h // pool, w // pool
        hp = hi + 2 * p
        ho, wo = (hp - k) // s + 1, (wi + 2 * p - k) // s + 1
        hy = k - 1 if s == 1 else 1
        y = HB.empty(n, ho, wo, cout, hy, 0, rows_t.dtype, rows_t.device)
        be = L.TC if RT.tc_ok(64, cout) else L.SIMT
        RT.launches += 1
        pl = P.plan_first_conv_fwd(rows_t, n, hp, wo, ho, k, s, layer.packed_rows(rows_t.dtype, 3), cout, layer.bias_f32(),
                                   y, be)
        stats = _attach_stats(pl, layer, n, cout, rows_t.device)
        if stats is None:
            stats = _attach_act_out(pl, layer, epi, n, ho, wo, cout, rows_t.dtype, rows_t.device)
        pl.launch()
        ctx.layer, ctx.meta = layer, (n, c, h, w, pool, hp, ho, wo, hy)
        ctx.save_for_backward(rows_t)
        if stats is None:
            return y.t
        ctx.mark_non_differentiable(stats)
        return y.t, stats

    @staticmethod
    def backward(ctx, dy_t, dstats=None):
        layer = ctx.layer
        (rows_t,) = ctx.saved_tensors
        n, c, h, w, pool, hp, ho, wo, hy = ctx.meta
        k, s, p, cout = layer.k, layer.stride, layer.padding, layer.cout
        dy = HB(dy_t.contiguous(), n, ho, wo, cout, hy, 0)
        dimg = None
        if ctx.needs_input_grad[0]:
            layout = 1 if s == 2 else 0
            dxp = HB.empty(n, h // pool, w // pool, c, p, layout, dy_t.dtype, dy_t.device)
            tc = RT.tc_ok(cout)
            if conv7_few_ok(dy_t.dtype, k, s, cout, c) and hy == 6 and layout == 0 and RT.conv7_which != "h":
                # image gradient = valid 7x7 conv of the zero-haloed dy with the flipped filter, 64 -> c channels:
                # element (o = ci, ky, kx, i = co) of that filter is W[co][6-ky][6-kx][ci] of the master [64][7][7][c]
                _, wm, _ = layer._raw_weight()
                kk = k * k * c
                conv7_few(dy.t, n, dy.hp, dy.wp, wm, 6 * k * c + 6 * c, 1, -k * c, -c, kk, None, c, dxp.t,
                          (c, dxp.wp * c, dxp.hp * dxp.wp * c))
            else:
                wd, rows_p = layer.packed_dgrad(dy_t.dtype, pad_rows=tc)
                for q in P.plan_conv_dgrad(dy, wd, dxp, k, s, L.TC if tc else L.SIMT, cin_padded=rows_p):
                    RT.launches += 1
                    q.launch()
            dimg = torch.empty(n, c, h, w, dtype=torch.float32, device=dy_t.device)
            ds = dxp.struct()
            _call("dwc_image_pad_bwd", C.byref(ds), pool, L.ptr(dimg), n, c, h, w, 0, L.stream())
        if ctx.needs_input_grad[1]:
            gw, gb = layer.grad_buffers()
            if not layer.bias_grad_needed():
                gb = None
            be = L.TC if RT.tc_ok(64, cout) else L.SIMT
            RT.launches += 3
            wp = P.plan_first_conv_wgrad(dy, rows_t, n, hp, wo, k, s, c, gw, gb, be)
            side_launch(lambda: wp.launch(lambda nbytes: RT.workspace(nbytes, dy_t.device)), (dy.t, rows_t),
                        name=layer._names()[1])
        return dimg, None, None, None, None, None


def first_conv(img, rows_t, layer, pool, epi=None) -> HB:
    n, c, h, w = img.shape
    k, s, p = layer.k, layer.stride, layer.padding
    ho, wo = (h // pool + 2 * p - k) // s + 1, (w // pool + 2 * p - k) // s + 1
    out = FirstConvFn.apply(img.contiguous().float(), layer.weight_param, layer, rows_t, pool, epi)
    return _hb_with_stats(out, n, ho, wo, layer.cout, k - 1 if s == 1 else 1,
                          epi if isinstance(out, tuple) and layer.norm_kind == NORM_NONE else None)


def conv7_few(x_t, n, hin, win, w_master, w_base, s_o, s_ky, s_kx, s_i, bias, cout, out_t, out_str):
    """7x7 stride-1 valid conv, 64 -> cout <= 4 channels, on the dedicated tcgen05 kernel (csrc/conv7few.cu).
    x_t: bf16 [n, hin, win, 64] contiguous; out_t: bf16 buffer addressed by out_str = (x, y, n) element strides."""
    assert x_t.dtype == torch.bfloat16 and x_t.is_contiguous() and out_t.dtype == torch.bfloat16
    istr = (C.c_int64 * 3)(64, win * 64, hin * win * 64)
    ostr = (C.c_int64 * 3)(*out_str)
    _call("dwc_conv7_few", L.ptr(x_t), n, hin, win, istr, L.ptr(w_master), w_base, s_o, s_ky, s_kx, s_i, L.ptr(bias),
          cout, L.ptr(out_t), ostr, L.stream())
    if os.environ.get("DWC_CONV7_CHECK"):          # diagnostics: the kernel must be bit-reproducible in situ
        first = out_t.clone()
        for rep in range(3):
            _call("dwc_conv7_few", L.ptr(x_t), n, hin, win, istr, L.ptr(w_master), w_base, s_o, s_ky, s_kx, s_i,
                  L.ptr(bias), cout, L.ptr(out_t), ostr, L.stream())
            if not torch.equal(first, out_t):
                d = (first.float() - out_t.float()).abs()
                print("conv7 NOT reproducible: n %d hin %d cout %d rep %d: %d elements differ, max %.4g, nan %d/%d" % (
                    n, hin, cout, rep, int((d > 0).sum()), float(d.max()), int(torch.isnan(first).sum()),
                    int(torch.isnan(out_t).sum())), flush=True)


def conv7_few_ok(dtype, k, stride, c64, cfew):
    return RT.use_conv7 and RT.use_tc and dtype == torch.bfloat16 and k == 7 and stride == 1 and c64 == 64 and cfew <= 4


class HeadsConvFn(torch.autograd.Function):
    """Decoder heads (networks_v2.py:162-169): the two 7x7 convolutions as one 4-channel gconv + tanh / sigmoid.
    Backward builds 64-wide window buffers of the 4-channel gradient so that dgrad and wgrad run on the tensor cores."""

    @staticmethod
    def forward(ctx, xp_t, weight, layer, xp: HB, on_att_grad):
        ctx.set_materialize_grads(False)
        k, cout = layer.k, layer.total_cout()
        n, h, w = xp.n, xp.h, xp.w
        y = HB.empty(n, h, w, cout, 0, 0, xp_t.dtype, xp_t.device)
        cin_real = xp.c
        if RT.use_tc and xp_t.dtype == torch.bfloat16 and xp.c % 64 != 0 and xp.c % 8 == 0 and xp.layout == 0:
            # 32 input channels (256x256 variant): zero-extend input and weights to 64 channels, see _pad_channels
            xp = _pad_channels(HB(xp_t, n, h, w, xp.c, xp.halo, 0), _pad64(xp.c))
            xp_t = xp.t
        if conv7_few_ok(xp_t.dtype, k, 1, xp.c, cout) and xp.layout == 0 and xp.halo == 3 and RT.conv7_which != "d":
            _, wm, bm = layer._raw_weight()                    # fp32 master [cout][7][7][cin]
            if cin_real != xp.c:
                wm = torch.nn.functional.pad(wm.view(cout * k * k, cin_real), (0, xp.c - cin_real)).reshape(-1)
            conv7_few(xp_t, n, xp.hp, xp.wp, wm, 0, 49 * 64, 7 * 64, 64, 1, bm, cout, y.t, (cout, w * cout, h * w * cout))
        else:
            wf, rows_p = layer.packed_fwd(xp_t.dtype)
            tc = RT.tc_ok(xp.c) and rows_p == 16
            RT.launches += 1
            P.plan_conv_fwd(HB(xp_t, n, h, w, xp.c, xp.halo, 0), wf, cout, rows_p, layer.bias_f32(), y, k, 1,
                            L.TC if tc else L.SIMT).launch()
        img = torch.empty(n, cout - 1, h, w, dtype=torch.float32, device=xp_t.device)
        att = torch.empty(n, 1, h, w, dtype=torch.float32, device=xp_t.device)
        ys = y.struct()
        _call("dwc_heads_fwd", C.byref(ys), L.ptr(img), L.ptr(att), L.stream())
        ctx.layer, ctx.meta, ctx.on_att_grad = layer, (n, h, w, xp.c, xp.halo, cin_real), on_att_grad
        ctx.save_for_backward(xp_t, img, att)
        return img, att

    @staticmethod
    def backward(ctx, dimg, datt):
        layer = ctx.layer
        xp_t, img, att = ctx.saved_tensors
        n, h, w, cin, halo_in, cin_real = ctx.meta              # cin: channels of the (possibly zero-extended) saved input
        k, cout = layer.k, layer.total_cout()
        dev, dtype = xp_t.device, xp_t.dtype
        if datt is not None and ctx.on_att_grad is not None:
            ctx.on_att_grad()
        dimg = dimg.contiguous() if dimg is not None else None
        datt = datt.contiguous() if datt is not None else None
        halo = k - 1
        hh, wh, wu = h + 2 * halo, w + 2 * halo, w + halo
        assert wu == w + 2 * halo_in, "window buffer must span the padded input width"
        rows_d = torch.empty(n, hh, wh, 64, dtype=dtype, device=dev)
        win = torch.empty(n, h, wu, 64, dtype=dtype, device=dev)
        part = torch.empty(1024 * 4, dtype=torch.float32, device=dev)
        nblk = C.c_int32(0)
        _call("dwc_heads_bwd_rows", L.ptr(dimg), L.ptr(datt), L.ptr(img), L.ptr(att), n, h, w, halo, L.ptr(rows_d),
              L.ptr(win), L.dt(dtype), L.ptr(part), C.byref(nblk), L.stream())
        be = L.TC if RT.tc_ok(64, cin) else L.SIMT
        dxp_t = None
        if ctx.needs_input_grad[0]:
            dxp = HB.empty(n, h, w, cin, halo_in, 0, dtype, dev)
            RT.launches += 1
            if cin_real != cin:                               # zero rows for the padding channels
                _, wm, _ = layer._raw_weight()
                wr = _pack(wm, cout, k, cin_real, 4, cin, (cin, k * 64), dtype)
            else:
                wr = layer.packed_rows(dtype, 4)
            P.plan_heads_dgrad(rows_d, n, hh, wh, wr, dxp, k, be).launch()
            dxp_t = dxp.t if cin_real == cin else _pad_channels(dxp, cin_real).t
        if ctx.needs_input_grad[1]:
            gw, gb = layer.grad_buffers()
            RT.launches += 2
            if cin_real != cin:
                gwp = torch.zeros(cout * k * k * cin, dtype=torch.float32, device=dev)
                P.plan_heads_wgrad(win, HB(xp_t, n, h, w, cin, halo_in, 0), gwp, k, cout, be).launch(
                    lambda nbytes: RT.workspace(nbytes, dev))
                gw.view(cout * k * k, cin_real).add_(gwp.view(cout * k * k, cin)[:, :cin_real])
            else:
                P.plan_heads_wgrad(win, HB(xp_t, n, h, w, cin, halo_in, 0), gw, k, cout, be).launch(
                    lambda nbytes: RT.workspace(nbytes, dev))
            _call("dwc_colsum", nblk.value, cout, L.ptr(part), cout, 1, L.ptr(gb), 1, L.stream())
        return dxp_t, None, None, None, None


def heads_conv(xp: HB, layer, on_att_grad=None):
    return HeadsConvFn.apply(xp.t, layer.weight_param, layer, xp, on_att_grad)


# --------------------------------------------------------------------------------------------
# norm + activation + residual + reflect pad
# --------------------------------------------------------------------------------------------
def _prefold(d: HB) -> int:
    """Fold the gradient of a reflect halo into the interior, in place (dwc_fold_halo); returns the `prefolded` flag
    for the backward passes.  Tiny images keep the gathering path."""
    if d.halo > 0 and d.h >= 2 * d.halo + 2 and d.w >= 2 * d.halo + 2 and d.c % 8 == 0:
        ds = d.struct()
        _call("dwc_fold_halo", C.byref(ds), L.stream())
        return 1
    return 0


class AdainSplitFn(torch.autograd.Function):
    """[N, L*2*F] AdaIN parameters -> 2L contiguous [N*F] vectors (mean_0, std_0, mean_1, ...) with one copy each way."""

    @staticmethod
    def forward(ctx, params, nl, f):
        n = params.shape[0]
        ctx.meta = (n, nl, f, params.dtype)
        ctx.set_materialize_grads(False)
        buf = params.detach().float().reshape(n, nl, 2, f).permute(1, 2, 0, 3).contiguous()   # [L, 2, N, F]
        return tuple(buf[l, j].reshape(-1) for l in range(nl) for j in range(2))

    @staticmethod
    def backward(ctx, *grads):
        n, nl, f, dtype = ctx.meta
        ref = next((g for g in grads if g is not None), None)
        if ref is None:
            return None, None, None
        gs = [g.reshape(n, f).float() if g is not None else torch.zeros(n, f, dtype=torch.float32, device=ref.device)
              for g in grads]
        g = torch.stack(gs, 0).view(nl, 2, n, f).permute(2, 0, 1, 3).reshape(n, nl * 2 * f)
        return g.to(dtype), None, None


class WeightedSumFn(torch.autograd.Function):
    """sum_i w_i * term_i over scalar loss terms with constant weights: one stack + one multiply + one sum (and one
    scaling in backward) instead of a multiply and an add kernel per term in each direction."""
    _wcache = {}

    @staticmethod
    def forward(ctx, wkey, *terms):
        dev = terms[0].device
        w = WeightedSumFn._wcache.get((wkey, dev))
        if w is None:                                     # built in the eager warm-up steps, before any graph capture
            w = WeightedSumFn._wcache[(wkey, dev)] = torch.tensor(list(wkey), dtype=torch.float32, device=dev)
        ctx.save_for_backward(w)
        ctx.k = len(terms)
        st = torch.stack([t.reshape(()).float() for t in terms])
        return (st * w).sum()                             # (not torch.dot: that would be a cuBLAS call)

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        gv = g * w
        return (None,) + tuple(gv.unbind(0))


def weighted_sum(pairs):
    """pairs: [(scalar tensor, python float weight), ...] -> scalar tensor."""
    wkey = tuple(float(w) for _, w in pairs)
    return WeightedSumFn.apply(wkey, *[t for t, _ in pairs])


NORM_NONE, NORM_IN, NORM_ADAIN, NORM_LN = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2


class PostFn(torch.autograd.Function):
    """out = reflect_pad(act(norm(y)) + residual): InstanceNorm2d / AdaIN / MUNIT LayerNorm, ReLU /
    LeakyReLU(0.1), ResBlock residual and the next block's ReflectionPad2d in one pass
    (networks.py:514-522,531,545,580-585,706-719,736-752)."""

    @staticmethod
    def forward(ctx, y_t, nw, nb, res_t, anchor, y: HB, kind, act, res: Optional[HB], out_halo, out_layout, ln_mod,
                eps, skip_box=None):
        _require_cuda(y_t)
        ctx.skip_box = skip_box
        yh = HB(y_t, y.n, y.h, y.w, y.c, y.halo, 0)
        n, c, hw = y.n, y.c, y.h * y.w
        dev = y_t.device
        coef = None
        splits = _stats_splits(hw, n)
        fused = kind in (NORM_IN, NORM_ADAIN) and RT.use_fused_norm and \
            bool(L.lib().dwc_post_fused_ok(c, hw, L.dt(y_t)))
        ready = y.act_out is not None and kind == NORM_NONE and res is None and \
            tuple(y.act_out[1:]) == (act, out_halo, out_layout)
        if ready:
            # the producing convolution's epilogue already wrote act(y) + reflect halo (+ parity planes): nothing to
            # launch; backward is the usual activation backward on the saved pre-activation y
            ctx.fused = False
            ctx.meta = (y.n, y.h, y.w, y.c, y.halo, kind, act, out_halo, out_layout, eps, splits, None)
            ctx.ln_mod = ln_mod
            ctx.save_for_backward(y_t, None, None)
            return y.act_out[0]
        out = HB.empty(n, y.h, y.w, c, out_halo, out_layout, y_t.dtype, dev)
        ys, os_ = yh.struct(), out.struct()
        rs = HB(res_t, res.n, res.h, res.w, res.c, res.halo, res.layout).struct() if res is not None else None
        if kind == NORM_ADAIN:
            nw = nw.contiguous().float()
            nb = nb.contiguous().float()
        if fused:
            # small feature map: statistics, coefficients and the normalise / activation / pad pass in one kernel
            coef = torch.empty(n * c * 4, dtype=torch.float32, device=dev)
            _call("dwc_post_fused_fwd", C.byref(ys), kind, L.ptr(nw), L.ptr(nb), L.f32(eps), act,
                  C.byref(rs) if rs is not None else None, C.byref(os_), L.ptr(coef), L.stream())
        else:
            if kind != NORM_NONE:
                if y.stats is not None:                   # partial sums written by the producing conv's epilogue
                    stats, fsplits = y.stats
                else:
                    stats, fsplits = torch.empty(n * splits * c * 2, dtype=torch.float32, device=dev), splits
                    _call("dwc_nc_stats", C.byref(ys), splits, L.ptr(stats), L.stream())
                coef = torch.empty(n * c * 4, dtype=torch.float32, device=dev)
                # statistics -> coefficients inside the normalise kernel (or a separate finalize pass, see the C side)
                _call("dwc_post_fwd_norm", C.byref(ys), kind, L.ptr(stats), fsplits, L.f32(eps), L.ptr(nw), L.ptr(nb),
                      act, C.byref(rs) if rs is not None else None, C.byref(os_), L.ptr(coef), L.stream())
            else:
                _call("dwc_post_fwd", C.byref(ys), None, act, C.byref(rs) if rs is not None else None, C.byref(os_),
                      L.stream())
        ctx.fused = fused
        ctx.meta = (y.n, y.h, y.w, y.c, y.halo, kind, act, out_halo, out_layout, eps, splits,
                    (res.n, res.h, res.w, res.c, res.halo, res.layout) if res is not None else None)
        ctx.ln_mod = ln_mod
        ctx.save_for_backward(y_t, coef, nw if kind in (NORM_ADAIN, NORM_LN) else None)
        return out.t

    @staticmethod
    def backward(ctx, dout_t):
        n, h, w, c, yhalo, kind, act, out_halo, out_layout, eps, splits, res_meta = ctx.meta
        y_t, coef, nw = ctx.saved_tensors
        dev = dout_t.device
        dout = HB(dout_t.contiguous(), n, h, w, c, out_halo, out_layout)
        yh = HB(y_t, n, h, w, c, yhalo, 0)
        ds, ys = dout.struct(), yh.struct()
        dnw = dnb = None
        dy = HB.empty(n, h, w, c, yhalo, 0, dout_t.dtype, dev)
        dys = dy.struct()
        dres = drs = None
        if res_meta is not None and ctx.needs_input_grad[3]:
            dres = HB.empty(*res_meta[:5], res_meta[5], dout_t.dtype, dev)
            assert dres.layout == 0
            drs = dres.struct()
        if ctx.fused:
            if kind == NORM_ADAIN:
                dnw = torch.empty(n, c, dtype=torch.float32, device=dev)
                dnb = torch.empty(n, c, dtype=torch.float32, device=dev)
            _call("dwc_post_fused_bwd", C.byref(ds), C.byref(ys), L.ptr(coef), kind, act, L.ptr(nw), L.ptr(dnw),
                  L.ptr(dnb), C.byref(dys), C.byref(drs) if drs is not None else None, L.stream())
        elif RT.norm_cluster and kind in (NORM_IN, NORM_ADAIN) and hasattr(L.lib(), "dwc_post_bwd_cluster") and \
                L.lib().dwc_post_bwd_cluster_ok(
                C.byref(ds), C.byref(ys), kind, C.byref(dys), C.byref(drs) if drs is not None else None):
            # a whole sample fits the shared memory of an 8-CTA cluster (the 256-channel 32x32 residual blocks): halo
            # fold, reductions, coefficients and apply in ONE launch and one read of dout and y
            if kind == NORM_ADAIN:
                dnw = torch.empty(n, c, dtype=torch.float32, device=dev)
                dnb = torch.empty(n, c, dtype=torch.float32, device=dev)
            _call("dwc_post_bwd_cluster", C.byref(ds), C.byref(ys), L.ptr(coef), kind, act, L.ptr(nw), L.ptr(dnw),
                  L.ptr(dnb), C.byref(dys), C.byref(drs) if drs is not None else None, L.stream())
        else:
            # fold the reflect-halo gradient once, in place: the streaming passes then read whole interior rows
            # (norm sites whose reduction streams whole padded rows fold it inside that pass: no separate launch)
            in_reduce = kind != NORM_NONE and bool(L.lib().dwc_post_bwd_reduce_folds(C.byref(ds), C.byref(ys)))
            pre = 1 if in_reduce else _prefold(dout)
            if kind != NORM_NONE:
                red = torch.empty(n * splits * c * 2, dtype=torch.float32, device=dev)
                _call("dwc_post_bwd_reduce", C.byref(ds), C.byref(ys), L.ptr(coef), act, splits, L.ptr(red),
                      2 if in_reduce else pre, L.stream())
                bco = torch.empty(n * c * 4, dtype=torch.float32, device=dev)
                if kind == NORM_ADAIN:
                    dnw = torch.empty(n, c, dtype=torch.float32, device=dev)
                    dnb = torch.empty(n, c, dtype=torch.float32, device=dev)
                    gw, gb = dnw, dnb
                elif kind == NORM_LN:
                    gw, gb = ctx.ln_mod.grad_buffers()
                else:
                    gw = gb = None
                # reductions -> coefficients (+ AdaIN / LayerNorm parameter gradients) -> apply
                _call("dwc_post_bwd_apply_norm", C.byref(ds), C.byref(ys), L.ptr(coef), kind, L.ptr(red), splits,
                      L.f32(eps), L.ptr(nw), L.ptr(gw), L.ptr(gb), L.ptr(bco), act, C.byref(dys),
                      C.byref(drs) if drs is not None else None, pre, L.stream())
            else:
                _call("dwc_post_bwd_apply", C.byref(ds), C.byref(ys), None, None, act, C.byref(dys),
                      C.byref(drs) if drs is not None else None, pre, L.stream())
        dres_t = dres.t if dres is not None else None
        if dres is not None and ctx.skip_box is not None:
            ctx.skip_box["dres"] = dres          # handed to ConvFn.backward of the block's first conv (see there)
            dres_t = None
        return (dy.t, dnw, dnb, dres_t, None, None, None, None, None, None, None, None, None, None)


def post(y: HB, kind=NORM_NONE, act=ACT_NONE, nw=None, nb=None, res: Optional[HB] = None, out_halo=0, out_layout=0,
         ln_mod=None, eps=1e-5, anchor=None, skip_box=None) -> HB:
    t = PostFn.apply(y.t, nw, nb, res.t if res is not None else None, anchor, y, kind, act, res, out_halo, out_layout,
                     ln_mod, eps, skip_box)
    return HB(t, y.n, y.h, y.w, y.c, out_halo, out_layout)


class UpsamplePadFn(torch.autograd.Function):
    """nn.Upsample(2, bilinear) + ReflectionPad2d of the next block (networks_v2.py:154)."""

    @staticmethod
    def forward(ctx, x_t, x: HB, out_halo):
        xh = HB(x_t, x.n, x.h, x.w, x.c, x.halo, x.layout)
        out = HB.empty(x.n, 2 * x.h, 2 * x.w, x.c, out_halo, 0, x_t.dtype, x_t.device)
        xs, os_ = xh.struct(), out.struct()
        _call("dwc_upsample_pad_fwd", C.byref(xs), C.byref(os_), L.stream())
        ctx.meta = (x.n, x.h, x.w, x.c, x.halo, x.layout, out_halo)
        return out.t

    @staticmethod
    def backward(ctx, dout_t):
        n, h, w, c, halo, layout, out_halo = ctx.meta
        dout = HB(dout_t.contiguous(), n, 2 * h, 2 * w, c, out_halo, 0)
        dx = HB.empty(n, h, w, c, halo, layout, dout_t.dtype, dout_t.device)
        ds, xs = dout.struct(), dx.struct()
        pre = _prefold(dout)
        _call("dwc_upsample_pad_bwd", C.byref(ds), C.byref(xs), pre, L.stream())
        return dx.t, None, None


def upsample_pad(x: HB, out_halo) -> HB:
    t = UpsamplePadFn.apply(x.t, x, out_halo)
    return HB(t, x.n, 2 * x.h, 2 * x.w, x.c, out_halo, 0)


class HeadsFn(torch.autograd.Function):
    """fused decoder heads: tanh(image_content) | sigmoid(image_attention) (networks_v2.py:162-169)."""

    @staticmethod
    def forward(ctx, y_t, y: HB, on_att_grad):
        ctx.set_materialize_grads(False)
        ctx.on_att_grad = on_att_grad
        yh = HB(y_t, y.n, y.h, y.w, y.c, y.halo, 0)
        img = torch.empty(y.n, y.c - 1, y.h, y.w, dtype=torch.float32, device=y_t.device)
        att = torch.empty(y.n, 1, y.h, y.w, dtype=torch.float32, device=y_t.device)
        ys = yh.struct()
        _call("dwc_heads_fwd", C.byref(ys), L.ptr(img), L.ptr(att), L.stream())
        ctx.meta = (y.n, y.h, y.w, y.c, y.halo, y_t.dtype)
        ctx.save_for_backward(img, att)
        return img, att

    @staticmethod
    def backward(ctx, dimg, datt):
        n, h, w, c, halo, dtype = ctx.meta
        img, att = ctx.saved_tensors
        dy = HB.empty(n, h, w, c, halo, 0, dtype, img.device)
        dys = dy.struct()
        dimg = dimg.contiguous() if dimg is not None else None
        datt = datt.contiguous() if datt is not None else None
        if datt is not None and ctx.on_att_grad is not None:
            ctx.on_att_grad()
        _call("dwc_heads_bwd", L.ptr(dimg), L.ptr(datt), L.ptr(img), L.ptr(att), C.byref(dys), L.stream())
        return dy.t, None, None


def heads(y: HB, on_att_grad=None):
    return HeadsFn.apply(y.t, y, on_att_grad)


class BlendFn(torch.autograd.Function):
    """x = img*att + real*(1-att)  (solver.py:160-161)."""

    @staticmethod
    def forward(ctx, img, att, real):
        _require_cuda(img)
        img, att, real = img.contiguous(), att.contiguous(), real.contiguous().float()
        n, c, h, w = img.shape
        out = torch.empty_like(img)
        _call("dwc_blend_fwd", L.ptr(img), L.ptr(att), L.ptr(real), L.ptr(out), n, c, h * w, L.stream())
        ctx.save_for_backward(img, att, real)
        return out

    @staticmethod
    def backward(ctx, dout):
        img, att, real = ctx.saved_tensors
        n, c, h, w = img.shape
        dimg, datt = torch.empty_like(img), torch.empty_like(att)
        _call("dwc_blend_bwd", L.ptr(dout.contiguous()), L.ptr(img), L.ptr(att), L.ptr(real), L.ptr(dimg), L.ptr(datt),
              n, c, h * w, L.stream())
        return dimg, datt, None


def blend(img, att, real):
    return BlendFn.apply(img, att, real)


class ReluGapFn(torch.autograd.Function):
    """ReLU + AdaptiveAvgPool2d(1) at the end of the style encoder (networks_v2.py:113)."""

    @staticmethod
    def forward(ctx, y_t, y: HB):
        yh = HB(y_t, y.n, y.h, y.w, y.c, y.halo, 0)
        out = torch.empty(y.n, y.c, dtype=torch.float32, device=y_t.device)
        ys = yh.struct()
        _call("dwc_relu_gap_fwd", C.byref(ys), L.ptr(out), L.stream())
        ctx.meta = (y.n, y.h, y.w, y.c, y.halo)
        ctx.save_for_backward(y_t)
        return out

    @staticmethod
    def backward(ctx, dout):
        n, h, w, c, halo = ctx.meta
        (y_t,) = ctx.saved_tensors
        yh = HB(y_t, n, h, w, c, halo, 0)
        dy = HB.empty(n, h, w, c, halo, 0, y_t.dtype, y_t.device)
        ys, dys = yh.struct(), dy.struct()
        _call("dwc_relu_gap_bwd", L.ptr(dout.contiguous()), C.byref(ys), C.byref(dys), L.stream())
        return dy.t, None


def relu_gap(y: HB):
    return ReluGapFn.apply(y.t, y)


# --------------------------------------------------------------------------------------------
# dense layers
# --------------------------------------------------------------------------------------------

def sgemm(m, n, k, alpha, a, a_sm, a_sk, b, b_sk, b_sn, beta, c, c_sm, c_sn, bias=None, act=0):
    need = int(L.lib().dwc_sgemm_workspace_bytes(m, n, k))
    ws = RT.workspace(need, c.device) if need else None
    _call("dwc_sgemm_ws", m, n, k, alpha, L.ptr(a), L.dt(a), a_sm, a_sk, L.ptr(b), b_sk, b_sn, beta, L.ptr(c), c_sm,
          c_sn, L.ptr(bias), act, L.ptr(ws), need, L.stream())


class LinearFn(torch.autograd.Function):
    """out = act(x @ W^T + b) with fp32 accumulation (nn.Linear: networks.py:496-499, networks_v2.py:117-127).
    x may be fp32 or the compute dtype; W [N,K], b [N] are fp32 (views of the flat parameter buffer)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, owner, anchor):
        _require_cuda(x)
        x = x.contiguous()
        m, k = x.shape
        n = weight.shape[0]
        out = torch.empty(m, n, dtype=torch.float32, device=x.device)
        sgemm(m, n, k, 1.0, x, k, 1, weight, 1, k, 0.0, out, n, 1, bias, act)
        ctx.act, ctx.owner = act, owner
        ctx.save_for_backward(x, weight, out if act else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, out = ctx.saved_tensors
        m, k = x.shape
        n = weight.shape[0]
        dout = dout.contiguous()
        if ctx.act:
            dz = torch.empty_like(dout)
            _call("dwc_relu_bwd", L.ptr(dout), L.ptr(out), L.ptr(dz), L.i64(dz.numel()), L.stream())
            dout = dz
        dx = None
        if ctx.needs_input_grad[0]:
            dxf = torch.empty(m, k, dtype=torch.float32, device=x.device)
            sgemm(m, k, n, 1.0, dout, n, 1, weight, k, 1, 0.0, dxf, k, 1)
            dx = dxf if x.dtype == torch.float32 else _cast(dxf, x.dtype)
        if ctx.needs_input_grad[5]:
            gw, gb = ctx.owner()
            # dW[n,k] += sum_m dout[m,n] x[m,k]   ->  A = x^T? use A = dout^T (fp32): C[n,k] = sum_m dout[m,n] x[m,k]
            xf = x if x.dtype == torch.float32 else _cast(x, torch.float32)
            sgemm(n, k, m, 1.0, dout, 1, n, xf, k, 1, 1.0, gw, k, 1)
            if gb is not None:
                _call("dwc_colsum", m, n, L.ptr(dout), n, 1, L.ptr(gb), 1, L.stream())
        return dx, None, None, None, None, None


def _cast(t, dtype):
    out = torch.empty(t.shape, dtype=dtype, device=t.device)
    _call("dwc_cast", L.ptr(t.contiguous()), L.dt(t), L.ptr(out), L.dt(dtype), L.i64(t.numel()), L.stream())
    return out


def linear(x, weight, bias, act, grad_owner, anchor):
    """grad_owner() -> (weight.grad view, bias.grad view or None); anchor: the nn.Parameter behind `weight`."""
    return LinearFn.apply(x, weight, bias, act, grad_owner, anchor)


class MulFn(torch.autograd.Function):
    """element-wise product with a constant mask (dropout)."""

    @staticmethod
    def forward(ctx, x, mask):
        out = torch.empty_like(x)
        _call("dwc_mul", L.ptr(x.contiguous()), L.ptr(mask), L.ptr(out), L.i64(x.numel()), L.stream())
        ctx.save_for_backward(mask)
        return out

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        dx = torch.empty_like(dout)
        _call("dwc_mul", L.ptr(dout.contiguous()), L.ptr(mask), L.ptr(dx), L.i64(dx.numel()), L.stream())
        return dx, None


# --------------------------------------------------------------------------------------------
# GMM + losses
# --------------------------------------------------------------------------------------------

def gmm_sample(mu, eps, stddev, c_dim):
    """z[b, j*c_dim+k] = mu[b,j] + stddev*eps[0,k,b,j]  (tools.py:65-70); no gradient."""
    _require_cuda(mu)
    b, ncls = mu.shape
    z = torch.empty(b, ncls * c_dim, dtype=torch.float32, device=mu.device)
    _call("dwc_gmm_sample", L.ptr(mu.contiguous().float()), L.ptr(eps.contiguous().float()), stddev, L.ptr(z), b, ncls,
          c_dim, L.stream())
    return z


class GmmKlFn(torch.autograd.Function):
    """gmm_kl_distance_sp (gmm.py:13-22) on concatenated [B, ncls*cdim] mu / logvar."""

    @staticmethod
    def forward(ctx, mu, lv, c, sigma, ncls, cdim):
        _require_cuda(mu)
        mu, lv, c = mu.contiguous().float(), lv.contiguous().float(), c.contiguous().float()
        loss = torch.empty(1, dtype=torch.float32, device=mu.device)
        dmu, dlv = torch.empty_like(mu), torch.empty_like(lv)
        _call("dwc_gmm_kl", L.ptr(mu), L.ptr(lv), L.ptr(c), sigma, L.ptr(loss), L.ptr(dmu), L.ptr(dlv), mu.shape[0],
              ncls, cdim, L.stream())
        ctx.save_for_backward(dmu, dlv)
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        dmu, dlv = ctx.saved_tensors
        return ScaleFn_apply(dmu, g), ScaleFn_apply(dlv, g), None, None, None, None


def ScaleFn_apply(t, g):
    """t * g for a 0-dim device scalar g, through dwc_mul on an expanded copy-free path."""
    out = torch.empty_like(t)
    gg = g.reshape(1).float().expand(t.numel()).contiguous() if t.numel() <= 4096 else None
    if gg is None:
        raise RuntimeError("ScaleFn_apply is for small tensors only")
    _call("dwc_mul", L.ptr(t), L.ptr(gg), L.ptr(out), L.i64(t.numel()), L.stream())
    return out


def gmm_kl(mu_cat, lv_cat, c, sigma, ncls, cdim):
    return GmmKlFn.apply(mu_cat, lv_cat, c, float(sigma), ncls, cdim)


class L1Fn(torch.autograd.Function):
    """mean |a - b| (solver.py:113-114,127-132), a/b fp32 or compute dtype, any (matching) layout."""

    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a)
        assert a.shape == b.shape and _same_layout(a, b), (a.shape, b.shape, a.stride(), b.stride())
        loss = torch.zeros(600, dtype=torch.float32, device=a.device)      # DWC_L1_SCRATCH: result + ordered partials
        _call("dwc_l1_loss_fwd", L.ptr(a), L.dt(a), L.ptr(b), L.dt(b), L.i64(a.numel()), L.ptr(loss), L.stream())
        ctx.save_for_backward(a, b)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        gs = g.reshape(1).float().contiguous()
        _call("dwc_l1_loss_bwd", L.ptr(a), L.dt(a), L.ptr(b), L.dt(b), L.i64(a.numel()), L.ptr(gs), L.ptr(da), L.ptr(db),
              L.stream())
        return da, db


def _same_layout(a, b):
    """Equal strides on every dimension that has more than one element (size-1 dimensions carry arbitrary strides)."""
    return all(sa == sb for sa, sb, n in zip(a.stride(), b.stride(), a.shape) if n > 1)


def _dense(t):
    """The tensor itself if it is a dense (possibly permuted) block, else a contiguous copy."""
    return t if t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last) else t.contiguous()


def l1_loss(a, b):
    a, b = _dense(a), _dense(b)
    if not _same_layout(a, b):
        a, b = a.contiguous(), b.contiguous()
    return L1Fn.apply(a, b)


class MseConstFn(torch.autograd.Function):
    """mean (x - target)^2 : LSGAN terms (networks.py:131,158)."""

    @staticmethod
    def forward(ctx, x, target):
        _require_cuda(x)
        x = x.contiguous().float()
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        _call("dwc_mse_const_loss_fwd", L.ptr(x), target, L.i64(x.numel()), L.ptr(loss), L.stream())
        ctx.target = target
        ctx.save_for_backward(x)
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        gs = g.reshape(1).float().contiguous()
        _call("dwc_mse_const_loss_bwd", L.ptr(x), ctx.target, L.i64(x.numel()), L.ptr(gs), L.ptr(dx), L.stream())
        return dx, None


def mse_const(x, target):
    return MseConstFn.apply(x, float(target))


class BceLogitsFn(torch.autograd.Function):
    """F.binary_cross_entropy_with_logits(x, y, 'mean') (networks.py:83)."""

    @staticmethod
    def forward(ctx, x, y):
        _require_cuda(x)
        x, y = x.contiguous().float(), y.contiguous().float()
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        _call("dwc_bce_logits_loss_fwd", L.ptr(x), L.ptr(y), L.i64(x.numel()), L.ptr(loss), L.stream())
        ctx.save_for_backward(x, y)
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        dx = torch.empty_like(x)
        gs = g.reshape(1).float().contiguous()
        _call("dwc_bce_logits_loss_bwd", L.ptr(x), L.ptr(y), L.i64(x.numel()), L.ptr(gs), L.ptr(dx), L.stream())
        return dx, None


def bce_logits(x, y):
    return BceLogitsFn.apply(x, y)


class AdvLossFn(torch.autograd.Function):
    """All adversarial terms of one discriminator scale in one kernel per direction (dwc_adv_loss_fwd/bwd): spec is a
    tuple of (kind, row0, row1, target, weight) - kind 0: weight * mean (src[row0:row1] - target)^2, kind 1: weight *
    mean BCE-with-logits(cls[row0:row1], labels) (networks.py:116-170 as the Solver batches the passes)."""

    @staticmethod
    def forward(ctx, src, cls, labels, spec):
        _require_cuda(src)
        assert src.dtype == torch.float32 and cls.dtype == torch.float32 and src.is_contiguous() and cls.is_contiguous()
        n = src.shape[0]
        labels = labels.contiguous().float()
        terms = (L.AdvTerm * len(spec))(*[L.AdvTerm(int(k), int(a), int(b), float(t), float(w)) for k, a, b, t, w in spec])
        loss = torch.empty(1, dtype=torch.float32, device=src.device)
        _call("dwc_adv_loss_fwd", L.ptr(src), n, src.numel() // n, L.ptr(cls), cls.shape[0], cls.numel() // cls.shape[0],
              L.ptr(labels), terms, len(spec), L.ptr(loss), L.stream())
        ctx.spec = spec
        ctx.save_for_backward(src, cls, labels)
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        src, cls, labels = ctx.saved_tensors
        spec = ctx.spec
        n = src.shape[0]
        terms = (L.AdvTerm * len(spec))(*[L.AdvTerm(int(k), int(a), int(b), float(t), float(w)) for k, a, b, t, w in spec])
        dsrc, dcls = torch.empty_like(src), torch.empty_like(cls)
        gs = g.reshape(1).float().contiguous()
        _call("dwc_adv_loss_bwd", L.ptr(src), n, src.numel() // n, L.ptr(cls), cls.shape[0], cls.numel() // cls.shape[0],
              L.ptr(labels), terms, len(spec), L.ptr(gs), L.ptr(dsrc), L.ptr(dcls), L.stream())
        return dsrc, dcls, None, None


def adv_loss(src, cls, labels, spec):
    """src: the discriminator's patch output [N, 1, h, w] (any shape with N first), cls [N, num_cls]."""
    return AdvLossFn.apply(src, cls, labels, tuple(spec))
