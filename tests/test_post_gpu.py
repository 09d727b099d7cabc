"""GPU parity of the norm/activation/residual/pad pass (dwc_post_* kernels), bilinear upsample+pad, image pad and
the decoder-head / blend kernels against plain torch ops (networks.py:514-522,531,545,706-752; networks_v2.py:154)."""
import pytest
import torch
import torch.nn.functional as F

import dwc_gan_b200
from dwc_gan_b200 import ops
from dwc_gan_b200.plan import HB
from oracle import dwc_oracle as O

pytestmark = pytest.mark.gpu


def to_hb(x_nchw, halo, dtype):
    """NCHW cpu tensor -> HB on cuda with garbage halo, requires grad on its storage tensor."""
    n, c, h, w = x_nchw.shape
    t = torch.randn(n, h + 2 * halo, w + 2 * halo, c)
    t[:, halo:halo + h, halo:halo + w, :] = x_nchw.permute(0, 2, 3, 1)
    t = t.to(dtype).cuda().requires_grad_(True)
    return HB(t, n, h, w, c, halo, 0)


def padded_to_nchw(hb: HB, t):
    return hb.like(t).padded_nhwc().permute(0, 3, 1, 2)


CASES = [
    # n, c, h, w, kind, act, res, out_halo, out_layout, y_halo
    (2, 64, 16, 16, 1, 1, False, 1, 0, 2),
    (2, 64, 16, 16, 1, 0, True, 1, 0, 2),
    (3, 512, 8, 8, 0, 2, False, 1, 1, 1),
    (3, 256, 8, 8, 0, 2, False, 1, 1, 1),
    (3, 512, 4, 4, 0, 2, False, 0, 0, 1),
    (2, 256, 32, 32, 2, 1, False, 1, 0, 2),
    (2, 256, 32, 32, 2, 0, True, 0, 0, 2),
    (2, 128, 16, 16, 3, 1, False, 0, 0, 4),
    (2, 64, 32, 32, 3, 1, False, 3, 0, 4),
    (2, 64, 16, 16, 1, 1, False, 1, 1, 6),
    (2, 8, 2, 2, 0, 1, False, 1, 1, 1),
    # rows of 16 KB (the row-streaming kernels at their in-network shapes), parity-plane output / gradient
    (2, 64, 128, 128, 1, 1, False, 3, 1, 4),
    (1, 128, 64, 64, 3, 1, False, 2, 0, 4),
    (2, 256, 32, 32, 1, 1, True, 1, 0, 2),
    # 32 KB rows: two segments per row
    (1, 64, 8, 256, 1, 1, True, 1, 0, 2),
    (1, 64, 6, 256, 3, 2, False, 2, 0, 1),
]


@pytest.mark.parametrize("n,c,h,w,kind,act,use_res,oh,ol,yh", CASES)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_post(n, c, h, w, kind, act, use_res, oh, ol, yh, mode):
    dwc_gan_b200.set_mode(mode)
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = 2e-5 if mode == "fp32" else 2e-2
    torch.manual_seed(0)
    y = (torch.randn(n, c, h, w) * 1.5 + 0.3).to(dtype).float()
    res = torch.randn(n, c, h, w).to(dtype).float()
    nw = torch.randn(n, c) if kind == 2 else (torch.rand(c) if kind == 3 else None)
    nb = torch.randn(n, c) if kind == 2 else (torch.randn(c) if kind == 3 else None)
    # ---- reference (double)
    yr = y.double().requires_grad_(True)
    rr = res.double().requires_grad_(True)
    nwr = nw.double().requires_grad_(True) if nw is not None else None
    nbr = nb.double().requires_grad_(True) if nb is not None else None
    if kind == 1:
        z = O.inst_norm(yr)
    elif kind == 2:
        z = O.adain(yr, nwr, nbr)
    elif kind == 3:
        z = O.layer_norm_munit(yr, nwr, nbr)
    else:
        z = yr
    z = torch.relu(z) if act == 1 else (F.leaky_relu(z, 0.1) if act == 2 else z)
    if use_res:
        z = z + rr
    out_ref = F.pad(z, (oh, oh, oh, oh), mode="reflect") if oh else z
    dout = torch.randn_like(out_ref).to(dtype).double()
    out_ref.backward(dout)
    # ---- ours
    yhb = to_hb(y, yh, dtype)
    rhb = to_hb(res, 1, dtype) if use_res else None

    class LN:  # stand-in for networks.LayerNorm.grad_buffers
        gw = torch.zeros(c, device="cuda")
        gb = torch.zeros(c, device="cuda")

        def grad_buffers(self):
            return self.gw, self.gb
    nwc = nw.cuda().requires_grad_(kind == 2) if nw is not None else None
    nbc = nb.cuda().requires_grad_(kind == 2) if nb is not None else None
    out = ops.post(yhb, kind, act, nwc, nbc, rhb, oh, ol, LN() if kind == 3 else None, 1e-5)
    got = padded_to_nchw(out, out.t.detach()).double().cpu()
    assert (got - out_ref.detach()).abs().max() < tol * out_ref.abs().max(), (got - out_ref.detach()).abs().max()
    # upstream gradient in the physical layout of `out`
    d_nhwc = dout.permute(0, 2, 3, 1).to(dtype)
    if ol == 0:
        d_phys = d_nhwc.contiguous()
    else:
        d_phys = torch.stack([d_nhwc[:, py::2, px::2, :] for py in range(2) for px in range(2)], dim=1).contiguous()
    out.t.backward(d_phys.cuda())
    gy = yhb.t.grad
    gi = gy[:, yh:yh + h, yh:yh + w, :].permute(0, 3, 1, 2).double().cpu()
    scale = yr.grad.abs().max()
    assert (gi - yr.grad).abs().max() < tol * 4 * scale, ((gi - yr.grad).abs().max(), scale)
    halo_sum = gy.double().abs().sum() - gy[:, yh:yh + h, yh:yh + w, :].double().abs().sum()
    assert float(halo_sum) == 0.0, "gradient halo must be zero"
    if use_res:
        gr = rhb.t.grad
        assert (gr[:, 1:1 + h, 1:1 + w, :].permute(0, 3, 1, 2).double().cpu() - rr.grad).abs().max() < tol * 4 * rr.grad.abs().max()
        assert float(gr.double().abs().sum() - gr[:, 1:1 + h, 1:1 + w, :].double().abs().sum()) == 0.0
    if kind == 2:
        assert (nwc.grad.double().cpu() - nwr.grad).abs().max() < tol * 8 * nwr.grad.abs().max()
        assert (nbc.grad.double().cpu() - nbr.grad).abs().max() < tol * 8 * nbr.grad.abs().max()
    if kind == 3:
        assert (LN.gw.double().cpu() - nwr.grad).abs().max() < tol * 8 * nwr.grad.abs().max()
        assert (LN.gb.double().cpu() - nbr.grad).abs().max() < tol * 8 * nbr.grad.abs().max()


@pytest.mark.parametrize("n,c,h,w,kind,act,use_res,oh,ol,yh", [c for c in CASES if c[4] in (1, 2)])
def test_post_three_pass(n, c, h, w, kind, act, use_res, oh, ol, yh):
    """The same sites with the single-kernel small-map path switched off: statistics / finalize / apply passes."""
    old = ops.RT.use_fused_norm
    ops.RT.use_fused_norm = False
    try:
        test_post(n, c, h, w, kind, act, use_res, oh, ol, yh, "bf16")
    finally:
        ops.RT.use_fused_norm = old


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("n,c,h,w,oh", [(2, 64, 8, 8, 2), (1, 128, 16, 32, 2), (2, 8, 4, 4, 0),
                                       # row-streaming forward at the in-network shapes, odd heights, one row
                                       (2, 256, 32, 32, 2), (1, 128, 64, 64, 2), (3, 64, 5, 16, 1), (2, 64, 1, 16, 0),
                                       (1, 64, 7, 256, 3)])
def test_upsample_pad(mode, n, c, h, w, oh):
    dwc_gan_b200.set_mode(mode)
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = 2e-5 if mode == "fp32" else 2e-2
    torch.manual_seed(0)
    x = torch.randn(n, c, h, w).to(dtype).float()
    xr = x.double().requires_grad_(True)
    up = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    ref = F.pad(up, (oh, oh, oh, oh), mode="reflect") if oh else up
    dout = torch.randn_like(ref).to(dtype).double()
    ref.backward(dout)
    xh = to_hb(x, 0, dtype)
    out = ops.upsample_pad(xh, oh)
    got = out.t.detach().permute(0, 3, 1, 2).double().cpu()
    assert (got - ref.detach()).abs().max() < tol * ref.abs().max()
    out.t.backward(dout.permute(0, 2, 3, 1).to(dtype).contiguous().cuda())
    g = xh.t.grad.permute(0, 3, 1, 2).double().cpu()
    assert (g - xr.grad).abs().max() < tol * 4 * xr.grad.abs().max()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("pool,pad,layout", [(1, 3, 0), (1, 1, 1), (2, 1, 1)])
def test_image_pad(mode, pool, pad, layout):
    dwc_gan_b200.set_mode(mode)
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = 1e-6 if mode == "fp32" else 1e-2
    torch.manual_seed(0)
    x = torch.rand(2, 3, 16, 16) * 2 - 1
    xr = x.double().requires_grad_(True)
    pooled = F.avg_pool2d(xr, pool) if pool > 1 else xr
    ref = F.pad(pooled, (pad,) * 4, mode="reflect")
    dout = torch.randn_like(ref).to(dtype).double()
    ref.backward(dout)
    xc = x.cuda().requires_grad_(True)
    out = ops.image_pad(xc, pool, pad, layout)
    got = padded_to_nchw(out, out.t.detach()).double().cpu()
    assert (got - ref.detach()).abs().max() < tol * 2
    d_nhwc = dout.permute(0, 2, 3, 1).to(dtype)
    d_phys = d_nhwc.contiguous() if layout == 0 else torch.stack(
        [d_nhwc[:, py::2, px::2, :] for py in range(2) for px in range(2)], dim=1).contiguous()
    out.t.backward(d_phys.cuda())
    assert (xc.grad.double().cpu() - xr.grad).abs().max() < 1e-5 * xr.grad.abs().max() + 1e-6


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_heads_blend_gap(mode):
    dwc_gan_b200.set_mode(mode)
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = 2e-5 if mode == "fp32" else 2e-2
    torch.manual_seed(0)
    n, h, w = 2, 16, 16
    y = torch.randn(n, 4, h, w).to(dtype).float()
    real = torch.rand(n, 3, h, w) * 2 - 1
    yr = y.double().requires_grad_(True)
    img_r, att_r = torch.tanh(yr[:, :3]), torch.sigmoid(yr[:, 3:])
    out_r = img_r * att_r + real.double() * (1 - att_r)
    d = torch.randn_like(out_r)
    out_r.backward(d)
    yh = to_hb(y, 6, dtype)
    img, att = ops.heads(yh)
    out = ops.blend(img, att, real.cuda())
    assert (out.double().cpu() - out_r.detach()).abs().max() < tol
    out.backward(d.float().cuda())
    g = yh.t.grad[:, 6:6 + h, 6:6 + w, :].permute(0, 3, 1, 2).double().cpu()
    assert (g - yr.grad).abs().max() < tol * 4 * yr.grad.abs().max()
    # relu + global average pool
    y2 = torch.randn(3, 64, 4, 4).to(dtype).float()
    y2r = y2.double().requires_grad_(True)
    ref = torch.relu(y2r).mean((2, 3))
    dd = torch.randn_like(ref)
    ref.backward(dd)
    y2h = to_hb(y2, 1, dtype)
    gp = ops.relu_gap(y2h)
    assert (gp.double().cpu() - ref.detach()).abs().max() < tol
    gp.backward(dd.float().cuda())
    g2 = y2h.t.grad[:, 1:5, 1:5, :].permute(0, 3, 1, 2).double().cpu()
    assert (g2 - y2r.grad).abs().max() < tol * 4 * y2r.grad.abs().max()


@pytest.mark.parametrize("n,h,w,c,halo", [(3, 32, 32, 256, 1), (2, 64, 64, 128, 2), (2, 128, 128, 64, 3), (2, 20, 24, 128, 1)])
@pytest.mark.parametrize("act", [0, 1])
def test_halo_fold_inside_the_backward_reduction(n, h, w, c, halo, act):
    """dwc_post_bwd_reduce(prefolded=2) folds the reflect-halo gradient while it streams dout: the partial sums and the
    folded buffer equal, bit for bit, what dwc_fold_halo followed by the plain reduction leaves."""
    import ctypes as C
    from dwc_gan_b200 import _lib as L
    from dwc_gan_b200.plan import HB
    torch.manual_seed(7)
    dt = torch.bfloat16
    y = HB.empty(n, h, w, c, 0, 0, dt, "cuda")
    y.t.copy_(torch.randn(y.t.shape, device="cuda"))
    d1 = HB.empty(n, h, w, c, halo, 0, dt, "cuda")
    d1.t.copy_(torch.randn(d1.t.shape, device="cuda"))
    d2 = d1.like(d1.t.clone())
    coef = torch.randn(n * c, 4, device="cuda")
    splits = 8
    ys, s1, s2 = y.struct(), d1.struct(), d2.struct()
    assert L.lib().dwc_post_bwd_reduce_can_fold(C.byref(s2), C.byref(ys))
    r1 = torch.full((n * splits * c * 2,), float("nan"), device="cuda")
    r2 = torch.full((n * splits * c * 2,), float("nan"), device="cuda")
    L.check(L.lib().dwc_fold_halo(C.byref(s1), L.stream()), "fold")
    L.check(L.lib().dwc_post_bwd_reduce(C.byref(s1), C.byref(ys), L.ptr(coef), act, splits, L.ptr(r1), 1, L.stream()), "red")
    L.check(L.lib().dwc_post_bwd_reduce(C.byref(s2), C.byref(ys), L.ptr(coef), act, splits, L.ptr(r2), 2, L.stream()), "red")
    torch.cuda.synchronize()
    assert torch.equal(d1.t, d2.t)
    assert torch.equal(r1, r2)
