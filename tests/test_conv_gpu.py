"""GPU parity of the gconv / wgrad kernels (SIMT fp32, SIMT bf16, tcgen05 bf16) against
F.conv2d(F.pad(x, reflect)) and its autograd gradients, through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

import dwc_gan_b200
from dwc_gan_b200 import _lib as L
from dwc_gan_b200 import plan as P
from dwc_gan_b200.plan import HB
from tests import emu

pytestmark = pytest.mark.gpu

_ws = {}


def workspace(nbytes):
    t = _ws.get("t")
    if t is None or t.numel() * 4 < nbytes:
        t = torch.empty((nbytes + 3) // 4 + 1024, dtype=torch.float32, device="cuda")
        _ws["t"] = t
    return t


def pack(w_krsc_dev, mode, dtype, rows_padded, cout, k, cin):
    if mode == 0:
        out = torch.empty(rows_padded, k * k * cin, dtype=dtype, device="cuda")
    elif mode == 1:
        out = torch.empty(rows_padded, k * k * cout, dtype=dtype, device="cuda")
    else:
        out = torch.empty(4, rows_padded, 4 * cout, dtype=dtype, device="cuda")
    L.check(L.lib().dwc_pack_weights(L.ptr(w_krsc_dev), cout, k, k, cin, mode, L.ptr(out), L.dt(dtype), rows_padded,
                                     L.stream()), "pack")
    return out


CASES = [
    # n, h, w, cin, cout, k, s, p
    (2, 16, 16, 64, 128, 3, 1, 1),
    (2, 16, 16, 128, 256, 3, 1, 1),
    (1, 32, 32, 64, 64, 5, 1, 2),
    (2, 16, 16, 64, 4, 7, 1, 3),
    (2, 16, 16, 64, 128, 4, 2, 1),
    (3, 8, 8, 128, 64, 4, 2, 1),
    (9, 4, 4, 256, 512, 4, 2, 1),
    (2, 16, 16, 3, 64, 7, 1, 3),
    (2, 16, 16, 3, 64, 4, 2, 1),
    (1, 32, 32, 256, 256, 3, 1, 1),
    (3, 8, 8, 512, 512, 4, 2, 1),
    (3, 16, 16, 256, 512, 4, 2, 1),
    (2, 16, 16, 64, 4, 7, 1, 3),
    (2, 16, 16, 128, 64, 5, 1, 2),      # G9 geometry: tap-grouped weight gradient (N = 4 taps x 64 channels)
    (3, 20, 12, 128, 64, 3, 1, 1),
    # few tiles, long K: the one-wave kernel splits K over a cluster (DSMEM reduction), 8 / 8 / 4 / 4 / 2 ways
    (16, 8, 8, 512, 512, 4, 2, 1),
    (4, 16, 16, 128, 128, 4, 2, 1),
    (4, 16, 16, 64, 64, 4, 2, 1),
    (16, 32, 32, 256, 256, 4, 2, 1),
    (8, 16, 16, 256, 256, 3, 1, 1),
]
MODES = [("simt", torch.float32), ("simt", torch.bfloat16), ("tc", torch.bfloat16)]


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,p", CASES)
@pytest.mark.parametrize("backend,dtype", MODES)
def test_conv_fwd_dgrad_wgrad(n, h, w, cin, cout, k, s, p, backend, dtype):
    tc = backend == "tc"
    be = L.TC if tc else L.SIMT
    if tc and not L.lib().dwc_tc_available():
        pytest.fail("tcgen05 path unavailable on this device")
    torch.manual_seed(1)
    x = torch.randn(n, cin, h, w).to(dtype).double()
    wt = (torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)).to(dtype).double()
    bias = torch.randn(cout)
    xpad = F.pad(x, (p, p, p, p), mode="reflect").requires_grad_(True)
    wt_r = wt.clone().requires_grad_(True)
    y_ref = F.conv2d(xpad, wt_r, bias.double(), stride=s)
    ho, wo = y_ref.shape[2:]
    dy = torch.randn(n, cout, ho, wo).to(dtype).double()
    y_ref.backward(dy)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    layout = 0 if s == 1 else 1
    hy = k - 1 if s == 1 else 1
    w_krsc = wt.permute(0, 2, 3, 1).contiguous().float().cuda()

    # forward
    fwd_tc = tc and cin % 64 == 0
    rows_p = cout if cout % 64 == 0 else 16
    xp = emu.make_padded(x, p, layout, dtype)
    xp = xp.like(xp.t.cuda())
    y = HB.empty(n, ho, wo, cout, hy, 0, dtype, "cuda", zero=True)
    wf = pack(w_krsc, 0, dtype, rows_p, cout, k, cin)
    assert torch.equal(wf.cpu().double(), emu.pack_fwd(w_krsc.cpu(), rows_p).to(dtype).double())
    pl = P.plan_conv_fwd(xp, wf, cout, rows_p, bias.cuda(), y, k, s, L.TC if fwd_tc else L.SIMT)
    pl.launch()
    torch.cuda.synchronize()
    got = y.interior().permute(0, 3, 1, 2).double().cpu()
    err = (got - y_ref.detach()).abs().max().item() / y_ref.abs().max().item()
    assert err < tol, "fwd rel err %g" % err

    # dgrad
    dg_tc = tc and cout % 64 == 0 and cin % 64 == 0
    dyz = emu.make_zero_haloed(dy, hy, dtype)
    dyz = dyz.like(dyz.t.cuda())
    dxp = HB.empty(n, h, w, cin, p, layout, dtype, "cuda")
    dxp.t.fill_(float("nan"))
    dg_tc = tc and cout % 64 == 0 and (cin % 64 == 0 or cin <= 16)
    rows_d = cin if cin % 64 == 0 or not dg_tc else 16
    wd = pack(w_krsc, 1 if s == 1 else 2, dtype, rows_d, cout, k, cin)
    ref_pack = emu.pack_dgrad_s1(w_krsc.cpu(), rows_d) if s == 1 else emu.pack_dgrad_s2(w_krsc.cpu(), rows_d)
    assert torch.equal(wd.cpu().double(), ref_pack.to(dtype).double())
    for q in P.plan_conv_dgrad(dyz, wd, dxp, k, s, L.TC if dg_tc else L.SIMT, cin_padded=rows_d):
        q.launch()
    torch.cuda.synchronize()
    got = dxp.padded_nhwc().permute(0, 3, 1, 2).double().cpu()
    err = (got - xpad.grad).abs().max().item() / xpad.grad.abs().max().item()
    assert err < tol, "dgrad rel err %g" % err

    # wgrad + dbias (accumulate on top of ones)
    wg_tc = tc and cin % 64 == 0 and cout % 64 == 0
    dw = torch.ones(cout, k, k, cin, device="cuda")
    db = torch.ones(cout, device="cuda")
    wp = P.plan_conv_wgrad(dyz, xp, dw, db, k, s, L.TC if wg_tc else L.SIMT)
    if not wg_tc:
        wp.box = P.choose_box(wo, ho, n, 64)
    wp.launch(workspace)
    torch.cuda.synchronize()
    got = (dw.cpu().double() - 1).permute(0, 3, 1, 2)
    err = (got - wt_r.grad).abs().max().item() / wt_r.grad.abs().max().item()
    assert err < tol, "wgrad rel err %g" % err
    errb = ((db.cpu().double() - 1) - dy.sum((0, 2, 3))).abs().max().item() / dy.sum((0, 2, 3)).abs().max().item()
    assert errb < tol, "dbias rel err %g" % errb


@pytest.mark.parametrize("n,hin,win,cout,flip", [(2, 134, 134, 4, False), (1, 31, 45, 3, False), (2, 140, 140, 3, True),
                                                 (1, 13, 58, 4, True), (3, 7, 7, 1, False)])
def test_conv7_few(n, hin, win, cout, flip):
    """dwc_conv7_few (decoder heads / first-conv image gradient) against F.conv2d on the same bf16-rounded operands."""
    import torch.nn.functional as F
    from dwc_gan_b200 import ops
    dwc_gan_b200.set_mode("bf16")
    torch.manual_seed(3)
    x = torch.randn(n, hin, win, 64, device="cuda").to(torch.bfloat16)
    hout, wout = hin - 6, win - 6
    out = torch.full((n, hout, wout, cout), 7.0, device="cuda", dtype=torch.bfloat16)
    if not flip:
        w = torch.randn(cout, 7, 7, 64, device="cuda") * 0.05          # [o][ky][kx][i]
        bias = torch.randn(cout, device="cuda")
        ops.conv7_few(x, n, hin, win, w.reshape(-1), 0, 49 * 64, 7 * 64, 64, 1, bias, cout, out,
                      (cout, wout * cout, hout * wout * cout))
        w_oihw = w.permute(0, 3, 1, 2)
    else:
        wf = torch.randn(64, 7, 7, cout, device="cuda") * 0.05          # forward weights [co][ky][kx][ci]
        bias = None
        ops.conv7_few(x, n, hin, win, wf.reshape(-1), 6 * 7 * cout + 6 * cout, 1, -7 * cout, -cout, 49 * cout, None,
                      cout, out, (cout, wout * cout, hout * wout * cout))
        w_oihw = wf.flip(1, 2).permute(3, 0, 1, 2)                      # [o = ci][i = co][ky][kx], taps reversed
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w_oihw.to(torch.bfloat16).float(), bias).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, (err, ref.abs().max().item())


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,p", [(2, 32, 32, 256, 256, 3, 1, 1),      # one-wave kernel, BN 256
                                                  (40, 32, 32, 256, 256, 3, 1, 1),     # persistent kernel, BN 256
                                                  (2, 64, 64, 256, 128, 5, 1, 2),      # BN 128
                                                  (24, 64, 64, 64, 128, 4, 2, 1),      # persistent, BN 128, stride 2
                                                  (3, 20, 12, 128, 64, 3, 1, 1),       # partial tiles, BN 64
                                                  (2, 32, 32, 256, 256, 4, 2, 1),      # one-wave, K split 8 ways
                                                  (12, 64, 64, 128, 64, 5, 1, 2)])     # persistent, BN 64
def test_conv_fused_statistics(n, h, w, cin, cout, k, s, p):
    """The tcgen05 epilogue's per-(image, tile, channel) {sum, sum of squares} of the STORED bf16 outputs
    (dwc_gconv_t.stats): the statistics pass of InstanceNorm / AdaIN / LayerNorm without re-reading y."""
    dtype = torch.bfloat16
    torch.manual_seed(2)
    x = torch.randn(n, cin, h, w).to(dtype).double()
    wt = (torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)).to(dtype).double()
    bias = torch.randn(cout)
    layout, hy = (0, k - 1) if s == 1 else (1, 1)
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    w_krsc = wt.permute(0, 2, 3, 1).contiguous().float().cuda()
    xp = emu.make_padded(x, p, layout, dtype)
    xp = xp.like(xp.t.cuda())
    wf = pack(w_krsc, 0, dtype, cout, cout, k, cin)
    y = HB.empty(n, ho, wo, cout, hy, 0, dtype, "cuda", zero=True)
    pl = P.plan_conv_fwd(xp, wf, cout, cout, bias.cuda(), y, k, s, L.TC)
    splits = P.stats_splits(pl)
    assert splits == pl.tiles[0] * pl.tiles[1] > 0
    pl.stats = torch.full((n * splits * cout * 2,), float("nan"), device="cuda")
    pl.launch()
    y2 = HB.empty(n, ho, wo, cout, hy, 0, dtype, "cuda", zero=True)
    P.plan_conv_fwd(xp, wf, cout, cout, bias.cuda(), y2, k, s, L.TC).launch()
    torch.cuda.synchronize()
    assert torch.equal(y.t, y2.t)                                  # the output itself is unchanged by the option
    st = pl.stats.view(n, splits, cout, 2).double().cpu()
    yi = y.interior().double().cpu()                               # [n, ho, wo, cout], the stored values
    bx, by, _ = pl.box
    for ty in range(pl.tiles[1]):
        for tx in range(pl.tiles[0]):
            blk = yi[:, ty * by:(ty + 1) * by, tx * bx:(tx + 1) * bx, :]
            ref = torch.stack([blk.sum((1, 2)), (blk * blk).sum((1, 2))], -1)
            got = st[:, ty * pl.tiles[0] + tx]
            assert float((got - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), (tx, ty)


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,p", [(2, 16, 16, 64, 64, 3, 1, 1),        # one-wave kernel, BN 64
                                                  (4, 16, 16, 128, 128, 4, 2, 1),      # one-wave, K split 8 ways
                                                  (3, 20, 12, 128, 128, 3, 1, 1),      # partial tiles, BN 128
                                                  (2, 32, 32, 64, 256, 4, 2, 1),       # stride 2, BN 256
                                                  (24, 64, 64, 64, 128, 4, 2, 1),      # persistent, BN 128, stride 2
                                                  (12, 64, 64, 128, 64, 5, 1, 2),      # persistent, BN 64
                                                  (40, 32, 32, 256, 512, 3, 1, 1)])    # persistent, BN 256
@pytest.mark.parametrize("act,halo,layout", [(1, 1, 1), (2, 1, 1), (2, 0, 0), (1, 3, 0), (2, 2, 1)])
def test_conv_activated_second_output(n, h, w, cin, cout, k, s, p, act, halo, layout):
    """dwc_gconv_t.out2: the epilogue's activated, reflect-haloed (parity-plane) copy of the output equals, bit for
    bit, what the separate activation + pad pass (dwc_post_fwd) makes of the stored output - the fusion of the
    norm-less Conv2dBlocks of the style encoder / discriminator (networks.py:531,556-567)."""
    import ctypes as C
    dtype = torch.bfloat16
    torch.manual_seed(4)
    x = torch.randn(n, cin, h, w).to(dtype).double()
    wt = (torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)).to(dtype).double()
    bias = torch.randn(cout)
    in_layout, hy = (0, k - 1) if s == 1 else (1, 1)
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    w_krsc = wt.permute(0, 2, 3, 1).contiguous().float().cuda()
    xp = emu.make_padded(x, p, in_layout, dtype)
    xp = xp.like(xp.t.cuda())
    wf = pack(w_krsc, 0, dtype, cout, cout, k, cin)
    y = HB.empty(n, ho, wo, cout, hy, 0, dtype, "cuda", zero=True)
    pl = P.plan_conv_fwd(xp, wf, cout, cout, bias.cuda(), y, k, s, L.TC)
    o2 = HB.empty(n, ho, wo, cout, halo, layout, dtype, "cuda")
    o2.t.fill_(float("nan"))
    pl.out2 = (o2.t, act, halo, layout)
    pl.launch()
    y2 = HB.empty(n, ho, wo, cout, hy, 0, dtype, "cuda", zero=True)
    P.plan_conv_fwd(xp, wf, cout, cout, bias.cuda(), y2, k, s, L.TC).launch()
    ref = HB.empty(n, ho, wo, cout, halo, layout, dtype, "cuda")
    ys, rs = y2.struct(), ref.struct()
    L.check(L.lib().dwc_post_fwd(C.byref(ys), None, act, None, C.byref(rs), L.stream()), "post_fwd")
    torch.cuda.synchronize()
    assert torch.equal(y.t, y2.t)                                  # the primary output is unchanged by the option
    assert not bool(torch.isnan(o2.t.float()).any())               # every halo / plane element was written
    assert torch.equal(o2.t, ref.t)
