"""Each bf16 kernel on its own IS "fp32 arithmetic on bf16-stored operands, rounded once on store".  The whole-step
comparison cannot show that (bf16 storage is chaotic at the 1e-2 level, profiles/r02_bf16_inherent_error.md), so
it is pinned here, per kernel, against fp64 on identical bf16 operands:
  * tcgen05 conv forward / data gradient with fp32 OUTPUT, and the weight gradient: true fp32 accumulation;
  * the same with bf16 output, and the norm + activation + residual + reflect-pad pass: equal to round_bf16(fp64 result)
    except for a handful of elements whose fp32 value sits on a rounding boundary."""
import pytest
import torch
import torch.nn.functional as F

import dwc_gan_b200
from dwc_gan_b200 import _lib as L
from dwc_gan_b200 import ops
from dwc_gan_b200 import plan as P
from dwc_gan_b200.plan import HB
from oracle import dwc_oracle as O
from tests import emu
from tests.test_conv_gpu import pack, workspace
from tests.test_post_gpu import padded_to_nchw, to_hb

pytestmark = pytest.mark.gpu
BT = torch.bfloat16


def _vs(got, ref):
    """(relative L2 error vs fp64, fraction of elements that differ from round_bf16(fp64), relative L2 vs that)"""
    got, ref = got.double().cpu(), ref.double().cpu()
    refq = ref.to(BT).double()
    return (float((got - ref).norm() / ref.norm()), float((got != refq).double().mean()),
            float((got - refq).norm() / refq.norm()))


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,p", [(2, 32, 32, 256, 256, 3, 1, 1), (2, 64, 64, 64, 128, 4, 2, 1),
                                                  (1, 64, 64, 256, 128, 5, 1, 2), (2, 32, 32, 128, 64, 5, 1, 2)])
def test_tcgen05_conv_is_fp32_accumulation_rounded_once(n, h, w, cin, cout, k, s, p):
    dwc_gan_b200.set_mode("bf16")
    torch.manual_seed(1)
    x = torch.randn(n, cin, h, w).to(BT).double()
    wt = (torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)).to(BT).double()
    bias = torch.randn(cout)
    xpad = F.pad(x, (p, p, p, p), mode="reflect").requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    y_ref = F.conv2d(xpad, wr, bias.double(), stride=s)
    ho, wo = y_ref.shape[2:]
    dy = torch.randn(n, cout, ho, wo).to(BT).double()
    y_ref.backward(dy)
    layout, hy = (0, k - 1) if s == 1 else (1, 1)
    w_krsc = wt.permute(0, 2, 3, 1).contiguous().float().cuda()
    xp = emu.make_padded(x, p, layout, BT)
    xp = xp.like(xp.t.cuda())
    wf = pack(w_krsc, 0, BT, cout, cout, k, cin)
    dyz = emu.make_zero_haloed(dy, hy, BT)
    dyz = dyz.like(dyz.t.cuda())
    wd = pack(w_krsc, 1 if s == 1 else 2, BT, cin, cout, k, cin)
    for odt in (torch.float32, BT):
        y = HB.empty(n, ho, wo, cout, hy, 0, odt, "cuda", zero=True)
        P.plan_conv_fwd(xp, wf, cout, cout, bias.cuda(), y, k, s, L.TC).launch()
        dxp = HB.empty(n, h, w, cin, p, layout, odt, "cuda")
        for q in P.plan_conv_dgrad(dyz, wd, dxp, k, s, L.TC, cin_padded=cin):
            q.launch()
        torch.cuda.synchronize()
        for name, got, ref in (("fprop", y.interior().permute(0, 3, 1, 2), y_ref.detach()),
                               ("dgrad", dxp.padded_nhwc().permute(0, 3, 1, 2), xpad.grad)):
            rel, mism, relq = _vs(got, ref)
            if odt == torch.float32:
                assert rel < 2e-5, (name, rel)                   # measured 5e-7 ... 5e-6: fp32 accumulation over K <= 6400
            else:
                assert mism < 1e-2 and relq < 5e-4, (name, mism, relq)   # measured <= 0.37 % of the elements, <= 1.4e-4
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    db = torch.zeros(cout, device="cuda")
    P.plan_conv_wgrad(dyz, xp, dw, db, k, s, L.TC).launch(workspace)
    torch.cuda.synchronize()
    rel, _, _ = _vs(dw.permute(0, 3, 1, 2), wr.grad)
    assert rel < 5e-6, ("wgrad", rel)                            # measured 2e-7 ... 8e-7


@pytest.mark.parametrize("n,c,h,w,kind,act,use_res,oh", [(2, 256, 32, 32, 1, 1, False, 1), (2, 256, 32, 32, 1, 0, True, 0),
                                                         (2, 256, 32, 32, 2, 1, False, 0), (2, 64, 128, 128, 1, 1, False, 1),
                                                         (2, 128, 64, 64, 3, 1, False, 2)])
def test_norm_pass_is_fp32_arithmetic_rounded_once(n, c, h, w, kind, act, use_res, oh):
    dwc_gan_b200.set_mode("bf16")
    torch.manual_seed(0)
    y = (torch.randn(n, c, h, w) * 1.5 + 0.3).to(BT).float()
    res = torch.randn(n, c, h, w).to(BT).float()
    nw = torch.randn(n, c) if kind == 2 else (torch.rand(c) if kind == 3 else None)
    nb = torch.randn(n, c) if kind == 2 else (torch.randn(c) if kind == 3 else None)
    yr, rr = y.double().requires_grad_(True), res.double().requires_grad_(True)
    if kind == 1:
        z = O.inst_norm(yr)
    elif kind == 2:
        z = O.adain(yr, nw.double(), nb.double())
    else:
        z = O.layer_norm_munit(yr, nw.double(), nb.double())
    z = torch.relu(z) if act == 1 else z
    if use_res:
        z = z + rr
    out_ref = F.pad(z, (oh, oh, oh, oh), mode="reflect") if oh else z
    dout = torch.randn_like(out_ref).to(BT).double()
    out_ref.backward(dout)

    class _LN:                                                # gradient sinks of the LayerNorm parameters
        def __init__(self):
            self.g, self.b = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")

        def grad_buffers(self):
            return self.g, self.b
    yh = to_hb(y, 2, BT)
    rh = to_hb(res, 1, BT) if use_res else None
    nwc = nw.cuda().requires_grad_(kind == 2) if nw is not None else None
    nbc = nb.cuda().requires_grad_(kind == 2) if nb is not None else None
    out = ops.post(yh, kind, act, nwc, nbc, rh, oh, 0, _LN() if kind == 3 else None, 1e-5)
    rel, mism, _ = _vs(padded_to_nchw(out, out.t.detach()), out_ref.detach())
    assert mism < 5e-4, ("forward", mism)                      # measured <= 0.004 % of the elements
    out.t.backward(dout.permute(0, 2, 3, 1).to(BT).cuda().contiguous())
    gy = yh.t.grad[:, 2:2 + h, 2:2 + w, :].permute(0, 3, 1, 2)
    rel, mism, _ = _vs(gy, yr.grad)
    one_rounding = 2.0 ** -9 / 3 ** 0.5 * 1.5                  # rms of one bf16 rounding (mantissa-averaged), relative
    if oh == 0:
        assert mism < 5e-4, ("backward", mism)                 # measured <= 0.008 %
    else:
        # behind a reflect halo the halo gradient is folded in bf16 first (one more rounding on the border pixels; the
        # per-channel means move by ~1e-5, which flips many near-zero elements by one ulp): bounded in L2 instead
        assert rel < 1.5 * one_rounding, ("backward", rel, one_rounding)     # measured 1.15 x
