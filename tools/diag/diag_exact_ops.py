"""Are the bf16 kernels the SAME computation as 'fp32 math on bf16-stored operands, rounded once on store'?
conv (tc) with fp32 output vs fp64; conv bf16 output vs round(fp64); post (IN+ReLU) bf16 vs round(fp64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
import dwc_gan_b200
from dwc_gan_b200 import _lib as L, plan as P, ops
from dwc_gan_b200.plan import HB
from tests import emu
from tests.test_conv_gpu import pack, workspace
from oracle import dwc_oracle as O

dwc_gan_b200.set_mode("bf16")
bt = torch.bfloat16
def stats(name, got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    rel = float((got - ref).norm() / ref.norm())
    refq = ref.to(bt).double()
    mism = float((got != refq).double().mean())
    relq = float((got - refq).norm() / refq.norm())
    print("%-34s rel-vs-fp64 %.3e | vs round_bf16(fp64): rel %.3e mismatching elements %.4f%%" % (name, rel, relq, 100 * mism))

for (n, h, w, cin, cout, k, s, p) in [(2, 32, 32, 256, 256, 3, 1, 1), (2, 64, 64, 64, 128, 4, 2, 1), (1, 64, 64, 256, 128, 5, 1, 2)]:
    torch.manual_seed(1)
    x = torch.randn(n, cin, h, w).to(bt).double()
    wt = (torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)).to(bt).double()
    bias = torch.randn(cout)
    xpad = F.pad(x, (p, p, p, p), mode="reflect").requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    y_ref = F.conv2d(xpad, wr, bias.double(), stride=s)
    ho, wo = y_ref.shape[2:]
    dy = torch.randn(n, cout, ho, wo).to(bt).double()
    y_ref.backward(dy)
    layout = 0 if s == 1 else 1
    hy = k - 1 if s == 1 else 1
    w_krsc = wt.permute(0, 2, 3, 1).contiguous().float().cuda()
    xp = emu.make_padded(x, p, layout, bt); xp = xp.like(xp.t.cuda())
    wf = pack(w_krsc, 0, bt, cout, cout, k, cin)
    tag = "conv %dx%d %d->%d k%d s%d" % (h, w, cin, cout, k, s)
    for odt in (torch.float32, bt):
        y = HB.empty(n, ho, wo, cout, hy, 0, odt, "cuda", zero=True)
        P.plan_conv_fwd(xp, wf, cout, cout, bias.cuda(), y, k, s, L.TC).launch()
        torch.cuda.synchronize()
        stats(tag + " fwd out=" + ("f32" if odt == torch.float32 else "bf16"), y.interior().permute(0, 3, 1, 2), y_ref.detach())
    dyz = emu.make_zero_haloed(dy, hy, bt); dyz = dyz.like(dyz.t.cuda())
    wd = pack(w_krsc, 1 if s == 1 else 2, bt, cin, cout, k, cin)
    for odt in (torch.float32, bt):
        dxp = HB.empty(n, h, w, cin, p, layout, odt, "cuda")
        for q in P.plan_conv_dgrad(dyz, wd, dxp, k, s, L.TC, cin_padded=cin):
            q.launch()
        torch.cuda.synchronize()
        stats(tag + " dgrad out=" + ("f32" if odt == torch.float32 else "bf16"), dxp.padded_nhwc().permute(0, 3, 1, 2), xpad.grad)
    dw = torch.zeros(cout, k, k, cin, device="cuda"); db = torch.zeros(cout, device="cuda")
    P.plan_conv_wgrad(dyz, xp, dw, db, k, s, L.TC).launch(workspace)
    torch.cuda.synchronize()
    stats(tag + " wgrad (f32)", dw.permute(0, 3, 1, 2), wr.grad)

# post: IN + ReLU (+res) + reflect pad, bf16 in / bf16 out
from tests.test_post_gpu import to_hb, padded_to_nchw
for (n, c, h, w, kind, act, use_res, oh) in [(2, 256, 32, 32, 1, 1, False, 1), (2, 256, 32, 32, 1, 0, True, 1), (2, 64, 128, 128, 1, 1, False, 1), (2, 128, 64, 64, 3, 1, False, 2)]:
    torch.manual_seed(0)
    y = (torch.randn(n, c, h, w) * 1.5 + 0.3).to(bt).float()
    res = torch.randn(n, c, h, w).to(bt).float()
    nw = torch.rand(c) if kind == 3 else None
    nb = torch.randn(c) if kind == 3 else None
    yr = y.double().requires_grad_(True); rr = res.double().requires_grad_(True)
    z = O.inst_norm(yr) if kind == 1 else O.layer_norm_munit(yr, nw.double(), nb.double())
    z = torch.relu(z) if act == 1 else z
    if use_res: z = z + rr
    out_ref = F.pad(z, (oh, oh, oh, oh), mode="reflect")
    dout = torch.randn_like(out_ref).to(bt).double()
    out_ref.backward(dout)
    yh = to_hb(y, 2, bt); rh = to_hb(res, 1, bt) if use_res else None
    class LN:  # minimal ln_mod
        def __init__(s_): s_.g = torch.zeros(c, device="cuda"); s_.b = torch.zeros(c, device="cuda")
        def grad_buffers(s_): return s_.g, s_.b
    out = ops.post(yh, kind, act, nw.cuda() if nw is not None else None, nb.cuda() if nb is not None else None, rh, oh, 0, LN() if kind == 3 else None, 1e-5)
    stats("post kind%d act%d res%d %dx%dx%d fwd" % (kind, act, use_res, c, h, w), padded_to_nchw(out, out.t.detach()), out_ref.detach())
    dt = torch.zeros_like(out.t)
    dnhwc = dout.permute(0, 2, 3, 1).to(bt).cuda()
    out.t.backward(dnhwc.contiguous())
    gy = yh.t.grad[:, 2:2 + h, 2:2 + w, :].permute(0, 3, 1, 2)
    stats("post kind%d act%d res%d %dx%dx%d bwd dy" % (kind, act, use_res, c, h, w), gy, yr.grad)
