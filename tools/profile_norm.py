"""Norm / post sites and the wgrad + reduce pair at their in-step shapes, each launched once between
cudaProfilerStart/Stop (for `ncu --profile-from-start off --set full`).
  python tools/profile_norm.py [batch=48]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dwc_gan_b200
from dwc_gan_b200 import ops, plan as P, _lib as L
from dwc_gan_b200.plan import HB

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import microbench as MB  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
dwc_gan_b200.set_mode("bf16")
bt = torch.bfloat16
SITES = [(256, 32, ops.NORM_IN, True), (256, 32, ops.NORM_ADAIN, False), (128, 64, ops.NORM_LN, False),
         (64, 128, ops.NORM_LN, False), (128, 64, ops.NORM_IN, False), (64, 128, ops.NORM_IN, False)]


def site(c, hw, kind, with_res):
    y = torch.randn(B, hw, hw, c, device="cuda").to(bt).requires_grad_(True)
    res = torch.randn(B, hw, hw, c, device="cuda").to(bt).requires_grad_(True) if with_res else None
    if kind == ops.NORM_ADAIN:
        nw = torch.rand(B, c, device="cuda", requires_grad=True)
        nb = torch.rand(B, c, device="cuda", requires_grad=True)
    else:
        nw = nb = None
    yh = HB(y, B, hw, hw, c, 0, 0)
    rh = HB(res, B, hw, hw, c, 0, 0) if with_res else None
    ln = None
    if kind == ops.NORM_LN:
        class _LN:
            def __init__(self):
                self.gw, self.gb = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")

            def grad_buffers(self):
                return self.gw, self.gb
        ln = _LN()
        nw, nb = torch.rand(c, device="cuda"), torch.rand(c, device="cuda")
    out = ops.post(yh, kind=kind, act=ops.ACT_RELU, nw=nw, nb=nb, res=rh, out_halo=2 if hw > 32 else 1, ln_mod=ln)
    g = torch.randn_like(out.t)
    out.t.backward(g)


def wgrad_pair(name):
    g = [x for x in MB.GEOMS if x[0] == name][0]
    _, cin, cout, k, s, p, hw = g
    ho = (hw + 2 * p - k) // s + 1
    layout = 0 if s == 1 else 1
    hy = k - 1 if s == 1 else 1
    x = HB(torch.randn(HB.shape_of(B, hw, hw, cin, p, layout), device="cuda").to(bt), B, hw, hw, cin, p, layout)
    y = HB(torch.randn(HB.shape_of(B, ho, ho, cout, hy, 0), device="cuda").to(bt), B, ho, ho, cout, hy, 0)
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    db = torch.zeros(cout, device="cuda")
    pl = P.plan_conv_wgrad(y, x, dw, db, k, s, L.TC)
    return lambda: pl.launch(MB.workspace)


def run():
    for (c, hw, kind, r) in SITES:
        site(c, hw, kind, r)
    for f in wg:
        f()


wg = [wgrad_pair("G7"), wgrad_pair("G8"), wgrad_pair("G9")]
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled norm sites + wgrad pairs, batch", B)
