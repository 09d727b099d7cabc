"""New kernels of this round at their in-step shapes, once each between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --set full -k regex:row_kernel|conv7few`: the four row-streaming norm passes on the
32x32x256 (AdaIN), 64x64x128 (LN) and 128x128x64 (IN, parity-plane gradient) sites at batch 48, and conv7few at both
geometries."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dwc_gan_b200
from dwc_gan_b200 import ops
from dwc_gan_b200.plan import HB

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
dwc_gan_b200.set_mode("bf16")
bt = torch.bfloat16


class _LN:
    def __init__(self, c):
        self.gw, self.gb = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")

    def grad_buffers(self):
        return self.gw, self.gb


def site(c, hw, kind, with_res, out_halo, out_layout):
    y = torch.randn(B, hw, hw, c, device="cuda").to(bt).requires_grad_(True)
    res = torch.randn(B, hw + 2, hw + 2, c, device="cuda").to(bt).requires_grad_(True) if with_res else None
    nw = nb = ln = None
    if kind == ops.NORM_ADAIN:
        nw = torch.rand(B, c, device="cuda", requires_grad=True)
        nb = torch.rand(B, c, device="cuda", requires_grad=True)
    elif kind == ops.NORM_LN:
        ln = _LN(c)
        nw, nb = torch.rand(c, device="cuda"), torch.rand(c, device="cuda")
    out = ops.post(HB(y, B, hw, hw, c, 0, 0), kind=kind, act=ops.ACT_RELU, nw=nw, nb=nb,
                   res=HB(res, B, hw, hw, c, 1, 0) if with_res else None, out_halo=out_halo, out_layout=out_layout,
                   ln_mod=ln)
    out.t.backward(torch.randn_like(out.t))


def conv7(hin, cout, flip):
    x = torch.randn(B, hin, hin, 64, device="cuda").to(bt)
    ho = hin - 6
    out = torch.empty(B, ho, ho, cout, device="cuda", dtype=bt)
    if not flip:
        w = torch.randn(cout * 49 * 64, device="cuda") * 0.05
        a = (0, 49 * 64, 7 * 64, 64, 1)
    else:
        w = torch.randn(64 * 49 * cout, device="cuda") * 0.05
        a = (6 * 7 * cout + 6 * cout, 1, -7 * cout, -cout, 49 * cout)
    ops.conv7_few(x, B, hin, hin, w, *a, None, cout, out, (cout, ho * cout, ho * ho * cout))


def run():
    site(256, 32, ops.NORM_ADAIN, True, 1, 0)
    site(128, 64, ops.NORM_LN, False, 0, 0)
    site(64, 128, ops.NORM_IN, False, 1, 1)
    conv7(134, 4, False)
    conv7(140, 3, True)


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled, batch", B)
