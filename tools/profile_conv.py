"""A few launches of one conv geometry (for `ncu --set full -k regex:gconv`): python tools/profile_conv.py G7 fprop 16"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dwc_gan_b200 import _lib as L, plan as P
from dwc_gan_b200.plan import HB
import tools.microbench as mb

name, op, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
g = [x for x in mb.GEOMS if x[0] == name][0]
_, cin, cout, k, s, p, hw = g
ho = (hw + 2 * p - k) // s + 1
layout = 0 if s == 1 else 1
hy = k - 1 if s == 1 else 1
bt = torch.bfloat16
x = HB(torch.randn(HB.shape_of(n, hw, hw, cin, p, layout), device="cuda").to(bt), n, hw, hw, cin, p, layout)
y = HB(torch.randn(HB.shape_of(n, ho, ho, cout, hy, 0), device="cuda").to(bt), n, ho, ho, cout, hy, 0)
w = torch.randn(cout, k, k, cin, device="cuda") * 0.02
bias = torch.zeros(cout, device="cuda")
wf = mb.pack(w, 0, cout, cout, k, cin)
wd = mb.pack(w, 1 if s == 1 else 2, cin, cout, k, cin)
dw = torch.zeros(cout, k, k, cin, device="cuda")
for _ in range(4):
    if op == "fprop":
        P.plan_conv_fwd(x, wf, cout, cout, bias, y, k, s, L.TC).launch()
    elif op == "dgrad":
        for q in P.plan_conv_dgrad(y, wd, x, k, s, L.TC):
            q.launch()
    else:
        P.plan_conv_wgrad(y, x, dw, None, k, s, L.TC).launch(mb.workspace)
torch.cuda.synchronize()
print("done")
