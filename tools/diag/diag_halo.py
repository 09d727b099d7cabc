"""Diagnostic for the halo-tile conv kernel: identity weights on a single tap reveal which input pixel / channel
each output element was read from."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dwc_gan_b200 import _lib as L, plan as P
from dwc_gan_b200.plan import HB

n, h, w, c, k, p = 1, 16, 16, 64, 3, 1
hp, wp = h + 2 * p, w + 2 * p
dev = "cuda"
bt = torch.bfloat16
Y, X, Cc = torch.meshgrid(torch.arange(hp), torch.arange(wp), torch.arange(c), indexing="ij")
enc = {"x": X.float(), "y": Y.float(), "c": Cc.float()}
for ty, tx in ((0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 2)):
    wt = torch.zeros(c, k, k, c)
    for i in range(c):
        wt[i, ty, tx, i] = 1.0
    wf = wt.reshape(c, k * k * c).to(bt).to(dev)
    res = {}
    for name, val in enc.items():
        xp = HB(val.reshape(1, hp, wp, c).to(bt).to(dev).contiguous(), n, h, w, c, p, 0)
        y = HB.empty(n, h, w, c, 0, 0, torch.float32, dev, zero=True)
        pl = P.plan_conv_fwd(xp, wf, c, c, None, y, k, 1, L.TC)
        pl.launch()
        torch.cuda.synchronize()
        res[name] = y.t.cpu()[0]
    oy, ox, oc = torch.meshgrid(torch.arange(h), torch.arange(w), torch.arange(c), indexing="ij")
    ok = (res["x"] == ox + tx) & (res["y"] == oy + ty) & (res["c"] == oc)
    print("tap (dy=%d,dx=%d): %.3f of elements correct" % (ty, tx, ok.float().mean().item()))
    if not ok.all():
        bad = (~ok).nonzero()[:6]
        for b in bad:
            yy, xx, cc = [int(v) for v in b]
            print("   out(y=%d,x=%d,c=%d) got src (y=%g,x=%g,c=%g) want (y=%d,x=%d,c=%d)" % (
                yy, xx, cc, res["y"][yy, xx, cc], res["x"][yy, xx, cc], res["c"][yy, xx, cc], yy + ty, xx + tx, cc))
        # systematic view: for output pixel x positions, which source x was read (channel 0 and channel 8)
        print("   src x by out x (row 0, ch 0):", [int(v) for v in res["x"][0, :, 0]])
        print("   src c by out x (row 0, ch 0):", [int(v) for v in res["c"][0, :, 0]])
        print("   src c by out c (pixel 0,0):", [int(v) for v in res["c"][0, 0, ::8]])
        print("   src c by out c (pixel 0,1):", [int(v) for v in res["c"][0, 1, ::8]])
