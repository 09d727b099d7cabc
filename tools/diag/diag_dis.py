"""Diagnostics (not a pytest): discriminator forward/backward vs the fp64 oracle, with spatial error structure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import dwc_oracle as O
from tests.util_gpu import build_solver, cpu_state, to_cuda, rel, grads_of

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
s, cfg = build_solver(mode)
D64 = {k: v.double().requires_grad_(True) for k, v in cpu_state(s.dis).items()}
batch = O.synthetic_batch(3, 128, seed=4)
x64 = batch["x_real"].double().clone().requires_grad_(True)
outs_ref = O.dis_forward(D64, x64)
b = to_cuda(batch)
xc = b["x_real"].clone().requires_grad_(True)
outs = s.dis(xc)
for i in range(2):
    print("scale", i, "src rel", rel(outs[i][0], outs_ref[i][0]), "cls rel", rel(outs[i][1], outs_ref[i][1]))
ws = [[torch.randn_like(o[0]), torch.randn_like(o[1])] for o in outs_ref]
which = sys.argv[2] if len(sys.argv) > 2 else "all"
def total(outs, ws, dev):
    t = 0
    for i, (o, w) in enumerate(zip(outs, ws)):
        if which in ("all", "s%d" % i):
            t = t + (o[0] * w[0].to(dev)).sum() + (o[1] * w[1].to(dev)).sum()
    return t
total(outs_ref, ws, "cpu").backward()
s.dis_opt.zero_grad()
total(outs, [[w[0].float(), w[1].float()] for w in ws], "cuda").backward()
g, gr = xc.grad.double().cpu(), x64.grad
print("x.grad rel", rel(g, gr), "norms", float(g.norm()), float(gr.norm()))
err = (g - gr).abs().sum((0, 1))
ref = gr.abs().sum((0, 1))
def reg(name, m):
    print("  %-10s err/ref = %.3e" % (name, float(err[m].sum() / ref[m].sum())))
H = 128
yy, xx = torch.meshgrid(torch.arange(H), torch.arange(H), indexing="ij")
border = lambda k: (yy < k) | (yy >= H - k) | (xx < k) | (xx >= H - k)
reg("interior", ~border(8)); reg("border1", border(1)); reg("border2", border(2) & ~border(1)); reg("border4", border(4) & ~border(2)); reg("border8", border(8) & ~border(4))
reg("odd rows", (yy % 2 == 1)); reg("even rows", (yy % 2 == 0))
mine = grads_of(s.dis)
for k, v in D64.items():
    if v.grad is not None:
        print("  grad %-28s rel %.3e" % (k, rel(mine[k], v.grad)))
