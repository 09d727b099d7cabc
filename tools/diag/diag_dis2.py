"""Diagnostics: discriminator scale-0 chain, every intermediate gradient vs a torch fp64 graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from dwc_gan_b200 import ops
from oracle import dwc_oracle as O
from tests.util_gpu import build_solver, cpu_state, rel

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
s, cfg = build_solver(mode)
D = {k: v.double().cuda().requires_grad_(True) for k, v in cpu_state(s.dis).items()}
batch = O.synthetic_batch(3, 128, seed=4)
x = batch["x_real"].cuda()
s.dis.ensure_flat()
# ---- ours, keeping every intermediate
h = ops.image_pad(x, 1, 1, 1)
mine_pad, mine_y = [], []
convs = list(s.dis.cnns_feat[0])
for j, cv in enumerate(convs):
    last = j == len(convs) - 1
    y = ops.conv(h, cv)
    y.t.retain_grad()
    mine_y.append(y)
    h = ops.post(y, ops.NORM_NONE, ops.ACT_LRELU, out_halo=0 if last else 1, out_layout=0 if last else 1)
    h.t.retain_grad()
    mine_pad.append(h)
wt = torch.randn(h.t.shape, device="cuda")
(h.t.float() * wt).sum().backward()
# ---- reference fp64
ref_pad, ref_y = [], []
xr = x.double()
hp = F.pad(xr, (1, 1, 1, 1), mode="reflect")
for j in range(5):
    y = F.conv2d(hp, D[f"cnns_feat.0.{j}.conv.weight"], D[f"cnns_feat.0.{j}.conv.bias"], stride=2)
    y.retain_grad(); ref_y.append(y)
    a = F.leaky_relu(y, 0.1)
    hp = a if j == 4 else F.pad(a, (1, 1, 1, 1), mode="reflect")
    hp.retain_grad(); ref_pad.append(hp)
(hp * wt.double().permute(0, 3, 1, 2)).sum().backward()
for j in range(4, -1, -1):
    my, ry = mine_y[j], ref_y[j]
    gy = my.t.grad[:, my.halo:my.halo + my.h, my.halo:my.halo + my.w, :].permute(0, 3, 1, 2)
    print("layer %d: y fwd rel %.2e | dY rel %.2e" % (j, rel(my.interior().permute(0, 3, 1, 2), ry), rel(gy, ry.grad)), end="")
    if j > 0:
        mp, rp = mine_pad[j - 1], ref_pad[j - 1]
        g = mp.like(mp.t.grad).padded_nhwc().permute(0, 3, 1, 2)
        e = (g.double() - rp.grad).abs()
        print(" | d(padded input) rel %.2e  max-err at %s" % (rel(g, rp.grad), str(torch.nonzero(e == e.max())[0].tolist())), end="")
        # per-parity-plane error
        pl = [float((g.double() - rp.grad)[:, :, py::2, px::2].norm() / rp.grad[:, :, py::2, px::2].norm()) for py in range(2) for px in range(2)]
        print(" planes", ["%.1e" % v for v in pl], end="")
    print()
