import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
import dwc_gan_b200
from dwc_gan_b200 import ops
from oracle import dwc_oracle as O
from tests.test_post_gpu import to_hb
dwc_gan_b200.set_mode("bf16")
bt = torch.bfloat16
for (n, c, h, w, act, oh) in [(2, 256, 32, 32, 1, 1), (2, 256, 32, 32, 1, 0), (2, 256, 32, 32, 0, 0), (2, 64, 128, 128, 1, 1)]:
    torch.manual_seed(0)
    y = (torch.randn(n, c, h, w) * 1.5 + 0.3).to(bt).float()
    yr = y.double().requires_grad_(True)
    z = O.inst_norm(yr)
    z = torch.relu(z) if act == 1 else z
    out_ref = F.pad(z, (oh, oh, oh, oh), mode="reflect") if oh else z
    dout = torch.randn_like(out_ref).to(bt).double()
    out_ref.backward(dout)
    yh = to_hb(y, 2, bt)
    out = ops.post(yh, 1, act, None, None, None, oh, 0, None, 1e-5)
    out.t.backward(dout.permute(0, 2, 3, 1).to(bt).cuda().contiguous())
    gy = yh.t.grad[:, 2:2 + h, 2:2 + w, :].permute(0, 3, 1, 2).double().cpu()
    ref = yr.grad
    err = gy - ref.to(bt).double()
    e_exact = gy - ref
    xhat = O.inst_norm(y.double())
    border = torch.zeros(h, w, dtype=torch.bool); 
    if oh: border[1, :] = border[h - 2, :] = True; border[:, 1] = border[:, w - 2] = True
    print("case act%d halo%d %dx%dx%d: mismatch %.3f%%  (border px %.3f%%, interior px %.3f%%)" % (act, oh, c, h, w,
          100 * float((err != 0).double().mean()), 100 * float((err[:, :, border] != 0).double().mean()) if oh else 0,
          100 * float((err[:, :, ~border] != 0).double().mean())))
    pc = e_exact.mean(dim=(2, 3))                # per (n,c) mean of the exact error
    px = (e_exact * xhat).mean(dim=(2, 3))
    print("   |dy| rms %.3e  exact-err rms %.3e  per-(n,c) mean err rms %.3e  per-(n,c) <err*xhat> rms %.3e  (bf16 half-ulp rms ~ %.3e)" % (
          float(ref.pow(2).mean().sqrt()), float(e_exact.pow(2).mean().sqrt()), float(pc.pow(2).mean().sqrt()), float(px.pow(2).mean().sqrt()),
          float(ref.abs().mean()) * 2 ** -9 / 3 ** 0.5))
    m = (z.detach() > 0) if act == 1 else torch.ones_like(z, dtype=torch.bool)
    if oh == 0:
        print("   mismatch where mask=1: %.3f%%, where mask=0: %.3f%%" % (100 * float((err[m] != 0).double().mean()), 100 * float((err[~m] != 0).double().mean()) if act else 0))
