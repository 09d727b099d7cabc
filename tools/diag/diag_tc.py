"""Diagnostics for the tcgen05 kernels (not a pytest): prints error structure for simple cases."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from dwc_gan_b200 import _lib as L, plan as P
from dwc_gan_b200.plan import HB
from tests import emu
from tests.test_conv_gpu import pack, workspace


def run(n, h, w, cin, cout, k, s, p, what):
    dtype = torch.bfloat16
    torch.manual_seed(1)
    x = torch.randn(n, cin, h, w).to(dtype).double()
    wt = (torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)).to(dtype).double()
    xpad = F.pad(x, (p, p, p, p), mode="reflect").requires_grad_(True)
    wt_r = wt.clone().requires_grad_(True)
    y_ref = F.conv2d(xpad, wt_r, None, stride=s)
    ho, wo = y_ref.shape[2:]
    dy = torch.randn(n, cout, ho, wo).to(dtype).double()
    y_ref.backward(dy)
    layout = 0 if s == 1 else 1
    hy = k - 1 if s == 1 else 1
    w_krsc = wt.permute(0, 2, 3, 1).contiguous().float().cuda()
    xp = emu.make_padded(x, p, layout, dtype); xp = xp.like(xp.t.cuda())
    dyz = emu.make_zero_haloed(dy, hy, dtype); dyz = dyz.like(dyz.t.cuda())
    try:
        if what == "fwd":
            rows_p = cout if cout % 64 == 0 else 16
            y = HB.empty(n, ho, wo, cout, hy, 0, dtype, "cuda", zero=True)
            wf = pack(w_krsc, 0, dtype, rows_p, cout, k, cin)
            P.plan_conv_fwd(xp, wf, cout, rows_p, None, y, k, s, L.TC).launch()
            torch.cuda.synchronize()
            got = y.interior().permute(0, 3, 1, 2).double().cpu(); ref = y_ref.detach()
        elif what == "wgrad":
            dw = torch.zeros(cout, k, k, cin, device="cuda")
            P.plan_conv_wgrad(dyz, xp, dw, None, k, s, L.TC).launch(workspace)
            torch.cuda.synchronize()
            got = dw.cpu().double().permute(0, 3, 1, 2); ref = wt_r.grad
        else:
            dxp = HB.empty(n, h, w, cin, p, layout, dtype, "cuda"); dxp.t.zero_()
            wd = pack(w_krsc, 1 if s == 1 else 2, dtype, cin, cout, k, cin)
            for q in P.plan_conv_dgrad(dyz, wd, dxp, k, s, L.TC):
                q.launch()
            torch.cuda.synchronize()
            got = dxp.padded_nhwc().permute(0, 3, 1, 2).double().cpu(); ref = xpad.grad
    except Exception as e:  # noqa
        print("CASE", (n, h, w, cin, cout, k, s, p), what, "EXC", repr(e)[:300])
        return
    err = (got - ref).abs()
    rel = err.max().item() / ref.abs().max().item()
    print("CASE", (n, h, w, cin, cout, k, s, p), what, "rel err %.3e" % rel, "ref max %.3f got max %.3f nan %d" % (
        ref.abs().max().item(), got.abs().max().item() if not torch.isnan(got).all() else float("nan"),
        int(torch.isnan(got).sum())))
    if rel > 2e-2:
        e2 = err.flatten(2).mean(2) if err.dim() == 4 else err
        print("  per-(dim0,dim1) mean err, first 8x16:\n", (e2[:8, :16] * 100).round() / 100)
        g = got.flatten(); r = ref.flatten()
        print("  corr with ref: %.4f" % float(torch.corrcoef(torch.stack([g, r]))[0, 1]))


if __name__ == "__main__":
    print("tc available:", L.lib().dwc_tc_available())
    cases = [(1, 8, 16, 64, 64, 1, 1, 0), (1, 8, 16, 128, 64, 1, 1, 0), (1, 8, 16, 64, 128, 1, 1, 0),
             (1, 8, 16, 64, 256, 1, 1, 0), (2, 16, 16, 64, 128, 3, 1, 1), (2, 16, 16, 64, 128, 4, 2, 1)]
    jobs = [(c, w) for c in cases for w in ("fwd", "dgrad", "wgrad")]
    if len(sys.argv) > 1:                       # one job per process: a trap poisons the CUDA context
        c, w = jobs[int(sys.argv[1])]
        run(*c, w)
    else:
        import subprocess
        for i in range(len(jobs)):
            r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=300)
            out = [l for l in (r.stdout + r.stderr).splitlines() if l.strip() and "Warning" not in l]
            print("\n".join(out[-14:]))
