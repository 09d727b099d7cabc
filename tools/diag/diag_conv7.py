"""Diagnostic: is dwc_conv7_few bit-reproducible launch to launch (alone and under a concurrent stream)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import dwc_gan_b200
from dwc_gan_b200 import ops

dwc_gan_b200.set_mode("bf16")
torch.manual_seed(0)
for (n, hin, cout, flip) in ((16, 134, 4, False), (48, 140, 3, True)):
    x = torch.randn(n, hin, hin, 64, device="cuda").to(torch.bfloat16)
    ho = hin - 6
    if not flip:
        w = torch.randn(cout * 49 * 64, device="cuda") * 0.05
        args = (0, 49 * 64, 7 * 64, 64, 1)
    else:
        w = torch.randn(64 * 49 * cout, device="cuda") * 0.05
        args = (6 * 7 * cout + 6 * cout, 1, -7 * cout, -cout, 49 * cout)

    def run():
        out = torch.empty(n, ho, ho, cout, device="cuda", dtype=torch.bfloat16)
        ops.conv7_few(x, n, hin, hin, w, *args, None, cout, out, (cout, ho * cout, ho * ho * cout))
        return out
    ref = run()
    torch.cuda.synchronize()
    bad = 0
    for i in range(40):
        o = run()
        if not torch.equal(o, ref):
            bad += 1
            d = (o.float() - ref.float()).abs()
            idx = d.flatten().argmax().item()
            print("  alone: launch", i, "differs: max", d.max().item(), "count", int((d > 0).sum()), "at flat", idx)
    side = torch.cuda.Stream()
    a = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
    for i in range(40):
        with torch.cuda.stream(side):
            for _ in range(3):
                a @ a
        o = run()
        if not torch.equal(o, ref):
            bad += 1
            d = (o.float() - ref.float()).abs()
            print("  concurrent: launch", i, "differs: max", d.max().item(), "count", int((d > 0).sum()))
    torch.cuda.synchronize()
    print("geometry", (n, hin, cout, flip), "mismatching launches:", bad, flush=True)
