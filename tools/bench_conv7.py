"""Times dwc_conv7_few at the two in-network geometries (decoder heads forward, first-conv image gradient)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dwc_gan_b200
from dwc_gan_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import microbench as MB

dwc_gan_b200.set_mode("bf16")
for B in (16, 48):
    for name, hin, cout, flip in (("heads fwd 64->4 @128", 134, 4, False), ("first-conv dgrad 64->3 @128", 140, 3, True)):
        nbuf = 6
        xs = [torch.randn(B, hin, hin, 64, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
        ho = hin - 6
        outs = [torch.empty(B, ho, ho, cout, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
        if not flip:
            w = torch.randn(cout * 49 * 64, device="cuda") * 0.05
            args = (0, 49 * 64, 7 * 64, 64, 1)
        else:
            w = torch.randn(64 * 49 * cout, device="cuda") * 0.05
            args = (6 * 7 * cout + 6 * cout, 1, -7 * cout, -cout, 49 * cout)
        fns = [(lambda i=i: ops.conv7_few(xs[i], B, hin, hin, w, *args, None, cout, outs[i],
                                          (cout, ho * cout, ho * ho * cout))) for i in range(nbuf)]
        t = MB.timeit(fns)
        byts = xs[0].numel() * 2 + outs[0].numel() * 2
        print("| %s | B=%d | %.1f us | %.1f us/image | %.0f GB/s of input+output |" % (name, B, t * 1e6, t * 1e6 / B, byts / t / 1e9),
              flush=True)
