"""Times dwc_upsample_pad_fwd / bwd at the two decoder geometries (CUDA-graph replay over rotating buffers)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dwc_gan_b200
from dwc_gan_b200 import _lib as L
from dwc_gan_b200.plan import HB
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import microbench as MB

dwc_gan_b200.set_mode("bf16")
lib = L.lib()
bt = torch.bfloat16
for B in (16, 48):
    for (c, hw) in ((256, 32), (128, 64)):
        nbuf = 4
        xs = [HB(torch.randn(B, hw, hw, c, device="cuda").to(bt), B, hw, hw, c, 0, 0) for _ in range(nbuf)]
        outs = [HB.empty(B, 2 * hw, 2 * hw, c, 2, 0, bt, "cuda") for _ in range(nbuf)]
        douts = [HB(torch.randn(B, 2 * hw + 4, 2 * hw + 4, c, device="cuda").to(bt), B, 2 * hw, 2 * hw, c, 2, 0) for _ in range(nbuf)]
        dxs = [HB.empty(B, hw, hw, c, 0, 0, bt, "cuda") for _ in range(nbuf)]
        S = lambda hb: C.byref(hb.struct())
        tf = MB.timeit([(lambda i=i: L.check(lib.dwc_upsample_pad_fwd(S(xs[i]), S(outs[i]), L.stream()))) for i in range(nbuf)])
        tb = MB.timeit([(lambda i=i: L.check(lib.dwc_upsample_pad_bwd(S(douts[i]), S(dxs[i]), 1, L.stream()))) for i in range(nbuf)])
        bf = xs[0].t.numel() * 2 + outs[0].t.numel() * 2
        print("| upsample %dx%dx%d B=%d | fwd %.1f us (%.0f GB/s) | bwd %.1f us (%.0f GB/s) |" % (
            c, hw, hw, B, tf * 1e6, bf / tf / 1e9, tb * 1e6, bf / tb / 1e9), flush=True)
