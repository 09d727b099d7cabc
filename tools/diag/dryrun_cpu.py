"""Plumbing dry run on CPU (no arithmetic): the C library is replaced by a stub that returns success, so that the
Python/autograd wiring of Solver.dis_update / gen_update can be exercised without a GPU.  Not a pytest."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from dwc_gan_b200 import _lib as L
from dwc_gan_b200 import ops


class FakeLib:
    calls = {}

    def __getattr__(self, name):
        def f(*a):
            FakeLib.calls[name] = FakeLib.calls.get(name, 0) + 1
            if name == "dwc_wgrad_workspace_bytes":
                return 1024
            return 0
        return f


L._lib = FakeLib()
L.stream = lambda: C.c_void_p(0)
ops._require_cuda = lambda t: None
ops.RT.workspace = lambda nbytes, device=None: torch.empty(max(nbytes // 4, 1) + 16)
import dwc_gan_b200.flat as flat

from dwc_gan_b200.solver import Solver
from dwc_gan_b200.utils import get_config
from oracle import dwc_oracle as O

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
import dwc_gan_b200
dwc_gan_b200.set_mode(mode)
cfg = get_config(os.path.join(os.path.dirname(__file__), "golden", "celeba_faces.yaml"))
cfg["vgg_w"] = 0
torch.manual_seed(0)
s = Solver(cfg, torch.device("cpu"), None)
flat._DRYRUN[0] = True
s.copy_nets()
b = O.synthetic_batch(2, 128, seed=0)
args = (b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg)
for it in range(2):
    s.dis_update(*args, it)
    print("dis ok; touched", len(s.dis.flat.touched), "of", len(s.dis.flat.names))
    s.gen_update(*args, it)
    untouched = [n for n in s.gen.flat.names if n not in s.gen.flat.touched]
    print("gen ok; untouched:", untouched)
    s.smooth_moving()
    s.update_learning_rate()
    s.update_attention_status(it)
out = s.forward(b["x_real"], b["txt"], b["txt_lens"])
print("forward", out.shape)
outs = s.sample(b["x_real"], b["txt"], b["txt_lens"])
print("sample", [o.shape for o in outs])
print("kernel calls per 2 steps:", sum(FakeLib.calls.values()))
print(sorted(FakeLib.calls.items(), key=lambda kv: -kv[1])[:12])
