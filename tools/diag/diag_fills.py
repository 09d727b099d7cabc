"""Where do the aten::fill_ / copy_ / cat launches of an eager training step come from?  Python-level call sites are counted
by wrapping the torch entry points; the remainder is issued from C++ (autograd: zero-materialised gradients, slice /
unbind backward, gradient accumulation)."""
import collections
import os
import sys
import traceback

os.environ["DWC_CUDA_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import bench  # noqa: E402

counts = collections.Counter()
ON = [False]


def site():
    for f in reversed(traceback.extract_stack()[:-2]):
        if "dwc_gan_b200" in f.filename or f.filename.endswith("bench.py"):
            return "%s:%d" % (os.path.basename(f.filename), f.lineno)
    return "?"


def wrap(mod, name):
    orig = getattr(mod, name)

    def f(*a, **k):
        if ON[0]:
            counts[(name, site())] += 1
        return orig(*a, **k)
    setattr(mod, name, f)


for n in ("zeros", "zeros_like", "ones", "ones_like", "full", "cat", "stack", "empty_like"):
    wrap(torch, n)
for n in ("zero_", "fill_", "copy_", "contiguous", "float", "clone"):
    wrap(torch.Tensor, n)

dev = torch.device("cuda", 0)
s, cfg = bench.build_solver(dev, "bf16")
s.use_cuda_graphs = False
b = {k: v.to(dev) for k, v in bench.make_host_batch(16, 128, 0).items()}
for it in range(3):
    bench.one_step(s, cfg, b, it)
torch.cuda.synchronize()
ON[0] = True
bench.one_step(s, cfg, b, 3)
torch.cuda.synchronize()
ON[0] = False
for (name, where), c in counts.most_common(60):
    print("%5d  %-12s %s" % (c, name, where))
