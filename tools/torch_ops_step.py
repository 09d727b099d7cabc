"""Which torch-native (non-dwc) kernels does an eager training step launch, and from where?  (DWC_CUDA_GRAPHS=0)"""
import os, sys, collections
os.environ["DWC_CUDA_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
s, cfg = bench.build_solver(dev, "bf16")
s.use_cuda_graphs = False
b = {k: v.to(dev) for k, v in bench.make_host_batch(16, 128, 0).items()}
for it in range(3):
    bench.one_step(s, cfg, b, it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    bench.one_step(s, cfg, b, 3)
    torch.cuda.synchronize()
agg = collections.Counter()
for e in prof.events():
    if not e.name.startswith("aten::") or e.device_time_total <= 0 or not e.kernels:
        continue
    if e.cpu_children and any(c.kernels for c in e.cpu_children):
        continue                                          # count the innermost op only
    st = [f for f in (e.stack or []) if "dwc_gan_b200" in f or "solver" in f]
    key = (e.name, st[0].split("/")[-1][:70] if st else "?")
    agg[key] += len(e.kernels)
tot = sum(agg.values())
print("torch-native kernel launches in one eager step:", tot)
for (name, where), c in agg.most_common(45):
    print("%5d  %-22s %s" % (c, name, where))
