"""Torch-native kernels of one eager training step with input shapes and device time, largest first."""
import collections
import os
import sys

os.environ["DWC_CUDA_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import bench  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

dev = torch.device("cuda", 0)
s, cfg = bench.build_solver(dev, "bf16")
s.use_cuda_graphs = False
b = {k: v.to(dev) for k, v in bench.make_host_batch(16, 128, 0).items()}
for it in range(3):
    bench.one_step(s, cfg, b, it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    bench.one_step(s, cfg, b, 3)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if not e.name.startswith("aten::") or not e.kernels:
        continue
    if e.cpu_children and any(c.kernels for c in e.cpu_children):
        continue
    key = (e.name, str(e.input_shapes)[:90])
    agg[key][0] += len(e.kernels)
    agg[key][1] += sum(k.duration for k in e.kernels)
tot = sum(v[1] for v in agg.values())
print("torch-native kernels: %d launches, %.1f us of device time" % (sum(v[0] for v in agg.values()), tot))
for (name, shp), (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:50]:
    print("%4d x %8.1f us  %-16s %s" % (c, d, name, shp))
