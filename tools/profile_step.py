"""One training step bracketed by cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
s, cfg = bench.build_solver(dev, "bf16")
b = {k: v.to(dev) for k, v in bench.make_host_batch(B, 128, 0).items()}
for it in range(5):
    bench.one_step(s, cfg, b, it)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
bench.one_step(s, cfg, b, 5)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step, batch", B)
