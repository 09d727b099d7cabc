timeout 14 python - <<'PY'
import torch
from dwc_gan_b200 import ops
ts = [torch.randn((), device="cuda", requires_grad=True) for _ in range(5)]
ws = [1.0, 0.5, 10.0, 0.0, 2.5]
out = ops.weighted_sum(list(zip(ts, ws))); out.backward()      # eager warm-up (builds the weight vector)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for t in ts: t.grad = None
    with torch.cuda.graph(g):
        o2 = ops.weighted_sum(list(zip(ts, ws)))
        o2.backward()
    g.replay()
torch.cuda.synchronize()
ref = sum(float(t) * w for t, w in zip(ts, ws))
print("weighted_sum graph ok", abs(float(o2) - ref) < 1e-4, [round(float(t.grad), 3) for t in ts])
PY
