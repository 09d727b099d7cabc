"""Diagnostic: every torch.empty buffer is pre-filled with NaN; any kernel that reads memory it (or a producer) never
wrote turns losses / gradients into NaN."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

_empty = torch.empty


def poisoned_empty(*a, **k):
    t = _empty(*a, **k)
    if t.is_floating_point() and t.is_cuda:
        t.fill_(float("nan"))
    return t


torch.empty = poisoned_empty
from tests.util_gpu import build_solver, to_cuda  # noqa: E402
from oracle import dwc_oracle as O  # noqa: E402

s, cfg = build_solver("bf16", deterministic=False)
s.use_cuda_graphs = False
s.copy_nets()
b = to_cuda(O.synthetic_batch(4, 128, seed=3))
for it in range(2):
    torch.manual_seed(500 + it)
    args = (b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, it)
    s.dis_update(*args)
    s.gen_update(*args)
    torch.cuda.synchronize()
    print(os.environ.get("TAG", ""), "step", it, "loss_dis", float(s.loss_dis), "loss_gen", float(s.loss_gen_total),
          "nan params gen", sum(int(torch.isnan(p).any()) for p in s.gen.parameters()),
          "dis", sum(int(torch.isnan(p).any()) for p in s.dis.parameters()), flush=True)
bad = [k for k, p in s.gen.named_parameters() if torch.isnan(p).any()]
print("first NaN gen params:", bad[:8])
