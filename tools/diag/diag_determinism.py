"""Runs the same D+G step twice from the same seeds and reports which gradient tensors differ bit-wise (none should)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from tests.util_gpu import build_solver, to_cuda  # noqa: E402
from oracle import dwc_oracle as O  # noqa: E402


def run(B):
    s, cfg = build_solver("bf16", deterministic=True)
    s.use_cuda_graphs = False
    s.copy_nets()
    b = to_cuda(O.synthetic_batch(B, 128, seed=5))
    torch.manual_seed(11)
    args = (b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, 0)
    s.dis_update(*args)
    gd = {k: p.grad.clone() for k, p in s.dis.named_parameters() if p.grad is not None}
    s.gen_update(*args)
    gg = {k: p.grad.clone() for k, p in s.gen.named_parameters() if p.grad is not None}
    torch.cuda.synchronize()
    return float(s.loss_dis), float(s.loss_gen_total), gd, gg


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    ref = run(B)
    for i in range(reps):
        cur = run(B)
        bad = [("dis." + k) for k in ref[2] if not torch.equal(ref[2][k], cur[2][k])]
        bad += [("gen." + k) for k in ref[3] if not torch.equal(ref[3][k], cur[3][k])]
        print("run %d: losses %s vs %s, %d tensors differ: %s" % (i, ref[:2], cur[:2], len(bad), bad[:12]), flush=True)


if __name__ == "__main__":
    main()
