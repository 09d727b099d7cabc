"""Record the UNMODIFIED reference's loss curve over N training iterations (default 200) for the loss-curve parity
test (tests/test_curve_gpu.py).  Build container only (needs /root/reference); deterministic mode (all dropouts
off), batch 1 of oracle.dwc_oracle.synthetic_batch(seed=7), GMM noise from torch.manual_seed(100+it) /
(200+it) exactly as tests/golden/make_golden.py does.

Usage:  python tests/golden/make_curve.py [steps [batch]]      -> tests/golden/ref_curve_b<batch>.json
"""
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (import stubs + sys.path for the reference)
from oracle import dwc_oracle as O  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    torch.set_num_threads(os.cpu_count())
    solver, cfg = MG.build_reference()
    solver.copy_nets()
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    batch = O.synthetic_batch(B, 128, seed=7)
    names = ["loss_dis", "loss_gen_total", "loss_gen_adv", "loss_gen_recon_x", "loss_gen_recon_c_real",
             "loss_gen_recon_s_fake", "loss_gen_cycrecon_x", "loss_kl_x", "loss_kl_trg", "loss_ds"]
    curve = {n: [] for n in names}
    t0 = time.time()
    for it in range(steps):
        args = (batch["x_real"], batch["c_src"], batch["c_trg"], batch["txt"], batch["txt_lens"], batch["label_src"],
                batch["label_trg"], cfg, it)
        torch.manual_seed(100 + it)
        solver.dis_update(*args)
        torch.manual_seed(200 + it)
        solver.gen_update(*args)
        solver.smooth_moving()
        solver.update_learning_rate()
        solver.update_attention_status(it)
        for n in names:
            curve[n].append(float(getattr(solver, n)))
        if it % 10 == 0:
            print(it, "%.1fs" % (time.time() - t0), curve["loss_dis"][-1], curve["loss_gen_total"][-1], flush=True)
    out = {"B": B, "batch_seed": 7, "steps": steps, "curve": curve,
           "note": "reference run: torch %s CPU fp32, deterministic mode, vgg_w=0, seed 1234" % torch.__version__}
    name = "ref_curve_b%d.json" % B
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f)
    print("wrote", name)


if __name__ == "__main__":
    main()
