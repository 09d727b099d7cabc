"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  It
  1. imports the reference Solver through two import stubs (torchfile, tensorboardX)
     with vgg_w=0, seeds torch with 1234 as train.py:23 does,
  2. records per-parameter checksums of the reference's random init,
  3. runs two deterministic-mode training iterations (all dropouts off) on the
     synthetic batch of oracle.dwc_oracle.synthetic_batch(B=2, seed=0),
  4. runs oracle/dwc_oracle.py from the same initial weights and asserts it agrees
     with the reference (losses, every gradient, every post-step parameter),
  5. writes tests/golden/ref_step_b2.json (+ ref_init_seed1234.json).

Usage:  python tests/golden/make_golden.py
"""
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

for name in ("torchfile", "tensorboardX"):
    m = types.ModuleType(name)
    m.load = lambda *a, **k: None
    m.SummaryWriter = object
    sys.modules[name] = m
sys.path.insert(0, REF)

from oracle import dwc_oracle as O  # noqa: E402


def build_reference(seed=1234):
    import utils as ref_utils
    from solver import Solver
    cfg = ref_utils.get_config(os.path.join(REF, "configs/celeba_faces.yaml"))
    cfg["vgg_w"] = 0
    torch.manual_seed(seed)
    s = Solver(cfg, torch.device("cpu"), None)
    # deterministic mode
    s.gen.enc_style.mapping[2].p = 0.0
    s.gen.enc_txt.dropout_in = 0.0
    s.gen.enc_txt.dropout_out = 0.0
    s.gen.enc_txt.lstm.dropout = 0.0
    return s, cfg


def checksum(t):
    t = t.detach().double()
    return [float(t.sum()), float(t.abs().sum()), float((t * t).sum().sqrt())]


def main():
    torch.set_num_threads(os.cpu_count())
    solver, cfg = build_reference()
    gen_sd = {k: v.clone() for k, v in solver.gen.state_dict().items()}
    dis_sd = {k: v.clone() for k, v in solver.dis.state_dict().items()}
    init = {"gen": {k: {"shape": list(v.shape), "ck": checksum(v), "head": v.flatten()[:4].tolist()}
                    for k, v in gen_sd.items()},
            "dis": {k: {"shape": list(v.shape), "ck": checksum(v), "head": v.flatten()[:4].tolist()}
                    for k, v in dis_sd.items()},
            "seed": 1234, "torch": torch.__version__}
    with open(os.path.join(HERE, "ref_init_seed1234.json"), "w") as f:
        json.dump(init, f)

    B = 2
    batch = O.synthetic_batch(B, 128, seed=0)
    solver.copy_nets()
    orc = O.OracleSolver(gen_sd, dis_sd)

    # inference (Solver.sample semantics) on the initial weights (raw decoder heads, no blend)
    with torch.no_grad():
        solver.eval()
        c, mus, _ = solver.gen.encode(batch["x_real"])
        st, _ = solver.gen.encode_txt(torch.cat(mus, 1), batch["txt"], batch["txt_lens"])
        img, att = solver.gen.decode(c, torch.cat(st, 1))
        solver.train()
        o_img = O.translate(orc.G, batch["x_real"], batch["txt"], batch["txt_lens"], use_attention=False)
    infer_err = float((img - o_img).abs().max())
    print("inference max abs diff ref vs oracle:", infer_err)
    assert infer_err < 1e-4
    steps = []
    for it in range(2):
        rec = {"iter": it, "use_attention": bool(solver.use_attention)}
        # ---- D phase
        torch.manual_seed(100 + it)
        solver.dis_update(batch["x_real"], batch["c_src"], batch["c_trg"], batch["txt"], batch["txt_lens"],
                          batch["label_src"], batch["label_trg"], cfg, it)
        torch.manual_seed(100 + it)
        eps1 = torch.randn(1, 8, B, 8)
        orc.dis_update(batch, eps1)
        rec["loss_dis"] = float(solver.loss_dis)
        dg = {k: p.grad for k, p in solver.dis.named_parameters()}
        rec["dis_grad"] = {k: checksum(g) for k, g in dg.items()}
        worst = 0.0
        for k, g in dg.items():
            e = float((g - orc.last_dis_grads[k]).norm() / (g.norm() + 1e-12))
            worst = max(worst, e)
        print(f"it{it} D: ref {rec['loss_dis']:.6f} oracle {orc.losses['loss_dis']:.6f} worst grad rel err {worst:.2e}")
        assert abs(rec["loss_dis"] - orc.losses["loss_dis"]) < 1e-4 * abs(rec["loss_dis"])
        # the reference's own fp32 noise floor (8 vs 3 threads) is ~3e-3 per tensor at iteration 0 and
        # grows after the first sign-like Adam step, so iteration 1 is only checked loosely
        assert worst < (1e-2 if it == 0 else 1e-1), worst
        # ---- G phase
        torch.manual_seed(200 + it)
        solver.gen_update(batch["x_real"], batch["c_src"], batch["c_trg"], batch["txt"], batch["txt_lens"],
                          batch["label_src"], batch["label_trg"], cfg, it)
        torch.manual_seed(200 + it)
        e1 = torch.randn(1, 8, B, 8)
        e2 = torch.randn(1, 8, B, 8)
        orc.gen_update(batch, e1, e2)
        names = [n for n in dir(solver) if "loss" in n and not callable(getattr(solver, n))]
        rec["losses"] = {n: float(getattr(solver, n)) for n in names}
        rec["init_ds_w"] = solver.init_ds_w
        gg = {k: p.grad for k, p in solver.gen.named_parameters()}
        rec["gen_grad"] = {k: (checksum(g) if g is not None else None) for k, g in gg.items()}
        worst, worst_k = 0.0, None
        for k, g in gg.items():
            og = orc.last_gen_grads[k]
            assert (g is None) == (og is None), k
            if g is None:
                continue
            if float(g.norm()) < 1e-6:        # biases in front of IN/AdaIN: pure round-off
                continue
            e = float((g - og).norm() / g.norm())
            if e > worst:
                worst, worst_k = e, k
        for n in names:
            if n in orc.losses:
                r, o = float(getattr(solver, n)), orc.losses[n]
                assert abs(r - o) <= 2e-4 * max(1.0, abs(r)), (n, r, o)
        print(f"it{it} G: ref {rec['losses']['loss_gen_total']:.6f} oracle {orc.losses['loss_gen_total']:.6f} "
              f"worst grad rel err {worst:.2e} ({worst_k})")
        gnum = sum(float((g - orc.last_gen_grads[k]).norm()) ** 2 for k, g in gg.items() if g is not None) ** 0.5
        gden = sum(float(g.norm()) ** 2 for g in gg.values() if g is not None) ** 0.5
        rec["gen_grad_global_norm"] = gden
        print(f"      global gen grad rel err {gnum / gden:.2e}")
        assert worst < (1e-2 if it == 0 else 3e-1), (worst, worst_k)
        assert gnum / gden < (5e-4 if it == 0 else 5e-2)
        solver.smooth_moving()
        orc.smooth_moving()
        solver.update_learning_rate()
        solver.update_attention_status(it)
        orc.update_attention_status(it)
        rec["gen_param"] = {k: checksum(p) for k, p in solver.gen.named_parameters()}
        rec["dis_param"] = {k: checksum(p) for k, p in solver.dis.named_parameters()}
        rec["gen_avg_param_sum"] = float(sum(p.double().sum() for p in solver.gen_copy.parameters()))
        steps.append(rec)

    out = {"B": B, "batch_seed": 0, "steps": steps,
           "infer_img_ck": checksum(img), "infer_att_ck": checksum(att),
           "note": "reference run: torch %s CPU fp32, deterministic mode, vgg_w=0" % torch.__version__}
    with open(os.path.join(HERE, "ref_step_b2.json"), "w") as f:
        json.dump(out, f)
    print("wrote goldens")


if __name__ == "__main__":
    main()
