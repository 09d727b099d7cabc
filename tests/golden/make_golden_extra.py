"""Extra golden fixtures recorded from the UNMODIFIED reference (build container only, needs /root/reference):

  gmm_sample : tools.dist_sampling_split (tools.py:65-70) on seeded inputs - the exact floats, for the bit-exact
               component-selection / layout test of dwc_gmm_sample;
  gmm_em     : gmm.gmm_earth_mover_distance_sp (gmm.py:33-41) value + gradient on seeded inputs;
  sample     : Solver.sample (solver.py:249-289) on the initial weights (seed 1234), batch 2 of
               synthetic_batch(seed=0), GMM noise from torch.manual_seed(300): per-output checksums and a 16x16
               average-pooled copy of every output image;
  resume_lr  : learning rates after Solver.resume (solver.py:359-381) from checkpoints named for iterations 1
               and 150000 (the reference re-steps its schedulers `iterations` more times: a quirk we mirror).

Usage:  python tests/golden/make_golden_extra.py      -> tests/golden/ref_extra.json
"""
import json
import os
import sys
import tempfile

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (import stubs + sys.path for the reference)
from oracle import dwc_oracle as O  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    import gmm as ref_gmm
    import tools as ref_tools
    out = {"torch": torch.__version__}

    # ---- GMM sampling: exact values
    cases = []
    for B, seed in ((1, 7), (4, 8), (16, 9)):
        g = torch.Generator().manual_seed(seed)
        mu = (torch.rand(B, 8, generator=g) > 0.5).float() * 2 - 1
        torch.manual_seed(1000 + seed)
        z = ref_tools.dist_sampling_split(mu, 8, 0.5, torch.device("cpu"))
        cases.append({"B": B, "seed": 1000 + seed, "mu": mu.tolist(), "z": z.tolist()})
    out["gmm_sample"] = cases

    # ---- earth-mover alternative
    g = torch.Generator().manual_seed(11)
    mus = [torch.randn(4, 8, generator=g).requires_grad_(True) for _ in range(8)]
    c = (torch.rand(4, 8, generator=g) > 0.5).float() * 2 - 1
    em = ref_gmm.gmm_earth_mover_distance_sp(mus, c)
    em.backward()
    out["gmm_em"] = {"seed": 11, "value": float(em), "grad": torch.cat([m.grad for m in mus], 1).tolist()}

    # ---- Solver.sample on the initial weights
    solver, cfg = MG.build_reference()
    batch = O.synthetic_batch(2, 128, seed=0)
    torch.manual_seed(300)
    with torch.no_grad():
        outs = solver.sample(batch["x_real"], batch["txt"], batch["txt_lens"])
    assert solver.training
    out["sample"] = {"B": 2, "batch_seed": 0, "noise_seed": 300, "n_outputs": len(outs),
                     "ck": [MG.checksum(o) for o in outs],
                     "pooled16": [F.adaptive_avg_pool2d(o, 16).tolist() for o in outs]}

    # ---- learning rate after resume
    lrs = {}
    with tempfile.TemporaryDirectory() as d:
        solver.copy_nets()
        for it in (0, 149999):
            for f in os.listdir(d):
                os.remove(os.path.join(d, f))
            solver.save(d, it)
            s2, cfg2 = MG.build_reference(seed=99)
            got = s2.resume(d, cfg2)
            assert got == it + 1
            lrs[str(it + 1)] = [s2.gen_opt.param_groups[0]["lr"], s2.dis_opt.param_groups[0]["lr"]]
            # the newest "gen" file is the EMA copy (utils.get_model_list sorts names; '_avg' sorts last)
            assert all(torch.equal(a, b) for a, b in zip(s2.gen.state_dict().values(),
                                                         solver.gen_copy.state_dict().values()))
    out["resume_lr"] = lrs
    with open(os.path.join(HERE, "ref_extra.json"), "w") as f:
        json.dump(out, f)
    print("wrote ref_extra.json", {k: (v if k in ("resume_lr", "torch") else "...") for k, v in out.items()})


if __name__ == "__main__":
    main()
