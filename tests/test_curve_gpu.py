"""Loss-curve parity over 200 training iterations against the UNMODIFIED reference (north_star: "loss curves must
agree over 200 steps").  The reference curves are committed fixtures (tests/golden/ref_curve_b{1,2}.json, recorded by
tests/golden/make_curve.py from /root/reference: CPU fp32, deterministic mode); the CUDA path runs the same
iterations in the bf16 product mode from the same seed / batch / GMM noise.

A GAN step is not contractive, so bf16 round-off moves individual iterations; the test bounds the per-iteration
deviation loosely and the windowed means tightly."""
import json
import os

import pytest
import torch

from oracle import dwc_oracle as O
from tests.util_gpu import build_solver, to_cuda

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDS = {b: json.load(open(os.path.join(HERE, "golden", "ref_curve_b%d.json" % b))) for b in (1, 2)}


def _window_means(v, w=20):
    return [sum(v[i:i + w]) / len(v[i:i + w]) for i in range(0, len(v), w)]


@pytest.mark.parametrize("mode,batch", [("bf16", 2), ("bf16", 1)])
def test_loss_curve_matches_reference(mode, batch):
    """batch 2: the text encoder's batch-row mixing is active over the whole curve; batch 1: the
    reference's own default batch size (configs/celeba_faces.yaml:13)."""
    GOLD = GOLDS[batch]
    steps = min(int(os.environ.get("DWC_CURVE_STEPS", GOLD["steps"])), GOLD["steps"])
    s, cfg = build_solver(mode)
    s.copy_nets()
    B = GOLD["B"]
    b = to_cuda(O.synthetic_batch(B, 128, seed=GOLD["batch_seed"]))
    eps = {}
    s.noise_hook = lambda tag: eps[tag]           # device tensors prepared per iteration (also keeps the run eager)
    names = list(GOLD["curve"].keys())
    mine = {n: [] for n in names}
    for it in range(steps):
        args = (b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, it)
        torch.manual_seed(100 + it)
        eps["dis1"] = torch.randn(1, 8, B, 8).cuda()
        s.dis_update(*args)
        torch.manual_seed(200 + it)
        eps["gen1"] = torch.randn(1, 8, B, 8).cuda()
        eps["gen2"] = torch.randn(1, 8, B, 8).cuda()
        s.gen_update(*args)
        s.smooth_moving()
        s.update_learning_rate()
        s.update_attention_status(it)
        for n in names:
            mine[n].append(float(getattr(s, n)))
    report = {}
    for n in names:
        ref = GOLD["curve"][n][:steps]
        scale = max(1e-3, max(abs(v) for v in ref))
        dev = max(abs(a - r) for a, r in zip(mine[n], ref)) / scale
        wdev = max(abs(a - r) for a, r in zip(_window_means(mine[n]), _window_means(ref))) / scale
        report[n] = (round(dev, 4), round(wdev, 4))
    print("max |ours - reference| / max|reference| per loss (per iteration, 20-iteration window means):", report)
    print("final: ours", {n: round(mine[n][-1], 4) for n in names}, "ref", {n: round(GOLD["curve"][n][steps - 1], 4) for n in names})
    # smooth losses: every iteration within 10 % of the curve's range, 20-iteration means within 5 %.
    # loss_dis is the adversarial term of a single over-fitted image: in the first ~25 iterations it jumps by several
    # units from one iteration to the next IN THE REFERENCE ITSELF (6.7, 6.3, ..., 1.55, 3.11, 2.0, ...), and which
    # iteration a jump lands on moves with summation order / bf16 round-off.  It is therefore held to the windowed
    # bound only (observed: 0.6 % - 3.4 % across kernel revisions).
    for n in ("loss_gen_total", "loss_dis", "loss_kl_x", "loss_kl_trg", "loss_gen_recon_x"):
        dev, wdev = report[n]
        if n != "loss_dis":
            assert dev < 0.10, (n, report[n])
        assert wdev < 0.05, (n, report[n])
    # the curve actually moved (training happened) and stayed finite
    assert all(torch.isfinite(torch.tensor(mine[n])).all() for n in names)
    assert abs(mine["loss_gen_total"][-1] - mine["loss_gen_total"][0]) > 1.0
