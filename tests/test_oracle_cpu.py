"""The CPU oracle (oracle/dwc_oracle.py) is pinned against goldens recorded from the UNMODIFIED reference
(tests/golden/make_golden.py imports /root/reference in the build container): same seed -> same initial weights
(test_init_cpu.py), and one full D+G step on the same synthetic batch reproduces the reference's losses,
gradient checksums and inference checksums."""
import json
import os

import torch

from oracle import dwc_oracle as O
from tests.test_init_cpu import make_solver

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_step_b2.json")))


def _ck(t):
    t = t.double()
    return [float(t.sum()), float(t.abs().sum())]


def test_oracle_step_matches_reference_goldens():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    s, cfg = make_solver()                 # parameter container: bit-identical init to the reference (golden-checked)
    gen_sd = {k: v.detach().clone().contiguous() for k, v in s.gen.state_dict().items()}
    dis_sd = {k: v.detach().clone().contiguous() for k, v in s.dis.state_dict().items()}
    orc = O.OracleSolver(gen_sd, dis_sd)
    B = GOLD["B"]
    batch = O.synthetic_batch(B, 128, seed=GOLD["batch_seed"])
    gold = GOLD["steps"][0]

    # inference on the initial weights (raw decoder image head, no attention blend), as recorded from the reference
    with torch.no_grad():
        img = O.translate(orc.G, batch["x_real"], batch["txt"], batch["txt_lens"], use_attention=False)
    for a, b in zip(_ck(img), GOLD["infer_img_ck"][:2]):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), ("infer img", a, b)

    torch.manual_seed(100)
    orc.dis_update(batch, torch.randn(1, 8, B, 8))
    assert abs(orc.losses["loss_dis"] - gold["loss_dis"]) < 1e-4 * abs(gold["loss_dis"])
    for k, ref in gold["dis_grad"].items():
        g = orc.last_dis_grads[k]
        got = _ck(g)
        assert abs(got[1] - ref[1]) <= 2e-2 * max(1e-6, abs(ref[1])), (k, got, ref)      # |grad| mass per tensor

    torch.manual_seed(200)
    e1, e2 = torch.randn(1, 8, B, 8), torch.randn(1, 8, B, 8)
    orc.gen_update(batch, e1, e2)
    for name, ref in gold["losses"].items():
        if name in orc.losses:
            assert abs(orc.losses[name] - ref) <= 2e-4 * max(1.0, abs(ref)), (name, orc.losses[name], ref)
    num = den = 0.0
    for k, ref in gold["gen_grad"].items():
        g = orc.last_gen_grads[k]
        assert (g is None) == (ref is None), k
        if g is None:
            continue
        got = _ck(g)
        num += (got[1] - ref[1]) ** 2
        den += ref[1] ** 2
    assert (num / den) ** 0.5 < 5e-3, (num, den)


def test_bf16_storage_rounding_is_chaotic_at_the_1e2_level():
    """Why the bf16 parity bound cannot be tight (profiles/r02_bf16_inherent_error.md): with bf16 storage rounding a
    perturbation of 1e-6 - the size of a different fp32 summation order - reaches ~1e-2 in the content code, about as
    far as bf16 storage is from fp32 in the first place; the fp32 oracle moves by ~1e-6."""
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    s, _ = make_solver()
    G = O.trainable({k: v.detach().clone().contiguous() for k, v in s.gen.state_dict().items()})
    x = O.synthetic_batch(2, 128, seed=3)["x_real"]
    noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(0)) * 1e-6
    rel = lambda a, b: float((a - b).norm() / b.norm())
    out = {}
    for mode in ("fp32", "bf16"):
        with torch.no_grad(), O.storage_rounding(mode):
            out[mode] = (O.content_encoder(G, x), O.content_encoder(G, x + noise))
    assert rel(out["fp32"][1], out["fp32"][0]) < 1e-5
    moved = rel(out["bf16"][1], out["bf16"][0])
    inherent = rel(out["bf16"][0], out["fp32"][0])
    assert 2e-3 < moved < 3e-2 and 5e-3 < inherent < 3e-2, (moved, inherent)
    assert moved > 0.3 * inherent, (moved, inherent)
    # rounding happens where the CUDA path stores: the rounded forward is idempotent under a second rounding
    assert torch.equal(out["bf16"][0], out["bf16"][0].to(torch.bfloat16).float())
