"""Solver.sample (solver.py:249-289) and the `em` style distance (gmm.py:33-41) against goldens recorded
from the unmodified reference (tests/golden/make_golden_extra.py)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from dwc_gan_b200.gmm import gmm_earth_mover_distance_sp
from oracle import dwc_oracle as O
from tests.util_gpu import build_solver, to_cuda

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXTRA = json.load(open(os.path.join(HERE, "golden", "ref_extra.json")))


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_sample_matches_reference(mode, tol):
    g = EXTRA["sample"]
    s, cfg = build_solver(mode, deterministic=False)      # sample() switches to eval() itself, as the reference does
    B = g["B"]
    b = to_cuda(O.synthetic_batch(B, 128, seed=g["batch_seed"]))
    torch.manual_seed(g["noise_seed"])
    draws = [torch.randn(1, 8, 1, 8) for _ in range(B)]   # one Normal.sample((1, 8)) per image, in image order
    it = iter(draws)
    s.noise_hook = lambda tag: next(it).cuda()
    outs = s.sample(b["x_real"], b["txt"], b["txt_lens"])
    assert s.training                                      # back in train mode (solver.py:288)
    assert len(outs) == g["n_outputs"] == 5                # x_real, reconstruction, text-driven, sampled, attention
    for i, (o, ck, pooled) in enumerate(zip(outs, g["ck"], g["pooled16"])):
        assert o.shape == (B, 3, 128, 128)
        o = o.float().cpu()
        want = torch.tensor(pooled)
        got = F.adaptive_avg_pool2d(o, 16)
        err = float((got - want).norm() / want.norm())
        print("sample output", i, mode, "pooled rel err", err)
        assert err < tol, (i, err)
        for img in range(B):                               # per image
            e = float((got[img] - want[img]).norm() / want[img].norm())
            assert e < 2 * tol, (i, img, e)
        assert abs(float(o.double().abs().sum()) - ck[1]) <= tol * ck[1], (i, ck)


def test_earth_mover_distance_matches_reference():
    g = EXTRA["gmm_em"]
    gen = torch.Generator().manual_seed(g["seed"])
    mus = [torch.randn(4, 8, generator=gen) for _ in range(8)]
    c = (torch.rand(4, 8, generator=gen) > 0.5).float() * 2 - 1
    leaves = [m.cuda().requires_grad_(True) for m in mus]
    loss = gmm_earth_mover_distance_sp(leaves, c.cuda())
    assert abs(float(loss) - g["value"]) <= 1e-5 * abs(g["value"])
    loss.backward()
    grad = torch.cat([m.grad for m in leaves], 1).cpu()
    assert float((grad - torch.tensor(g["grad"])).abs().max()) <= 1e-6
