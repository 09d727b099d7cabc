"""Bit-exact pieces of the path (north_star: "GMM component selection and token indexing must be bit-exact"):
GMM style sampling against the reference's own draws (goldens recorded from tools.dist_sampling_split,
tools.py:65-70), the token -> embedding gather + style concat (networks_v2.py:217-223) and the batch-row mixing of
the final LSTM states (networks_v2.py:248-249)."""
import ctypes as C
import json
import os

import pytest
import torch
import torch.nn.functional as F

from dwc_gan_b200 import _lib as L
from dwc_gan_b200.tools import asign_label, dist_sampling_split
from oracle import dwc_oracle as O
from tests.util_gpu import build_solver, cpu_state, to_cuda

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXTRA = json.load(open(os.path.join(HERE, "golden", "ref_extra.json")))


@pytest.mark.parametrize("case", EXTRA["gmm_sample"], ids=lambda c: "B%d" % c["B"])
def test_gmm_sample_bit_exact_vs_reference(case):
    B = case["B"]
    mu = torch.tensor(case["mu"], dtype=torch.float32)
    want = torch.tensor(case["z"], dtype=torch.float32)
    torch.manual_seed(case["seed"])
    eps = torch.randn(1, 8, B, 8)                       # the stream Normal(mu, 0.5).sample((1, 8)) consumed
    z = dist_sampling_split(mu.cuda(), 8, 0.5, torch.device("cuda"), eps=eps.cuda())
    assert z.shape == (B, 64) and not z.requires_grad
    assert torch.equal(z.cpu(), want)                   # component selection, layout and arithmetic: exact
    assert torch.equal(z.cpu(), O.gmm_sample(mu, eps, 0.5))
    # a standard deviation that is not a power of two: still the two-rounding arithmetic of torch.normal
    z3 = dist_sampling_split(mu.cuda(), 8, 0.3, torch.device("cuda"), eps=eps.cuda())
    assert torch.equal(z3.cpu(), O.gmm_sample(mu, eps, 0.3))
    # labels {0,1} -> component means {-1,+1}
    lab = (mu + 1) / 2
    assert torch.equal(asign_label(lab.cuda()).cpu(), mu)


def test_embedding_gather_and_style_concat_bit_exact():
    torch.manual_seed(3)
    B, T, E, S = 5, 80, 300, 64
    batch = O.synthetic_batch(B, 128, seed=11)
    tokens = batch["txt"]
    table = torch.randn(102, E)
    table[0] = 0                                         # padding row (nn.Embedding(padding_idx=0))
    style = torch.randn(B, S)
    x = torch.empty(T, B, E + S, dtype=torch.float32, device="cuda")
    L.check(L.lib().dwc_embed_concat_fwd(L.ptr(tokens.cuda()), L.ptr(table.cuda()), L.ptr(style.cuda()), None,
                                         L.ptr(x), B, T, E, S, L.stream()))
    want = torch.cat([F.embedding(tokens.t(), table, padding_idx=0), style.unsqueeze(0).expand(T, -1, -1)], dim=-1)
    assert torch.equal(x.cpu(), want)


def test_text_encoder_rows_depend_on_other_samples_like_the_reference():
    """The quirk end to end: at B > 1 a row of the text feature is built from OTHER samples' final states, so changing
    only sample 1's tokens changes sample 0's style code - by exactly what the oracle says."""
    s, _ = build_solver("fp32")
    G = O.trainable(cpu_state(s.gen))
    B = 4
    batch = O.synthetic_batch(B, 128, seed=12)
    style = torch.randn(B, 64)
    txt2 = batch["txt"].clone()
    n1 = int(batch["txt_lens"][1])
    txt2[1, 1:n1 - 1] = (txt2[1, 1:n1 - 1] + 7 - 4) % 98 + 4
    outs = []
    for txt in (batch["txt"], txt2):
        with torch.no_grad():
            mu, _ = s.gen.encode_txt(style.cuda(), txt.cuda(), batch["txt_lens"].cuda())
            mu_ref, _ = O.text_encoder(G, style, txt, batch["txt_lens"])
        mu, mu_ref = torch.cat(mu, 1).cpu(), torch.cat(mu_ref, 1)
        assert float((mu - mu_ref).abs().max()) <= 1e-4 * float(mu_ref.abs().max())
        outs.append((mu, mu_ref))
    d_mine = (outs[1][0] - outs[0][0])[0]
    d_ref = (outs[1][1] - outs[0][1])[0]
    assert float(d_ref.abs().max()) > 1e-3              # row 0 moved although sample 0's tokens did not change
    assert float((d_mine - d_ref).abs().max()) <= 1e-4 * max(1.0, float(d_ref.abs().max()))
