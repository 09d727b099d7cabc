"""CPU checks of the drop-in boundary: state_dict keys/shapes and seed-for-seed initial weights must equal the
reference's (golden checksums recorded from the real reference by tests/golden/make_golden.py)."""
import copy
import json
import os

import torch

from dwc_gan_b200.solver import Solver
from dwc_gan_b200.utils import get_config

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = os.path.join(HERE, "golden", "celeba_faces.yaml")


def make_solver(seed=1234):
    cfg = get_config(CFG)
    cfg["vgg_w"] = 0
    torch.manual_seed(seed)
    return Solver(cfg, torch.device("cpu"), None), cfg


def test_init_matches_reference_golden():
    s, _ = make_solver()
    g = json.load(open(os.path.join(HERE, "golden", "ref_init_seed1234.json")))
    for net, key in ((s.gen, "gen"), (s.dis, "dis")):
        sd = net.state_dict()
        assert list(sd.keys()) == list(g[key].keys())
        for k, v in sd.items():
            ref = g[key][k]
            assert list(v.shape) == ref["shape"], k
            got = [float(v.double().sum()), float(v.double().abs().sum())]
            for a, b in zip(got, ref["ck"][:2]):
                assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (k, got, ref["ck"])
            assert torch.allclose(v.flatten()[:4].float(), torch.tensor(ref["head"]), atol=0, rtol=0), k


def test_flat_buffers_and_deepcopy():
    s, _ = make_solver()
    assert s.gen.flat.ok() and s.dis.flat.ok()
    n_gen = sum(p.numel() for p in s.gen.parameters())
    n_dis = sum(p.numel() for p in s.dis.parameters())
    assert (n_gen, n_dis) == (20356044, 13985666)          # parameter counts of the reference
    # conv weights are channels_last views of the flat buffer
    w = s.gen.enc_content.model[0].conv.weight
    assert w.shape == (64, 3, 7, 7) and w.permute(0, 2, 3, 1).is_contiguous()
    # fused head groups are contiguous in the flat buffer
    f = s.gen.flat
    o0 = f.offsets["enc_style.fcs.0.weight"]
    assert f.offsets["enc_style.fcvars.7.weight"] == o0 + 15 * 8 * 256
    assert f.offsets["dec.image_attention.conv.weight"] == f.offsets["dec.image_content.conv.weight"] + 3 * 49 * 64
    s.copy_nets()
    assert s.gen_copy.flat.ok() and s.gen_copy.flat.data.data_ptr() != f.data.data_ptr()
    sd, sd2 = s.gen.state_dict(), s.gen_copy.state_dict()
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    # state dict round trip through plain contiguous tensors (reference checkpoint format)
    plain = {k: v.contiguous().clone() for k, v in sd.items()}
    s.gen_copy.load_state_dict(plain)
    assert s.gen_copy.flat.ok()


def test_hot_path_refuses_cpu():
    s, cfg = make_solver()
    import pytest
    with pytest.raises(RuntimeError):
        s.gen.encode(torch.zeros(1, 3, 128, 128))


def test_adain_split_matches_slicing():
    """ops.AdainSplitFn (one transposing copy) = the reference's slice-by-slice split (networks.py:463-472), values and
    gradients, including parts that receive no gradient."""
    import torch
    from dwc_gan_b200 import ops
    torch.manual_seed(0)
    n, nl, f = 3, 4, 8
    p1 = torch.randn(n, nl * 2 * f, requires_grad=True)
    p2 = p1.detach().clone().requires_grad_(True)
    parts = ops.AdainSplitFn.apply(p1, nl, f)
    ref, rest = [], p2
    for _ in range(nl):
        ref += [rest[:, :f].contiguous().view(-1), rest[:, f:2 * f].contiguous().view(-1)]
        rest = rest[:, 2 * f:]
    assert len(parts) == 2 * nl
    for a, b in zip(parts, ref):
        assert torch.equal(a, b)
    w = [torch.randn(n * f) for _ in range(2 * nl)]
    used = [0, 1, 3, 6]                                   # the other parts get no gradient at all
    sum((parts[i] * w[i]).sum() for i in used).backward()
    sum((ref[i] * w[i]).sum() for i in used).backward()
    assert torch.allclose(p1.grad, p2.grad)


def test_weighted_sum_matches_python_sum():
    import torch
    from dwc_gan_b200 import ops
    torch.manual_seed(1)
    ts = [torch.randn((), requires_grad=True) for _ in range(5)]
    ws = [1.0, 0.5, 10.0, 0.0, 2.5]
    out = ops.weighted_sum(list(zip(ts, ws)))
    ref = sum(t.detach() * w for t, w in zip(ts, ws))
    assert torch.allclose(out.detach(), ref)
    out.backward()
    for t, w in zip(ts, ws):
        assert abs(float(t.grad) - w) < 1e-6
