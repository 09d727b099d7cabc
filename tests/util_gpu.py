"""Shared helpers of the GPU parity tests."""
import os

import torch

import dwc_gan_b200
from dwc_gan_b200.solver import Solver
from dwc_gan_b200.utils import get_config
from oracle import dwc_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = os.path.join(HERE, "golden", "celeba_faces.yaml")


def build_solver(mode, seed=1234, deterministic=True, overrides=None):
    dwc_gan_b200.set_mode(mode)
    cfg = get_config(CFG)
    cfg["vgg_w"] = 0
    if overrides:
        for k, v in overrides.items():
            if isinstance(v, dict):
                cfg[k].update(v)
            else:
                cfg[k] = v
    torch.manual_seed(seed)
    s = Solver(cfg, torch.device("cuda"), None).to("cuda")
    if deterministic:
        s.gen.enc_style.mapping[2].p = 0.0
        s.gen.enc_txt.dropout_in = 0.0
        s.gen.enc_txt.dropout_out = 0.0
        s.gen.enc_txt.lstm.dropout = 0.0
    return s, cfg


def cpu_state(net):
    return {k: v.detach().float().cpu().contiguous().clone() for k, v in net.state_dict().items()}


def to_cuda(batch):
    return {k: v.cuda() for k, v in batch.items()}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def grads_of(net):
    return {k: (p.grad.detach().float().cpu().contiguous().clone() if p.grad is not None else None)
            for k, p in net.named_parameters()}


def cancelled_bias(k):
    """Conv biases that feed InstanceNorm / AdaIN: the mean subtraction cancels them, so their gradient is exactly
    zero in exact arithmetic.  The CUDA path writes that zero; the reference / oracle compute round-off noise there
    (fp32: ~1e-8; with bf16 storage rounding: rounding noise of the summed gradient) - not comparable."""
    return k.endswith("conv.bias") and (k.startswith("enc_content.") or k.startswith("dec.model.0."))


def compare_grads(mine, ref, skip_rel=1e-5):
    """(worst per-tensor rel err, its key, global rel err).  Tensors whose reference gradient is pure round-off
    (biases in front of InstanceNorm/AdaIN: exactly zero in exact arithmetic) are left out."""
    worst, wk, num, den = 0.0, None, 0.0, 0.0
    gmax = max(float(g.double().norm()) for g in ref.values() if g is not None)
    skip_tiny = skip_rel * gmax
    for k, g in ref.items():
        if g is None or cancelled_bias(k):
            continue
        m = mine[k]
        assert m is not None, k
        gn = float(g.double().norm())
        num += float((m.double() - g.double()).norm()) ** 2
        den += gn ** 2
        if gn < skip_tiny:
            continue
        e = float((m.double() - g.double()).norm()) / gn
        if e > worst:
            worst, wk = e, k
    return worst, wk, (num ** 0.5) / (den ** 0.5 + 1e-30)


def per_tensor_errs(mine, ref, skip_rel=1e-5):
    """{key: relative L2 error} over tensors whose reference gradient is not pure round-off."""
    gmax = max(float(g.double().norm()) for g in ref.values() if g is not None)
    out = {}
    for k, g in ref.items():
        if g is None or cancelled_bias(k):
            continue
        gn = float(g.double().norm())
        if gn < skip_rel * gmax:
            continue
        out[k] = float((mine[k].double() - g.double()).norm()) / gn
    return out


def params_of(net):
    return {k: p.detach().float().cpu().contiguous().clone() for k, p in net.named_parameters()}


def update_errs(p0, p1_mine, p1_ref):
    """Per tensor: ||dp_mine - dp_ref|| / ||dp_ref|| for the parameter update dp = p1 - p0 of one optimizer step, and
    the largest absolute parameter difference."""
    out = {}
    for k, ref in p1_ref.items():
        d_ref = ref.double() - p0[k].double()
        d_mine = p1_mine[k].double() - p0[k].double()
        dn = float(d_ref.norm())
        if dn == 0.0:
            assert float(d_mine.norm()) == 0.0, ("parameter moved that the reference left alone", k)
            continue
        out[k] = (float((d_mine - d_ref).norm()) / dn, float((p1_mine[k].double() - ref.double()).abs().max()))
    return out


def top_errs(mine, ref, n=6):
    """The n worst tensors as 'key:err' strings (diagnostics printed by the parity tests)."""
    e = sorted(per_tensor_errs(mine, ref).items(), key=lambda kv: -kv[1])[:n]
    return ", ".join("%s:%.2e" % kv for kv in e)


def assert_grads(tag, mine, ref32, refq, mode, glob_fp32=2e-3, worst_fp32=2e-2):
    """Gradient check shared by the parity tests.  fp32 validation mode: absolute bounds (global / worst tensor)
    against the fp32 oracle.  bf16: against the fp32 oracle, bounded per tensor and globally by 1.5 x the deviation of the
    oracle's own bf16-storage-rounding run (`refq`), see tests/test_step_gpu.py."""
    worst, wk, glob = compare_grads(mine, ref32)
    if mode == "fp32":
        print("%s [fp32]: grads vs fp32 oracle worst %.3e (%s) global %.3e" % (tag, worst, wk, glob))
        assert glob < glob_fp32 and worst < worst_fp32, (tag, worst, wk, glob)
        return
    iworst, ik, iglob = compare_grads(refq, ref32)
    print("%s [bf16]: grads vs fp32 oracle worst %.3e (%s) global %.3e | rounding oracle vs fp32 worst %.3e (%s) "
          "global %.3e" % (tag, worst, wk, glob, iworst, ik, iglob))
    assert glob < 1.5 * iglob + 1e-3, (tag, glob, iglob)
    inh = per_tensor_errs(refq, ref32)
    for k, e in per_tensor_errs(mine, ref32).items():
        assert e < 1.5 * inh.get(k, 0.0) + 2e-2, (tag, k, e, inh.get(k))
