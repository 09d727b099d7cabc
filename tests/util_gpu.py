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


def compare_grads(mine, ref, skip_rel=1e-5):
    """(worst per-tensor rel err, its key, global rel err).  Tensors whose reference gradient is pure round-off
    (biases in front of InstanceNorm/AdaIN: exactly zero in exact arithmetic) are left out of the per-tensor figure."""
    worst, wk, num, den = 0.0, None, 0.0, 0.0
    gmax = max(float(g.double().norm()) for g in ref.values() if g is not None)
    skip_tiny = skip_rel * gmax
    for k, g in ref.items():
        if g is None:
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
