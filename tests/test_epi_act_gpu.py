"""Activation + reflect halo written by the conv epilogue (norm-less Conv2dBlocks of the style encoder and the
discriminator, networks.py:531,556-567) is the same program as the separate pass: identical losses and gradients."""
import pytest
import torch

from dwc_gan_b200 import ops
from tests.util_gpu import build_solver, to_cuda
from oracle import dwc_oracle as O

pytestmark = pytest.mark.gpu


def _grads(fused):
    old = ops.RT.epi_act
    ops.RT.epi_act = fused
    try:
        s, cfg = build_solver("bf16", deterministic=True)
        s.use_cuda_graphs = False
        s.copy_nets()
        b = to_cuda(O.synthetic_batch(4, 128, seed=5))
        torch.manual_seed(11)
        args = (b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, 0)
        n0 = ops.RT.launches
        s.dis_update(*args)
        gd = s.dis.flat.grad.clone()
        s.gen_update(*args)
        gg = s.gen.flat.grad.clone()
        torch.cuda.synchronize()
        return (float(s.loss_dis), float(s.loss_gen_total)), gd, gg, ops.RT.launches - n0
    finally:
        ops.RT.epi_act = old


def test_epilogue_activation_matches_separate_pass():
    la, gda, gga, na = _grads(True)
    lb, gdb, ggb, nb = _grads(False)
    assert la == lb, (la, lb)
    assert torch.equal(gda, gdb)
    assert torch.equal(gga, ggb)


def test_step_is_bit_reproducible():
    """Two runs of the same D+G step from the same seeds: identical losses and gradients, bit for bit (every reduction
    - split-K weight gradients, cluster split-K convolutions, loss sums - has a fixed order)."""
    la, gda, gga, _ = _grads(False)
    lb, gdb, ggb, _ = _grads(False)
    assert la == lb, (la, lb)
    assert torch.equal(gda, gdb)
    assert torch.equal(gga, ggb)
