"""The other BASELINE.json configurations on the CUDA path: generator-only inference (configs[1]) and the scaled
256x256 variant with an extra down/up-sampling stage (configs[3]), against the CPU oracle."""
import pytest
import torch

from oracle import dwc_oracle as O
from tests.util_gpu import assert_grads, build_solver, cpu_state, grads_of, rel, to_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_inference_matches_oracle(mode, tol):
    s, cfg = build_solver(mode)
    s.eval()
    B = 3
    batch = O.synthetic_batch(B, 128, seed=5)
    b = to_cuda(batch)
    G = O.trainable(cpu_state(s.gen))
    with torch.no_grad():
        ref = O.translate(G, batch["x_real"], batch["txt"], batch["txt_lens"], use_attention=True)
        out = s(b["x_real"], b["txt"], b["txt_lens"])
    assert out.shape == (B, 3, 128, 128)
    assert rel(out.float(), ref) < tol, rel(out.float(), ref)


def test_inference_batch64_bf16():
    s, cfg = build_solver("bf16")
    s.eval()
    b = to_cuda(O.synthetic_batch(64, 128, seed=6))
    with torch.no_grad():
        out = s(b["x_real"], b["txt"], b["txt_lens"])
        out2 = s(b["x_real"], b["txt"], b["txt_lens"])
    assert out.shape == (64, 3, 128, 128) and torch.isfinite(out).all()
    assert torch.equal(out, out2)                      # the kernels are deterministic
    assert float(out.abs().max()) <= 1.0 + 1e-3        # attention blend of tanh output and the input image


@pytest.mark.parametrize("mode,ltol", [("fp32", 2e-4), ("bf16", 2e-2)])
def test_256_variant_training_step(mode, ltol):
    """BASELINE configs[3]: image_size 256, dis.image_size 256, gen.content_downsample 3, batch 2 (the
    text encoder's row mixing is active), one full D+G step; gradient bounds as in tests/test_step_gpu.py."""
    over = {"image_size": 256, "dis": {"image_size": 256}, "gen": {"content_downsample": 3}}
    s, cfg = build_solver(mode, overrides=over)
    ocfg = dict(O.DEFAULT_CFG, image_size=256, content_downsample=3)
    B = 2
    batch = O.synthetic_batch(B, 256, seed=2)
    b = to_cuda(batch)
    orc = O.OracleSolver(cpu_state(s.gen), cpu_state(s.dis), cfg=ocfg)
    orcq = O.OracleSolver(cpu_state(s.gen), cpu_state(s.dis), cfg=ocfg) if mode == "bf16" else None
    s.copy_nets()
    eps = {}
    s.noise_hook = lambda tag: eps[tag].cuda()
    args = (b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, 0)
    torch.manual_seed(100)
    eps["dis1"] = torch.randn(1, 8, B, 8)
    s.dis_update(*args)
    orc.dis_update(batch, eps["dis1"])
    if orcq is not None:
        with O.storage_rounding("bf16"):
            orcq.dis_update(batch, eps["dis1"])
    ld = float(s.loss_dis)
    assert abs(ld - orc.losses["loss_dis"]) < ltol * abs(ld), (ld, orc.losses["loss_dis"])
    assert_grads("256 dis", grads_of(s.dis), orc.last_dis_grads, orcq.last_dis_grads if orcq else None, mode)
    torch.manual_seed(200)
    eps["gen1"], eps["gen2"] = torch.randn(1, 8, B, 8), torch.randn(1, 8, B, 8)
    s.gen_update(*args)
    orc.gen_update(batch, eps["gen1"], eps["gen2"])
    if orcq is not None:
        with O.storage_rounding("bf16"):
            orcq.gen_update(batch, eps["gen1"], eps["gen2"])
    for name in ("loss_gen_total", "loss_gen_adv", "loss_gen_recon_x", "loss_kl_x", "loss_kl_trg", "loss_ds"):
        mine = float(getattr(s, name))
        assert abs(mine - orc.losses[name]) <= ltol * max(1.0, abs(orc.losses[name])), (name, mine, orc.losses[name])
    assert_grads("256 gen", grads_of(s.gen), orc.last_gen_grads, orcq.last_gen_grads if orcq else None, mode)
