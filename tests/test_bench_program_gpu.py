"""The BENCHMARKED program against the oracle: batch 16, bf16 product mode, CUDA-graph replay (with its 3B pass
batching and forked streams), dropout off so that the oracle can follow.  Three iterations warm the graphs up (one
with attention, two eager without); the fourth is a pure graph replay and is compared - losses and every gradient -
with the oracle started from the weights the GPU holds at that point, run twice: plain fp32 and with bf16 storage
rounding (see tests/test_step_gpu.py for why the bound is "no worse than 1.5 x what bf16 storage costs the oracle
itself").  GMM noise reaches the captured graph through persistent device buffers (Solver.noise_buffers)."""
import pytest
import torch

from oracle import dwc_oracle as O
from tests.util_gpu import build_solver, compare_grads, cpu_state, grads_of, per_tensor_errs, to_cuda, top_errs

pytestmark = pytest.mark.gpu


def _bounded(tag, mine, ref32, refq):
    """Gradients against the fp32 oracle, bounded per tensor and globally by what bf16 storage costs the oracle."""
    worst, wk, glob = compare_grads(mine, ref32)
    iworst, ik, iglob = compare_grads(refq, ref32)
    qworst, qk, qglob = compare_grads(mine, refq)
    print("B=16 graph replay, %s grads: vs fp32 oracle worst %.3e (%s) global %.3e | rounding oracle vs fp32 worst %.3e "
          "(%s) global %.3e | vs rounding oracle worst %.3e (%s) global %.3e" % (tag, worst, wk, glob, iworst, ik, iglob,
                                                                                 qworst, qk, qglob))
    print("   worst tensors vs fp32:", top_errs(mine, ref32))
    assert glob < 1.5 * iglob + 1e-3, (tag, glob, iglob)
    inh = per_tensor_errs(refq, ref32)
    for k, e in per_tensor_errs(mine, ref32).items():
        assert e < 1.5 * inh[k] + 2e-2, (tag, k, e, inh[k])
    assert qglob < iglob, (tag, qglob, iglob)


def test_batch16_graph_replay_matches_oracle():
    B = 16
    s, cfg = build_solver("bf16")
    assert s.use_cuda_graphs
    s.copy_nets()
    batch = O.synthetic_batch(B, 128, seed=21)
    b = to_cuda(batch)
    s.noise_buffers = {k: torch.zeros(1, 8, B, 8, device="cuda") for k in ("dis1", "gen1", "gen2")}
    args = lambda it: (b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, it)

    def noise(it):
        torch.manual_seed(900 + it)
        e = {k: torch.randn(1, 8, B, 8) for k in ("dis1", "gen1", "gen2")}
        for k, v in e.items():
            s.noise_buffers[k].copy_(v)
        return e

    for it in range(3):
        noise(it)
        s.dis_update(*args(it))
        s.gen_update(*args(it))
        s.smooth_moving()
        s.update_learning_rate()
        s.update_attention_status(it)
    it = 3
    assert not s.use_attention
    e = noise(it)
    orc = O.OracleSolver(cpu_state(s.gen), cpu_state(s.dis))             # bf16 storage rounding
    orc32 = O.OracleSolver(cpu_state(s.gen), cpu_state(s.dis))           # plain fp32
    for o in (orc, orc32):
        o.use_attention = False
        o.ds_w = s.init_ds_w
    launches0 = len([g for g in s._graphs.values() if g["graph"] is not None])
    s.dis_update(*args(it))
    with O.storage_rounding("bf16"):
        orc.dis_update(batch, e["dis1"])
    orc32.dis_update(batch, e["dis1"])
    ld = float(s.loss_dis)
    assert abs(ld - orc32.losses["loss_dis"]) <= 2e-2 * abs(ld), (ld, orc32.losses["loss_dis"])
    assert abs(ld - orc.losses["loss_dis"]) <= 2e-3 * abs(ld), (ld, orc.losses["loss_dis"])
    _bounded("dis", grads_of(s.dis), orc32.last_dis_grads, orc.last_dis_grads)
    for o in (orc, orc32):                                               # the G phase sees the D this GPU step produced
        o.D = {k: v.clone() for k, v in cpu_state(s.dis).items()}
    s.gen_update(*args(it))
    with O.storage_rounding("bf16"):
        orc.gen_update(batch, e["gen1"], e["gen2"])
    orc32.gen_update(batch, e["gen1"], e["gen2"])
    captured = [k[0] for k, g in s._graphs.items() if g["graph"] is not None]
    assert sorted(captured) == ["dis", "gen"] and launches0 == 0, (captured, launches0)   # iteration 3 was the replay
    for name in ("loss_gen_total", "loss_gen_adv", "loss_gen_recon_x", "loss_gen_recon_c_real", "loss_gen_recon_s_fake",
                 "loss_gen_cycrecon_x", "loss_kl_x", "loss_kl_trg", "loss_ds"):
        mine, ref, ref32 = float(getattr(s, name)), orc.losses[name], orc32.losses[name]
        assert abs(mine - ref32) <= 2e-2 * max(1.0, abs(ref32)), (name, mine, ref32)
        assert abs(mine - ref) <= 5e-3 * max(1.0, abs(ref)), (name, mine, ref)
    assert abs(s.init_ds_w - orc.ds_w) < 1e-12
    _bounded("gen", grads_of(s.gen), orc32.last_gen_grads, orc.last_gen_grads)
