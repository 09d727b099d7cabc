"""CUDA-graph replay of the update phases is the same program as the eager calls: identical losses and
parameters after several steps (same seeds, dropout on, GMM noise from torch's generator)."""
import pytest
import torch

from tests.util_gpu import build_solver, to_cuda
from oracle import dwc_oracle as O

pytestmark = pytest.mark.gpu


def _run(use_graphs, steps, mode):
    s, cfg = build_solver(mode, deterministic=False)
    s.use_cuda_graphs = use_graphs
    s.copy_nets()
    b = to_cuda(O.synthetic_batch(4, 128, seed=3))
    losses = []
    for it in range(steps):
        torch.manual_seed(500 + it)
        args = (b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, it)
        s.dis_update(*args)
        s.gen_update(*args)
        s.smooth_moving()
        s.update_learning_rate()
        s.update_attention_status(it)
        losses.append((float(s.loss_dis), float(s.loss_gen_total), float(s.loss_kl_trg), float(s.loss_ds)))
    torch.cuda.synchronize()
    return s, losses


@pytest.mark.parametrize("mode", ["bf16"])
def test_graph_replay_matches_eager(mode):
    steps = 7       # iteration 0 with attention, 1-2 eager without, 3.. replayed
    se, le = _run(False, steps, mode)
    sg, lg = _run(True, steps, mode)
    assert any(e["graph"] is not None for e in sg._graphs.values()), "no phase was captured"
    for it, (a, b) in enumerate(zip(le, lg)):
        for x, y in zip(a, b):
            assert abs(x - y) <= 2e-3 * max(1.0, abs(x)), (it, a, b)
    for (k, p), (_, q) in zip(se.gen.named_parameters(), sg.gen.named_parameters()):
        assert float((p - q).abs().max()) <= 5e-4, k       # 7 Adam steps of lr 1e-4 move a weight by <= 7e-4
    for (k, p), (_, q) in zip(se.gen_copy.named_parameters(), sg.gen_copy.named_parameters()):
        assert float((p - q).abs().max()) <= 5e-4, k
    # step counters advanced identically (attention head skipped while attention is off)
    assert se.gen_opt.steps == sg.gen_opt.steps and se.dis_opt.steps == sg.dis_opt.steps
