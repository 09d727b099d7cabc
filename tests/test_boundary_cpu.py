"""CPU checks of the Solver-level drop-in boundary that need no kernels: the reference's write_loss attribute
filter, checkpoint save / resume / init_network - against the real reference when /root/reference is
present, against goldens recorded from it otherwise -, the optimizer state layout, and the text encoder's
batch-row mixing (networks_v2.py:248-249) as exact data movement."""
import json
import os
import sys
import tempfile
import types

import pytest
import torch

from tests.test_init_cpu import make_solver

HERE = os.path.dirname(os.path.abspath(__file__))
EXTRA = json.load(open(os.path.join(HERE, "golden", "ref_extra.json")))
REF = "/root/reference"


def _write_loss_members(trainer):
    # the filter of the reference's utils.write_loss (utils.py:132-136), restated
    return [a for a in dir(trainer) if not callable(getattr(trainer, a)) and not a.startswith("__")
            and ("loss" in a or "grad" in a or "nwd" in a)]


def test_write_loss_filter_only_sees_loss_scalars():
    s, _ = make_solver()
    assert _write_loss_members(s) == []                 # fresh trainer: nothing to log, nothing that would break add_scalar
    s.loss_gen_total = torch.tensor(1.0)
    s.loss_gen_vgg = 0
    s.copy_nets()
    members = _write_loss_members(s)
    assert members == ["loss_gen_total", "loss_gen_vgg"]
    for m in members:
        v = getattr(s, m)
        assert isinstance(v, (int, float)) or (isinstance(v, torch.Tensor) and v.numel() == 1)


def test_final_state_row_mixing_is_the_reference_layout():
    from dwc_gan_b200.text import final_state_rows, final_state_rows_bwd
    L, H = 2, 5
    for B in (1, 2, 3, 4, 16):
        g = torch.Generator().manual_seed(B)
        fh = [torch.randn(B, 2 * H, generator=g) for _ in range(L)]
        fc = [torch.randn(B, 2 * H, generator=g) for _ in range(L)]
        got = final_state_rows(fh, fc)
        # the reference's expression, verbatim semantics (networks_v2.py:248-249)
        ref = torch.cat([torch.stack(fh, 0), torch.stack(fc, 0)], dim=1).view(B, -1)
        assert torch.equal(got, ref)
        # independent restatement by index arithmetic: the flat sequence of 2H-vectors is
        # layer 0: h_0..h_{B-1}, c_0..c_{B-1}; layer 1: ...; row r takes vectors 4r..4r+3 of it
        seq = []
        for l in range(L):
            seq += [fh[l][b] for b in range(B)] + [fc[l][b] for b in range(B)]
        for r in range(B):
            assert torch.equal(got[r], torch.cat(seq[4 * r:4 * r + 4]))
        if B == 1:                                        # the intended layout only at batch 1
            assert torch.equal(got[0], torch.cat([fh[0][0], fc[0][0], fh[1][0], fc[1][0]]))
        # backward is the exact adjoint
        leaves = [t.clone().requires_grad_(True) for t in fh + fc]
        w = torch.randn(B, 4 * L * H, generator=g)
        (final_state_rows(leaves[:L], leaves[L:]) * w).sum().backward()
        dfh, dfc = final_state_rows_bwd(w, L, B, H)
        for l in range(L):
            assert torch.equal(dfh[l], leaves[l].grad) and torch.equal(dfc[l], leaves[L + l].grad)


def test_optimizer_state_dict_is_in_parameters_order():
    s, cfg = make_solver()
    opt = s.gen_opt
    opt._ensure_state()
    named = dict(s.gen.named_parameters())
    g = torch.Generator().manual_seed(0)
    opt.m.copy_(torch.randn(opt.m.shape, generator=g))
    opt.v.copy_(torch.rand(opt.v.shape, generator=g))
    for i, n in enumerate(opt.steps):
        opt.steps[n] = 1 + (i % 3)
    sd = opt.state_dict()
    # the reference's optimizer over the same module (solver.py:63-68) accepts it and lands every moment on its parameter
    ref = torch.optim.Adam([p for p in s.gen.parameters() if p.requires_grad], lr=cfg["lr"],
                           betas=(cfg["beta1"], cfg["beta2"]), weight_decay=cfg["weight_decay"])
    ref.load_state_dict(sd)
    order = [n for n, p in s.gen.named_parameters() if p.requires_grad]
    assert len(sd["state"]) == len(order) == len(sd["param_groups"][0]["params"])
    for n in order:
        st = ref.state[named[n]]
        assert int(st["step"]) == opt.steps[n], n
        assert torch.equal(st["exp_avg"], s.gen.flat._view(opt.m, n, named[n].shape)), n
        assert torch.equal(st["exp_avg_sq"], s.gen.flat._view(opt.v, n, named[n].shape)), n
    # and back: a torch.optim.Adam state dict loads into the fused optimizer
    s2, _ = make_solver(seed=5)
    s2.gen_opt.load_state_dict(ref.state_dict())
    assert s2.gen_opt.steps == opt.steps
    named2 = dict(s2.gen.named_parameters())
    for n in order:                                       # (the flat buffers also hold alignment padding)
        assert torch.equal(s2.gen.flat._view(s2.gen_opt.m, n, named2[n].shape), s.gen.flat._view(opt.m, n, named[n].shape)), n
        assert torch.equal(s2.gen.flat._view(s2.gen_opt.v, n, named2[n].shape), s.gen.flat._view(opt.v, n, named[n].shape)), n


def test_save_resume_round_trip_and_lr_quirk():
    s, cfg = make_solver()
    s.copy_nets()
    with torch.no_grad():                                 # make the EMA copy differ from the live weights
        for p in s.gen_copy.parameters():
            p.mul_(0.5)
        for p in s.dis_copy.parameters():
            p.add_(0.25)
    with tempfile.TemporaryDirectory() as d:
        for it, want in ((0, EXTRA["resume_lr"]["1"]), (149999, EXTRA["resume_lr"]["150000"])):
            for f in os.listdir(d):
                os.remove(os.path.join(d, f))
            s.save(d, it)
            assert sorted(os.listdir(d)) == sorted(["gen_%08d.pt" % (it + 1), "dis_%08d.pt" % (it + 1),
                                                    "gen_%08d_avg.pt" % (it + 1), "dis_%08d_avg.pt" % (it + 1),
                                                    "optimizer.pt"])
            ck = torch.load(os.path.join(d, "gen_%08d.pt" % (it + 1)))
            assert list(ck.keys()) == ["a"] and list(ck["a"].keys()) == list(s.gen.state_dict().keys())
            s2, cfg2 = make_solver(seed=77)
            assert s2.resume(d, cfg2) == it + 1
            # like the reference, the newest file of a kind is the '_avg' one (utils.py:169-178 sorts the names)
            for (k, a), b in zip(s2.gen.state_dict().items(), s.gen_copy.state_dict().values()):
                assert torch.equal(a, b), k
            for (k, a), b in zip(s2.dis.state_dict().items(), s.dis_copy.state_dict().values()):
                assert torch.equal(a, b), k
            assert s2.gen.flat.ok() and s2.dis.flat.ok()
            # learning rate after resume: the reference's value, recorded from it (re-stepped schedulers)
            got = [s2.gen_opt.param_groups[0]["lr"], s2.dis_opt.param_groups[0]["lr"]]
            assert got == pytest.approx(want, rel=1e-12), (it, got, want)
        # init_network: everything but the token embedding is taken from the files
        s3, _ = make_solver(seed=78)
        emb0 = s3.gen.enc_txt.embed_tokens.weight.detach().clone()
        s3.init_network(os.path.join(d, "gen_%08d.pt" % 150000), os.path.join(d, "dis_%08d.pt" % 150000))
        for (k, a), b in zip(s3.gen.state_dict().items(), s.gen.state_dict().values()):
            assert torch.equal(a, emb0 if "embed_tokens" in k else b), k
        for (k, a), b in zip(s3.dis.state_dict().items(), s.dis.state_dict().values()):
            assert torch.equal(a, b), k


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_checkpoints_cross_load_with_the_real_reference():
    for name in ("torchfile", "tensorboardX"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.load = lambda *a, **k: None
            m.SummaryWriter = object
            sys.modules[name] = m
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ("utils", "solver", "networks", "tools",
                                                                              "gmm", "vocab")}
    sys.path.insert(0, REF)
    try:
        import utils as ref_utils
        from solver import Solver as RefSolver
        rcfg = ref_utils.get_config(os.path.join(REF, "configs/celeba_faces.yaml"))
        rcfg["vgg_w"] = 0
        torch.manual_seed(4321)
        ref = RefSolver(rcfg, torch.device("cpu"), None)
        ref.copy_nets()
        ours, cfg = make_solver(seed=9)
        with tempfile.TemporaryDirectory() as d:
            ref.save(d, 41)                               # reference writes ...
            assert ours.resume(d, cfg) == 42              # ... we read
            for (k, a), (k2, b) in zip(ours.gen.state_dict().items(), ref.gen_copy.state_dict().items()):
                assert k == k2 and torch.equal(a, b), k
            for (k, a), (k2, b) in zip(ours.dis.state_dict().items(), ref.dis_copy.state_dict().items()):
                assert k == k2 and torch.equal(a, b), k
            ours.init_network(os.path.join(d, "gen_%08d.pt" % 42), os.path.join(d, "dis_%08d.pt" % 42))
        ours2, _ = make_solver(seed=10)
        ours2.copy_nets()
        with tempfile.TemporaryDirectory() as d:
            ours2.save(d, 6)                              # we write ...
            torch.manual_seed(1)
            ref2 = RefSolver(rcfg, torch.device("cpu"), None)
            assert ref2.resume(d, rcfg) == 7              # ... the reference reads
            for (k, a), (k2, b) in zip(ref2.gen.state_dict().items(), ours2.gen_copy.state_dict().items()):
                assert k == k2 and torch.equal(a, b), k
            for (k, a), (k2, b) in zip(ref2.dis.state_dict().items(), ours2.dis_copy.state_dict().items()):
                assert k == k2 and torch.equal(a, b), k
            ref2.init_network(os.path.join(d, "gen_%08d.pt" % 7), os.path.join(d, "dis_%08d.pt" % 7))
            # the optimizer file the reference writes (and never reads) has the shape its Adam would accept
            osd = torch.load(os.path.join(d, "optimizer.pt"))
            ref2.gen_opt.load_state_dict(osd["gen"])
            ref2.dis_opt.load_state_dict(osd["dis"])
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
