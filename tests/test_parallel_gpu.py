"""Data-parallel parity on the GPU: two ranks, each with its own shard, must equal two oracle replicas
stepping their own shards with averaged gradients (the text encoder's row mixing makes the forward depend on the
LOCAL batch, so this - not one oracle run at the global batch - is the definition).  NCCL over two GPUs when the
box has them; otherwise both ranks share cuda:0 and the same GradSync / fused-Adam code runs over gloo."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ngpu, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank if ngpu >= world else 0))
        import torch.distributed as dist
        from dwc_gan_b200 import parallel
        from oracle import dwc_oracle as O
        from tests.util_gpu import build_solver, cancelled_bias, compare_grads, cpu_state, grads_of, params_of, update_errs
        backend = "nccl" if ngpu >= world else "gloo"
        torch.cuda.set_device(rank if ngpu >= world else 0)
        parallel.init_from_env(backend=backend)
        torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
        s, cfg = build_solver("fp32", seed=1234 + rank)          # different weights per rank until attach() broadcasts
        parallel.attach(s)
        s.copy_nets()
        B = 2
        batch = O.synthetic_batch(B, 128, seed=40 + rank)        # this rank's shard
        b = {k: v.cuda() for k, v in batch.items()}
        orc = O.OracleSolver(cpu_state(s.gen), cpu_state(s.dis))
        p0 = {"dis": params_of(s.dis), "gen": params_of(s.gen)}
        eps = {}
        s.noise_hook = lambda tag: eps[tag].cuda()
        args = (b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, 0)

        def averaged(grads):
            keys = [k for k, g in grads.items() if g is not None]
            flat = torch.cat([grads[k].reshape(-1) for k in keys]).cuda()
            dist.all_reduce(flat)
            flat = (flat / world).cpu()
            out, o = {k: None for k in grads}, 0
            for k in keys:
                n = grads[k].numel()
                out[k] = flat[o:o + n].view_as(grads[k])
                o += n
            return out

        res = {}
        for phase in ("dis", "gen"):
            torch.manual_seed(100 * (rank + 1) + (0 if phase == "dis" else 1))
            if phase == "dis":
                eps["dis1"] = torch.randn(1, 8, B, 8)
                s.dis_update(*args)
            else:
                eps["gen1"], eps["gen2"] = torch.randn(1, 8, B, 8), torch.randn(1, 8, B, 8)
                s.gen_update(*args)
            net = s.dis if phase == "dis" else s.gen
            # the flat gradient buffer holds the SUM over ranks; the mean is folded into the fused Adam
            mine = {k: (g / world if g is not None else None) for k, g in grads_of(net).items()}
            # oracle replica of THIS rank on its own shard: gradients only, then the cross-rank average, then Adam
            if phase == "dis":
                D = orc._leaf(orc.D)
                with torch.no_grad():
                    x = batch["x_real"]
                    content, mus, _ = O.encode(orc.G, x)
                    sr = torch.cat(mus, 1)
                    st1 = O.gmm_sample(batch["c_trg"], eps["dis1"], 0.5)
                    mt, _ = O.text_encoder(orc.G, sr, batch["txt"], batch["txt_lens"])
                    f0, a0 = O.decode(orc.G, content, torch.cat(mt, 1))
                    f1, a1 = O.decode(orc.G, content, st1)
                    f0, f1 = O.blend(f0, a0, x, True), O.blend(f1, a1, x, True)
                loss = O.dis_loss(D, f0, x, batch["label_src"]) + O.dis_loss(D, f1, x, batch["label_src"])
                loss.backward()
                g_local = {k: v.grad for k, v in D.items()}
                state, params = orc.d_state, orc.D
            else:
                G = orc._leaf(orc.G)
                Ls = O.gen_phase_losses(G, orc.D, batch, eps["gen1"], eps["gen2"], True, orc.ds_w)
                Ls["loss_gen_total"].backward()
                loss = Ls["loss_gen_total"]
                g_local = {k: v.grad for k, v in G.items()}
                state, params = orc.g_state, orc.G
            mine_loss = float(s.loss_dis if phase == "dis" else s.loss_gen_total)
            assert abs(mine_loss - float(loss)) <= 1e-4 * abs(float(loss)), (phase, mine_loss, float(loss))
            g_avg = averaged(g_local)
            worst, wk, glob = compare_grads(mine, g_avg)
            assert glob < 1e-3 and worst < 1e-2, (phase, worst, wk, glob)
            O.adam_step(params, g_avg, state, orc.lr)
            errs = update_errs(p0[phase], params_of(net), params)
            wu = max(v[0] for k, v in errs.items() if not cancelled_bias(k))     # see tests/test_step_gpu.py
            assert wu < 0.5, (phase, wu)
            # every rank holds the same parameters after the step
            flat = net.flat.data.clone()
            ref = flat.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(flat, ref), phase
            res[phase] = (worst, glob, wu)
        # ---- a second step with the same attention status runs the BUCKETED schedule learned in the first one: early
        # buckets are reduced from inside backward, the rest at the end.  A range that missed its all-reduce would be
        # off by the other rank's gradient (~70 % per tensor); the trajectories themselves have separated by ~1e-2.
        sync = s._dp_sync
        for phase in ("dis", "gen"):
            c0 = sync.collectives
            torch.manual_seed(300 * (rank + 1) + (0 if phase == "dis" else 1))
            if phase == "dis":
                eps["dis1"] = torch.randn(1, 8, B, 8)
                s.dis_update(*args)
                D = orc._leaf(orc.D)
                with torch.no_grad():
                    x = batch["x_real"]
                    content, mus, _ = O.encode(orc.G, x)
                    st1 = O.gmm_sample(batch["c_trg"], eps["dis1"], 0.5)
                    mt, _ = O.text_encoder(orc.G, torch.cat(mus, 1), batch["txt"], batch["txt_lens"])
                    f0, a0 = O.decode(orc.G, content, torch.cat(mt, 1))
                    f1, a1 = O.decode(orc.G, content, st1)
                    f0, f1 = O.blend(f0, a0, x, True), O.blend(f1, a1, x, True)
                (O.dis_loss(D, f0, x, batch["label_src"]) + O.dis_loss(D, f1, x, batch["label_src"])).backward()
                g_local, state, params, net = {k: v.grad for k, v in D.items()}, orc.d_state, orc.D, s.dis
                want = 2 + 1                      # two early buckets + the pieces in front of, between and behind them as ONE collective
            else:
                eps["gen1"], eps["gen2"] = torch.randn(1, 8, B, 8), torch.randn(1, 8, B, 8)
                s.gen_update(*args)
                G = orc._leaf(orc.G)
                O.gen_phase_losses(G, orc.D, batch, eps["gen1"], eps["gen2"], True, orc.ds_w)["loss_gen_total"].backward()
                g_local, state, params, net = {k: v.grad for k, v in G.items()}, orc.g_state, orc.G, s.gen
                want = 1 + 1                      # the decoder bucket + (encoders in front, text encoder / MLP behind) coalesced
            assert sync.collectives - c0 == want, (phase, sync.collectives - c0, want)
            mine = {k: (g / world if g is not None else None) for k, g in grads_of(net).items()}
            g_avg = averaged(g_local)
            worst, wk, glob = compare_grads(mine, g_avg)
            assert glob < 5e-2 and worst < 0.3, ("bucketed", phase, worst, wk, glob)
            O.adam_step(params, g_avg, state, orc.lr)
            flat = net.flat.data.clone()
            ref = flat.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(flat, ref), phase
            res["bucketed_" + phase] = (worst, glob)
        q.put((rank, "ok", res, backend))
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc(), None))
        raise


def test_two_ranks_equal_two_oracle_replicas_with_averaged_gradients():
    world, port = 2, _free_port()
    ngpu = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ngpu, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(120)
    for rank, status, res, backend in sorted(out):
        assert status == "ok", res
        print("rank", rank, backend, res)
    assert all(p.exitcode == 0 for p in procs)
