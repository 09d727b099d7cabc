#!/usr/bin/env python
"""Headline benchmark: full G+D training step of DWC-GAN at 128x128, bf16, batch 16 per GPU (BASELINE.json
configs[2]), data-parallel over N GPUs of one node.

  python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # reference's CPU path (oracle port) on the host cores

One JSON line on stdout (rank 0).  A "step" = dis_update + gen_update + smooth_moving + update_learning_rate +
update_attention_status (train.py:102-111) on one synthetic batch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

GFLOP_PER_IMAGE_STEP = 569.7          # required algorithmic conv+linear work of one G+D step
METRIC = "G+D train images/sec @128^2"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops", 1590.0), tflops_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), src="measured")
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def make_host_batch(B, size, seed):
    from oracle import dwc_oracle as O          # synthetic input generator only (no arithmetic of the hot path)
    return O.synthetic_batch(B, size, seed=seed)


def build_solver(device, mode="bf16"):
    import dwc_gan_b200
    from dwc_gan_b200.solver import Solver
    from dwc_gan_b200.utils import get_config
    dwc_gan_b200.set_mode(mode)
    cfg = get_config(os.path.join(ROOT, "tests", "golden", "celeba_faces.yaml"))
    cfg["vgg_w"] = 0
    torch.manual_seed(1234)                      # train.py:23
    s = Solver(cfg, device, None).to(device)
    s.copy_nets()
    return s, cfg


def one_step(s, cfg, b, it):
    s.dis_update(b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, it)
    s.gen_update(b["x_real"], b["c_src"], b["c_trg"], b["txt"], b["txt_lens"], b["label_src"], b["label_trg"], cfg, it)
    s.smooth_moving()
    s.update_learning_rate()
    s.update_attention_status(it)


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at its in-step shape, from the
# `ncu --set full` capture summarised in profiles/r01g_ncu_g7_b48.md (None until measured for another batch)
# n = 48: 29.71 MB read + 0.08 MB written (input 28.4 MB + weights 1.2 MB read exactly once; the 31.8 MB output was
# still in the 126 MB L2 when the capture ended)
G7_TRAFFIC_BYTES = {48: 29794560}


def conv_roofline(peaks, B):
    """Dominant kernel of the step: the tcgen05 implicit-GEMM conv at the G7 geometry (3x3, 256->256 @32x32: 16 of
    the 27 generator convs) at the batch it runs with inside the step - the three gradient-carrying decodes / the three
    re-encodes of gen_update are one 3B batch, which the persistent kernel variant serves (64 launches x ~50 us per
    step against 67 x ~25 us for the single-batch launches of the same geometry).  Timed alone with CUDA events,
    rotating over enough buffers to exceed the 126 MB L2."""
    from dwc_gan_b200 import _lib as L, plan as P
    from dwc_gan_b200.plan import HB
    n, c, hw = 3 * B, 256, 32
    nbuf = 8
    xs = [HB(torch.randn(n, hw + 2, hw + 2, c, device="cuda").to(torch.bfloat16), n, hw, hw, c, 1, 0) for _ in range(nbuf)]
    ys = [HB.empty(n, hw, hw, c, 2, 0, torch.bfloat16, "cuda") for _ in range(nbuf)]
    w = (torch.randn(c, 9 * c, device="cuda") * 0.02).to(torch.bfloat16)
    bias = torch.zeros(c, device="cuda")
    plans = [P.plan_conv_fwd(xs[i], w, c, c, bias, ys[i], 3, 1, L.TC) for i in range(nbuf)]
    for p in plans[:4]:
        p.launch()
    torch.cuda.synchronize()
    reps = 6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for p in plans:
            p.launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * nbuf)
    flops = 2.0 * n * hw * hw * c * 9 * c
    achieved = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "gconv_tcp_kernel<256> (G7 3x3 256->256 @32x32, 3 x batch %d = %d images)" % (B, n),
            "achieved": round(achieved, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
            "frac": round(achieved / peaks["tflops"], 4), "traffic": G7_TRAFFIC_BYTES.get(n),
            "algorithmic_flops": flops, "algorithmic_bytes": int(xs[0].t.numel() * 2 + ys[0].t.numel() * 2 + w.numel() * 2),
            "peak_source": peaks["src"] + " burst", "us_per_launch": round(ms * 1e3, 2)}


def cpu_baseline(sample_b=2):
    """Oracle port (the reference's algorithm on torch CPU fp32) timed on this box's host cores: 1 G+D step."""
    from oracle import dwc_oracle as O
    import dwc_gan_b200  # noqa: F401
    from dwc_gan_b200.solver import Solver
    from dwc_gan_b200.utils import get_config
    cfg = get_config(os.path.join(ROOT, "tests", "golden", "celeba_faces.yaml"))
    cfg["vgg_w"] = 0
    torch.manual_seed(1234)
    s = Solver(cfg, torch.device("cpu"), None)            # parameter container only (init parity with the reference)
    orc = O.OracleSolver({k: v.contiguous() for k, v in s.gen.state_dict().items()},
                         {k: v.contiguous() for k, v in s.dis.state_dict().items()})
    batch = O.synthetic_batch(sample_b, 128, seed=0)
    return orc, batch


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import dwc_oracle as O
    torch.set_num_threads(os.cpu_count())
    sample_b = 2 if args.steps <= 12 else 1
    orc, batch = cpu_baseline(sample_b)
    B = sample_b

    def step(it):
        torch.manual_seed(it)
        orc.dis_update(batch, torch.randn(1, 8, B, 8))
        orc.gen_update(batch, torch.randn(1, 8, B, 8), torch.randn(1, 8, B, 8))
        orc.smooth_moving()
        orc.update_attention_status(it)
    for it in range(args.warmup):
        step(it)
    t0 = time.perf_counter()
    for it in range(args.steps):
        step(args.warmup + it)
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "full G+D training step, 128x128, configs/celeba_faces.yaml, vgg_w=0",
                       "per_gpu_batch": 16, "sample_batch": B},
            "cpu_baseline": {"value": round(v, 4), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d G+D steps of batch %d (reference algorithm, oracle port, torch CPU fp32)" % (
                                 args.steps, B)},
            "e2e": {"value": round(v, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    from dwc_gan_b200 import parallel
    import dwc_gan_b200
    import torch.distributed as dist
    rank, world, local = parallel.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    s, cfg = build_solver(dev, args.mode)
    if world > 1:
        parallel.attach(s)
    B = args.batch
    host = make_host_batch(B, 128, seed=rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    it = 0
    for _ in range(args.warmup):
        one_step(s, cfg, resident, it)
        it += 1
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = dwc_gan_b200.RT.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step(s, cfg, resident, it)
        it += 1
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = dwc_gan_b200.RT.launches - l0
    sampler.stop_flag = True
    # ---- end to end: pinned host batch -> device every step, losses read back every step
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    d2h = 0
    for _ in range(args.steps):
        b = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        one_step(s, cfg, b, it)
        it += 1
        losses = torch.stack([s.loss_gen_total.detach().float(), s.loss_dis_all.detach().float()]).cpu()
        d2h = losses.numel() * 4
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    assert torch.isfinite(losses).all(), losses
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        _leave(world)
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    sampler.join(timeout=2)
    roof = conv_roofline(peaks, B)
    roof["step_tensor_frac_sustained"] = round(value / world * GFLOP_PER_IMAGE_STEP * 1e9 / (peaks["tflops_sustained"] * 1e12), 4)
    line = {"metric": METRIC, "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.mode.startswith("bf16") else "f32",
            "data": "synthetic",
            "config": {"workload": "full G+D training step, 128x128, configs/celeba_faces.yaml, vgg_w=0 (BASELINE configs[2])",
                       "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                       "l2": "inputs larger than L2: one step streams > 2 GB of activations per GPU",
                       "mode": args.mode},
            "e2e": {"value": round(e2e, 2), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof}
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        orc, cb = cpu_baseline(2)
        t0 = time.perf_counter()
        orc.dis_update(cb, torch.randn(1, 8, 2, 8))
        orc.gen_update(cb, torch.randn(1, 8, 2, 8), torch.randn(1, 8, 2, 8))
        orc.smooth_moving()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": round(2 / dt, 4), "unit": "images/s", "cores": torch.get_num_threads(),
                                "kind": "port", "sample": "1 G+D step of batch 2 (oracle port, torch CPU fp32), %.1f s" % dt}
    print(json.dumps(line), flush=True)
    _leave(world)


def _leave(world):
    """End the process without tearing NCCL down: captured CUDA graphs still reference the communicator, and a
    destroy_process_group() that one rank reaches long before the other (rank 0 goes on to the roofline and CPU
    legs) has been seen to block at exit."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
