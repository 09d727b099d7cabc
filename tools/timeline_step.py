"""Kernel timeline of one replayed training step through torch.profiler (CUPTI): in-situ kernel durations (warm
caches, real overlap between the main and the side stream), idle gaps, per-kernel totals.
  python tools/timeline_step.py [batch [config]] > profiles/xxx.txt        (config: train128 | train256 | infer64)"""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

conf = bench.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "train128"]
B = int(sys.argv[1]) if len(sys.argv) > 1 and int(sys.argv[1]) > 0 else conf["batch"]
from dwc_gan_b200 import parallel
rank, world, local = parallel.init_from_env()        # under torchrun: data-parallel step, rank 0 reports
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
s, cfg = bench.build_solver(dev, "bf16", conf["overrides"])
if world > 1:
    parallel.attach(s)
b = {k: v.to(dev) for k, v in bench.make_host_batch(B, conf["size"], 0).items()}
if conf["train"]:
    step = lambda it: bench.one_step(s, cfg, b, it)
else:
    s.eval()

    def step(it):
        with torch.no_grad():
            return s(b["x_real"], b["txt"], b["txt_lens"])
for it in range(6):
    step(it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(6)
    torch.cuda.synchronize()
if rank != 0:
    torch.cuda.synchronize()
    os._exit(0)
path = "/tmp/trace.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
t0, t1 = ks[0]["ts"], max(e["ts"] + e["dur"] for e in ks)
print("step span %.2f ms, %d GPU activities" % ((t1 - t0) / 1e3, len(ks)))
streams = collections.defaultdict(list)
for e in ks:
    streams[e["args"].get("stream")].append(e)
for sid, lst in streams.items():
    busy = sum(e["dur"] for e in lst)
    print("stream %s: %d activities, busy %.2f ms" % (sid, len(lst), busy / 1e3))
# union busy time
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ks)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for a, c in iv[1:]:
    if a > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = a, c
    else:
        cur_e = max(cur_e, c)
busy += cur_e - cur_s
print("GPU busy (union) %.2f ms, idle %.2f ms" % (busy / 1e3, (t1 - t0 - busy) / 1e3))
# gaps on the main stream
main = max(streams.values(), key=len)
gaps = [main[i + 1]["ts"] - (main[i]["ts"] + main[i]["dur"]) for i in range(len(main) - 1)]
gaps = [g for g in gaps if g > 0]
print("main stream: mean gap %.2f us, total gaps %.2f ms" % (sum(gaps) / max(1, len(gaps)), sum(gaps) / 1e3))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ks:
    nm = e["name"][:70]
    agg[nm][0] += 1
    agg[nm][1] += e["dur"]
tot = sum(v[1] for v in agg.values())
print("sum of activity durations %.2f ms" % (tot / 1e3))
for nm, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%6.2f%% %9.1f us %5d x %8.1f us  %s" % (100 * d / tot, d, c, d / c, nm))
if os.environ.get("DWC_TL_TAIL"):
    # the last N activities of the step: what the optimizer is waiting for
    for e in sorted(ks, key=lambda e: e["ts"] + e["dur"])[-int(os.environ["DWC_TL_TAIL"]):]:
        print("   tail %8.3f .. %8.3f ms  stream %-4s %7.1f us  %s" % ((e["ts"] - t0) / 1e3, (e["ts"] + e["dur"] - t0) / 1e3,
                                                                   e["args"].get("stream"), e["dur"], e["name"][:60]))
# Time during which nothing substantial runs (only kernels shorter than 10 us, or nothing at all): loss glue, launch chains
big = sorted((e["ts"], e["ts"] + e["dur"]) for e in ks if e["dur"] >= 10.0)
cov, cs, ce = 0.0, big[0][0], big[0][1]
holes = []
for a, c in big[1:]:
    if a > ce:
        cov += ce - cs
        holes.append((cs if False else ce, a))
        cs, ce = a, c
    else:
        ce = max(ce, c)
cov += ce - cs
print("time with no kernel >= 10 us in flight: %.2f ms of %.2f ms" % ((t1 - t0 - cov) / 1e3, (t1 - t0) / 1e3))
for a, c in sorted(holes, key=lambda h: h[0] - h[1])[:12]:
    inside = [e for e in ks if e["ts"] >= a - 0.5 and e["ts"] + e["dur"] <= c + 0.5]
    names = collections.Counter(e["name"][:40] for e in inside)
    print("   hole at %8.3f ms, %7.1f us, %3d small kernels: %s" % ((a - t0) / 1e3, c - a, len(inside),
                                                                ", ".join("%dx %s" % (v, k) for k, v in names.most_common(4))))
# Approximate critical path: walk back from the last activity; the predecessor of an activity is the one (any stream)
# that ended last before it started - the dependency that released it, or the previous kernel of its own stream.
ends = sorted(ks, key=lambda e: e["ts"] + e["dur"])
import bisect
end_ts = [e["ts"] + e["dur"] for e in ends]
cur = ends[-1]
path = []
while cur is not None:
    path.append(cur)
    i = bisect.bisect_right(end_ts, cur["ts"] + 0.5) - 1          # last activity that ended before `cur` started
    while i >= 0 and ends[i] is cur:
        i -= 1
    cur = ends[i] if i >= 0 else None
cp = collections.defaultdict(lambda: [0, 0.0])
busy_cp = 0.0
for e in path:
    cp[e["name"][:70]][0] += 1
    cp[e["name"][:70]][1] += e["dur"]
    busy_cp += e["dur"]
print("critical path (approximate): %d activities, %.2f ms of kernels + %.2f ms of gaps between them" % (
    len(path), busy_cp / 1e3, (t1 - t0 - busy_cp) / 1e3))
for nm, (c, d) in sorted(cp.items(), key=lambda kv: -kv[1][1])[:25]:
    print("   cp %6.2f%% %9.1f us %5d x %8.1f us  %s" % (100 * d / max(busy_cp, 1e-9), d, c, d / c, nm))
if world > 1:
    # where the collectives sit: start / end relative to the step, and what else runs at that time
    print("collectives (ms from the start of the step):")
    for e in ks:
        if "nccl" in e["name"].lower():
            a, c = e["ts"], e["ts"] + e["dur"]
            others = sum(min(c, o["ts"] + o["dur"]) - max(a, o["ts"]) for o in ks
                         if o is not e and o["ts"] < c and o["ts"] + o["dur"] > a)
            print("  %8.3f .. %8.3f  (%7.1f us)  other kernels in flight meanwhile: %7.1f us  %s" % (
                (a - t0) / 1e3, (c - t0) / 1e3, e["dur"], others, e["name"][:60]))
    os._exit(0)
