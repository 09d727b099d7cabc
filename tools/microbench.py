"""Per-layer microbenchmark (BASELINE.json configs[4]): every conv geometry of SURVEY 8(a)-2 x {fprop, dgrad,
wgrad} on the tcgen05 kernels, and the norm / post passes, each against its own roofline bound
min(tensor peak, arithmetic intensity x HBM bandwidth).  Timed with CUDA events on the launching stream over a
rotation of buffers larger than the 126 MB L2.

  python tools/microbench.py [--batch 16] [--out profiles/xxx.md]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dwc_gan_b200 import _lib as L  # noqa: E402
from dwc_gan_b200 import plan as P  # noqa: E402
from dwc_gan_b200.plan import HB  # noqa: E402

# id, cin, cout, k, stride, pad, input H=W (128x128 network), how many times the layer runs forward per G+D step
GEOMS = [
    ("G2", 64, 128, 4, 2, 1, 128), ("G3", 128, 256, 4, 2, 1, 64), ("G4", 256, 256, 4, 2, 1, 32),
    ("G5", 256, 256, 4, 2, 1, 16), ("G6", 256, 256, 4, 2, 1, 8), ("G7", 256, 256, 3, 1, 1, 32),
    ("G8", 256, 128, 5, 1, 2, 64), ("G9", 128, 64, 5, 1, 2, 128),
    ("D2", 64, 128, 4, 2, 1, 64), ("D3", 128, 256, 4, 2, 1, 32), ("D4", 256, 512, 4, 2, 1, 16),
    ("D5", 512, 512, 4, 2, 1, 8), ("D2s", 64, 128, 4, 2, 1, 32), ("D3s", 128, 256, 4, 2, 1, 16),
    ("D4s", 256, 512, 4, 2, 1, 8), ("D5s", 512, 512, 4, 2, 1, 4),
    ("G10", 64, 4, 7, 1, 3, 128),       # decoder heads (3+1 channels): forward only (backward uses window buffers)
    ("G1d", 3, 64, 7, 1, 3, 128),       # first encoder conv: data gradient to the generated image only
]

_ws = {}


def workspace(nbytes):
    t = _ws.get("t")
    if t is None or t.numel() * 4 < nbytes:
        t = torch.empty((nbytes + 3) // 4 + 1024, dtype=torch.float32, device="cuda")
        _ws["t"] = t
    return t


def timeit(fns, reps=3):
    """Seconds per call: the calls are captured once into a CUDA graph and the replay is timed, so that the figure is
    device time (the product path replays graphs too) and not Python / ctypes launch overhead."""
    for f in fns[:2]:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (reps * len(fns))


def pack(w_krsc, mode, rows, cout, k, cin):
    if mode == 0:
        out = torch.empty(rows, k * k * cin, dtype=torch.bfloat16, device="cuda")
    elif mode == 1:
        out = torch.empty(rows, k * k * cout, dtype=torch.bfloat16, device="cuda")
    else:
        out = torch.empty(4, rows, 4 * cout, dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().dwc_pack_weights(L.ptr(w_krsc), cout, k, k, cin, mode, L.ptr(out), L.BF16, rows, L.stream()), "pack")
    return out


def bench_geom(g, n, peaks):
    name, cin, cout, k, s, p, hw = g
    ho = (hw + 2 * p - k) // s + 1
    layout = 0 if s == 1 else 1
    hy = k - 1 if s == 1 else 1
    bt = torch.bfloat16
    in_bytes = n * (hw + 2 * p) ** 2 * cin * 2
    out_bytes = n * (ho + 2 * hy) ** 2 * cout * 2
    nbuf = max(2, min(24, int(300e6 // max(1, in_bytes + out_bytes)) + 1))
    xs = [HB(torch.randn(HB.shape_of(n, hw, hw, cin, p, layout), device="cuda").to(bt), n, hw, hw, cin, p, layout)
          for _ in range(nbuf)]
    ys = [HB(torch.randn(HB.shape_of(n, ho, ho, cout, hy, 0), device="cuda").to(bt), n, ho, ho, cout, hy, 0)
          for _ in range(nbuf)]
    w = (torch.randn(cout, k, k, cin, device="cuda") * 0.02)
    bias = torch.zeros(cout, device="cuda")
    rows_f = cout if cout % 64 == 0 else 16
    rows_d = cin if cin % 64 == 0 else 16
    wf = pack(w, 0, rows_f, cout, k, cin)
    wd = pack(w, 1 if s == 1 else 2, rows_d, cout, k, cin)
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    db = torch.zeros(cout, device="cuda")
    flops = 2.0 * n * ho * ho * cout * k * k * cin
    res = {}
    if cin % 64 == 0:
        fw = [P.plan_conv_fwd(xs[i], wf, cout, rows_f, bias, ys[i], k, s, L.TC) for i in range(nbuf)]
        res["fprop"] = timeit([pl.launch for pl in fw])
    if cout % 64 == 0:
        dg = [P.plan_conv_dgrad(ys[i], wd, xs[i], k, s, L.TC, cin_padded=rows_d) for i in range(nbuf)]
        res["dgrad"] = timeit([(lambda ps=ps: [q.launch() for q in ps]) for ps in dg])
    if cin % 64 == 0 and cout % 64 == 0:
        wg = [P.plan_conv_wgrad(ys[i], xs[i], dw, db, k, s, L.TC) for i in range(nbuf)]
        res["wgrad"] = timeit([(lambda pl=pl: pl.launch(workspace)) for pl in wg])
    w_bytes = cout * k * k * cin * 2
    ai = flops / (in_bytes + out_bytes + w_bytes)
    bound = min(peaks["tflops"] * 1e12, ai * peaks["hbm"] * 1e9)
    rows = []
    for op in ("fprop", "dgrad", "wgrad"):
        if op not in res:
            continue
        t = res[op]
        rows.append(dict(layer=name, op=op, n=n, cin=cin, cout=cout, k=k, stride=s, hw_in=hw, us=round(t * 1e6, 1),
                         tflops=round(flops / t / 1e12, 1), bound_tflops=round(bound / 1e12, 1),
                         frac_of_bound=round(flops / t / bound, 3), frac_of_tensor_peak=round(flops / t / 1e12 / peaks["tflops"], 3)))
    return rows


# first layers (3 input channels) through the 8-pixel x 8-channel row-im2col buffers: id, cout, k, stride, pad, H=W, pool
FIRST = [("G1", 64, 7, 1, 3, 128, 1), ("D1", 64, 4, 2, 1, 128, 1), ("D1s", 64, 4, 2, 1, 128, 2)]


def bench_first(g, n, peaks):
    """3 -> 64 first convolution: forward and weight gradient on the tensor cores over the row-im2col operand, plus the
    kernel that builds that operand from the fp32 NCHW image (HBM-bound: the layer's bound is AI x HBM)."""
    from dwc_gan_b200 import ops
    import ctypes as C
    name, cout, k, s, p, hw, pool = g
    bt = torch.bfloat16
    hi = hw // pool
    hp = hi + 2 * p
    ho = (hp - k) // s + 1
    hy = k - 1 if s == 1 else 1
    nbuf = 6
    imgs = [torch.rand(n, 3, hw, hw, device="cuda") * 2 - 1 for _ in range(nbuf)]
    rows = [torch.empty(P.rows_shape(n, hp, ho, s), dtype=bt, device="cuda") for _ in range(nbuf)]
    ys = [HB(torch.randn(HB.shape_of(n, ho, ho, cout, hy, 0), device="cuda").to(bt), n, ho, ho, cout, hy, 0) for _ in range(nbuf)]
    w = torch.randn(cout, k, k, 3, device="cuda") * 0.1
    wr = torch.empty(cout, k * 64, dtype=bt, device="cuda")
    L.check(L.lib().dwc_pack_weights(L.ptr(w), cout, k, k, 3, 3, L.ptr(wr), L.BF16, cout, L.stream()), "pack")
    bias = torch.zeros(cout, device="cuda")
    dw, db = torch.zeros(cout, k, k, 3, device="cuda"), torch.zeros(cout, device="cuda")

    def f_rows(i):
        L.check(L.lib().dwc_image_rows_fwd(L.ptr(imgs[i]), n, 3, hw, hw, pool, p, s, s, ho, L.ptr(rows[i]), L.BF16, L.stream()))
    for i in range(nbuf):
        f_rows(i)
    fw = [P.plan_first_conv_fwd(rows[i], n, hp, ho, ho, k, s, wr, cout, bias, ys[i], L.TC) for i in range(nbuf)]
    wg = [P.plan_first_conv_wgrad(ys[i], rows[i], n, hp, ho, k, s, 3, dw, db, L.TC) for i in range(nbuf)]
    res = {"rows (image -> im2col operand)": timeit([(lambda i=i: f_rows(i)) for i in range(nbuf)]),
           "fprop": timeit([pl.launch for pl in fw]),
           "wgrad": timeit([(lambda pl=pl: pl.launch(workspace)) for pl in wg])}
    flops = 2.0 * n * ho * ho * cout * k * k * 3
    in_bytes, out_bytes = n * 3 * hw * hw * 4, n * ho * ho * cout * 2
    bound = min(peaks["tflops"] * 1e12, flops / (in_bytes + out_bytes) * peaks["hbm"] * 1e9)
    out = []
    for op, t in res.items():
        out.append("| %s | %s | 3->%d k%d/s%d @%d | %.1f | %.1f | %.1f | %.3f | %.3f |" % (
            name, op, cout, k, s, hi, t * 1e6, flops / t / 1e12, bound / 1e12, flops / t / bound, flops / t / 1e12 / peaks["tflops"]))
    return out


def bench_post(n, c, hw, kind, peaks, halo_out=1, halo_dy=2):
    """norm site through the C ABI: forward = stats + finalize + fused normalise/act/reflect-pad pass; backward =
    reduce + finalize + apply.  Each kernel is timed on its own (graph replay over rotating buffers)."""
    import ctypes as C
    from dwc_gan_b200 import ops
    lib = L.lib()
    bt = torch.bfloat16
    st = L.stream
    nbuf = max(2, min(16, int(400e6 // (n * hw * hw * c * 8)) + 1))
    E = n * hw * hw * c
    splits = ops._stats_splits(hw * hw, n)
    ys = [HB(torch.randn(n, hw, hw, c, device="cuda").to(bt), n, hw, hw, c, 0, 0) for _ in range(nbuf)]
    outs = [HB.empty(n, hw, hw, c, halo_out, 0, bt, "cuda") for _ in range(nbuf)]
    douts = [HB(torch.randn(HB.shape_of(n, hw, hw, c, halo_out, 0), device="cuda").to(bt), n, hw, hw, c, halo_out, 0)
             for _ in range(nbuf)]
    dys = [HB.empty(n, hw, hw, c, halo_dy, 0, bt, "cuda") for _ in range(nbuf)]
    stats = torch.empty(n * splits * c * 2, device="cuda")
    red = torch.empty(n * splits * c * 2, device="cuda")
    coef = torch.empty(n * c * 4, device="cuda")
    bco = torch.empty(n * c * 4, device="cuda")
    nw = torch.rand(n, c, device="cuda") if kind == ops.NORM_ADAIN else torch.rand(c, device="cuda")
    nb = torch.rand(n, c, device="cuda") if kind == ops.NORM_ADAIN else torch.rand(c, device="cuda")
    gw, gb = torch.zeros(n * c, device="cuda"), torch.zeros(n * c, device="cuda")
    S = lambda hb: C.byref(hb.struct())

    def f_stats(i):
        L.check(lib.dwc_nc_stats(S(ys[i]), splits, L.ptr(stats), st()))

    def f_fin(i):
        L.check(lib.dwc_norm_finalize(kind, L.ptr(stats), splits, n, c, hw * hw, L.f32(1e-5), L.ptr(nw), L.ptr(nb),
                                      L.ptr(coef), st()))

    def f_post(i):
        L.check(lib.dwc_post_fwd(S(ys[i]), L.ptr(coef), 1, None, S(outs[i]), st()))

    def b_fold(i):
        L.check(lib.dwc_fold_halo(S(douts[i]), st()))

    def b_red(i):
        L.check(lib.dwc_post_bwd_reduce(S(douts[i]), S(ys[i]), L.ptr(coef), 1, splits, L.ptr(red), 1, st()))

    def b_fin(i):
        L.check(lib.dwc_norm_bwd_finalize(kind, L.ptr(red), splits, L.ptr(coef), n, c, hw * hw, L.f32(1e-5), L.ptr(nw),
                                          L.ptr(gw), L.ptr(gb), L.ptr(bco), st()))

    def b_app(i):
        L.check(lib.dwc_post_bwd_apply(S(douts[i]), S(ys[i]), L.ptr(coef), L.ptr(bco), 1, S(dys[i]), None, 1, st()))
    f_stats(0); f_fin(0); b_red(0); b_fin(0)
    dres = [HB.empty(n, hw, hw, c, halo_out, 0, bt, "cuda") for _ in range(nbuf)]
    cluster_ok = kind in (1, 2) and hasattr(lib, "dwc_post_bwd_cluster") and \
        bool(lib.dwc_post_bwd_cluster_ok(S(douts[0]), S(ys[0]), kind, S(dys[0]), None))       # experimental builds only

    def b_cluster(i):
        L.check(lib.dwc_post_bwd_cluster(S(douts[i]), S(ys[i]), L.ptr(coef), kind, 1, L.ptr(nw), L.ptr(gw), L.ptr(gb),
                                         S(dys[i]), None, st()))
    rows = []
    name = "%s %dx%dx%d" % ({1: "IN", 2: "AdaIN", 3: "LN"}[kind], c, hw, hw)
    tot = {"fwd": 0.0, "bwd": 0.0}
    for op, fn, byts, grp in (("stats", f_stats, E * 2, "fwd"), ("finalize", f_fin, 0, "fwd"),
                              ("post_fwd", f_post, 2 * E * 2, "fwd"), ("bwd_fold_halo", b_fold, 0, "bwd"),
                              ("bwd_reduce", b_red, 2 * E * 2, "bwd"),
                              ("bwd_finalize", b_fin, 0, "bwd"), ("bwd_apply", b_app, 3 * E * 2, "bwd")):
        t = timeit([(lambda i=i: fn(i)) for i in range(nbuf)])
        tot[grp] += t
        rows.append(dict(layer=name, op=op, n=n, us=round(t * 1e6, 1), gbs=round(byts / t / 1e9, 1),
                         frac_of_hbm=round(byts / t / 1e9 / peaks["hbm"], 3)))
    if cluster_ok:
        t = timeit([(lambda i=i: b_cluster(i)) for i in range(nbuf)])
        rows.append(dict(layer=name, op="SITE bwd, one-pass cluster kernel (parked experiment)", n=n,
                         us=round(t * 1e6, 1), gbs=round(3 * E * 2 / t / 1e9, 1),
                         frac_of_hbm=round(3 * E * 2 / t / 1e9 / peaks["hbm"], 3)))
    # product path since round 2: the statistics come out of the producing convolution's epilogue (dwc_gconv_t.stats) and
    # the coefficients are computed inside the apply kernel, so a forward site is the post_fwd pass alone
    t_fused = [r for r in rows if r["op"] == "post_fwd"][0]["us"] * 1e-6
    rows.append(dict(layer=name, op="SITE fwd, statistics from the conv epilogue (product path)", n=n,
                     us=round(t_fused * 1e6, 1), gbs=round(2 * E * 2 / t_fused / 1e9, 1),
                     frac_of_hbm=round(2 * E * 2 / t_fused / 1e9 / peaks["hbm"], 3)))
    for grp, byts in (("fwd", 2 * E * 2), ("bwd", 3 * E * 2)):
        t = tot[grp]
        rows.append(dict(layer=name, op="SITE " + grp + (" with a separate statistics pass" if grp == "fwd" else "") +
                         " (algorithmic bytes)", n=n, us=round(t * 1e6, 1),
                         gbs=round(byts / t / 1e9, 1), frac_of_hbm=round(byts / t / 1e9 / peaks["hbm"], 3)))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    d = json.load(open(pk)) if os.path.exists(pk) else {}
    peaks = dict(tflops=d.get("bf16_tflops", 1590.0), hbm=d.get("hbm_gbs", 6650.0))
    from dwc_gan_b200 import ops
    ops.RT.set_mode("bf16")
    lines = ["# conv microbench, batch %d, bf16 tcgen05 kernels (peaks: %.0f TFLOP/s, %.0f GB/s %s)" % (
        args.batch, peaks["tflops"], peaks["hbm"], "measured" if d else "fallback"),
        "| layer | op | Cin->Cout k/s @in | us | TFLOP/s | bound TFLOP/s | frac of bound | frac of tensor peak |", "|---|---|---|---|---|---|---|---|"]
    for g in GEOMS:
        if args.only and g[0] not in args.only.split(","):
            continue
        for r in bench_geom(g, args.batch, peaks):
            lines.append("| %s | %s | %d->%d k%d/s%d @%d | %.1f | %.1f | %.1f | %.3f | %.3f |" % (
                r["layer"], r["op"], r["cin"], r["cout"], r["k"], r["stride"], r["hw_in"], r["us"], r["tflops"],
                r["bound_tflops"], r["frac_of_bound"], r["frac_of_tensor_peak"]))
            print(lines[-1], flush=True)
    for g in FIRST:
        if args.only and g[0] not in args.only.split(","):
            continue
        for ln in bench_first(g, args.batch, peaks):
            lines.append(ln)
            print(ln, flush=True)
    lines += ["", "| norm site | op | us | GB/s (algorithmic) | frac of HBM peak |", "|---|---|---|---|---|"]
    if not args.only or args.only == "NONE":
        for (c, hw, kind) in ((64, 128, 1), (128, 64, 1), (256, 32, 1), (256, 32, 2), (128, 64, 3), (64, 128, 3)):
            for r in bench_post(args.batch, c, hw, kind, peaks):
                lines.append("| %s | %s | %.1f | %.1f | %.3f |" % (r["layer"], r["op"], r["us"], r["gbs"], r["frac_of_hbm"]))
                print(lines[-1], flush=True)
    if not args.only:
        # the dedicated few-channel 7x7 kernel and the upsample passes (same measurements as tools/bench_conv7.py and
        # tools/bench_upsample.py, which print more detail)
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        for tool in ("bench_conv7.py", "bench_upsample.py"):
            r = subprocess.run([sys.executable, os.path.join(here, tool)], capture_output=True, text=True)
            for ln in r.stdout.splitlines():
                if ln.startswith("|"):
                    lines.append(ln)
                    print(ln, flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
