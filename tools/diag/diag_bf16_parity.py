"""Where does the bf16 CUDA path differ from the bf16-storage-rounding oracle?  Sub-network forward/backward with
fixed upstream gradients, CUDA (bf16) vs oracle(fp32) vs oracle(bf16 storage)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import dwc_oracle as O
from tests.util_gpu import build_solver, compare_grads, cpu_state, grads_of, rel, top_errs

def leaf(P):
    return {k: v.clone().requires_grad_(True) for k, v in P.items()}

B = int(os.environ.get("B", "2"))
s, _ = build_solver("bf16")
state = O.trainable(cpu_state(s.gen))
batch = O.synthetic_batch(B, 128, seed=3)
x = batch["x_real"]
torch.manual_seed(0)

# ---------------- encode
res = {}
wc = wm = wl = None
for mode in ("fp32", "bf16"):
    G = leaf(state)
    xl = x.clone().requires_grad_(True)
    with O.storage_rounding(mode):
        c_ref, mus, lvs = O.encode(G, xl)
        mu_ref, lv_ref = torch.cat(mus, 1), torch.cat(lvs, 1)
        if wc is None:
            wc, wm, wl = torch.randn_like(c_ref), torch.randn_like(mu_ref), torch.randn_like(lv_ref)
        ((c_ref * wc).sum() / 100 + (mu_ref * wm).sum() + (lv_ref * wl).sum()).backward()
    res[mode] = (c_ref.detach(), mu_ref.detach(), {k: v.grad for k, v in G.items()}, xl.grad)
s.gen_opt.zero_grad()
xc = x.cuda().requires_grad_(True)
content, mu_l, lv_l = s.gen.encode(xc)
mu, lv = torch.cat(mu_l, 1), torch.cat(lv_l, 1)
((content.float() * wc.cuda()).sum() / 100 + (mu * wm.cuda()).sum() + (lv * wl.cuda()).sum()).backward()
mine = grads_of(s.gen)
for mode in ("fp32", "bf16"):
    c_ref, mu_ref, gr, xg = res[mode]
    keys = {k: g for k, g in gr.items() if g is not None}
    w, wk, gl = compare_grads(mine, keys)
    print("ENCODE vs oracle[%s]: content %.3e mu %.3e | grads worst %.3e (%s) global %.3e | dx %.3e" % (
        mode, rel(content.float(), c_ref), rel(mu, mu_ref), w, wk, gl, rel(xc.grad, xg)))
    print("    ", top_errs(mine, keys))
w, wk, gl = compare_grads(res["bf16"][2], {k: g for k, g in res["fp32"][2].items() if g is not None})
print("ENCODE oracle[bf16] vs oracle[fp32]: grads worst %.3e (%s) global %.3e" % (w, wk, gl))

# ---------------- decode
torch.manual_seed(5)
cont = torch.randn(B, 256, 32, 32).to(torch.bfloat16).float()
style = torch.randn(B, 64)
res = {}
wi = wa = None
for mode in ("fp32", "bf16"):
    G = leaf(state)
    cl, sl = cont.clone().requires_grad_(True), style.clone().requires_grad_(True)
    with O.storage_rounding(mode):
        img, att = O.decode(G, cl, sl)
        if wi is None:
            wi, wa = torch.randn_like(img), torch.randn_like(att)
        ((img * wi).sum() + (att * wa).sum()).backward()
    res[mode] = (img.detach(), att.detach(), {k: v.grad for k, v in G.items()}, cl.grad, sl.grad)
s.gen_opt.zero_grad()
cc = cont.cuda().to(torch.bfloat16).requires_grad_(True)
sc = style.cuda().requires_grad_(True)
img, att = s.gen.decode(cc, sc)
((img * wi.cuda()).sum() + (att * wa.cuda()).sum()).backward()
mine = grads_of(s.gen)
for mode in ("fp32", "bf16"):
    i_ref, a_ref, gr, cg, sg = res[mode]
    keys = {k: g for k, g in gr.items() if g is not None}
    w, wk, gl = compare_grads(mine, keys)
    print("DECODE vs oracle[%s]: img %.3e att %.3e | grads worst %.3e (%s) global %.3e | dcontent %.3e dstyle %.3e" % (
        mode, rel(img, i_ref), rel(att, a_ref), w, wk, gl, rel(cc.grad.float(), cg), rel(sc.grad, sg)))
    print("    ", top_errs(mine, keys))
w, wk, gl = compare_grads(res["bf16"][2], {k: g for k, g in res["fp32"][2].items() if g is not None})
print("DECODE oracle[bf16] vs oracle[fp32]: grads worst %.3e (%s) global %.3e dcontent %.3e" % (
    w, wk, gl, rel(res["bf16"][3], res["fp32"][3])))
