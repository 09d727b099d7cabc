"""CPU oracle for the DWC-GAN generator+discriminator training step.

TEST INFRASTRUCTURE ONLY.  This file is a plain PyTorch (fp32, CPU) functional
restatement of the reference's hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product package (``dwc_gan_b200``) never
does.

Parity pinning: the reference ships no golden vectors, so this
restatement is pinned against the reference itself, imported from
``/root/reference`` in the build container by ``tests/golden/make_golden.py``;
the resulting losses / gradient norms / parameter checksums are committed under
``tests/golden/`` and re-checked on every run by ``tests/test_oracle_cpu.py``.

Everything here is written as pure functions over a flat ``{state_dict key:
tensor}`` parameter dictionary (same keys/shapes as the reference checkpoints,
SURVEY.md 8b).  Reference lines each function follows are cited inline
(paths relative to the reference repo root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# --------------------------------------------------------------------------
# configuration constants (configs/celeba_faces.yaml)
# --------------------------------------------------------------------------

DEFAULT_CFG = dict(
    image_size=128, input_dim=3, c_dim=8, num_cls=8, stddev=0.5,
    gen_dim=64, mlp_dim=256, n_res=4, style_downsample=5, content_downsample=2,
    hidden_size=300, num_layers=2,
    dis_dim=64, dis_n_layer=5, dis_num_scales=2,
    gan_w=1.0, cls_w=1.0, ds_w=1.0, kl_w=0.1, recon_x_w=10.0, recon_s_w=1.0,
    recon_c_w=1.0, recon_x_cyc_w=10.0,
    lr=1e-4, beta1=0.5, beta2=0.999, weight_decay=1e-4, adam_eps=1e-8,
    ema_beta=0.999,
)


# --------------------------------------------------------------------------
# storage-rounding mode (bf16 product-mode emulation)
# --------------------------------------------------------------------------
# The CUDA path keeps activations and activation gradients in bf16 HBM buffers and
# feeds the tensor cores bf16 weights (fp32 accumulation, fp32 statistics, fp32
# master weights / weight gradients / dense layers).  ``storage_rounding("bf16")``
# makes this oracle round to bf16 at exactly those storage points - forward values
# AND the gradients flowing back through them - while all arithmetic stays fp32.
# It answers "how far from the fp32 reference can a bf16-storage implementation be
# before anything is wrong with it": tests compare the CUDA path against both.

class _RoundBf16(torch.autograd.Function):
    """y = bf16(x) in forward, dx = bf16(dy) in backward (an activation buffer and its gradient buffer)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundBf16Fwd(torch.autograd.Function):
    """bf16 copy of an fp32 master weight: rounded operand, fp32 gradient."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


_STORAGE = {"mode": "fp32"}


class storage_rounding:
    """Context manager: ``with storage_rounding("bf16"): ...`` (default "fp32" = no rounding)."""

    def __init__(self, mode: str):
        assert mode in ("fp32", "bf16"), mode
        self.mode = mode

    def __enter__(self):
        self.prev = _STORAGE["mode"]
        _STORAGE["mode"] = self.mode
        return self

    def __exit__(self, *a):
        _STORAGE["mode"] = self.prev


def _q(x: Tensor) -> Tensor:
    """An activation the CUDA path stores in HBM (and whose gradient it stores too)."""
    return _RoundBf16.apply(x) if _STORAGE["mode"] == "bf16" else x


def _qw(w: Tensor) -> Tensor:
    """A convolution weight as the tensor cores see it."""
    return _RoundBf16Fwd.apply(w) if _STORAGE["mode"] == "bf16" else w


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------

def _act(x: Tensor, kind: str) -> Tensor:
    # networks/networks.py:556-571 (LeakyReLU slope is 0.1 in Conv2dBlock)
    if kind == "relu":
        return torch.relu(x)
    if kind == "lrelu":
        return F.leaky_relu(x, 0.1)
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "sigmoid":
        return torch.sigmoid(x)
    assert kind == "none", kind
    return x


def conv_reflect(P: Params, key: str, x: Tensor, k: int, s: int, p: int) -> Tensor:
    """reflect-pad + conv + bias: networks/networks.py:531,577-580."""
    if p > 0:
        x = F.pad(x, (p, p, p, p), mode="reflect")
    return _q(F.conv2d(x, _qw(P[key + ".weight"]), P[key + ".bias"], stride=s))


def inst_norm(x: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.InstanceNorm2d(affine=False): networks/networks.py:545 (biased variance)."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def adain(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    """AdaptiveInstanceNorm2d.forward: networks/networks.py:706-719.

    ``weight``/``bias`` are [B, C] (the reference flattens them to [B*C] and runs
    batch_norm in training mode on a (1, B*C, H, W) view: same thing)."""
    b, c = x.shape[:2]
    return inst_norm(x, eps) * weight.view(b, c, 1, 1) + bias.view(b, c, 1, 1)


def layer_norm_munit(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    """Custom LayerNorm: networks/networks.py:736-752.

    Per-sample mean and UNBIASED std over C*H*W, eps added to std (not var)."""
    b = x.shape[0]
    flat = x.reshape(b, -1)
    mu = flat.mean(1).view(b, 1, 1, 1)
    sd = flat.std(1).view(b, 1, 1, 1)
    y = (x - mu) / (sd + eps)
    return y * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def linear(P: Params, key: str, x: Tensor) -> Tensor:
    return F.linear(x, P[key + ".weight"], P[key + ".bias"])


# --------------------------------------------------------------------------
# generator sub-networks
# --------------------------------------------------------------------------

def style_encoder(P: Params, x: Tensor, cfg=DEFAULT_CFG, pre="enc_style.",
                  drop_mask: Optional[Tensor] = None) -> Tuple[List[Tensor], List[Tensor]]:
    """StyleEncoder (v2): networks/networks_v2.py:98-141."""
    h = _q(_act(conv_reflect(P, pre + "model.0.conv", _q(x), 7, 1, 3), "relu"))
    for i in range(1, 1 + cfg["style_downsample"]):
        h = _act(conv_reflect(P, pre + f"model.{i}.conv", h, 4, 2, 1), "relu")
        if i < cfg["style_downsample"]:
            h = _q(h)                                        # the last layer's ReLU is fused into the pooling
    h = h.mean(dim=(2, 3))                                   # AdaptiveAvgPool2d(1)
    h = torch.relu(linear(P, pre + "mapping.0", h))
    if drop_mask is not None:                                # Dropout(0.1), :119
        h = h * drop_mask
    h = torch.relu(linear(P, pre + "mapping.3", h))
    mus = [linear(P, pre + f"fcs.{i}", h) for i in range(cfg["num_cls"])]
    lvs = [linear(P, pre + f"fcvars.{i}", h) for i in range(cfg["num_cls"])]
    return mus, lvs


def content_encoder(P: Params, x: Tensor, cfg=DEFAULT_CFG, pre="enc_content.") -> Tensor:
    """ContentEncoder + ResBlocks: networks/networks.py:428-446, 480-489, 509-522."""
    h = _q(torch.relu(inst_norm(conv_reflect(P, pre + "model.0.conv", _q(x), 7, 1, 3))))
    nd = cfg["content_downsample"]
    for i in range(1, 1 + nd):
        h = _q(torch.relu(inst_norm(conv_reflect(P, pre + f"model.{i}.conv", h, 4, 2, 1))))
    rb = pre + f"model.{nd + 1}.model."
    for j in range(cfg["n_res"]):
        r = h
        h = _q(torch.relu(inst_norm(conv_reflect(P, rb + f"{j}.model.0.conv", h, 3, 1, 1))))
        h = inst_norm(conv_reflect(P, rb + f"{j}.model.1.conv", h, 3, 1, 1))
        h = _q(h + r)                                        # second block: norm, no act, += residual
    return h


def mlp(P: Params, style: Tensor, pre="mlp.") -> Tensor:
    """MLP 64->256->256->4096: networks/networks.py:491-503."""
    h = torch.relu(linear(P, pre + "model.0.fc", style.reshape(style.shape[0], -1)))
    h = torch.relu(linear(P, pre + "model.1.fc", h))
    return linear(P, pre + "model.2.fc", h)


def upsample2x(x: Tensor) -> Tensor:
    # nn.Upsample(scale_factor=2, mode='bilinear'), align_corners=False: networks_v2.py:154
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def decoder(P: Params, content: Tensor, adain_params: Tensor, cfg=DEFAULT_CFG,
            pre="dec.") -> Tuple[Tensor, Tensor]:
    """Decoder (v2): networks/networks_v2.py:144-169 with AdaIN parameters consumed
    in module order, 2*C per layer, bias("mean") first then weight("std"):
    networks/networks_v2.py:78-87.  The attention head is always evaluated."""
    h = _q(content)
    c = content.shape[1]
    off = 0
    rb = pre + "model.0.model."
    for j in range(cfg["n_res"]):
        r = h
        for t, act in ((0, "relu"), (1, "none")):
            bias = adain_params[:, off:off + c]
            weight = adain_params[:, off + c:off + 2 * c]
            off += 2 * c
            h = conv_reflect(P, rb + f"{j}.model.{t}.conv", h, 3, 1, 1)
            h = _act(adain(h, weight, bias), act)
            if t == 0:
                h = _q(h)
        h = _q(h + r)
    idx = 2
    for _ in range(cfg["content_downsample"]):
        h = _q(upsample2x(h))
        h = conv_reflect(P, pre + f"model.{idx}.conv", h, 5, 1, 2)
        h = _q(torch.relu(layer_norm_munit(h, P[pre + f"model.{idx}.norm.gamma"],
                                           P[pre + f"model.{idx}.norm.beta"])))
        idx += 2
    img = torch.tanh(conv_reflect(P, pre + "image_content.conv", h, 7, 1, 3))
    att = torch.sigmoid(conv_reflect(P, pre + "image_attention.conv", h, 7, 1, 3))
    return img, att


def decode(P: Params, content: Tensor, style: Tensor, cfg=DEFAULT_CFG) -> Tuple[Tensor, Tensor]:
    """AdaINGen_v2.decode: networks/networks_v2.py:71-76."""
    return decoder(P, content, mlp(P, style), cfg)


def _lstm_direction(x: Tensor, lens: Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """One direction of one LSTM layer over a padded [T,B,I] batch with per-sample
    lengths (the semantics pack_padded_sequence gives nn.LSTM,
    networks/networks_v2.py:224-233).  Gate order i,f,g,o.  Returns (out[T,B,H]
    zero at padded steps, h_final[B,H], c_final[B,H])."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = [None] * T
    xp = F.linear(x, w_ih, b_ih)                              # [T,B,4H]
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        m = (lens > t).to(x.dtype).view(B, 1)                 # sample still active at t
        g = xp[t] + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        c = m * c_new + (1 - m) * c
        h = m * h_new + (1 - m) * h
        outs[t] = m * h_new
    return torch.stack(outs, 0), h, c


def text_encoder(P: Params, style: Tensor, tokens: Tensor, lens: Tensor, cfg=DEFAULT_CFG,
                 pre="enc_txt.") -> Tuple[List[Tensor], List[Tensor]]:
    """TxtEncoder (v2).forward: networks/networks_v2.py:213-254, deterministic mode
    (all dropouts off).  The sort/unsort by length is a no-op on the result
     and is omitted; the batch-dimension cat+view quirk
    (:248-249, gotcha #1) is reproduced exactly."""
    B, T = tokens.shape
    emb = F.embedding(tokens.t(), P[pre + "embed_tokens.weight"], padding_idx=0)  # [T,B,E]
    x = torch.cat([emb, style.unsqueeze(0).expand(T, -1, -1)], dim=-1)
    lens = lens.to(torch.long)
    fin_h, fin_c = [], []
    for layer in range(cfg["num_layers"]):
        outs = []
        hs, cs = [], []
        for d, suffix in enumerate(("", "_reverse")):
            o, h, c = _lstm_direction(
                x, lens,
                P[pre + f"lstm.weight_ih_l{layer}{suffix}"], P[pre + f"lstm.weight_hh_l{layer}{suffix}"],
                P[pre + f"lstm.bias_ih_l{layer}{suffix}"], P[pre + f"lstm.bias_hh_l{layer}{suffix}"],
                reverse=(d == 1))
            outs.append(o); hs.append(h); cs.append(c)
        x = torch.cat(outs, dim=-1)                           # input of next layer [T,B,2H]
        fin_h.append(torch.cat(hs, dim=-1))                   # combine_bidir: [B,2H] per layer
        fin_c.append(torch.cat(cs, dim=-1))
    final_h = torch.stack(fin_h, 0)                           # [L,B,2H]
    final_c = torch.stack(fin_c, 0)
    out = torch.cat([final_h, final_c], dim=1).reshape(B, -1)  # (L, 2B, 2H) -> (B, 4*L*H): the quirk
    mus = [linear(P, pre + f"fcs.{i}", out) for i in range(cfg["num_cls"])]
    lvs = [linear(P, pre + f"fcvars.{i}", out) for i in range(cfg["num_cls"])]
    return mus, lvs


def encode(P: Params, x: Tensor, cfg=DEFAULT_CFG, drop_mask=None):
    """AdaINGen_v2.encode: networks/networks_v2.py:61-65."""
    mus, lvs = style_encoder(P, x, cfg, drop_mask=drop_mask)
    return content_encoder(P, x, cfg), mus, lvs


# --------------------------------------------------------------------------
# discriminator
# --------------------------------------------------------------------------

def dis_forward(D: Params, x: Tensor, cfg=DEFAULT_CFG):
    """MsImageDis.forward: networks/networks.py:102-114."""
    outs = []
    for s in range(cfg["dis_num_scales"]):
        h = _q(x)
        for i in range(cfg["dis_n_layer"]):
            h = _q(_act(conv_reflect(D, f"cnns_feat.{s}.{i}.conv", h, 4, 2, 1), "lrelu"))
        src = F.conv2d(h, D[f"cnns_src.{s}.weight"], D[f"cnns_src.{s}.bias"])
        cls = F.conv2d(h, D[f"cnns_cls.{s}.weight"]).reshape(x.shape[0], -1)
        outs.append((src, cls))
        x = F.avg_pool2d(x, 2)            # == F.interpolate(scale_factor=0.5, bilinear)
    return outs


def dis_loss(D: Params, fake: Tensor, real: Tensor, real_cls: Tensor, cfg=DEFAULT_CFG) -> Tensor:
    """MsImageDis.calc_dis_loss, lsgan: networks/networks.py:116-146."""
    loss = 0.0
    for (sf, _), (sr, cr) in zip(dis_forward(D, fake, cfg), dis_forward(D, real, cfg)):
        loss = loss + (torch.mean(sf ** 2) + torch.mean((sr - 1) ** 2)) * cfg["gan_w"]
        loss = loss + F.binary_cross_entropy_with_logits(cr, real_cls) * cfg["cls_w"]
    return loss


def gen_adv_loss(D: Params, fake: Tensor, target_cls: Tensor, cfg=DEFAULT_CFG) -> Tensor:
    """MsImageDis.calc_gen_loss, lsgan: networks/networks.py:148-170."""
    loss = 0.0
    for sf, cf in dis_forward(D, fake, cfg):
        loss = loss + torch.mean((sf - 1) ** 2) * cfg["gan_w"]
        loss = loss + F.binary_cross_entropy_with_logits(cf, target_cls) * cfg["cls_w"]
    return loss


# --------------------------------------------------------------------------
# GMM helpers
# --------------------------------------------------------------------------

def assign_label(label: Tensor) -> Tensor:
    """tools.py:40-47 ('CelebA', normalize=True): {0,1} -> {-1,+1}."""
    return label * 2.0 - 1.0


def gmm_sample(mu: Tensor, eps: Tensor, stddev: float = 0.5) -> Tensor:
    """dist_sampling_split: tools.py:65-70.  ``eps`` is the standard-normal draw of
    shape (1, c_dim, B, num_cls) that Normal(mu, stddev).sample((1, c_dim))
    consumes; z[b, j*c_dim + k] = mu[b, j] + stddev * eps[0, k, b, j]."""
    smp = mu.unsqueeze(0).unsqueeze(0) + stddev * eps          # (1, c_dim, B, num_cls)
    return smp.transpose(2, 1).transpose(3, 2).contiguous().view(mu.shape[0], -1)


def gmm_kl(mus: List[Tensor], logvars: List[Tensor], c: Tensor, sigma: float = 0.25) -> Tensor:
    """gmm_kl_distance_sp: gmm.py:13-22 (sigma is the variance, stddev**2)."""
    sig = torch.tensor(sigma, dtype=c.dtype)
    tot = 0.0
    for i, (m, lv) in enumerate(zip(mus, logvars)):
        v = lv.exp()
        tot = tot + (0.5 * (torch.log(sig / v) + (v + (m - c[:, i:i + 1]) ** 2) / sig - 1.0)).sum(1).mean()
    return tot


def gmm_em(mus: List[Tensor], c: Tensor) -> Tensor:
    """gmm_earth_mover_distance_sp: gmm.py:33-41."""
    tot = 0.0
    for i, m in enumerate(mus):
        tot = tot + torch.abs(m - c[:, i:i + 1]).sum(1).mean()
    return tot


def blend(img: Tensor, att: Tensor, x_real: Tensor, use_attention: bool) -> Tensor:
    # solver.py:160-161
    return img * att + x_real * (1 - att) if use_attention else img


def l1(a: Tensor, b: Tensor) -> Tensor:
    return torch.mean(torch.abs(a - b))


# --------------------------------------------------------------------------
# the two phases of the training step
# --------------------------------------------------------------------------

def translate(G: Params, x: Tensor, tokens: Tensor, lens: Tensor, use_attention=True, cfg=DEFAULT_CFG):
    """Inference path with Solver.sample's semantics (cat then decode): solver.py:142-149, 255-260."""
    content, mus, _ = encode(G, x, cfg)
    mt, _ = text_encoder(G, torch.cat(mus, 1), tokens, lens, cfg)
    img, att = decode(G, content, torch.cat(mt, 1), cfg)
    return blend(img, att, x, use_attention)


def dis_phase_loss(G: Params, D: Params, batch: dict, eps1: Tensor, use_attention: bool,
                   cfg=DEFAULT_CFG) -> Tensor:
    """Solver.dis_update up to the loss: solver.py:317-336 (gp_w=0, use_r1=False)."""
    x = batch["x_real"]
    content, mus, _ = encode(G, x, cfg)
    style_real = torch.cat(mus, 1)
    style1 = gmm_sample(batch["c_trg"], eps1, cfg["stddev"])
    mt, _ = text_encoder(G, style_real, batch["txt"], batch["txt_lens"], cfg)
    f0, a0 = decode(G, content, torch.cat(mt, 1), cfg)
    f1, a1 = decode(G, content, style1, cfg)
    f0 = blend(f0, a0, x, use_attention)
    f1 = blend(f1, a1, x, use_attention)
    return dis_loss(D, f0, x, batch["label_src"], cfg) + dis_loss(D, f1, x, batch["label_src"], cfg)


def gen_phase_losses(G: Params, D: Params, batch: dict, eps1: Tensor, eps2: Tensor,
                     use_attention: bool, ds_w: float, cfg=DEFAULT_CFG) -> Dict[str, Tensor]:
    """Solver.gen_update up to the total loss: solver.py:151-238 (vgg_w=0, dist_mode kls)."""
    x = batch["x_real"]
    content, mus, lvs = encode(G, x, cfg)
    style_real = torch.cat(mus, 1)
    rec, rec_a = decode(G, content, style_real, cfg)
    rec = blend(rec, rec_a, x, use_attention)
    c_rec, s_rec, _ = encode(G, rec, cfg)
    mt, lvt = text_encoder(G, style_real, batch["txt"], batch["txt_lens"], cfg)
    style_txt = torch.cat(mt, 1)
    fake, fa = decode(G, content, style_txt, cfg)
    fake = blend(fake, fa, x, use_attention)
    style1 = gmm_sample(batch["c_trg"], eps1, cfg["stddev"])
    f1, a1 = decode(G, content, style1, cfg)
    style2 = gmm_sample(batch["c_trg"], eps2, cfg["stddev"])
    f2, a2 = decode(G, content, style2, cfg)
    f1 = blend(f1, a1, x, use_attention)
    f2 = blend(f2, a2, x, use_attention)
    L = {}
    L["loss_ds"] = l1(f1, f2.detach())
    c_rand, s_rand, _ = encode(G, f1, cfg)
    ds_w = max(ds_w - 1 / 1e5, 0.0)                            # solver.py:183
    c_fake, s_fake, _ = encode(G, fake, cfg)
    cyc, cyc_a = decode(G, c_fake, style_real, cfg)
    cyc = blend(cyc, cyc_a, x, use_attention)
    L["loss_gen_recon_x"] = l1(rec, x)
    L["loss_gen_recon_c_real"] = l1(c_rec, content)
    L["loss_gen_recon_c_fake"] = l1(c_fake, content)
    L["loss_gen_recon_c_rand"] = l1(c_rand, content)
    L["loss_gen_recon_s_real"] = l1(torch.cat(s_rec, 1), style_real)
    L["loss_gen_recon_s_fake"] = l1(torch.cat(s_fake, 1), style_txt)
    L["loss_gen_recon_s_rand"] = l1(torch.cat(s_rand, 1), style1)
    L["loss_gen_cycrecon_x"] = l1(cyc, x)
    L["loss_gen_adv"] = gen_adv_loss(D, fake, batch["label_trg"], cfg) + \
        gen_adv_loss(D, f1, batch["label_trg"], cfg)
    L["loss_kl_x"] = gmm_kl(mus, lvs, batch["c_src"], cfg["stddev"] ** 2)
    L["loss_kl_trg"] = gmm_kl(mt, lvt, batch["c_trg"], cfg["stddev"] ** 2)
    L["loss_gen_total"] = (L["loss_gen_adv"]
                           + cfg["recon_x_w"] * L["loss_gen_recon_x"]
                           + cfg["recon_c_w"] * (L["loss_gen_recon_c_real"] + L["loss_gen_recon_c_fake"]
                                                 + L["loss_gen_recon_c_rand"])
                           + cfg["recon_s_w"] * (L["loss_gen_recon_s_real"] + L["loss_gen_recon_s_fake"]
                                                 + L["loss_gen_recon_s_rand"])
                           + cfg["recon_x_cyc_w"] * L["loss_gen_cycrecon_x"]
                           + cfg["kl_w"] * (L["loss_kl_x"] + L["loss_kl_trg"])
                           - ds_w * L["loss_ds"])
    L["_ds_w"] = ds_w
    return L


# --------------------------------------------------------------------------
# optimizer / EMA
# --------------------------------------------------------------------------

def adam_step(P: Params, grads: Dict[str, Optional[Tensor]], state: dict, lr: float, cfg=DEFAULT_CFG):
    """torch.optim.Adam with coupled L2 (solver.py:65-68), torch>=2 semantics: entries
    whose grad is None are skipped entirely (no decay, no moment/step update)."""
    b1, b2, eps, wd = cfg["beta1"], cfg["beta2"], cfg["adam_eps"], cfg["weight_decay"]
    for k, p in P.items():
        g = grads.get(k)
        if g is None:
            continue
        st = state.setdefault(k, dict(step=0, m=torch.zeros_like(p), v=torch.zeros_like(p)))
        st["step"] += 1
        g = g + wd * p
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** st["step"]
        bc2 = 1 - b2 ** st["step"]
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(st["m"], denom, value=-lr / bc1)


def ema_step(P: Params, P_avg: Params, beta: float = 0.999):
    """moving_average: utils.py:52-54, p_avg = lerp(p, p_avg, beta)."""
    for k in P:
        P_avg[k] = torch.lerp(P[k], P_avg[k], beta)


GEN_BUFFER_SUFFIXES = ("running_mean", "running_var")


def trainable(P: Params) -> Params:
    """Parameters (as opposed to AdaIN's dummy buffers) of a state dict."""
    return {k: v for k, v in P.items() if not k.endswith(GEN_BUFFER_SUFFIXES)}


class OracleSolver:
    """Stateful wrapper mirroring Solver's step methods (solver.py:151-240, 317-357)
    on top of the functional oracle above.  Deterministic mode: dropout off; the
    GMM noise ``eps`` is passed in explicitly."""

    def __init__(self, gen_sd: Params, dis_sd: Params, cfg=None):
        self.cfg = dict(DEFAULT_CFG) if cfg is None else cfg
        self.G = {k: v.detach().clone().float() for k, v in trainable(gen_sd).items()}
        self.D = {k: v.detach().clone().float() for k, v in dis_sd.items()}
        self.G_avg = {k: v.clone() for k, v in self.G.items()}
        self.D_avg = {k: v.clone() for k, v in self.D.items()}
        self.g_state, self.d_state = {}, {}
        self.use_attention = True
        self.ds_w = self.cfg["ds_w"]
        self.lr = self.cfg["lr"]
        self.losses: Dict[str, float] = {}
        self.last_gen_grads: Dict[str, Optional[Tensor]] = {}
        self.last_dis_grads: Dict[str, Optional[Tensor]] = {}

    @staticmethod
    def _leaf(P):
        return {k: v.detach().requires_grad_(True) for k, v in P.items()}

    def dis_update(self, batch, eps1):
        D = self._leaf(self.D)
        with torch.no_grad():
            pass
        # the reference does not detach the fakes (solver.py:327-328); gradients to G are
        # discarded by gen_opt.zero_grad() at solver.py:153, so G is evaluated without a graph here
        with torch.no_grad():
            G = self.G
            x = batch["x_real"]
            content, mus, _ = encode(G, x, self.cfg)
            style_real = torch.cat(mus, 1)
            style1 = gmm_sample(batch["c_trg"], eps1, self.cfg["stddev"])
            mt, _ = text_encoder(G, style_real, batch["txt"], batch["txt_lens"], self.cfg)
            f0, a0 = decode(G, content, torch.cat(mt, 1), self.cfg)
            f1, a1 = decode(G, content, style1, self.cfg)
            f0 = blend(f0, a0, x, self.use_attention)
            f1 = blend(f1, a1, x, self.use_attention)
        loss = dis_loss(D, f0, x, batch["label_src"], self.cfg) + \
            dis_loss(D, f1, x, batch["label_src"], self.cfg)
        loss.backward()
        grads = {k: v.grad for k, v in D.items()}
        self.last_dis_grads = grads
        self.fakes = (f0, f1)
        adam_step(self.D, grads, self.d_state, self.lr, self.cfg)
        self.losses["loss_dis"] = float(loss)
        self.losses["loss_dis_all"] = float(loss)
        return float(loss)

    def gen_update(self, batch, eps1, eps2):
        G = self._leaf(self.G)
        L = gen_phase_losses(G, self.D, batch, eps1, eps2, self.use_attention, self.ds_w, self.cfg)
        self.ds_w = L.pop("_ds_w")
        L["loss_gen_total"].backward()
        grads = {k: v.grad for k, v in G.items()}
        self.last_gen_grads = grads
        adam_step(self.G, grads, self.g_state, self.lr, self.cfg)
        for k, v in L.items():
            self.losses[k] = float(v)
        return float(L["loss_gen_total"])

    def smooth_moving(self):
        ema_step(self.G, self.G_avg, self.cfg["ema_beta"])
        ema_step(self.D, self.D_avg, self.cfg["ema_beta"])

    def update_attention_status(self, iters):
        # solver.py:109-111 with att_status == True (configs gen.use_attention)
        self.use_attention = iters >= 10000


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------

def synthetic_batch(B: int, size: int = 128, seed: int = 0, T: int = 80, device="cpu") -> dict:
    """x_real U(-1,1), labels Bernoulli(1/2), uniform token rows [bos, U{4..101}.., eos, pad..]."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, size, size, generator=g) * 2 - 1
    label_src = (torch.rand(B, 8, generator=g) > 0.5).float()
    label_trg = (torch.rand(B, 8, generator=g) > 0.5).float()
    lens = torch.randint(3, 30, (B,), generator=g)
    txt = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        n = int(lens[b])
        txt[b, 0] = 1
        txt[b, 1:n - 1] = torch.randint(4, 102, (n - 2,), generator=g)
        txt[b, n - 1] = 2
    batch = dict(x_real=x, label_src=label_src, label_trg=label_trg,
                 c_src=assign_label(label_src), c_trg=assign_label(label_trg),
                 txt=txt, txt_lens=lens)
    return {k: v.to(device) for k, v in batch.items()}
