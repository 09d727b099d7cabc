"""Text encoder body: embedding + style concat, 2-layer bidirectional LSTM with per-sample lengths
(pack_padded_sequence semantics), final (h, c) assembled with the reference's batch-dimension
cat/view quirk (networks/networks_v2.py:213-249).

One autograd Function for the whole recurrent body: the input projections and all weight
gradients are single GEMMs over the packed [T*B] rows (dwc_sgemm); only the recurrence itself is
sequential: one persistent cooperative kernel per layer and pass (csrc/lstm.cu).
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from .ops import _call, sgemm


def _lstm_groups(num_layers):
    groups = []
    for l in range(num_layers):
        for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            groups.append([f"enc_txt.lstm.{kind}_l{l}", f"enc_txt.lstm.{kind}_l{l}_reverse"])
    return groups


def final_state_rows(finals_h, finals_c):
    """[B, 4*L*H] feature from the per-layer final states (lists of L tensors [B, 2H], directions concatenated).

    The reference concatenates final_h and final_c along dim=1 - the BATCH dimension of [L, B, 2H] - and then views
    the (L, 2B, 2H) result as (B, -1) (networks_v2.py:248-249).  For B > 1 row r therefore holds 4 consecutive
    2H-vectors of the flat sequence (layer 0: h_0..h_{B-1}, c_0..c_{B-1}; layer 1: ...), mixing samples (SURVEY
    8a-3 #1).  Pure data movement: reproduced bit for bit."""
    B = finals_h[0].shape[0]
    final_h = torch.stack(list(finals_h), 0)                           # [L, B, 2H]
    final_c = torch.stack(list(finals_c), 0)
    return torch.cat([final_h, final_c], dim=1).reshape(B, -1)


def final_state_rows_bwd(dres, num_layers, B, H):
    """Adjoint of final_state_rows: ([L, B, 2H] gradient of final_h, [L, B, 2H] gradient of final_c)."""
    d = dres.contiguous().float().reshape(num_layers, 2 * B, 2 * H)
    return d[:, :B], d[:, B:]


class TxtBodyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, style, anchor, enc, tokens, lens, mask_in, mask_mid, need_grad):
        ops._require_cuda(style)
        owner = enc.__dict__["_owner"]
        f = owner.flat
        dev = style.device
        B, T = tokens.shape
        E, S, H, NL = enc.embed_dim, enc.style_dim, enc.hidden_size, enc.num_layers
        tokens = tokens.contiguous()
        lens = lens.to(torch.int64).contiguous()
        style = style.contiguous().float()
        pre = owner.param_name_of(enc, "lstm.weight_ih_l0")[:-len("weight_ih_l0")]
        emb_name = owner.param_name_of(enc, "embed_tokens.weight")

        x = torch.empty(T, B, E + S, dtype=torch.float32, device=dev)
        _call("dwc_embed_concat_fwd", L.ptr(tokens), L.ptr(f.raw(emb_name)), L.ptr(style), L.ptr(mask_in), L.ptr(x), B,
              T, E, S, L.stream())
        ones = torch.ones(T * B, dtype=torch.float32, device=dev)
        saved = []
        finals_h, finals_c = [], []
        inp = x
        for l in range(NL):
            I = inp.shape[2]
            wih = f.raw(pre + f"weight_ih_l{l}", 2 * 4 * H * I)
            whh = f.raw(pre + f"weight_hh_l{l}", 2 * 4 * H * H)
            bih = f.raw(pre + f"bias_ih_l{l}", 2 * 4 * H)
            bhh = f.raw(pre + f"bias_hh_l{l}", 2 * 4 * H)
            xproj = torch.empty(T, B, 2, 4 * H, dtype=torch.float32, device=dev)
            # xproj = b_ih + b_hh (rank-1 GEMM) then += inp @ Wih^T
            sgemm(T * B, 8 * H, 1, 1.0, ones, 1, 1, bhh, 8 * H, 1, 0.0, xproj, 8 * H, 1, bias=bih)
            sgemm(T * B, 8 * H, I, 1.0, inp, I, 1, wih, 1, I, 1.0, xproj, 8 * H, 1)
            hf = torch.empty(2, B, H, dtype=torch.float32, device=dev)
            cf = torch.empty(2, B, H, dtype=torch.float32, device=dev)
            out = torch.empty(T, B, 2 * H, dtype=torch.float32, device=dev)
            gates = torch.empty(T, B, 2, 4 * H, dtype=torch.float32, device=dev) if need_grad else None
            csave = torch.empty(T, B, 2, H, dtype=torch.float32, device=dev) if need_grad else None
            ws = torch.empty(int(L.lib().dwc_lstm_workspace_bytes(B, H)), dtype=torch.uint8, device=dev)
            _call("dwc_lstm_layer_fwd", T, B, H, L.ptr(xproj), L.ptr(whh), L.ptr(lens), L.ptr(out), L.ptr(gates),
                  L.ptr(csave), L.ptr(hf), L.ptr(cf), L.ptr(ws), L.stream())
            finals_h.append(hf.permute(1, 0, 2).reshape(B, 2 * H))   # combine_bidir
            finals_c.append(cf.permute(1, 0, 2).reshape(B, 2 * H))
            nxt = out
            if l + 1 < NL and mask_mid is not None:
                nxt = torch.empty_like(out)
                _call("dwc_mul", L.ptr(out), L.ptr(mask_mid), L.ptr(nxt), L.i64(out.numel()), L.stream())
            saved.append((inp, out, gates, csave))
            inp = nxt
        res = final_state_rows(finals_h, finals_c)                     # the reference's (L, 2B, 2H) -> (B, -1) quirk
        ctx.enc, ctx.dims = enc, (B, T, E, S, H, NL)
        ctx.saved = saved
        ctx.aux = (tokens, lens, mask_in, mask_mid, pre, emb_name)
        return res

    @staticmethod
    def backward(ctx, dres):
        enc = ctx.enc
        owner = enc.__dict__["_owner"]
        f = owner.flat
        B, T, E, S, H, NL = ctx.dims
        tokens, lens, mask_in, mask_mid, pre, emb_name = ctx.aux
        dev = dres.device
        dfh, dfc = final_state_rows_bwd(dres, NL, B, H)                 # [L, B, 2H] each
        dseq = None
        for l in range(NL - 1, -1, -1):
            inp, out, gates, csave = ctx.saved[l]
            I = inp.shape[2]
            wih = f.raw(pre + f"weight_ih_l{l}", 2 * 4 * H * I)
            whh = f.raw(pre + f"weight_hh_l{l}", 2 * 4 * H * H)
            dh0 = dfh[l].reshape(B, 2, H).permute(1, 0, 2).contiguous()
            dc0 = dfc[l].reshape(B, 2, H).permute(1, 0, 2).contiguous()
            dgates = torch.empty(T, B, 2, 4 * H, dtype=torch.float32, device=dev)
            ws = torch.empty(int(L.lib().dwc_lstm_workspace_bytes(B, H)), dtype=torch.uint8, device=dev)
            _call("dwc_lstm_layer_bwd", T, B, H, L.ptr(whh), L.ptr(lens), L.ptr(dseq), L.ptr(gates), L.ptr(csave),
                  L.ptr(dh0), L.ptr(dc0), L.ptr(dgates), L.ptr(ws), L.stream())
            names = [pre + f"{k}_l{l}{sfx}" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")
                     for sfx in ("", "_reverse")]
            f.touch(*names)
            g_wih = f.raw_grad(pre + f"weight_ih_l{l}", 2 * 4 * H * I, touch=False)
            g_whh = f.raw_grad(pre + f"weight_hh_l{l}", 2 * 4 * H * H, touch=False)
            g_bih = f.raw_grad(pre + f"bias_ih_l{l}", 2 * 4 * H, touch=False)
            g_bhh = f.raw_grad(pre + f"bias_hh_l{l}", 2 * 4 * H, touch=False)
            # dW_ih[2*4H, I] += dgates^T @ inp
            sgemm(8 * H, I, T * B, 1.0, dgates, 1, 8 * H, inp, I, 1, 1.0, g_wih, I, 1)
            _call("dwc_colsum", T * B, 8 * H, L.ptr(dgates), 8 * H, 1, L.ptr(g_bih), 1, L.stream())
            _call("dwc_colsum", T * B, 8 * H, L.ptr(dgates), 8 * H, 1, L.ptr(g_bhh), 1, L.stream())
            # dW_hh[dir][4H, H] += sum_t dgates[t,:,dir,:]^T @ h_prev(t)
            if T > 1:
                rows = (T - 1) * B
                dg_flat = dgates.view(-1)
                out_flat = out.view(-1)
                # forward direction: h_prev(t) = out[t-1, :, 0:H], t = 1..T-1
                sgemm(4 * H, H, rows, 1.0, dg_flat[B * 8 * H:], 1, 8 * H, out_flat, 2 * H, 1, 1.0, g_whh, H, 1)
                # reverse direction: h_prev(t) = out[t+1, :, H:2H], t = 0..T-2
                sgemm(4 * H, H, rows, 1.0, dg_flat[4 * H:], 1, 8 * H, out_flat[B * 2 * H + H:], 2 * H, 1, 1.0,
                      g_whh[4 * H * H:], H, 1)
            # dinp[T*B, I] = dgates @ Wih
            dinp = torch.empty(T, B, I, dtype=torch.float32, device=dev)
            sgemm(T * B, I, 8 * H, 1.0, dgates, 8 * H, 1, wih, I, 1, 0.0, dinp, I, 1)
            if l > 0 and mask_mid is not None:
                dm = torch.empty_like(dinp)
                _call("dwc_mul", L.ptr(dinp), L.ptr(mask_mid), L.ptr(dm), L.i64(dinp.numel()), L.stream())
                dinp = dm
            dseq = dinp
        dstyle = torch.empty(B, S, dtype=torch.float32, device=dev)
        emb_param = enc.embed_tokens.weight
        demb = f.raw_grad(emb_name) if emb_param.requires_grad else None
        _call("dwc_embed_concat_bwd", L.ptr(tokens), L.ptr(dseq), L.ptr(mask_in), L.ptr(demb), L.ptr(dstyle), B, T, E, S,
              int(enc.embed_tokens.padding_idx), L.stream())
        ctx.saved = None
        return dstyle, None, None, None, None, None, None, None


def txt_encode(enc, style, tokens, lens):
    """[B, 4*L*H] feature the 16 linear heads consume."""
    B, T = tokens.shape
    dev = style.device
    mask_in = mask_mid = None
    if enc.training:
        if enc.dropout_in > 0:
            mask_in = (torch.rand(T, B, enc.embed_dim, device=dev) >= enc.dropout_in).float() / (1 - enc.dropout_in)
        p = enc.lstm.dropout
        if enc.num_layers > 1 and p > 0:
            mask_mid = (torch.rand(T, B, 2 * enc.hidden_size, device=dev) >= p).float() / (1 - p)
    anchor = enc.lstm.weight_hh_l0
    need_grad = torch.is_grad_enabled() and (style.requires_grad or anchor.requires_grad)
    return TxtBodyFn.apply(style, anchor, enc, tokens, lens, mask_in, mask_mid, need_grad)
