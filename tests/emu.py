"""CPU emulator of the abstract gconv / wgrad operations (test infrastructure).

Executes the launch plans of dwc_gan_b200.plan with plain torch indexing so that the geometry
(taps, offsets, weight packing, flat/box modes) can be validated without a GPU.  The CUDA kernels
implement exactly this semantics (include/dwc_b200.h).
"""
import torch


def _rows(plan):
    bx, by, bn = plan.box
    tx, ty, tn = plan.tiles
    r = torch.arange(bx * by * bn)
    rx, ry, rn = r % bx, (r // bx) % by, r // (bx * by)
    t = torch.arange(tx * ty * tn)
    ox, oy, on = (t % tx) * bx, ((t // tx) % ty) * by, (t // (tx * ty)) * bn
    x = (ox[:, None] + rx[None, :]).reshape(-1)
    y = (oy[:, None] + ry[None, :]).reshape(-1)
    n = (on[:, None] + rn[None, :]).reshape(-1)
    return x, y, n


def _gather(storage, off, dims, strs, x, y, z, n, cols):
    C, X, Y, Z, N = dims
    inb = (x >= 0) & (x < X) & (y >= 0) & (y < Y) & (n < N) & (z >= 0) & (z < Z)
    idx = off + x * strs[1] + y * strs[2] + z * strs[3] + n * strs[4]
    idx = torch.where(inb, idx, torch.zeros_like(idx))
    g = storage[idx[:, None] + cols[None, :]]
    return g * inb[:, None].to(g.dtype)


def emu_gconv(plan):
    import copy
    if getattr(plan, "nphase", 1) > 1:
        for ph in range(plan.nphase):
            q = copy.copy(plan)
            q.nphase = 1
            q.w_off = plan.w_off + ph * plan.phase_w_off
            q.out_off = plan.out_off + ph * plan.phase_out_off
            emu_gconv(q)
        return
    a = plan.a.reshape(-1).double()
    C = plan.a_dim[0]
    x, y, n = _rows(plan)
    W = plan.w.reshape(-1)[plan.w_off:plan.w_off + plan.ncols_padded * len(plan.taps) * C]
    W = W.reshape(plan.ncols_padded, len(plan.taps), C).double()
    acc = torch.zeros(x.numel(), plan.ncols_padded, dtype=torch.double)
    cols = torch.arange(C)
    for t, (dx, dy, z) in enumerate(plan.taps):
        A = _gather(a, plan.a_off, plan.a_dim, plan.a_str, x + dx, y + dy, torch.full_like(x, z), n, cols)
        acc += A @ W[:, t, :].t()
    flag, img, pitch, fh, fw = plan.flat
    if flag:
        nn = x // img
        rem = x % img
        yy, xx = rem // pitch, rem % pitch
        valid = (nn < plan.valid[2]) & (yy < fh) & (xx < fw)
    else:
        nn, yy, xx = n, y, x
        valid = (x < plan.valid[0]) & (y < plan.valid[1]) & (n < plan.valid[2])
    off = plan.out_off + nn * plan.o_str[2] + yy * plan.o_str[1] + xx * plan.o_str[0]
    out = plan.out.reshape(-1)
    res = acc[:, :plan.ncols]
    if plan.bias is not None:
        res = res + plan.bias.double()[None, :]
    idx = (off[valid][:, None] + torch.arange(plan.ncols)[None, :]).reshape(-1)
    vals = res[valid].reshape(-1)
    if plan.accumulate:
        vals = vals + out[idx].double()
    out[idx] = vals.to(out.dtype)


def emu_wgrad(plan):
    a = plan.a.reshape(-1).double()
    b = plan.b.reshape(-1).double()
    x, y, n = _rows(plan)
    A = _gather(a, plan.a_off, plan.a_dim, plan.a_str, x, y, torch.zeros_like(x), n, torch.arange(plan.ca))
    dw = plan.dw.reshape(-1)
    for t, (dx, dy, z) in enumerate(plan.taps):
        B = _gather(b, plan.b_off, plan.b_dim, plan.b_str, x + dx, y + dy, torch.full_like(x, z), n, torch.arange(plan.cb))
        g = A.t() @ B                                        # [ca, cb]
        ia, ib = torch.arange(plan.ca), torch.arange(plan.cb)
        oa, ob = ia * plan.s_a, ib * plan.s_b
        va, vb = torch.ones_like(ia, dtype=torch.bool), torch.ones_like(ib, dtype=torch.bool)
        if plan.remap is not None:
            axis, div, lo_lim, hi_lim, hi_s, lo_s = plan.remap
            if axis == 1:
                oa, va = (ia // div) * hi_s + (ia % div) * lo_s, ((ia % div) < lo_lim) & ((ia // div) < hi_lim)
            else:
                ob, vb = (ib // div) * hi_s + (ib % div) * lo_s, ((ib % div) < lo_lim) & ((ib // div) < hi_lim)
        idx = (oa[:, None] + t * plan.s_t + ob[None, :])
        ok = (va[:, None] & vb[None, :]).reshape(-1)
        idx, gv = idx.reshape(-1)[ok], g.reshape(-1)[ok].to(dw.dtype)
        if plan.accumulate:
            dw[idx] += gv
        else:
            dw[idx] = gv
    if plan.dbias is not None:
        s = A.sum(0).to(plan.dbias.dtype)
        if plan.accumulate:
            plan.dbias += s
        else:
            plan.dbias.copy_(s)


# ---- weight packing (mirrors dwc_pack_weights) : w is [Cout, KH, KW, Cin]
def pack_fwd(w, rows_padded=None):
    co = w.shape[0]
    rows_padded = rows_padded or co
    out = w.new_zeros(rows_padded, w[0].numel())
    out[:co] = w.reshape(co, -1)
    return out


def pack_dgrad_s1(w, rows_padded=None):
    co, kh, kw, ci = w.shape
    rows_padded = rows_padded or ci
    wf = w.flip(1, 2)                                         # tap reversed
    out = w.new_zeros(rows_padded, kh * kw * co)
    out[:ci] = wf.permute(3, 1, 2, 0).reshape(ci, -1)
    return out


def pack_dgrad_s2(w, rows_padded=None):
    co, kh, kw, ci = w.shape
    assert kh == 4 and kw == 4
    rows_padded = rows_padded or ci
    out = w.new_zeros(4, rows_padded, 4 * co)
    for py in range(2):
        for px in range(2):
            blocks = []
            for ip in range(2):
                for jp in range(2):
                    blocks.append(w[:, 2 * (1 - ip) + py, 2 * (1 - jp) + px, :].t())   # [ci, co]
            out[py * 2 + px, :ci] = torch.cat(blocks, dim=1)
    return out


# ---- buffer helpers
def make_padded(x_nchw, p, layout, dtype=torch.float32):
    """reflect-pad an NCHW tensor and store it as an HB (plain or parity planes)."""
    import torch.nn.functional as F
    from dwc_gan_b200.plan import HB
    n, c, h, w = x_nchw.shape
    xp = F.pad(x_nchw, (p, p, p, p), mode="reflect") if p > 0 else x_nchw
    nhwc = xp.permute(0, 2, 3, 1).contiguous().to(dtype)
    if layout == 0:
        return HB(nhwc, n, h, w, c, p, 0)
    planes = torch.stack([nhwc[:, py::2, px::2, :] for py in range(2) for px in range(2)], dim=1).contiguous()
    return HB(planes, n, h, w, c, p, 1)


def make_zero_haloed(y_nchw, halo, dtype=torch.float32):
    from dwc_gan_b200.plan import HB
    n, c, h, w = y_nchw.shape
    t = torch.zeros(n, h + 2 * halo, w + 2 * halo, c, dtype=dtype)
    t[:, halo:halo + h, halo:halo + w, :] = y_nchw.permute(0, 2, 3, 1).to(dtype)
    return HB(t, n, h, w, c, halo, 0)


def make_rows(img_nchw, pool, pad, sx, ys, wo, dtype=torch.float32):
    """torch version of dwc_image_rows_fwd."""
    import torch.nn.functional as F
    x = F.avg_pool2d(img_nchw, pool) if pool > 1 else img_nchw
    xp = F.pad(x, (pad,) * 4, mode="reflect")
    n, c, hp, wp = xp.shape
    xpp = F.pad(xp, (0, sx * wo + 8 - wp if sx * wo + 8 > wp else 0, 0, 0))
    rows = torch.zeros(n, hp, wo, 8, 8, dtype=dtype)
    for j in range(8):
        rows[:, :, :, j, :c] = xpp[:, :, :, j:j + sx * wo:sx][:, :, :, :wo].permute(0, 2, 3, 1).to(dtype)
    rows = rows.reshape(n, hp // ys, ys, wo, 64).permute(0, 2, 1, 3, 4).contiguous()
    return rows


def pack_rows_fwd(w):
    co, kh, kw, ci = w.shape
    out = w.new_zeros(co, kh, 8, 8)
    out[:, :, :kw, :ci] = w
    return out.reshape(co, kh * 64)


def pack_rows_dgrad(w):
    co, kh, kw, ci = w.shape
    out = w.new_zeros(ci, kh, 8, 8)
    out[:, :, :kw, :co] = w.flip(1, 2).permute(3, 1, 2, 0)
    return out.reshape(ci, kh * 64)


def make_heads_rows(dy_nchw, halo, dtype=torch.float32):
    """torch version of dwc_heads_bwd_rows' two outputs (rows_d, win) from dy [N, 4, H, W]."""
    import torch.nn.functional as F
    n, c, h, w = dy_nchw.shape
    dyz = F.pad(dy_nchw, (halo, halo + 8, halo, halo))                     # extra zeros right for the windows
    hh, wh = h + 2 * halo, w + 2 * halo
    rows_d = torch.zeros(n, hh, wh, 8, 8, dtype=dtype)
    for j in range(8):
        rows_d[:, :, :, j, :c] = dyz[:, :, :, j:j + wh].permute(0, 2, 3, 1).to(dtype)
    wu = w + halo
    dyl = F.pad(dy_nchw, (8, halo, 0, 0))                                  # dy[u - j]: zeros on the left
    win = torch.zeros(n, h, wu, 8, 8, dtype=dtype)
    for j in range(8):
        win[:, :, :, j, :c] = dyl[:, :, :, 8 - j:8 - j + wu].permute(0, 2, 3, 1).to(dtype)
    return rows_d.reshape(n, hh, wh, 64), win.reshape(n, h, wu, 64)
