"""CPU checks of the launch-plan geometry: the emulated gconv/wgrad plans must reproduce
F.conv2d(F.pad(x, reflect)) and its autograd gradients (networks/networks.py:531,577-580)."""
import pytest
import torch
import torch.nn.functional as F

from dwc_gan_b200 import plan as P
from dwc_gan_b200.plan import HB
from tests import emu

CASES = [
    # n, h, w, cin, cout, k, stride, pad
    (2, 8, 8, 8, 16, 3, 1, 1),
    (1, 16, 16, 8, 8, 5, 1, 2),
    (2, 8, 8, 3, 8, 7, 1, 3),
    (2, 8, 8, 8, 16, 4, 2, 1),
    (3, 4, 4, 8, 8, 4, 2, 1),
    (5, 2, 2, 8, 8, 4, 2, 1),
    (2, 16, 8, 8, 4, 7, 1, 3),
]


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,p", CASES)
def test_conv_plans_match_torch(n, h, w, cin, cout, k, s, p):
    torch.manual_seed(0)
    x = torch.randn(n, cin, h, w, dtype=torch.double)
    wt = torch.randn(cout, cin, k, k, dtype=torch.double) * 0.1
    bias = torch.randn(cout, dtype=torch.double)
    xpad = F.pad(x, (p, p, p, p), mode="reflect").requires_grad_(True)
    wt_r = wt.clone().requires_grad_(True)
    y_ref = F.conv2d(xpad, wt_r, bias, stride=s)
    ho, wo = y_ref.shape[2:]
    w_krsc = wt.permute(0, 2, 3, 1).contiguous()                   # [Cout, KH, KW, Cin]

    # ---- forward
    layout = 0 if s == 1 else 1
    xp = emu.make_padded(x, p, layout, torch.double)
    hy = k - 1 if s == 1 else 1
    y = HB.empty(n, ho, wo, cout, hy, 0, torch.double, "cpu", zero=True)
    pl = P.plan_conv_fwd(xp, emu.pack_fwd(w_krsc, 16 if cout < 16 else cout), cout, 16 if cout < 16 else cout,
                         bias, y, k, s, 0)
    emu.emu_gconv(pl)
    assert torch.allclose(y.interior().permute(0, 3, 1, 2), y_ref, atol=1e-10)

    # ---- backward reference
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    # ---- dgrad
    dyz = emu.make_zero_haloed(dy, hy, torch.double)
    dxp = HB.empty(n, h, w, cin, p, layout, torch.double, "cpu")
    dxp.t.fill_(float("nan"))
    wd = emu.pack_dgrad_s1(w_krsc) if s == 1 else emu.pack_dgrad_s2(w_krsc)
    for q in P.plan_conv_dgrad(dyz, wd, dxp, k, s, 0):
        emu.emu_gconv(q)
    got = dxp.padded_nhwc().permute(0, 3, 1, 2)
    assert torch.allclose(got, xpad.grad, atol=1e-10)
    if s == 1:
        # accumulate mode (ResBlock skip gradient already in the buffer): out += result
        skip = torch.randn_like(dxp.t)
        dxa = HB(skip.clone(), n, h, w, cin, p, layout)
        for q in P.plan_conv_dgrad(dyz, wd, dxa, k, s, 0, accumulate=True):
            emu.emu_gconv(q)
        assert torch.allclose(dxa.t, skip + dxp.t, atol=1e-10)

    # ---- wgrad (+bias)
    dw = torch.zeros(cout, k, k, cin, dtype=torch.double)
    db = torch.zeros(cout, dtype=torch.double)
    emu.emu_wgrad(P.plan_conv_wgrad(dyz, xp, dw, db, k, s, 0))
    assert torch.allclose(dw.permute(0, 3, 1, 2), wt_r.grad, atol=1e-9)
    assert torch.allclose(db, dy.sum((0, 2, 3)), atol=1e-9)


def test_choose_box():
    for w, h, n in [(32, 32, 16), (128, 128, 2), (4, 4, 48), (2, 2, 3), (64, 64, 1), (1, 1, 7)]:
        for rows in (64, 128):
            bx, by, bn = P.choose_box(w, h, n, rows)
            assert bx * by * bn == rows and bx <= max(w, 1) and by <= max(h, 1)


@pytest.mark.parametrize("k,s,p,pool", [(7, 1, 3, 1), (4, 2, 1, 1), (4, 2, 1, 2)])
def test_first_conv_row_im2col_plans(k, s, p, pool):
    """3-channel first layers through the 64-wide row-im2col buffer: forward and weight gradient."""
    torch.manual_seed(0)
    n, cin, cout, H = 2, 3, 16, 16
    img = torch.randn(n, cin, H, H, dtype=torch.double)
    wt = torch.randn(cout, cin, k, k, dtype=torch.double) * 0.1
    bias = torch.randn(cout, dtype=torch.double)
    x = F.avg_pool2d(img, pool) if pool > 1 else img
    xpad = F.pad(x, (p, p, p, p), mode="reflect")
    wt_r = wt.clone().requires_grad_(True)
    y_ref = F.conv2d(xpad, wt_r, bias, stride=s)
    ho, wo = y_ref.shape[2:]
    hp = xpad.shape[2]
    w_krsc = wt.permute(0, 2, 3, 1).contiguous()
    rows = emu.make_rows(img, pool, p, s, s, wo, torch.double)
    assert tuple(rows.shape) == P.rows_shape(n, hp, wo, s)
    hy = k - 1 if s == 1 else 1
    y = HB.empty(n, ho, wo, cout, hy, 0, torch.double, "cpu", zero=True)
    emu.emu_gconv(P.plan_first_conv_fwd(rows, n, hp, wo, ho, k, s, emu.pack_rows_fwd(w_krsc), cout, bias, y, 0))
    assert torch.allclose(y.interior().permute(0, 3, 1, 2), y_ref, atol=1e-10)
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    dyz = emu.make_zero_haloed(dy, hy, torch.double)
    dw = torch.zeros(cout, k, k, cin, dtype=torch.double)
    db = torch.zeros(cout, dtype=torch.double)
    emu.emu_wgrad(P.plan_first_conv_wgrad(dyz, rows, n, hp, wo, k, s, cin, dw, db, 0))
    assert torch.allclose(dw.permute(0, 3, 1, 2), wt_r.grad, atol=1e-9)
    assert torch.allclose(db, dy.sum((0, 2, 3)), atol=1e-9)


def test_heads_row_im2col_plans():
    """4-channel decoder heads backward through 64-wide window buffers: data gradient and weight gradient."""
    torch.manual_seed(0)
    n, cin, cout, H, k, p = 2, 8, 4, 16, 7, 3
    x = torch.randn(n, cin, H, H, dtype=torch.double)
    wt = torch.randn(cout, cin, k, k, dtype=torch.double) * 0.1
    xpad = F.pad(x, (p, p, p, p), mode="reflect").requires_grad_(True)
    wt_r = wt.clone().requires_grad_(True)
    y_ref = F.conv2d(xpad, wt_r, None)
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    w_krsc = wt.permute(0, 2, 3, 1).contiguous()
    rows_d, win = emu.make_heads_rows(dy, k - 1, torch.double)
    hh = H + 2 * (k - 1)
    dxp = HB.empty(n, H, H, cin, p, 0, torch.double, "cpu")
    dxp.t.fill_(float("nan"))
    emu.emu_gconv(P.plan_heads_dgrad(rows_d, n, hh, hh, emu.pack_rows_dgrad(w_krsc), dxp, k, 0))
    assert torch.allclose(dxp.t.permute(0, 3, 1, 2), xpad.grad, atol=1e-10)
    xp = emu.make_padded(x, p, 0, torch.double)
    dw = torch.zeros(cout, k, k, cin, dtype=torch.double)
    emu.emu_wgrad(P.plan_heads_wgrad(win, xp, dw, k, cout, 0))
    assert torch.allclose(dw.permute(0, 3, 1, 2), wt_r.grad, atol=1e-9)
