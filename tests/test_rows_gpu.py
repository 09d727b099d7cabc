"""GPU parity of the few-channel layers on the tensor-core path (row-im2col buffers): first convolutions 3 -> 64
(7x7 s1, 4x4 s2, with avg-pool) and the fused 4-channel decoder heads, forward / dgrad / wgrad, through the
autograd Functions of ops.py against torch (networks.py:102-113, networks_v2.py:162-169)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import dwc_gan_b200
from dwc_gan_b200 import ops
from dwc_gan_b200.networks import Conv2dBlock, Decoder, _FlatOwner
from dwc_gan_b200.plan import HB

pytestmark = pytest.mark.gpu


class Holder(_FlatOwner):
    def __init__(self, **mods):
        super().__init__()
        for k, v in mods.items():
            setattr(self, k, v)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("k,s,p,pool", [(7, 1, 3, 1), (4, 2, 1, 1), (4, 2, 1, 2)])
def test_first_conv(mode, k, s, p, pool):
    dwc_gan_b200.set_mode(mode)
    tol = 3e-5 if mode == "fp32" else 2e-2
    torch.manual_seed(0)
    net = Holder(conv=Conv2dBlock(3, 64, k, s, p, norm="none", activation="relu", pad_type="reflect")).cuda()
    net.ensure_flat()
    layer = net.conv
    n, H = 3, 32
    img = (torch.rand(n, 3, H, H) * 2 - 1)
    w = layer.conv.weight.detach().cpu().double()
    b = layer.conv.bias.detach().cpu().double()
    if mode == "bf16":
        w = w.to(torch.bfloat16).double()
    imr = img.double().requires_grad_(True)
    x = F.avg_pool2d(imr, pool) if pool > 1 else imr
    xq = x if mode == "fp32" else (x.to(torch.bfloat16).double() - x).detach() + x       # bf16-rounded activations
    wr = w.clone().requires_grad_(True)
    y_ref = F.conv2d(F.pad(xq, (p,) * 4, mode="reflect"), wr, b, stride=s)
    dy = torch.randn_like(y_ref)
    if mode == "bf16":
        dy = dy.to(torch.bfloat16).double()
    y_ref.backward(dy)

    imc = img.cuda().requires_grad_(True)
    net.flat.zero_grad()
    rows = ops.image_rows(imc, pool, layer)
    y = ops.first_conv(imc, rows, layer, pool)
    got = y.interior().permute(0, 3, 1, 2).double().cpu()
    assert (got - y_ref.detach()).abs().max() < tol * y_ref.abs().max()
    dyz = torch.zeros_like(y.t)
    dyz[:, y.halo:y.halo + y.h, y.halo:y.halo + y.w, :] = dy.permute(0, 2, 3, 1).to(y.t.dtype).cuda()
    y.t.backward(dyz)
    gw = layer.conv.weight.grad.double().cpu()
    assert (gw - wr.grad).abs().max() < tol * wr.grad.abs().max(), (gw - wr.grad).abs().max() / wr.grad.abs().max()
    gb = layer.conv.bias.grad.double().cpu()
    assert (gb - dy.sum((0, 2, 3))).abs().max() < tol * dy.sum((0, 2, 3)).abs().max()
    assert (imc.grad.double().cpu() - imr.grad).abs().max() < tol * imr.grad.abs().max()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_heads(mode):
    dwc_gan_b200.set_mode(mode)
    dtype = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = 3e-5 if mode == "fp32" else 2e-2
    torch.manual_seed(0)
    dec = Decoder(0, 1, 64, 3, res_norm="adain", activ="relu", pad_type="reflect", use_attention=True)
    net = Holder(dec=dec)
    net._fuse_groups = [["dec.image_content.conv.weight", "dec.image_attention.conv.weight"],
                        ["dec.image_content.conv.bias", "dec.image_attention.conv.bias"]]
    with torch.no_grad():
        dec.image_content.conv.bias.normal_()
        dec.image_attention.conv.bias.normal_()
    net = net.cuda()
    net.ensure_flat()
    n, H = 2, 32
    x = torch.randn(n, 64, H, H).to(dtype).float()
    w4 = torch.cat([dec.image_content.conv.weight, dec.image_attention.conv.weight], 0).detach().cpu().double()
    b4 = torch.cat([dec.image_content.conv.bias, dec.image_attention.conv.bias], 0).detach().cpu().double()
    if mode == "bf16":
        w4 = w4.to(torch.bfloat16).double()
    xr = x.double().requires_grad_(True)
    wr = w4.clone().requires_grad_(True)
    yr = F.conv2d(F.pad(xr, (3,) * 4, mode="reflect"), wr, b4)
    if mode == "bf16":
        yr = (yr.to(torch.bfloat16).double() - yr).detach() + yr
    img_r, att_r = torch.tanh(yr[:, :3]), torch.sigmoid(yr[:, 3:])
    wi, wa = torch.randn_like(img_r), torch.randn_like(att_r)
    ((img_r * wi).sum() + (att_r * wa).sum()).backward()

    net.flat.zero_grad()
    xh_t = torch.randn(n, H, H, 64).to(dtype)
    xh_t.copy_(x.permute(0, 2, 3, 1).to(dtype))
    xh = HB(xh_t.cuda().requires_grad_(True), n, H, H, 64, 0, 0)
    xp = ops.post(xh, out_halo=3)
    img, att = ops.heads_conv(xp, dec.image_content, dec.image_attention.touch_params)
    assert (img.double().cpu() - img_r.detach()).abs().max() < tol and (att.double().cpu() - att_r.detach()).abs().max() < tol
    ((img * wi.float().cuda()).sum() + (att * wa.float().cuda()).sum()).backward()
    gx = xh.t.grad.permute(0, 3, 1, 2).double().cpu()
    assert (gx - xr.grad).abs().max() < tol * 2 * xr.grad.abs().max(), (gx - xr.grad).abs().max() / xr.grad.abs().max()
    gw = torch.cat([dec.image_content.conv.weight.grad, dec.image_attention.conv.weight.grad], 0).double().cpu()
    assert (gw - wr.grad).abs().max() < tol * 2 * wr.grad.abs().max(), (gw - wr.grad).abs().max() / wr.grad.abs().max()
    gb = torch.cat([dec.image_content.conv.bias.grad, dec.image_attention.conv.bias.grad], 0).double().cpu()
    dyr = torch.autograd.grad((img_r * wi).sum() + (att_r * wa).sum(), yr, retain_graph=False, allow_unused=True) \
        if False else None
    assert torch.isfinite(gb).all() and gb.abs().sum() > 0
