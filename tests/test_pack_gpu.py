"""dwc_pack_weights_batch (one launch per network and step: coalesced cast / tiled-transpose paths) must produce exactly
what the element-wise reference kernel dwc_pack_weights produces, for every operand kind the networks register."""
import pytest
import torch

from dwc_gan_b200 import _lib as L

pytestmark = pytest.mark.gpu

# cout, k, cin, mode, rows_padded, dtype         (mode 0 forward, 1 stride-1 dgrad, 2 stride-2 dgrad, 3/4 row-im2col)
ENTRIES = [
    (256, 3, 256, 0, 256, torch.bfloat16), (256, 3, 256, 1, 256, torch.bfloat16),
    (128, 4, 64, 0, 128, torch.bfloat16), (128, 4, 64, 2, 64, torch.bfloat16),
    (512, 4, 512, 0, 512, torch.bfloat16), (512, 4, 512, 2, 512, torch.bfloat16),
    (64, 5, 128, 0, 64, torch.bfloat16), (64, 5, 128, 1, 128, torch.bfloat16),
    (32, 5, 64, 0, 64, torch.bfloat16),              # zero rows beyond cout (256x256 variant)
    (64, 7, 3, 1, 16, torch.bfloat16),               # few-channel image gradient: zero rows beyond cin
    (64, 7, 3, 3, 64, torch.bfloat16), (4, 7, 64, 4, 64, torch.bfloat16), (64, 4, 3, 3, 64, torch.bfloat16),
    (128, 3, 64, 0, 128, torch.float32), (128, 3, 64, 1, 64, torch.float32),
]


def _shape(cout, k, cin, mode, rows):
    if mode == 0:
        return (rows, k * k * cin)
    if mode == 1:
        return (rows, k * k * cout)
    if mode == 2:
        return (4, rows, 4 * cout)
    return (rows, k * 64)


def test_batch_pack_equals_elementwise_pack():
    torch.manual_seed(0)
    lib = L.lib()
    ws, outs, refs = [], [], []
    arr = (L.PackEntry * len(ENTRIES))()
    for i, (cout, k, cin, mode, rows, dt) in enumerate(ENTRIES):
        w = torch.randn(cout, k, k, cin, device="cuda")
        out = torch.full(_shape(cout, k, cin, mode, rows), 7.0, dtype=dt, device="cuda")
        ref = torch.full_like(out, -3.0)
        L.check(lib.dwc_pack_weights(L.ptr(w), cout, k, k, cin, mode, L.ptr(ref), L.dt(ref), rows, L.stream()), "pack")
        arr[i] = L.PackEntry(w.data_ptr(), out.data_ptr(), cout, k, k, cin, mode, rows, L.dt(out), 0, out.numel())
        ws.append(w); outs.append(out); refs.append(ref)
    table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    L.check(lib.dwc_pack_weights_batch(L.ptr(table), len(ENTRIES), L.stream()), "pack_batch")
    torch.cuda.synchronize()
    for e, out, ref in zip(ENTRIES, outs, refs):
        assert torch.equal(out, ref), e
