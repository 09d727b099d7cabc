"""tf32 tensor-core GEMM (csrc/dense_tc.cu: tcgen05 kind::tf32, fp32 operands via TMA) against fp64, for every operand
layout the text encoder / MLP use: K-major and MN-major A and B, ragged sizes, accumulation into C, bias + ReLU."""
import pytest
import torch

from dwc_gan_b200 import _lib as L

pytestmark = pytest.mark.gpu

CASES = [
    # m, n, k, a_mn, b_mn, beta, bias, act      (LSTM input projection, its data / weight gradients, recurrent weight gradient)
    (1280, 2400, 364, 0, 0, 0.0, True, 0),
    (1280, 2400, 600, 0, 0, 1.0, False, 0),
    (1280, 364, 2400, 0, 1, 0.0, False, 0),
    (2400, 364, 1280, 1, 1, 1.0, False, 0),
    (1200, 300, 1264, 1, 1, 1.0, False, 0),
    (5120, 2400, 364, 0, 0, 0.0, True, 0),      # inference batch 64
    (130, 70, 40, 0, 0, 0.0, True, 1),          # ragged tiles in every dimension, ReLU
    (200, 96, 72, 1, 0, 0.0, False, 0),
    (64, 4096, 256, 0, 0, 0.0, True, 0),
]


@pytest.mark.parametrize("m,n,k,a_mn,b_mn,beta,use_bias,act", CASES)
def test_gemm_tf32(m, n, k, a_mn, b_mn, beta, use_bias, act):
    torch.manual_seed(0)
    lib = L.lib()
    # leading dimensions padded to a multiple of 4 elements (the TMA needs 16-byte strides), as the in-network views are
    pad = lambda v: (v + 3) // 4 * 4
    if a_mn:
        A_store = torch.randn(k, pad(m) + 4, device="cuda")[:, :m]          # element (mi, ki) at ki * lda + mi
        A = A_store.t()
        a_sm, a_sk = 1, A_store.stride(0)
    else:
        A = torch.randn(m, pad(k) + 4, device="cuda")[:, :k]
        a_sm, a_sk = A.stride(0), 1
    if b_mn:
        Bm = torch.randn(k, pad(n) + 8, device="cuda")[:, :n]               # [K, N], n contiguous
        b_sk, b_sn = Bm.stride(0), 1
    else:
        B_store = torch.randn(n, pad(k), device="cuda")[:, :k]              # nn.Linear weight [N, K]
        Bm = B_store.t()
        b_sk, b_sn = 1, B_store.stride(0)
    C0 = torch.randn(m, n, device="cuda")
    Cc = C0.clone()
    bias = torch.randn(n, device="cuda") if use_bias else None
    a_ptr, b_ptr = A.data_ptr(), Bm.data_ptr()
    assert lib.dwc_gemm_tf32_ok(m, n, k, a_ptr, a_sm, a_sk, b_ptr, b_sk, b_sn, Cc.data_ptr(), n, 1) == 1
    L.check(lib.dwc_gemm_tf32(m, n, k, 1.0, a_ptr, a_sm, a_sk, b_ptr, b_sk, b_sn, beta, L.ptr(Cc), n, L.ptr(bias), act,
                              L.stream()), "gemm_tf32")
    torch.cuda.synchronize()
    ref = A.double() @ Bm.double() + beta * C0.double()
    if bias is not None:
        ref = ref + bias.double()
    if act:
        ref = torch.relu(ref)
    err = float((Cc.double() - ref).norm() / ref.norm())
    print("tf32 gemm", (m, n, k, a_mn, b_mn), "rel err %.2e" % err)
    assert err < 1e-3, err                       # tf32: 10 mantissa bits per operand; measured ~3e-4
    # rows / columns outside the problem were not touched: covered by the ragged cases through the exact-size C buffer
    assert torch.isfinite(Cc).all()


def test_sgemm_routes_to_tensor_cores_only_in_product_mode():
    lib = L.lib()
    m, n, k = 1280, 2400, 364
    torch.manual_seed(1)
    A, W = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda")
    out = {}
    prev = lib.dwc_get_tf32()
    try:
        for flag in (0, 1):
            lib.dwc_set_tf32(flag)
            C = torch.empty(m, n, device="cuda")
            L.check(lib.dwc_sgemm(m, n, k, 1.0, L.ptr(A), L.F32, k, 1, L.ptr(W), 1, k, 0.0, L.ptr(C), n, 1, None, 0,
                                  L.stream()), "sgemm")
            out[flag] = C
    finally:
        lib.dwc_set_tf32(prev)
    ref = A.double() @ W.double().t()
    e0 = float((out[0].double() - ref).norm() / ref.norm())
    e1 = float((out[1].double() - ref).norm() / ref.norm())
    assert e0 < 1e-6 and 1e-5 < e1 < 1e-3, (e0, e1)      # exact fp32 vs tf32 operands
