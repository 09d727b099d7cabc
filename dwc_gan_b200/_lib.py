"""ctypes binding of libdwc_b200.so (the C ABI declared in include/dwc_b200.h).

The product path has no CPU fallback: if the shared library is missing, or a kernel entry
point reports an error, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdwc_b200.so")

F32, BF16 = 0, 1
SIMT, TC, TC_HALO, TC_HALO1 = 0, 1, 2, 3
MAX_TAPS = 64

_lib = None


class AdvTerm(C.Structure):           # dwc_adv_term_t
    _fields_ = [("kind", C.c_int32), ("row0", C.c_int32), ("row1", C.c_int32), ("target", C.c_float),
                ("weight", C.c_float)]


class GConv(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("backend", C.c_int32),
        ("a", C.c_void_p),
        ("a_dim", C.c_int64 * 5), ("a_str", C.c_int64 * 5),
        ("box", C.c_int32 * 3), ("tiles", C.c_int32 * 3), ("valid", C.c_int32 * 3),
        ("flat", C.c_int32), ("flat_img", C.c_int32), ("flat_pitch", C.c_int32),
        ("flat_h", C.c_int32), ("flat_w", C.c_int32),
        ("ntaps", C.c_int32),
        ("taps", C.POINTER(C.c_int32)),
        ("w", C.c_void_p),
        ("ncols", C.c_int32), ("ncols_padded", C.c_int32),
        ("bias", C.c_void_p),
        ("out", C.c_void_p),
        ("o_str", C.c_int64 * 3),
        ("out_dtype", C.c_int32), ("accumulate", C.c_int32),
        ("nphase", C.c_int32), ("reserved0", C.c_int32), ("phase_w_off", C.c_int64), ("phase_out_off", C.c_int64),
        ("stats", C.c_void_p),
        ("out2", C.c_void_p), ("out2_halo", C.c_int32), ("out2_layout", C.c_int32), ("out2_act", C.c_int32),
        ("reserved1", C.c_int32),
    ]


class WGrad(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("backend", C.c_int32),
        ("a", C.c_void_p), ("a_dim", C.c_int64 * 5), ("a_str", C.c_int64 * 5),
        ("b", C.c_void_p), ("b_dim", C.c_int64 * 5), ("b_str", C.c_int64 * 5),
        ("box", C.c_int32 * 3), ("tiles", C.c_int32 * 3),
        ("ntaps", C.c_int32),
        ("taps", C.POINTER(C.c_int32)),
        ("ca", C.c_int32), ("cb", C.c_int32),
        ("dw", C.c_void_p), ("s_a", C.c_int64), ("s_t", C.c_int64), ("s_b", C.c_int64),
        ("dbias", C.c_void_p),
        ("accumulate", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("remap_axis", C.c_int32), ("remap_div", C.c_int32), ("remap_lo_limit", C.c_int32), ("remap_hi_limit", C.c_int32),
        ("remap_hi_stride", C.c_int64), ("remap_lo_stride", C.c_int64),
    ]


class HBuf(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("halo", C.c_int32), ("layout", C.c_int32), ("dtype", C.c_int32)]


class PackEntry(C.Structure):
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("cout", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("cin", C.c_int32), ("mode", C.c_int32), ("rows_padded", C.c_int32), ("out_dtype", C.c_int32),
                ("reserved", C.c_int32), ("total", C.c_int64)]


EXPORTS = [
    "dwc_last_error", "dwc_abi_version", "dwc_tc_available", "dwc_gconv", "dwc_wgrad_workspace_bytes", "dwc_wgrad",
    "dwc_nc_stats", "dwc_norm_finalize", "dwc_post_fwd", "dwc_post_bwd_reduce", "dwc_post_bwd_reduce_folds", "dwc_post_bwd_reduce_can_fold", "dwc_norm_bwd_finalize",
    "dwc_post_bwd_apply", "dwc_post_fwd_norm", "dwc_post_bwd_apply_norm", "dwc_fold_halo", "dwc_post_fused_ok", "dwc_post_fused_fwd", "dwc_post_fused_bwd", "dwc_upsample_pad_fwd", "dwc_upsample_pad_bwd", "dwc_image_pad_fwd", "dwc_image_pad_bwd",
    "dwc_heads_fwd", "dwc_heads_bwd", "dwc_image_rows_fwd", "dwc_heads_bwd_rows", "dwc_blend_fwd", "dwc_blend_bwd", "dwc_relu_gap_fwd", "dwc_relu_gap_bwd",
    "dwc_sgemm", "dwc_sgemm_ws", "dwc_sgemm_workspace_bytes", "dwc_gemm_tf32", "dwc_gemm_tf32_ok", "dwc_set_tf32", "dwc_get_tf32", "dwc_colsum", "dwc_relu_bwd", "dwc_mul", "dwc_embed_concat_fwd", "dwc_embed_concat_bwd",
    "dwc_lstm_workspace_bytes", "dwc_lstm_layer_fwd", "dwc_lstm_layer_bwd", "dwc_transpose", "dwc_gmm_sample", "dwc_gmm_kl", "dwc_l1_loss_fwd", "dwc_l1_loss_bwd", "dwc_adv_loss_fwd", "dwc_adv_loss_bwd",
    "dwc_mse_const_loss_fwd", "dwc_mse_const_loss_bwd", "dwc_bce_logits_loss_fwd", "dwc_bce_logits_loss_bwd",
    "dwc_adam_step", "dwc_ema_step", "dwc_pack_weights", "dwc_pack_weights_batch", "dwc_conv7_few", "dwc_cast", "dwc_fill",
]


def lib():
    """Load (once) and return the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "dwc_gan_b200: %s is missing - build it with `python -m dwc_gan_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.dwc_last_error.restype = C.c_char_p
        _lib.dwc_wgrad_workspace_bytes.restype = C.c_int64
        _lib.dwc_lstm_workspace_bytes.restype = C.c_int64
        _lib.dwc_sgemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_int, C.c_void_p]
        _lib.dwc_sgemm_ws.argtypes = list(_lib.dwc_sgemm.argtypes[:-1]) + [C.c_void_p, C.c_int64, C.c_void_p]
        _lib.dwc_sgemm_workspace_bytes.restype = C.c_int64
        if hasattr(_lib, "dwc_post_bwd_cluster"):          # parked kernel, only in a DWC_EXPERIMENTAL=1 build
            _lib.dwc_post_bwd_cluster_ok.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            _lib.dwc_post_bwd_cluster.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.dwc_gemm_tf32.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                       C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
        _lib.dwc_gemm_tf32_ok.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_int64, C.c_void_p, C.c_int64, C.c_int64]
        _lib.dwc_set_tf32.argtypes = [C.c_int]
        _lib.dwc_set_tf32.restype = None
        _lib.dwc_colsum.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
        _lib.dwc_conv7_few.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        _lib.dwc_post_fwd_norm.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.dwc_post_bwd_apply_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                 C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.dwc_norm_finalize.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.dwc_post_fused_fwd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.dwc_norm_bwd_finalize.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.dwc_gmm_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]
        _lib.dwc_gmm_kl.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.dwc_mse_const_loss_fwd.argtypes = [C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.dwc_mse_const_loss_bwd.argtypes = [C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.dwc_ema_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
        _lib.dwc_fill.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_void_p]
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("dwc_b200 %s failed: %s" % (what, lib().dwc_last_error().decode()))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t_or_dtype):
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise TypeError("unsupported dtype %s" % d)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def i64(v):
    return C.c_int64(int(v))


def i32(v):
    return C.c_int(int(v))


def f32(v):
    return C.c_float(float(v))
