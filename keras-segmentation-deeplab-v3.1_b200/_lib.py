"""ctypes binding of libdeeplab_b200.so (include/deeplab_b200.h).

The product path has NO CPU fallback: every op below raises if the CUDA library is missing or the call fails.
PyTorch tensors are only the buffer type (data_ptr + shapes); all arithmetic happens in the sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdeeplab_b200.so")

F16, BF16, F32 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
_DT = {torch.float16: F16, torch.bfloat16: BF16, torch.float32: F32}
TORCH_DT = {F16: torch.float16, BF16: torch.bfloat16, F32: torch.float32}

vp, i32, i64, f32, f64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64


class BnFin(C.Structure):
    """dlb_bn_fin: consumer-side BatchNorm finalisation (include/deeplab_b200.h)."""
    _fields_ = [
        ("sum", vp), ("sqs", vp), ("gamma", vp), ("beta", vp), ("eps", f32), ("momentum", f32), ("count", C.c_double),
        ("moving_mean", vp), ("moving_var", vp), ("scale", vp), ("shift", vp), ("mean", vp), ("rstd", vp),
    ]


class PwGemmParams(C.Structure):
    _fields_ = [
        ("M", i32), ("N", i32), ("K", i32), ("dtype", i32), ("out_dtype", i32),
        ("A", vp), ("lda", i32), ("Bt", vp), ("ldb", i32), ("C", vp), ("ldc", i32), ("n_store", i32),
        ("col_scale", vp), ("col_shift", vp), ("row_bias", vp), ("rows_per_img", i32), ("ld_row_bias", i32),
        ("act", i32), ("R", vp), ("ldr", i32), ("stat_sum", vp), ("stat_sqs", vp),
        ("shuffle_r", i32), ("shuffle_h", i32), ("shuffle_w", i32),
        ("a_scale", vp), ("a_shift", vp), ("a_act", i32), ("Bt_lo", vp), ("a_fin", C.POINTER(BnFin)),
    ]


class PwWgradParams(C.Structure):
    _fields_ = [
        ("M", i32), ("N", i32), ("K", i32), ("dtype", i32), ("A", vp), ("lda", i32), ("dY", vp), ("ldy", i32),
        ("dW", vp), ("ldw", i32), ("dbias", vp), ("beta", f32), ("workspace", vp), ("workspace_bytes", i64),
        ("a_scale", vp), ("a_shift", vp), ("a_act", i32),
    ]


class DwConvParams(C.Structure):
    _fields_ = [
        ("B", i32), ("H", i32), ("W", i32), ("C", i32), ("Ho", i32), ("Wo", i32),
        ("stride", i32), ("dilation", i32), ("pad_top", i32), ("pad_left", i32), ("dtype", i32),
        ("x", vp), ("y", vp), ("w", vp), ("in_scale", vp), ("in_shift", vp), ("in_act", i32),
        ("out_scale", vp), ("out_shift", vp), ("out_act", i32), ("stat_sum", vp), ("stat_sqs", vp),
        ("in_fin", C.POINTER(BnFin)),
    ]


class DwConvBwdParams(C.Structure):
    _fields_ = [
        ("B", i32), ("H", i32), ("W", i32), ("C", i32), ("Ho", i32), ("Wo", i32),
        ("stride", i32), ("dilation", i32), ("pad_top", i32), ("pad_left", i32), ("dtype", i32),
        ("x", vp), ("dy", vp), ("dx", vp), ("w", vp), ("dw", vp),
        ("in_scale", vp), ("in_shift", vp), ("in_act", i32),
    ]


class StemConvParams(C.Structure):
    _fields_ = [
        ("B", i32), ("H", i32), ("W", i32), ("Cout", i32), ("Ho", i32), ("Wo", i32), ("dtype", i32),
        ("x", vp), ("y", vp), ("w", vp), ("out_scale", vp), ("out_shift", vp), ("out_act", i32),
        ("stat_sum", vp), ("stat_sqs", vp),
    ]


class BnApplyParams(C.Structure):
    _fields_ = [
        ("M", i64), ("C", i32), ("dtype", i32), ("x", vp), ("y", vp), ("res", vp),
        ("scale", vp), ("shift", vp), ("act", i32), ("drop_rate", f32), ("drop_seed", u64), ("drop_seed_dev", vp),
        ("fin", C.POINTER(BnFin)),
    ]


class BnBwdParams(C.Structure):
    _fields_ = [
        ("M", i64), ("C", i32), ("dtype", i32), ("x", vp), ("da", vp), ("dx", vp),
        ("scale", vp), ("shift", vp), ("mean", vp), ("rstd", vp), ("act", i32),
        ("red", vp), ("dgamma", vp), ("dbeta", vp), ("drop_rate", f32), ("drop_seed", u64), ("frozen_stats", i32),
        ("drop_seed_dev", vp),
    ]


class SoftmaxCeParams(C.Structure):
    _fields_ = [
        ("B", i32), ("h", i32), ("w", i32), ("C", i32), ("ldl", i32), ("H", i32), ("W", i32),
        ("logits", vp), ("labels", vp), ("sample_w", vp), ("grad_scale_dev", vp),
        ("dlogits", vp), ("loss_sum", vp), ("wcount", vp), ("argmax", vp),
    ]


class SepconvFusedParams(C.Structure):
    _fields_ = [
        ("B", i32), ("H", i32), ("W", i32), ("C", i32), ("N", i32), ("dtype", i32), ("n_branches", i32),
        ("rates", i32 * 4), ("x", vp), ("w_pw", vp * 4), ("dw_pack", vp), ("pw_scale", vp * 4), ("pw_shift", vp * 4),
        ("out", vp * 4), ("ldc", i32), ("dw_act", i32), ("pw_act", i32), ("res", vp * 4), ("ldr", i32),
    ]


class AugParams(C.Structure):
    _fields_ = [("hflip", i32), ("vflip", i32), ("blur_ksize", i32), ("warp", i32), ("minv", f64 * 6)]


class CrfConfig(C.Structure):
    _fields_ = [
        ("H", i32), ("W", i32), ("M", i32), ("iters", i32), ("sxy_gauss", f32), ("compat_gauss", f32),
        ("sxy_bilat", f32), ("srgb_bilat", f32), ("compat_bilat", f32), ("unary_layout", i32),
    ]


# every symbol include/deeplab_b200.h declares (tests/test_abi.py checks the .so exports all of them)
EXPORTS = [
    "dlb_version", "dlb_last_error", "dlb_launch_count", "dlb_device_ok", "dlb_pw_gemm", "dlb_pw_wgrad",
    "dlb_pw_wgrad_workspace_bytes", "dlb_dw_conv_fwd", "dlb_dw_conv_bwd", "dlb_stem_conv_fwd",
    "dlb_stem_conv_wgrad", "dlb_bn_finalize", "dlb_bn_fold", "dlb_bn_act_apply", "dlb_bn_bwd_reduce",
    "dlb_bn_bwd_apply", "dlb_global_avgpool_fwd", "dlb_global_avgpool_bwd", "dlb_small_gemm",
    "dlb_resize_softmax_fwd", "dlb_resize_softmax_ce", "dlb_ce_grad_scale", "dlb_phase_shift", "dlb_adam_step",
    "dlb_cast_weight", "dlb_cast_weights_batched", "dlb_cast", "dlb_fill_zero", "dlb_confusion", "dlb_crf_workspace_bytes",
    "dlb_crf_inference", "dlb_conv3x3_fwd", "dlb_subsample", "dlb_resize_bilinear", "dlb_aspp_dw3_fwd",
    "dlb_sepconv_fused_fwd", "dlb_sepconv_pack_bytes", "dlb_sepconv_pack_dw", "dlb_pw_gemm_plan", "dlb_label_weights",
    "dlb_grad_finite_check", "dlb_crf_workspace_bytes_batched", "dlb_crf_inference_batched", "dlb_augment_batch", "dlb_subpixel_grad_gather",
]

_lib = None


def lib() -> C.CDLL:
    """Load the C-ABI library; fails loudly (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "deeplab_b200 has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        L.dlb_last_error.restype = C.c_char_p
        L.dlb_launch_count.restype = i64
        L.dlb_pw_wgrad_workspace_bytes.restype = i64
        L.dlb_crf_workspace_bytes.restype = i64
        L.dlb_bn_finalize.argtypes = [i32, f64, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp, vp, i32, vp]
        L.dlb_bn_fold.argtypes = [i32, vp, vp, vp, vp, f32, vp, vp, vp]
        L.dlb_global_avgpool_fwd.argtypes = [i32, i32, i32, i32, vp, vp, vp, i32, vp, vp]
        L.dlb_global_avgpool_bwd.argtypes = [i32, i32, i32, i32, vp, vp, i32, vp]
        L.dlb_small_gemm.argtypes = [i32, i32, i32, vp, i32, i32, vp, i32, i32, vp, i32, f32, f32, vp]
        L.dlb_resize_softmax_fwd.argtypes = [i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
        L.dlb_ce_grad_scale.argtypes = [i64, vp, vp, vp, f32, vp, vp]
        L.dlb_phase_shift.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp, i32, vp]
        L.dlb_adam_step.argtypes = [i64, vp, vp, vp, vp, vp, f32, f32, f32, f32, f32, f32, vp, vp, vp]
        L.dlb_grad_finite_check.argtypes = [i64, vp, vp, vp]
        L.dlb_cast_weight.argtypes = [i32, i32, vp, i32, vp, vp, vp]
        L.dlb_cast.argtypes = [i64, i32, vp, i32, vp, vp]
        L.dlb_cast_weights_batched.argtypes = [i32, vp, i64, vp]
        L.dlb_fill_zero.argtypes = [vp, i64, vp]
        L.dlb_confusion.argtypes = [i32, i64, i32, vp, vp, vp, vp]
        L.dlb_label_weights.argtypes = [i32, i64, i32, i32, vp, vp, vp, vp, vp]
        L.dlb_stem_conv_wgrad.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp]
        L.dlb_pw_wgrad_workspace_bytes.argtypes = [i32, i32, i32]
        L.dlb_conv3x3_fwd.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]
        L.dlb_subsample.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp, vp]
        L.dlb_resize_bilinear.argtypes = [i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
        L.dlb_aspp_dw3_fwd.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.dlb_sepconv_pack_bytes.restype = i64
        L.dlb_sepconv_pack_bytes.argtypes = [i32, i32]
        L.dlb_sepconv_pack_dw.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp]
        L.dlb_sepconv_fused_fwd.argtypes = [vp, vp]
        L.dlb_crf_workspace_bytes.argtypes = [vp]
        L.dlb_crf_inference.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp]
        L.dlb_subpixel_grad_gather.argtypes = [i32, i32, i32, i32, i32, vp, i32, vp, vp]
        L.dlb_augment_batch.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.dlb_crf_workspace_bytes_batched.restype = i64
        L.dlb_crf_workspace_bytes_batched.argtypes = [vp, i32]
        L.dlb_crf_inference_batched.argtypes = [vp, i32, vp, vp, vp, vp, vp, i64, vp]
        for name in ("dlb_pw_gemm", "dlb_pw_wgrad", "dlb_dw_conv_fwd", "dlb_dw_conv_bwd", "dlb_stem_conv_fwd",
                     "dlb_bn_act_apply", "dlb_bn_bwd_reduce", "dlb_bn_bwd_apply", "dlb_resize_softmax_ce"):
            getattr(L, name).argtypes = [vp, vp]
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"libdeeplab_b200 {what} failed ({rc}): {lib().dlb_last_error().decode()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def dt(t: torch.Tensor) -> int:
    return _DT[t.dtype]


def require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("deeplab_b200 ops need CUDA tensors (no CPU fallback)")


def launch_count() -> int:
    return int(lib().dlb_launch_count())
