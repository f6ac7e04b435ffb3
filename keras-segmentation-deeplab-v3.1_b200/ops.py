"""Thin tensor-level wrappers over the C ABI (one Python function per exported op).

Tensors are NHWC; a 1x1 convolution sees the [B*H*W, C] matrix view.  Nothing here computes on the host --
each function marshals pointers/sizes into the param struct of include/deeplab_b200.h and launches on the
current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L
from ._lib import ACT_NONE, ACT_RELU, ACT_RELU6, F16, BF16, F32  # noqa: F401


def tf_same_pad(in_size: int, k: int, stride: int, dilation: int) -> Tuple[int, int, int]:
    """TF 'SAME' geometry: returns (out_size, pad_before, pad_after)."""
    out = -(-in_size // stride)
    k_eff = (k - 1) * dilation + 1
    total = max((out - 1) * stride + k_eff - in_size, 0)
    return out, total // 2, total - total // 2


def _mat(t: torch.Tensor) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a tensor viewed as a row-major matrix over its last dim."""
    assert t.stride(-1) == 1
    cols = t.shape[-1]
    rows = t.numel() // cols
    ld = t.stride(-2) if t.dim() >= 2 else cols
    return rows, cols, ld


def pw_gemm(A: torch.Tensor, Bt: torch.Tensor, out: torch.Tensor, *, K: Optional[int] = None, N: Optional[int] = None,
            n_store: Optional[int] = None, col_scale=None, col_shift=None, row_bias=None, rows_per_img: int = 1,
            act: int = ACT_NONE, residual=None, stat_sum=None, stat_sqs=None, shuffle: Optional[Tuple[int, int, int]] = None,
            a_scale=None, a_shift=None, a_act: int = ACT_NONE, a_fin=None):
    """out[M, :n_store] = epilogue(A'[M, K] @ Bt[N, K]^T), A' = a_act(A * a_scale + a_shift) if given; see dlb_pw_gemm.
    a_fin (a `bn_fin(...)`) instead of a_scale / a_shift: the kernel finalises the producing BatchNorm itself.
    fp32 inference weights may be passed pre-split as a tuple (hi, lo) from `f32_split` (3xTF32, see dlb_pw_gemm_params)."""
    Bt_lo = None
    if isinstance(Bt, (tuple, list)):
        Bt, Bt_lo = Bt
    L.require_cuda(A, Bt, out)
    M, Ka, lda = _mat(A)
    Nb, Kb, ldb = _mat(Bt)
    K = K if K is not None else Ka
    N = N if N is not None else Nb
    p = L.PwGemmParams()
    p.M, p.N, p.K = M, N, K
    p.dtype, p.out_dtype = L.dt(A), L.dt(out)
    p.A, p.lda, p.Bt, p.ldb = A.data_ptr(), lda, Bt.data_ptr(), ldb
    if shuffle is None:
        _, oc, ldc = _mat(out)
        p.C, p.ldc = out.data_ptr(), ldc
        p.n_store = n_store if n_store is not None else oc
    else:
        p.C, p.ldc = out.data_ptr(), 0
        p.n_store = N
        p.shuffle_r, p.shuffle_h, p.shuffle_w = shuffle
    p.col_scale, p.col_shift = L.ptr(col_scale), L.ptr(col_shift)
    p.row_bias = L.ptr(row_bias)
    p.rows_per_img = rows_per_img
    p.ld_row_bias = row_bias.stride(0) if row_bias is not None else 0
    p.act = act
    if residual is not None:
        p.R, p.ldr = residual.data_ptr(), _mat(residual)[2]
    p.stat_sum, p.stat_sqs = L.ptr(stat_sum), L.ptr(stat_sqs)
    p.a_scale, p.a_shift, p.a_act = L.ptr(a_scale), L.ptr(a_shift), a_act
    p.Bt_lo = L.ptr(Bt_lo)
    if a_fin is not None:
        p.a_fin = C.pointer(a_fin)
    L.check(L.lib().dlb_pw_gemm(C.byref(p), L.stream_ptr()), "pw_gemm")
    return out


def f32_split(w: torch.Tensor):
    """fp32 tensor -> (hi, lo) with hi = w with the 13 low mantissa bits cleared (exactly representable in tf32) and
    lo = w - hi (exact).  Host-side preparation of constant inference weights for the 3xTF32 GEMM."""
    hi = (w.contiguous().view(torch.int32) & -8192).view(torch.float32)
    return hi, (w - hi).contiguous()


def pw_wgrad_workspace_bytes(M: int, N: int, K: int) -> int:
    return int(L.lib().dlb_pw_wgrad_workspace_bytes(M, N, K))


def pw_wgrad(A: torch.Tensor, dY: torch.Tensor, dW: torch.Tensor, *, K: Optional[int] = None, N: Optional[int] = None,
             dbias=None, beta: float = 0.0, workspace: Optional[torch.Tensor] = None, a_scale=None, a_shift=None,
             a_act: int = ACT_NONE):
    """dW[K, N] = A'[M, K]^T @ dY[M, N], A' = a_act(A * a_scale + a_shift) if given; see dlb_pw_wgrad."""
    L.require_cuda(A, dY, dW)
    M, Ka, lda = _mat(A)
    _, Nb, ldy = _mat(dY)
    p = L.PwWgradParams()
    p.M, p.K, p.N = M, (K if K is not None else Ka), (N if N is not None else Nb)
    p.dtype = L.dt(A)
    p.A, p.lda, p.dY, p.ldy = A.data_ptr(), lda, dY.data_ptr(), ldy
    p.dW, p.ldw = dW.data_ptr(), p.N
    p.dbias = L.ptr(dbias)
    p.beta = beta
    if workspace is not None:
        p.workspace, p.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    p.a_scale, p.a_shift, p.a_act = L.ptr(a_scale), L.ptr(a_shift), a_act
    L.check(L.lib().dlb_pw_wgrad(C.byref(p), L.stream_ptr()), "pw_wgrad")
    return dW


def dw_conv_fwd(x: torch.Tensor, w: torch.Tensor, y: torch.Tensor, *, stride: int, dilation: int, pad_top: int,
                pad_left: int, in_scale=None, in_shift=None, in_act: int = ACT_NONE, out_scale=None, out_shift=None,
                out_act: int = ACT_NONE, stat_sum=None, stat_sqs=None, in_fin=None):
    L.require_cuda(x, w, y)
    p = L.DwConvParams()
    p.B, p.H, p.W, p.C = x.shape
    p.Ho, p.Wo = y.shape[1], y.shape[2]
    p.stride, p.dilation, p.pad_top, p.pad_left = stride, dilation, pad_top, pad_left
    p.dtype = L.dt(x)
    p.x, p.y, p.w = x.data_ptr(), y.data_ptr(), w.data_ptr()
    p.in_scale, p.in_shift, p.in_act = L.ptr(in_scale), L.ptr(in_shift), in_act
    p.out_scale, p.out_shift, p.out_act = L.ptr(out_scale), L.ptr(out_shift), out_act
    p.stat_sum, p.stat_sqs = L.ptr(stat_sum), L.ptr(stat_sqs)
    if in_fin is not None:
        p.in_fin = C.pointer(in_fin)
    L.check(L.lib().dlb_dw_conv_fwd(C.byref(p), L.stream_ptr()), "dw_conv_fwd")
    return y


def dw_conv_bwd(x: Optional[torch.Tensor], dy: torch.Tensor, w: torch.Tensor, *, dx=None, dw=None, in_shape=None,
                stride: int, dilation: int, pad_top: int, pad_left: int, in_scale=None, in_shift=None,
                in_act: int = ACT_NONE):
    L.require_cuda(dy, w)
    p = L.DwConvBwdParams()
    shp = in_shape if in_shape is not None else (x.shape if x is not None else dx.shape)
    p.B, p.H, p.W, p.C = shp
    p.Ho, p.Wo = dy.shape[1], dy.shape[2]
    p.stride, p.dilation, p.pad_top, p.pad_left = stride, dilation, pad_top, pad_left
    p.dtype = L.dt(dy)
    p.x, p.dy, p.dx, p.w, p.dw = L.ptr(x), dy.data_ptr(), L.ptr(dx), w.data_ptr(), L.ptr(dw)
    p.in_scale, p.in_shift, p.in_act = L.ptr(in_scale), L.ptr(in_shift), in_act
    L.check(L.lib().dlb_dw_conv_bwd(C.byref(p), L.stream_ptr()), "dw_conv_bwd")


def stem_conv_fwd(x: torch.Tensor, w: torch.Tensor, y: torch.Tensor, *, out_scale=None, out_shift=None,
                  out_act: int = ACT_NONE, stat_sum=None, stat_sqs=None):
    L.require_cuda(x, w, y)
    assert x.dtype == torch.float32 and x.shape[-1] == 3
    p = L.StemConvParams()
    p.B, p.H, p.W = x.shape[0], x.shape[1], x.shape[2]
    p.Cout, p.Ho, p.Wo = y.shape[3], y.shape[1], y.shape[2]
    p.dtype = L.dt(y)
    p.x, p.y, p.w = x.data_ptr(), y.data_ptr(), w.data_ptr()
    p.out_scale, p.out_shift, p.out_act = L.ptr(out_scale), L.ptr(out_shift), out_act
    p.stat_sum, p.stat_sqs = L.ptr(stat_sum), L.ptr(stat_sqs)
    L.check(L.lib().dlb_stem_conv_fwd(C.byref(p), L.stream_ptr()), "stem_conv_fwd")
    return y


def stem_conv_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor):
    L.require_cuda(x, dy, dw)
    B, H, W, _ = x.shape
    L.check(L.lib().dlb_stem_conv_wgrad(B, H, W, dy.shape[3], L.dt(dy), x.data_ptr(), dy.data_ptr(), dw.data_ptr(),
                                        L.stream_ptr()), "stem_conv_wgrad")


def bn_finalize(count: float, s: torch.Tensor, q: torch.Tensor, gamma, beta, eps: float, momentum: float,
                moving_mean, moving_var, scale, shift, mean=None, rstd=None, reset: bool = True):
    L.check(L.lib().dlb_bn_finalize(s.numel(), float(count), s.data_ptr(), q.data_ptr(), gamma.data_ptr(),
                                    beta.data_ptr(), eps, momentum, L.ptr(moving_mean), L.ptr(moving_var),
                                    scale.data_ptr(), shift.data_ptr(), L.ptr(mean), L.ptr(rstd), int(reset),
                                    L.stream_ptr()), "bn_finalize")


def bn_fin(count: float, s: torch.Tensor, q: torch.Tensor, gamma, beta, eps: float, momentum: float, moving_mean,
           moving_var, scale, shift, mean=None, rstd=None):
    """dlb_bn_fin for the `fin=` / `a_fin=` / `in_fin=` arguments: the consuming kernel finalises the BatchNorm from the
    fp64 sums in its prologue (same arithmetic as `bn_finalize`, which this replaces launch for launch).  The sums are
    not cleared: zero them once per step."""
    f = L.BnFin()
    f.sum, f.sqs, f.gamma, f.beta = s.data_ptr(), q.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    f.eps, f.momentum, f.count = eps, momentum, float(count)
    f.moving_mean, f.moving_var = L.ptr(moving_mean), L.ptr(moving_var)
    f.scale, f.shift, f.mean, f.rstd = scale.data_ptr(), shift.data_ptr(), L.ptr(mean), L.ptr(rstd)
    return f


def bn_fold(gamma, beta, moving_mean, moving_var, eps: float, scale, shift):
    L.check(L.lib().dlb_bn_fold(gamma.numel(), gamma.data_ptr(), beta.data_ptr(), moving_mean.data_ptr(),
                                moving_var.data_ptr(), eps, scale.data_ptr(), shift.data_ptr(), L.stream_ptr()),
            "bn_fold")


def bn_act_apply(x: torch.Tensor, y: torch.Tensor, *, scale=None, shift=None, act: int = ACT_NONE, res=None,
                 drop_rate: float = 0.0, drop_seed: int = 0, drop_seed_dev=None, fin=None):
    L.require_cuda(x, y)
    p = L.BnApplyParams()
    p.C = x.shape[-1]
    p.M = x.numel() // p.C
    p.dtype = L.dt(x)
    p.x, p.y, p.res = x.data_ptr(), y.data_ptr(), L.ptr(res)
    p.scale, p.shift, p.act = L.ptr(scale), L.ptr(shift), act
    p.drop_rate, p.drop_seed, p.drop_seed_dev = drop_rate, drop_seed, L.ptr(drop_seed_dev)
    if fin is not None:
        p.fin = C.pointer(fin)
    L.check(L.lib().dlb_bn_act_apply(C.byref(p), L.stream_ptr()), "bn_act_apply")
    return y


def _bn_bwd_params(x, da, dx, scale, shift, mean, rstd, act, red, dgamma, dbeta, drop_rate, drop_seed, frozen,
                   drop_seed_dev=None):
    p = L.BnBwdParams()
    p.C = x.shape[-1]
    p.M = x.numel() // p.C
    p.dtype = L.dt(x)
    p.x, p.da, p.dx = x.data_ptr(), da.data_ptr(), L.ptr(dx)
    p.scale, p.shift, p.mean, p.rstd, p.act = scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), act
    p.red, p.dgamma, p.dbeta = red.data_ptr(), L.ptr(dgamma), L.ptr(dbeta)
    p.drop_rate, p.drop_seed, p.frozen_stats = drop_rate, drop_seed, int(frozen)
    p.drop_seed_dev = L.ptr(drop_seed_dev)
    return p


def bn_bwd(x, da, dx, *, scale, shift, mean, rstd, act, red, dgamma=None, dbeta=None, drop_rate=0.0, drop_seed=0,
           frozen=False, drop_seed_dev=None):
    """Both passes of the BatchNorm(+activation, +dropout) backward; `red` ([2C] fp64) must be zero on entry."""
    L.require_cuda(x, da, dx)
    p = _bn_bwd_params(x, da, dx, scale, shift, mean, rstd, act, red, dgamma, dbeta, drop_rate, drop_seed, frozen,
                       drop_seed_dev)
    L.check(L.lib().dlb_bn_bwd_reduce(C.byref(p), L.stream_ptr()), "bn_bwd_reduce")
    L.check(L.lib().dlb_bn_bwd_apply(C.byref(p), L.stream_ptr()), "bn_bwd_apply")
    return dx


def global_avgpool_fwd(x: torch.Tensor, out: torch.Tensor, *, in_scale=None, in_shift=None, in_act=ACT_NONE):
    B, C_ = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * C_)
    L.check(L.lib().dlb_global_avgpool_fwd(B, HW, C_, L.dt(x), x.data_ptr(), L.ptr(in_scale), L.ptr(in_shift),
                                           in_act, out.data_ptr(), L.stream_ptr()), "global_avgpool_fwd")
    return out


def global_avgpool_bwd(dout: torch.Tensor, dx: torch.Tensor, accumulate: bool):
    B, C_ = dx.shape[0], dx.shape[-1]
    HW = dx.numel() // (B * C_)
    L.check(L.lib().dlb_global_avgpool_bwd(B, HW, C_, L.dt(dx), dout.data_ptr(), dx.data_ptr(), int(accumulate),
                                           L.stream_ptr()), "global_avgpool_bwd")


def small_gemm(A, B, out, *, M, N, K, transA=False, transB=False, alpha=1.0, beta=0.0):
    L.check(L.lib().dlb_small_gemm(M, N, K, A.data_ptr(), A.stride(0), int(transA), B.data_ptr(), B.stride(0),
                                   int(transB), out.data_ptr(), out.stride(0), alpha, beta, L.stream_ptr()),
            "small_gemm")
    return out


def resize_softmax_fwd(logits: torch.Tensor, C_: int, H: int, W: int, probs=None, argmax=None):
    B, h, w, ldl = logits.shape
    L.check(L.lib().dlb_resize_softmax_fwd(B, h, w, C_, ldl, H, W, logits.data_ptr(), L.ptr(probs), L.ptr(argmax),
                                           L.stream_ptr()), "resize_softmax_fwd")


def resize_softmax_ce(logits, C_, H, W, labels, sample_w, grad_scale, dlogits, loss_sum, wcount, argmax=None):
    B, h, w, ldl = logits.shape
    p = L.SoftmaxCeParams()
    p.B, p.h, p.w, p.C, p.ldl, p.H, p.W = B, h, w, C_, ldl, H, W
    p.logits, p.labels, p.sample_w = logits.data_ptr(), labels.data_ptr(), L.ptr(sample_w)
    p.grad_scale_dev, p.dlogits = grad_scale.data_ptr(), dlogits.data_ptr()
    p.loss_sum, p.wcount, p.argmax = loss_sum.data_ptr(), L.ptr(wcount), L.ptr(argmax)
    L.check(L.lib().dlb_resize_softmax_ce(C.byref(p), L.stream_ptr()), "resize_softmax_ce")


def ce_grad_scale(n: int, sample_w, grad_scale, wcount, loss_scale: float = 1.0, loss_scale_state=None):
    L.check(L.lib().dlb_ce_grad_scale(n, L.ptr(sample_w), grad_scale.data_ptr(), wcount.data_ptr(), float(loss_scale),
                                      L.ptr(loss_scale_state), L.stream_ptr()), "ce_grad_scale")


def phase_shift(x: torch.Tensor, out: torch.Tensor, r: int, inverse: bool = False):
    """forward: x [B,h,w,Cs*r*r] -> out [B,h*r,w*r,Cs]; inverse: x [B,h*r,w*r,Cs] -> out [B,h,w,Cs*r*r]."""
    lo = out if inverse else x
    B, h, w, c = lo.shape
    L.check(L.lib().dlb_phase_shift(B, h, w, c // (r * r), r, L.dt(x), x.data_ptr(), out.data_ptr(), int(inverse),
                                    L.stream_ptr()), "phase_shift")
    return out


def subpixel_grad_gather(dlogits: torch.Tensor, dst: torch.Tensor, h: int, w: int, r: int):
    """[B, h*r, w*r, Cs] fp32 gradient -> [B, h, w, (jj, i, k)] in dst's dtype (inverse of the fused Subpixel store)."""
    B, Cs = dlogits.shape[0], dlogits.shape[-1]
    L.check(L.lib().dlb_subpixel_grad_gather(B, h, w, Cs, r, dlogits.data_ptr(), L.dt(dst), dst.data_ptr(), L.stream_ptr()),
            "subpixel_grad_gather")
    return dst


def adam_step(param, grad, m, v, step_dev, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, decay=0.0, grad_mult=1.0,
              train_mask=None, loss_scale_state=None):
    L.check(L.lib().dlb_adam_step(param.numel(), param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(),
                                  step_dev.data_ptr(), lr, beta1, beta2, eps, decay, grad_mult, L.ptr(train_mask),
                                  L.ptr(loss_scale_state), L.stream_ptr()), "adam_step")


def grad_finite_check(grad, loss_scale_state):
    """loss_scale_state[2] = 1 if any gradient is inf / NaN (dynamic fp16 loss scaling, see dlb_adam_step)."""
    L.check(L.lib().dlb_grad_finite_check(grad.numel(), grad.data_ptr(), loss_scale_state.data_ptr(), L.stream_ptr()),
            "grad_finite_check")


def cast_weight(w: torch.Tensor, K: int, N: int, w_kn=None, w_nk=None):
    t = w_kn if w_kn is not None else w_nk
    L.check(L.lib().dlb_cast_weight(K, N, w.data_ptr(), L.dt(t), L.ptr(w_kn), L.ptr(w_nk), L.stream_ptr()),
            "cast_weight")


def cast_weights_batched(table: torch.Tensor, total: int):
    """All 1x1-layer weight copies in one launch; `table` = device int64 [n, 8] (see dlb_cast_weights_batched)."""
    L.check(L.lib().dlb_cast_weights_batched(table.shape[0], table.data_ptr(), total, L.stream_ptr()),
            "cast_weights_batched")


def cast(src: torch.Tensor, dst: torch.Tensor):
    L.check(L.lib().dlb_cast(src.numel(), L.dt(src), src.data_ptr(), L.dt(dst), dst.data_ptr(), L.stream_ptr()), "cast")
    return dst


def fill_zero(t: torch.Tensor):
    L.check(L.lib().dlb_fill_zero(t.data_ptr(), t.numel() * t.element_size(), L.stream_ptr()), "fill_zero")


def confusion(labels: torch.Tensor, argmax: torch.Tensor, C_: int, conf: torch.Tensor):
    B = argmax.shape[0]
    npix = argmax.numel() // B
    L.check(L.lib().dlb_confusion(B, npix, C_, labels.data_ptr(), argmax.data_ptr(), conf.data_ptr(), L.stream_ptr()),
            "confusion")


def label_weights(labels: torch.Tensor, n_classes: int, y: Optional[torch.Tensor] = None,
                  sw: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None):
    """Generator label contract on the device (reference utils.py:360-399): void remap into y [B, P] f32 and the
    per-image balanced class weights sw [B, P] f32; see dlb_label_weights.  labels: [B, P] uint8 / int32 / float32."""
    L.require_cuda(labels)
    lt = {torch.uint8: 0, torch.int32: 1, torch.float32: 2}[labels.dtype]
    B = labels.shape[0]
    P = labels.numel() // B
    if counts is None:
        counts = torch.empty(B, n_classes + 1, device=labels.device, dtype=torch.int64)
    L.check(L.lib().dlb_label_weights(B, P, n_classes, lt, labels.data_ptr(), counts.data_ptr(), L.ptr(y), L.ptr(sw),
                                      L.stream_ptr()), "label_weights")
    return y, sw, counts


def conv3x3_fwd(x: torch.Tensor, w: torch.Tensor, y: torch.Tensor, *, out_scale=None, out_shift=None, out_act=ACT_NONE):
    B, H, W_, Cin = x.shape
    L.check(L.lib().dlb_conv3x3_fwd(B, H, W_, Cin, y.shape[-1], L.dt(x), x.data_ptr(), w.data_ptr(), y.data_ptr(),
                                    L.ptr(out_scale), L.ptr(out_shift), out_act, L.stream_ptr()), "conv3x3_fwd")
    return y


def subsample(x: torch.Tensor, y: torch.Tensor, step: int = 2):
    B, H, W_, C_ = x.shape
    L.check(L.lib().dlb_subsample(B, H, W_, C_, step, L.dt(x), x.data_ptr(), y.data_ptr(), L.stream_ptr()), "subsample")
    return y


def resize_bilinear(x: torch.Tensor, y: torch.Tensor, C_: Optional[int] = None):
    """x [B,h,w,C] -> y[..., :C] of a [B,H,W,ldo] buffer (legacy TF1 bilinear)."""
    B, h, w, C0 = x.shape
    C_ = C_ or C0
    L.check(L.lib().dlb_resize_bilinear(B, h, w, C_, y.shape[1], y.shape[2], y.stride(2), L.dt(x), x.data_ptr(),
                                        y.data_ptr(), L.stream_ptr()), "resize_bilinear")
    return y


def aspp_dw3_fwd(x: torch.Tensor, ws, rates, scales, shifts, ys):
    """Fused ASPP atrous depthwise stage: x [B,H,W,C] -> ys[0..2] (dw 3x3 at rates[i] + folded BN + ReLU)."""
    B, H, W_, C_ = x.shape
    P3 = C.c_void_p * 3
    wv = P3(*[t.data_ptr() for t in ws])
    sc = P3(*[t.data_ptr() for t in scales])
    sh = P3(*[t.data_ptr() for t in shifts])
    yv = P3(*[t.data_ptr() for t in ys])
    rt = (C.c_int * 3)(*rates)
    L.check(L.lib().dlb_aspp_dw3_fwd(B, H, W_, C_, L.dt(x), x.data_ptr(), wv, rt, sc, sh, yv, L.stream_ptr()),
            "aspp_dw3_fwd")
    return ys


def sepconv_pack_dw(ws, scales, shifts, dtype: torch.dtype) -> torch.Tensor:
    """Pack depthwise kernels [3,3,C] + folded BN (scale, shift) of the rate>0 branches for sepconv_fused_fwd."""
    n = len(ws)
    C_ = ws[0].shape[2]
    L.require_cuda(*ws, *scales, *shifts)
    nbytes = L.lib().dlb_sepconv_pack_bytes(C_, n)
    pack = torch.empty(nbytes, dtype=torch.uint8, device=ws[0].device)
    PN = C.c_void_p * n
    L.check(L.lib().dlb_sepconv_pack_dw(C_, L._DT[dtype], n, PN(*[t.data_ptr() for t in ws]),
                                        PN(*[t.data_ptr() for t in scales]), PN(*[t.data_ptr() for t in shifts]),
                                        pack.data_ptr(), L.stream_ptr()), "sepconv_pack_dw")
    return pack


def sepconv_fused_fwd(x: torch.Tensor, rates, w_pws, dw_pack, pw_scales, pw_shifts, outs, dw_act=L.ACT_RELU,
                      pw_act=L.ACT_RELU, residuals=None):
    """Fused [atrous depthwise 3x3 + BN + act] -> [1x1 + BN + act] branches sharing the input x (rate 0 = plain 1x1).

    x [B,H,W,C] f16/bf16; w_pws[i] [N,C]; outs[i]: [B,H,W,N] views (channel slices of a concat buffer are fine)."""
    B, H, W_, C_ = x.shape
    n = len(rates)
    L.require_cuda(x, *w_pws, *outs)
    p = L.SepconvFusedParams()
    p.B, p.H, p.W, p.C, p.N = B, H, W_, C_, w_pws[0].shape[0]
    p.dtype, p.n_branches = L.dt(x), n
    for i in range(n):
        assert x.is_contiguous() and w_pws[i].is_contiguous() and w_pws[i].shape[1] == C_ and outs[i].stride(3) == 1
        p.rates[i] = int(rates[i])
        p.w_pw[i] = w_pws[i].data_ptr()
        p.pw_scale[i] = L.ptr(pw_scales[i])
        p.pw_shift[i] = L.ptr(pw_shifts[i])
        p.out[i] = outs[i].data_ptr()
        if residuals is not None and residuals[i] is not None:
            p.res[i] = residuals[i].data_ptr()
            p.ldr = residuals[i].stride(2)
    p.x = x.data_ptr()
    p.dw_pack = L.ptr(dw_pack)
    p.ldc = outs[0].stride(2)
    p.dw_act, p.pw_act = dw_act, pw_act
    L.check(L.lib().dlb_sepconv_fused_fwd(C.byref(p), L.stream_ptr()), "sepconv_fused_fwd")
    return outs
