"""Kernel-level parity: every C-ABI op against a plain PyTorch fp32 reference of the same op (on the GPU).

The network-level parity tests against the oracle live in test_model_gpu.py; these isolate one kernel each so a
failure points at a descriptor / index bug directly.
"""
import math

import pytest
import numpy as np
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    import deeplab_b200  # noqa: F401
    from deeplab_b200 import ops
    return ops


def rel_err(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


GEMM_SHAPES = [
    # (M, K, N)  -- the MobileNetV2 / ASPP / head shapes at reduced M, plus ragged M
    (256, 64, 64), (384, 16, 96), (1024, 96, 24), (640, 24, 144), (512, 144, 32), (1024, 192, 64),
    (512, 384, 96), (384, 576, 160), (256, 960, 320), (512, 160, 960), (300, 320, 256), (256, 256, 1344),
    (1000, 256, 21), (128 * 37 + 5, 32, 192), (256, 512, 256),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,K,N", GEMM_SHAPES)
def test_pw_gemm_tc_plain(M, K, N, dtype):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + K * 3 + N)
    A = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    Bt = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(dtype)
    ldc = (N + 7) // 8 * 8
    out = torch.full((M, ldc), float("nan"), device="cuda", dtype=dtype)
    ops.pw_gemm(A, Bt, out)
    ref = A.float() @ Bt.float().t()
    torch.cuda.synchronize()
    tol = 4e-3 if dtype == torch.float16 else 2e-2
    assert rel_err(out[:, :N], ref) < tol
    if ldc > N:
        assert (out[:, N:] == 0).all()    # zero-filled pad columns (OOB weight rows)


@pytest.mark.parametrize("M,K,N", [(512, 96, 24), (384, 576, 160), (256, 160, 960), (640, 320, 256)])
def test_pw_gemm_tc_epilogue(M, K, N):
    """scale/shift + per-image bias + relu6 + residual + BN statistics, fp16."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    Bt = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
    sc = torch.rand(N, device="cuda", generator=g) + 0.5
    sh = torch.randn(N, device="cuda", generator=g)
    rows_per_img = M // 2
    rb = torch.randn(2, N, device="cuda", generator=g)
    R = torch.randn(M, N, device="cuda", generator=g).half()
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    ssum = torch.zeros(N, device="cuda", dtype=torch.float64)
    ssqs = torch.zeros(N, device="cuda", dtype=torch.float64)
    ops.pw_gemm(A, Bt, out, col_scale=sc, col_shift=sh, row_bias=rb, rows_per_img=rows_per_img,
                act=ops.ACT_RELU6, residual=R, stat_sum=ssum, stat_sqs=ssqs)
    pre = (A.float() @ Bt.float().t()) * sc + sh + rb.repeat_interleave(rows_per_img, 0)
    ref = pre.clamp(0, 6) + R.float()
    assert rel_err(out, ref) < 4e-3
    pre_r = pre.half().double()
    assert rel_err(ssum, pre_r.sum(0)) < 1e-3
    assert rel_err(ssqs, (pre_r * pre_r).sum(0)) < 1e-3


def test_pw_gemm_f32_out_and_f32_exact():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    M, K, N = 777, 256, 21
    A = torch.randn(M, K, device="cuda", generator=g)
    Bt = torch.randn(N, K, device="cuda", generator=g) / 16
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A @ Bt.t() + bias
    # 16-bit inputs, fp32 logits with padded pitch 32 (the head)
    out = torch.full((M, 32), float("nan"), device="cuda")
    ops.pw_gemm(A.half(), Bt.half(), out, col_shift=bias, n_store=32)
    assert rel_err(out[:, :N], A.half().float() @ Bt.half().float().t() + bias) < 1e-3
    # exact fp32 SIMT path
    out32 = torch.full((M, 32), float("nan"), device="cuda")
    ops.pw_gemm(A, Bt, out32, col_shift=bias, n_store=32)
    assert rel_err(out32[:, :N], ref) < 1e-5
    assert (out32[:, N:] == 0).all()


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_pw_gemm_subpixel_store(dtype):
    """phase-shift fused into the epilogue == subpixel.py:77-88 applied to the plain conv output."""
    ops = _ops()
    from oracle import ref_ops
    g = torch.Generator(device="cuda").manual_seed(5)
    B, h, w, K, Cs, r = 2, 8, 16, 64, 21, 8
    N = Cs * r * r
    A = torch.randn(B * h * w, K, device="cuda", generator=g).to(dtype)
    W_keras = (torch.randn(K, N, device="cuda", generator=g) / 8)       # Keras column order k*r*r + i*r + j
    bias_keras = torch.randn(N, device="cuda", generator=g)
    perm = torch.from_numpy(ref_ops.subpixel_column_perm(Cs, r)).cuda()  # internal column j' <- keras column perm[j']
    Wt = W_keras[:, perm].t().contiguous().to(dtype)
    bias = bias_keras[perm].contiguous()
    out = torch.full((B, h * r, w * r, Cs), float("nan"), device="cuda", dtype=torch.float32)
    ops.pw_gemm(A, Wt, out, col_shift=bias, shuffle=(r, h, w))
    conv = (A.float() @ W_keras.to(dtype).float() + bias_keras).view(B, h, w, N).cpu()
    ref = ref_ops.phase_shift(conv, r)
    assert rel_err(out.cpu(), ref) < (3e-3 if dtype == torch.float16 else 1e-5)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,K,N", [(4096, 96, 24), (2048, 16, 96), (1024, 960, 160), (1536, 160, 960),
                                   (8192 + 40, 320, 256), (1000, 256, 32), (2048, 384, 64), (512, 24, 144)])
def test_pw_wgrad(M, K, N, dtype):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    A = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    dY = torch.randn(M, N, device="cuda", generator=g).to(dtype)
    dW = torch.full((K, N), float("nan"), device="cuda")
    db = torch.empty(N, device="cuda")
    ws = torch.empty(ops.pw_wgrad_workspace_bytes(M, N, K) // 4, device="cuda")
    ops.pw_wgrad(A, dY, dW, dbias=db, workspace=ws)
    ref = A.float().t() @ dY.float()
    assert rel_err(dW, ref) < 1e-4
    assert rel_err(db, dY.float().sum(0)) < 1e-4
    ops.pw_wgrad(A, dY, dW, beta=1.0, workspace=ws)
    assert rel_err(dW, 2 * ref) < 1e-4


MNV2_KN = [(16, 96), (96, 24), (24, 144), (144, 24), (144, 32), (32, 192), (192, 32), (192, 64), (64, 384), (384, 64),
           (384, 96), (96, 576), (576, 96), (576, 160), (160, 960), (960, 160), (960, 320), (320, 256), (256, 256),
           (256, 24), (256, 1344), (32, 16)]


@pytest.mark.parametrize("K,N", MNV2_KN)
def test_all_mobilenet_layer_shapes_fwd_dgrad_wgrad(K, N):
    """every (Cin, Cout) pair of the network through the three tensor-core GEMM roles, fp16, ragged M."""
    ops = _ops()
    M = 128 * 21 + 64
    g = torch.Generator(device="cuda").manual_seed(K * 1000 + N)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    Wm = (torch.randn(K, N, device="cuda", generator=g) / math.sqrt(K)).half()      # [K, N] Keras layout
    dY = torch.randn(M, N, device="cuda", generator=g).half()
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    ops.pw_gemm(A, Wm.t().contiguous(), out)                                        # forward: Bt = W^T [N, K]
    assert rel_err(out, A.float() @ Wm.float()) < 4e-3
    dA = torch.empty(M, K, device="cuda", dtype=torch.float16)
    ops.pw_gemm(dY, Wm, dA)                                                         # dgrad: Bt = W [K, N]
    assert rel_err(dA, dY.float() @ Wm.float().t()) < 4e-3
    dW = torch.empty(K, N, device="cuda")
    ws = torch.empty(ops.pw_wgrad_workspace_bytes(M, N, K) // 4, device="cuda")
    ops.pw_wgrad(A, dY, dW, workspace=ws)
    assert rel_err(dW, A.float().t() @ dY.float()) < 1e-4


def _dw_ref(x, w, stride, dil, in_scale=None, in_shift=None, in_act=0):
    """NHWC fp32 reference with TF-SAME padding (oracle/ref_ops.py has the CPU twin)."""
    a = x.float()
    if in_scale is not None:
        a = a * in_scale + in_shift
        a = a.clamp(0, 6) if in_act == 2 else (a.clamp_min(0) if in_act == 1 else a)
    B, H, W, C = a.shape
    from deeplab_b200.ops import tf_same_pad
    Ho, pt, pb = tf_same_pad(H, 3, stride, dil)
    Wo, pl, pr = tf_same_pad(W, 3, stride, dil)
    a = F.pad(a.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(a, w.permute(2, 0, 1).unsqueeze(1), stride=stride, dilation=dil, groups=C)
    return y.permute(0, 2, 3, 1).contiguous(), (Ho, Wo, pt, pl)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("H,W,C,stride,dil", [(32, 32, 96, 1, 1), (32, 48, 144, 2, 1), (33, 31, 32, 2, 1),
                                               (16, 16, 384, 1, 2), (16, 16, 960, 1, 4), (24, 24, 64, 1, 12),
                                               (8, 8, 2048, 1, 36),
                                               # ragged sizes through the dilation-phase decomposition of the TMA kernels
                                               (35, 29, 96, 1, 2), (30, 44, 64, 1, 4), (41, 37, 72, 1, 3)])
def test_dw_conv_fwd_bwd(H, W, C, stride, dil, dtype):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(H * W + C)
    B = 2
    x = torch.randn(B, H, W, C, device="cuda", generator=g).to(dtype)
    w = torch.randn(3, 3, C, device="cuda", generator=g) / 3
    isc = torch.rand(C, device="cuda", generator=g) + 0.5
    ish = torch.randn(C, device="cuda", generator=g) * 0.5
    xr = x.float().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref, (Ho, Wo, pt, pl) = _dw_ref(xr, wr, stride, dil, isc, ish, 2)
    y = torch.empty(B, Ho, Wo, C, device="cuda", dtype=dtype)
    ssum = torch.zeros(C, device="cuda", dtype=torch.float64)
    ssqs = torch.zeros(C, device="cuda", dtype=torch.float64)
    ops.dw_conv_fwd(x, w, y, stride=stride, dilation=dil, pad_top=pt, pad_left=pl, in_scale=isc, in_shift=ish,
                    in_act=2, stat_sum=ssum, stat_sqs=ssqs)
    # 16 bit: the prologue and the 3-tap row sums run on packed half2 / bfloat162 (rows are added in fp32)
    tol = {torch.float16: 8e-3, torch.bfloat16: 5e-2, torch.float32: 1e-5}[dtype]
    stol = {torch.float16: 2e-3, torch.bfloat16: 1e-2, torch.float32: 1e-4}[dtype]
    assert rel_err(y, ref) < tol
    yr = y.double()
    assert rel_err(ssum, yr.sum((0, 1, 2))) < stol
    assert rel_err(ssqs, (yr * yr).sum((0, 1, 2))) < stol
    # backward: gradient w.r.t. the *activated* input a and w.r.t. the weights
    dy = torch.randn(B, Ho, Wo, C, device="cuda", generator=g).to(dtype)
    a = (x.float() * isc + ish).clamp(0, 6).requires_grad_(True)
    ref2, _ = _dw_ref(a, wr, stride, dil)
    ga, gw = torch.autograd.grad(ref2, [a, wr], dy.float())
    dx = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=dtype)
    dw = torch.zeros(3, 3, C, device="cuda")
    ops.dw_conv_bwd(x, dy, w, dx=dx, dw=dw, stride=stride, dilation=dil, pad_top=pt, pad_left=pl, in_scale=isc,
                    in_shift=ish, in_act=2)
    assert rel_err(dx, ga) < tol
    assert rel_err(dw, gw) < {torch.float16: 5e-3, torch.bfloat16: 3e-2, torch.float32: 1e-3}[dtype]


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("H,W", [(64, 64), (33, 47)])
def test_stem_conv(H, W, dtype):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(H)
    B = 2
    x = torch.randint(0, 256, (B, H, W, 3), device="cuda", generator=g).float()
    w = torch.randn(3, 3, 3, 32, device="cuda", generator=g) / 5
    Ho, pt, pb = ops.tf_same_pad(H, 3, 2, 1)
    Wo, pl, pr = ops.tf_same_pad(W, 3, 2, 1)
    xp = F.pad((x / 127.5 - 1).permute(0, 3, 1, 2), (pl, pr, pt, pb))
    wr = w.clone().requires_grad_(True)
    ref = F.conv2d(xp, wr.permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1)
    y = torch.empty(B, Ho, Wo, 32, device="cuda", dtype=dtype)
    ssum = torch.zeros(32, device="cuda", dtype=torch.float64)
    ssqs = torch.zeros(32, device="cuda", dtype=torch.float64)
    ops.stem_conv_fwd(x, w, y, stat_sum=ssum, stat_sqs=ssqs)
    tol = 3e-3 if dtype == torch.float16 else 1e-5
    assert rel_err(y, ref) < tol
    assert rel_err(ssqs, (y.double() ** 2).sum((0, 1, 2))) < 1e-4
    dy = torch.randn(B, Ho, Wo, 32, device="cuda", generator=g).to(dtype)
    (gw,) = torch.autograd.grad(ref, [wr], dy.float())
    dw = torch.zeros(3, 3, 3, 32, device="cuda")
    ops.stem_conv_wgrad(x, dy, dw)
    assert rel_err(dw, gw) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_bn_train_fwd_bwd(dtype):
    """finalize + apply (+residual) and the two-pass backward vs autograd through F.batch_norm."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(11)
    M, C = 4096, 96
    x = (torch.randn(M, C, device="cuda", generator=g) * 2 + 0.5).to(dtype)
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g)
    res = torch.randn(M, C, device="cuda", generator=g).to(dtype)
    xd = x.double()
    ssum, ssqs = xd.sum(0), (xd * xd).sum(0)
    mm, mv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    scale, shift, mean, rstd = (torch.empty(C, device="cuda") for _ in range(4))
    eps, mom = 1e-3, 0.999
    ops.bn_finalize(M, ssum, ssqs, gamma, beta, eps, mom, mm, mv, scale, shift, mean, rstd)
    assert (ssum == 0).all()
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.batch_norm(xr, None, None, gr, br, True, 0.0, eps)
    ref = z.clamp(0, 6) + res.float()
    y = torch.empty_like(x)
    ops.bn_act_apply(x, y, scale=scale, shift=shift, act=2, res=res)
    tol = {torch.float16: 3e-3, torch.bfloat16: 2e-2, torch.float32: 2e-5}[dtype]
    assert rel_err(y, ref) < tol
    var_u = x.float().var(0, unbiased=True)
    assert rel_err(mv, 0.999 + 0.001 * var_u) < 1e-5
    assert rel_err(mm, 0.001 * x.float().mean(0)) < 1e-4
    da = torch.randn(M, C, device="cuda", generator=g).to(dtype)
    gx, gg, gb = torch.autograd.grad(ref, [xr, gr, br], da.float())
    red = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    dx = torch.empty_like(x)
    dgamma, dbeta = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.bn_bwd(x, da, dx, scale=scale, shift=shift, mean=mean, rstd=rstd, act=2, red=red, dgamma=dgamma, dbeta=dbeta)
    if dtype != torch.float32:
        # the 16-bit kernels evaluate the ReLU6 mask on packed pairs: an element whose pre-activation lies within one
        # ulp of 0 or 6 may flip, so compare in the L2 norm instead of the max norm
        assert ((dx.float() - gx).norm() / gx.norm()).item() < (1e-2 if dtype == torch.float16 else 5e-2)
    else:
        assert rel_err(dx, gx) < 1e-4
    # 16 bit: the reduce pass runs on packed pairs with 4-row partial sums (x-hat from 16-bit mean / rstd)
    rtol = {torch.float16: 2e-2, torch.bfloat16: 8e-2, torch.float32: 1e-3}[dtype]
    assert rel_err(dgamma, gg) < rtol
    assert rel_err(dbeta, gb) < rtol


def test_dropout_apply_and_bwd_consistent():
    ops = _ops()
    M, C = 2048, 256
    x = torch.randn(M, C, device="cuda")
    one, zero = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    y = torch.empty_like(x)
    ops.bn_act_apply(x, y, scale=one, shift=zero, act=1, drop_rate=0.1, drop_seed=1234)
    kept = (y != 0) | (x <= 0)
    frac = ((y == 0) & (x > 0)).float().sum() / (x > 0).float().sum()
    assert abs(frac.item() - 0.1) < 0.01
    assert rel_err(y[kept & (x > 0)], x[kept & (x > 0)] / 0.9) < 1e-6
    # backward uses the same mask
    red = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    dx = torch.empty_like(x)
    da = torch.ones_like(x)
    ops.bn_bwd(x, da, dx, scale=one, shift=zero, mean=zero, rstd=one, act=1, red=red, drop_rate=0.1, drop_seed=1234,
               frozen=True)
    assert torch.equal(dx != 0, y != 0)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_global_avgpool(dtype):
    ops = _ops()
    B, HW, C = 3, 1024, 320
    x = torch.randn(B, HW, C, device="cuda").to(dtype)
    out = torch.empty(B, C, device="cuda")
    ops.global_avgpool_fwd(x, out)
    assert rel_err(out, x.float().mean(1)) < 1e-4
    d = torch.randn(B, C, device="cuda")
    dx = torch.randn(B, HW, C, device="cuda").to(dtype)
    base = dx.float().clone()
    ops.global_avgpool_bwd(d, dx, True)
    assert rel_err(dx, base + d[:, None, :] / HW) < (2e-3 if dtype == torch.float16 else 1e-6)


def test_small_gemm():
    ops = _ops()
    A = torch.randn(16, 320, device="cuda")
    Bm = torch.randn(320, 256, device="cuda")
    out = torch.empty(16, 256, device="cuda")
    ops.small_gemm(A, Bm, out, M=16, N=256, K=320)
    assert rel_err(out, A @ Bm) < 1e-5
    out2 = torch.empty(320, 256, device="cuda")
    ops.small_gemm(A, out, out2, M=320, N=256, K=16, transA=True, alpha=2.0)
    assert rel_err(out2, 2 * A.t() @ out) < 1e-5
    # both operands K-contiguous (the image-pooling dgrads): warp-per-column kernel, odd K, beta accumulation
    for M, N, K in [(16, 320, 256), (16, 256, 256), (3, 21, 77), (64, 40, 512)]:
        A2, B2 = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
        c0 = torch.randn(M, N, device="cuda")
        o = c0.clone()
        ops.small_gemm(A2, B2, o, M=M, N=N, K=K, transB=True, alpha=0.5, beta=2.0)
        assert rel_err(o, 0.5 * A2 @ B2.t() + 2.0 * c0) < 1e-5
        ops.small_gemm(A2, B2, o, M=M, N=N, K=K, transB=True)
        assert rel_err(o, A2 @ B2.t()) < 1e-5
    A3, B3 = torch.randn(8, 600, device="cuda"), torch.randn(24, 600, device="cuda")       # K > 512: the generic kernel
    o3 = torch.empty(8, 24, device="cuda")
    ops.small_gemm(A3, B3, o3, M=8, N=24, K=600, transB=True)
    assert rel_err(o3, A3 @ B3.t()) < 1e-5


@pytest.mark.parametrize("h,w,S,ldl", [(16, 16, 8, 32), (12, 20, 8, 32), (16, 24, 4, 32), (32, 32, 1, 21)])
def test_resize_softmax_and_ce(h, w, S, ldl):
    ops = _ops()
    from oracle import ref_ops
    g = torch.Generator(device="cuda").manual_seed(h * w + S)
    B, C = 2, 21
    H, W = h * S, w * S
    logits = torch.zeros(B, h, w, ldl, device="cuda")
    logits[..., :C] = torch.randn(B, h, w, C, device="cuda", generator=g) * 3
    probs = torch.empty(B, H * W, C, device="cuda")
    am = torch.empty(B, H * W, device="cuda", dtype=torch.uint8)
    ops.resize_softmax_fwd(logits, C, H, W, probs, am)
    lr = logits[..., :C].cpu().clone().requires_grad_(True)
    up = ref_ops.resize_bilinear_tf1(lr, H, W)
    pref = torch.softmax(up.reshape(B, H * W, C), -1)
    assert rel_err(probs.cpu(), pref.detach()) < 1e-5
    assert (am.cpu().long() == pref.argmax(-1)).float().mean() > 0.9999
    labels = torch.randint(0, C + 1, (B, H * W, 1), device="cuda", generator=g).float()
    sw = torch.rand(B, H * W, device="cuda", generator=g)
    sw[sw < 0.2] = 0
    loss_ref = ref_ops.keras_weighted_loss(labels.cpu(), pref, sw.cpu())
    (gref,) = torch.autograd.grad(loss_ref, [lr])
    gs = torch.zeros(1, device="cuda")
    wc = torch.zeros(1, device="cuda", dtype=torch.float64)
    ops.ce_grad_scale(B * H * W, sw, gs, wc)
    assert wc.item() == (sw != 0).sum().item()
    dlog = torch.zeros(B, h, w, ldl, device="cuda")
    loss_sum = torch.zeros(1, device="cuda", dtype=torch.float64)
    am2 = torch.empty_like(am)
    ops.resize_softmax_ce(logits, C, H, W, labels, sw, gs, dlog, loss_sum, wc, am2)
    loss = loss_sum.item() / wc.item()
    assert abs(loss - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    assert rel_err(dlog[..., :C].cpu(), gref) < 1e-4
    assert torch.equal(am, am2)


def test_phase_shift_roundtrip():
    ops = _ops()
    from oracle import ref_ops
    B, h, w, Cs, r = 2, 4, 6, 21, 8
    x = torch.randn(B, h, w, Cs * r * r, device="cuda")
    out = torch.empty(B, h * r, w * r, Cs, device="cuda")
    ops.phase_shift(x, out, r)
    assert torch.equal(out.cpu(), ref_ops.phase_shift(x.cpu(), r))
    back = torch.empty_like(x)
    ops.phase_shift(out, back, r, inverse=True)
    assert torch.equal(back, x)


def test_adam_matches_keras_rule():
    ops = _ops()
    from oracle import ref_ops
    n = 10007
    p = torch.randn(n, device="cuda")
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    step = torch.zeros(1, device="cuda", dtype=torch.int64)
    pr, mr, vr = p.cpu().clone(), m.cpu().clone(), v.cpu().clone()
    for it in range(3):
        gte = torch.randn(n, device="cuda")
        ops.adam_step(p, gte, m, v, step, lr=7e-4, eps=1e-8, decay=1e-6)
        pr, mr, vr = ref_ops.keras_adam(pr, gte.cpu(), mr, vr, it, lr=7e-4, eps=1e-8, decay=1e-6)
    assert step.item() == 3
    assert rel_err(p.cpu(), pr) < 1e-6


def test_cast_weight_and_confusion():
    ops = _ops()
    K, N = 96, 24
    w = torch.randn(K, N, device="cuda")
    wkn = torch.empty(K, N, device="cuda", dtype=torch.float16)
    wnk = torch.empty(N, K, device="cuda", dtype=torch.float16)
    ops.cast_weight(w, K, N, wkn, wnk)
    assert torch.equal(wkn, w.half()) and torch.equal(wnk, w.half().t())
    B, npix, C = 2, 5000, 21
    labels = torch.randint(0, C + 1, (B, npix, 1), device="cuda").float()
    am = torch.randint(0, C, (B, npix), device="cuda", dtype=torch.uint8)
    conf = torch.zeros(B, C + 1, C, device="cuda", dtype=torch.int64)
    ops.confusion(labels, am, C, conf)
    ref = torch.zeros(B, C + 1, C, dtype=torch.int64)
    for b in range(B):
        idx = labels[b, :, 0].long().cpu() * C + am[b].long().cpu()
        ref[b] = torch.bincount(idx, minlength=(C + 1) * C).view(C + 1, C)
    assert torch.equal(conf.cpu(), ref)


@pytest.mark.parametrize("ldtype", [torch.uint8, torch.int32, torch.float32])
def test_label_weights_bit_exact_vs_oracle(ldtype):
    """dlb_label_weights (generator contract, reference utils.py:360-399) vs the numpy restatement: exact."""
    from oracle import ref_ops as R
    ops = _ops()
    rng = np.random.RandomState(5)
    n_classes = 21
    labs = np.stack([rng.choice([0, 1, 5, 20, 21, 255], size=(64, 48), p=[.5, .1, .1, .1, .1, .1]),
                     rng.choice([0, 255], size=(64, 48)),
                     np.full((64, 48), 255),
                     rng.randint(0, 22, size=(64, 48))]).astype(np.int64)
    t = torch.from_numpy(labs).to(ldtype).cuda().view(4, -1)
    y = torch.empty(4, 64 * 48, device="cuda")
    sw = torch.empty(4, 64 * 48, device="cuda")
    ops.label_weights(t, n_classes, y, sw)
    for b in range(4):
        yr, swr = R.generator_labels_and_weights(labs[b], n_classes)
        assert np.array_equal(y[b].cpu().numpy(), yr.astype(np.float32))
        assert np.array_equal(sw[b].cpu().numpy(), swr)


# ------------------------------------------------------------------------------------------------------------------
# A-operand transform: the project conv (forward GEMM and weight gradient) consumes the RAW depthwise output and
# applies depthwise_BN + relu6 to the landed tiles (deeplabv3p.py:189-196)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,K,N", [(4096 + 37, 96, 24), (16384, 576, 96), (8192, 960, 160), (8192, 960, 320), (1024, 32, 16),
                                   (2048, 144, 32)])
def test_pw_gemm_and_wgrad_a_operand_transform(dt, M, K, N):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    raw = (torch.randn(M, K, device="cuda", generator=g) * 2 + 0.5).to(dt)
    sc = torch.rand(K, device="cuda", generator=g) + 0.5
    sh = torch.randn(K, device="cuda", generator=g)
    Bt = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(dt)
    a_ref = (raw.float() * sc + sh).clamp(0, 6)
    a_rnd = a_ref.to(dt).float()                       # what the tensor cores see (tile rounded back to the storage type)
    # forward GEMM + statistics of the (rounded) output
    out = torch.empty(M, N, device="cuda", dtype=dt)
    ssum = torch.zeros(N, device="cuda", dtype=torch.float64)
    ssqs = torch.zeros(N, device="cuda", dtype=torch.float64)
    ops.pw_gemm(raw, Bt, out, stat_sum=ssum, stat_sqs=ssqs, a_scale=sc, a_shift=sh, a_act=ops.ACT_RELU6)
    ref = a_rnd @ Bt.float().t()
    tol = {torch.float16: 2e-3, torch.bfloat16: 1.5e-2, torch.float32: 1e-5}[dt]
    assert rel_err(out, ref) < tol
    o = out.double()
    assert rel_err(ssum, o.sum(0)) < 1e-3 and rel_err(ssqs, (o * o).sum(0)) < 1e-3
    # weight gradient dW[K, N] = A'^T dY
    dY = (torch.randn(M, N, device="cuda", generator=g) / 8).to(dt)
    dW = torch.zeros(K, N, device="cuda")
    ops.pw_wgrad(raw, dY, dW, beta=0.0, a_scale=sc, a_shift=sh, a_act=ops.ACT_RELU6)
    dW_ref = a_rnd.double().t() @ dY.double()
    assert rel_err(dW, dW_ref) < {torch.float16: 1e-3, torch.bfloat16: 1e-3, torch.float32: 1e-4}[dt]
    # and without the transform the same kernels still see A as is
    dW2 = torch.zeros(K, N, device="cuda")
    ops.pw_wgrad(raw, dY, dW2, beta=0.0)
    assert rel_err(dW2, raw.double().t() @ dY.double()) < 1e-3


# ------------------------------------------------------------------------------------------------------------------
# second half of _inverted_res_block in one kernel (inference): depthwise 3x3 (rate r) + BN + relu6 -> project 1x1 + BN
# (+ add), deeplabv3p.py:186-206, for every stride-1 MobileNetV2 block shape at 64x64 / 128x128
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("HW,C,N,rate,skip", [(128, 144, 24, 1, True), (64, 192, 32, 1, True), (64, 192, 64, 1, False),
                                               (64, 384, 64, 2, True), (64, 384, 96, 2, False), (64, 576, 96, 2, True),
                                               (64, 576, 160, 2, False), (64, 960, 160, 4, True)])
def test_fused_depthwise_project_matches_unfused(dt, HW, C, N, rate, skip):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(C + N + rate)
    B = 2
    x = (torch.rand(B, HW, HW, C, device="cuda", generator=g) * 6).to(dt)            # relu6 output of the expand conv
    wd = torch.randn(3, 3, C, 1, device="cuda", generator=g) / 3
    dsc = torch.rand(C, device="cuda", generator=g) + 0.5
    dsh = torch.randn(C, device="cuda", generator=g) * 0.5
    wp = (torch.randn(N, C, device="cuda", generator=g) / math.sqrt(C)).to(dt)
    psc = torch.rand(N, device="cuda", generator=g) + 0.5
    psh = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(B, HW, HW, N, device="cuda", generator=g).to(dt) if skip else None
    out = torch.empty(B, HW, HW, N, device="cuda", dtype=dt)
    pk = ops.sepconv_pack_dw([wd], [dsc], [dsh], dt)
    ops.sepconv_fused_fwd(x, [rate], [wp], pk, [psc], [psh], [out], dw_act=ops.ACT_RELU6, pw_act=ops.ACT_NONE,
                          residuals=[res])
    # fp32 reference of the same chain on the stored (rounded) operands
    xn = x.float().permute(0, 3, 1, 2)
    y = F.conv2d(xn, wd.to(dt).float().reshape(3, 3, C).permute(2, 0, 1).unsqueeze(1), padding=rate, dilation=rate, groups=C)
    a = (y * dsc.to(dt).float().view(1, C, 1, 1) + dsh.to(dt).float().view(1, C, 1, 1)).clamp(0, 6).permute(0, 2, 3, 1)
    ref = (a.reshape(-1, C) @ wp.float().t()) * psc + psh
    if skip:
        ref = ref + res.float().reshape(-1, N)
    tol = 1e-2 if dt == torch.float16 else 6e-2          # the depthwise stage accumulates its 9 taps in 16 bit
    assert rel_err(out.reshape(-1, N), ref) < tol


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_dropout_in_streaming_bn_kernels_matches_fp32_mask(dt):
    """Dropout(0.1) behind concat_projection_BN (deeplabv3p.py:410) runs in the bulk-copy streaming BatchNorm kernels for
    16-bit storage (round 1 fell back to the register-staged kernels there: 134 us instead of 33).  The keep mask is a
    hash of (seed, element), so the 16-bit streaming kernels and the fp32 kernels must drop exactly the same elements,
    forward and backward, and scale the kept ones by 1 / (1 - p)."""
    ops = _ops()
    M, C = 16 * 64 * 64 // 4, 256
    g = torch.Generator(device="cuda").manual_seed(7)
    x32 = torch.randn(M, C, device="cuda", generator=g)
    x = x32.to(dt)
    sc = torch.rand(C, device="cuda", generator=g) + 0.5
    sh = torch.randn(C, device="cuda", generator=g) * 0.2
    seed_dev = torch.tensor([5], device="cuda", dtype=torch.int64)
    y32, y = torch.empty_like(x32), torch.empty_like(x)
    ops.bn_act_apply(x.float(), y32, scale=sc, shift=sh, act=1, drop_rate=0.1, drop_seed=99, drop_seed_dev=seed_dev)
    ops.bn_act_apply(x, y, scale=sc, shift=sh, act=1, drop_rate=0.1, drop_seed=99, drop_seed_dev=seed_dev)
    assert rel_err(y, y32) < (2e-3 if dt == torch.float16 else 1.5e-2)
    pos = y32 > 1e-2
    assert torch.equal((y != 0)[pos], torch.ones_like(pos)[pos])                  # kept elements are kept
    z = (x.float() * sc + sh)
    dropped32 = (y32 == 0) & (z > 1e-2)
    assert torch.equal((y == 0)[dropped32], torch.ones_like(pos)[dropped32])      # dropped elements are dropped
    assert abs(dropped32.float().sum().item() / (z > 1e-2).float().sum().item() - 0.1) < 0.01
    # backward: same mask, gradient scaled by 1 / (1 - p); statistics frozen so dx = scale * mask * da
    da32 = torch.randn(M, C, device="cuda", generator=g)
    da = da32.to(dt)
    mean, rstd = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    dx32, dx = torch.empty_like(x32), torch.empty_like(x)
    for (xx, dd, oo) in ((x.float(), da.float(), dx32), (x, da, dx)):
        red = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
        ops.bn_bwd(xx, dd, oo, scale=sc, shift=sh, mean=mean, rstd=rstd, act=1, red=red, drop_rate=0.1, drop_seed=99,
                   drop_seed_dev=seed_dev, frozen=True)
    clear = z.abs() > 1e-2                                                        # away from the ReLU kink
    assert rel_err(dx[clear], dx32[clear]) < (3e-3 if dt == torch.float16 else 2e-2)
    assert torch.equal((dx != 0)[clear & (da != 0)], (dx32 != 0)[clear & (da != 0)])


def _bn_state(C, g):
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g)
    mm, mv = torch.randn(C, device="cuda", generator=g), torch.rand(C, device="cuda", generator=g) + 0.5
    return gamma, beta, mm, mv


def _fin_pair(ops, x2d, gamma, beta, mm, mv, moving=True):
    """(stand-alone finalize results, a dlb_bn_fin over cloned state + its output tensors) for the same statistics."""
    C = x2d.shape[1]
    xd = x2d.double()
    ssum, ssqs = xd.sum(0), (xd * xd).sum(0)
    ref = dict(mm=mm.clone(), mv=mv.clone(), **{k: torch.empty(C, device="cuda") for k in ("scale", "shift", "mean", "rstd")})
    ops.bn_finalize(x2d.shape[0], ssum.clone(), ssqs.clone(), gamma, beta, 1e-3, 0.999, ref["mm"] if moving else None,
                    ref["mv"] if moving else None, ref["scale"], ref["shift"], ref["mean"], ref["rstd"], reset=False)
    out = dict(mm=mm.clone(), mv=mv.clone(), **{k: torch.full((C,), float("nan"), device="cuda") for k in ("scale", "shift", "mean", "rstd")})
    fin = ops.bn_fin(x2d.shape[0], ssum, ssqs, gamma, beta, 1e-3, 0.999, out["mm"] if moving else None,
                     out["mv"] if moving else None, out["scale"], out["shift"], out["mean"], out["rstd"])
    return ref, out, fin, (ssum, ssqs)


def _same_state(ref, out):
    for k in ref:
        assert torch.equal(ref[k], out[k]), k


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,C,res,moving", [(16384, 96, True, True), (4096 + 5, 32, False, True), (8192, 960, False, False),
                                            (16, 256, False, True)])
def test_bn_apply_consumer_side_finalize_is_bit_identical(dt, M, C, res, moving):
    """dlb_bn_fin through bn_act_apply (streaming kernel and the fallback shapes): output, scale / shift / mean / rstd
    and the moving statistics equal the stand-alone finalize + apply bit for bit; the sums are left untouched."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(C + M)
    x = (torch.randn(M, C, device="cuda", generator=g) * 1.5 + 0.3).to(dt)
    r = torch.randn(M, C, device="cuda", generator=g).to(dt) if res else None
    gamma, beta, mm, mv = _bn_state(C, g)
    ref, out, fin, (ssum, ssqs) = _fin_pair(ops, x, gamma, beta, mm, mv, moving)
    s0 = ssum.clone()
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    ops.bn_act_apply(x, y0, scale=ref["scale"], shift=ref["shift"], act=2, res=r)
    ops.bn_act_apply(x, y1, fin=fin, act=2, res=r)
    assert torch.equal(y0, y1)
    _same_state(ref, out)
    assert torch.equal(ssum, s0)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,K,N", [(16384, 576, 96), (8192, 960, 160), (4096 + 37, 96, 24), (1024, 32, 16)])
def test_pw_gemm_a_operand_consumer_side_finalize(dt, M, K, N):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(K + N)
    a = (torch.randn(M, K, device="cuda", generator=g) * 1.5 + 0.3).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(dt)
    gamma, beta, mm, mv = _bn_state(K, g)
    ref, out, fin, _ = _fin_pair(ops, a, gamma, beta, mm, mv)
    y0, y1 = torch.empty(M, N, device="cuda", dtype=dt), torch.empty(M, N, device="cuda", dtype=dt)
    s0, q0, s1, q1 = (torch.zeros(N, device="cuda", dtype=torch.float64) for _ in range(4))
    ops.pw_gemm(a, w, y0, stat_sum=s0, stat_sqs=q0, a_scale=ref["scale"], a_shift=ref["shift"], a_act=2)
    ops.pw_gemm(a, w, y1, stat_sum=s1, stat_sqs=q1, a_fin=fin, a_act=2)
    assert torch.equal(y0, y1)
    _same_state(ref, out)
    assert rel_err(s1, s0) < 1e-6 and rel_err(q1, q0) < 1e-6      # atomics: order differs run to run


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("H,C,stride,dil", [(64, 96, 1, 1), (64, 144, 2, 1), (32, 384, 1, 2), (32, 960, 1, 4), (16, 64, 1, 36)])
def test_dw_conv_prologue_consumer_side_finalize(dt, H, C, stride, dil):
    ops = _ops()
    B = 2
    g = torch.Generator(device="cuda").manual_seed(C + dil)
    x = (torch.randn(B, H, H, C, device="cuda", generator=g) * 1.5 + 0.3).to(dt)
    w = torch.randn(3, 3, C, 1, device="cuda", generator=g) * 0.3
    gamma, beta, mm, mv = _bn_state(C, g)
    ref, out, fin, _ = _fin_pair(ops, x.view(-1, C), gamma, beta, mm, mv)
    Ho = (H + stride - 1) // stride
    pad = max((Ho - 1) * stride + 2 * dil + 1 - H, 0) // 2
    y0, y1 = (torch.empty(B, Ho, Ho, C, device="cuda", dtype=dt) for _ in range(2))
    s0, q0, s1, q1 = (torch.zeros(C, device="cuda", dtype=torch.float64) for _ in range(4))
    kw = dict(stride=stride, dilation=dil, pad_top=pad, pad_left=pad, in_act=2)
    ops.dw_conv_fwd(x, w, y0, in_scale=ref["scale"], in_shift=ref["shift"], stat_sum=s0, stat_sqs=q0, **kw)
    ops.dw_conv_fwd(x, w, y1, in_fin=fin, stat_sum=s1, stat_sqs=q1, **kw)
    assert torch.equal(y0, y1)
    _same_state(ref, out)
    assert rel_err(s1, s0) < 1e-6 and rel_err(q1, q0) < 1e-6


def test_consumer_side_finalize_argument_checks():
    ops = _ops()
    C = 64
    x = torch.randn(2048, C, device="cuda").half()
    g = torch.Generator(device="cuda").manual_seed(0)
    gamma, beta, mm, mv = _bn_state(C, g)
    ref, out, fin, _ = _fin_pair(ops, x, gamma, beta, mm, mv)
    y = torch.empty_like(x)
    fin.count = 0.0
    with pytest.raises(RuntimeError, match="count"):
        ops.bn_act_apply(x, y, fin=fin, act=2)
    fin.count = 2048.0
    w = torch.randn(32, C, device="cuda").half()
    with pytest.raises(RuntimeError, match="not both"):
        ops.pw_gemm(x, w, torch.empty(2048, 32, device="cuda", dtype=torch.float16), a_fin=fin, a_scale=ref["scale"],
                    a_shift=ref["shift"], a_act=2)
