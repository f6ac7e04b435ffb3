"""Xception backbone (BASELINE config 3): Deeplabv3(backbone='xception', OS=8|16) inference vs the oracle restatement
of deeplabv3p.py:272-313 / :375-429 on seeded Keras-default-initialised weights with perturbed BN statistics.
The reference ships no Xception weights and its own Xception path raises NameError (deeplabv3p.py:147), so parity
here is against the restatement only (unpinned)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _push(model, W):
    for l in model.layers:
        if l.name in W:
            l.set_weights([w.detach().float().numpy() for w in W[l.name]])


@pytest.mark.parametrize("OS", [8, 16])
def test_xception_fp32_matches_oracle(OS):
    from deeplab_b200.deeplabv3p import Deeplabv3
    from oracle import network as N
    W = N.random_xception_weights(seed=OS)
    H, Wd, B = 128, 160, 2
    model = Deeplabv3(weights=None, input_shape=(H, Wd, 3), classes=21, backbone='xception', OS=OS, compute_dtype='float32')
    _push(model, W)
    x = np.random.RandomState(OS).randint(0, 256, (B, H, Wd, 3)).astype(np.float32)
    probs = model.predict(x, batch_size=B)
    with torch.no_grad():
        logits_ref, pref, _ = N.deeplabv3_forward(W, torch.from_numpy(x), backbone="xception", OS=OS)
    logits = model.engine.workspace(B, False)["logits"][..., :21].cpu()
    # 1e-3 relative (north_star fp32 tolerance)
    assert (logits - logits_ref).abs().max() <= 1e-3 * logits_ref.abs().max()
    assert np.abs(probs - pref.numpy()).max() < 2e-3
    assert (probs.argmax(-1) == pref.argmax(-1).numpy()).mean() > 0.999


def test_xception_fp16_tensor_core_path():
    from deeplab_b200.deeplabv3p import Deeplabv3
    from oracle import network as N
    W = N.random_xception_weights(seed=3)
    H, Wd, B = 128, 128, 1
    model = Deeplabv3(weights=None, input_shape=(H, Wd, 3), backbone='xception', OS=8, compute_dtype='float16')
    _push(model, W)
    x = np.random.RandomState(5).randint(0, 256, (B, H, Wd, 3)).astype(np.float32)
    probs = model.predict(x)
    with torch.no_grad():
        _, pref, _ = N.deeplabv3_forward(W, torch.from_numpy(x), backbone="xception", OS=8)
    assert np.abs(probs - pref.numpy()).max() < 5e-2
    assert (probs.argmax(-1) == pref.argmax(-1).numpy()).mean() > 0.97


def test_xception_helper_kernels():
    """dense 3x3 conv, subsample, feature resize vs torch / the oracle ops."""
    import torch.nn.functional as F
    from deeplab_b200 import ops
    from oracle import ref_ops as R
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(2, 20, 24, 32, device="cuda", generator=g)
    w = torch.randn(3, 3, 32, 64, device="cuda", generator=g) / 10
    sc = torch.rand(64, device="cuda", generator=g) + 0.5
    sh = torch.randn(64, device="cuda", generator=g)
    for dt, tol in ((torch.float32, 1e-5), (torch.float16, 4e-3)):
        y = torch.empty(2, 20, 24, 64, device="cuda", dtype=dt)
        ops.conv3x3_fwd(x.to(dt), w, y, out_scale=sc, out_shift=sh, out_act=1)
        # CPU reference: torch's CUDA conv2d may use TF32
        ref = F.conv2d(x.to(dt).float().cpu().permute(0, 3, 1, 2), w.cpu().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
        ref = (ref * sc.cpu() + sh.cpu()).clamp_min(0).cuda()
        assert ((y.float() - ref).abs().max() / ref.abs().max()).item() < tol
    sub = torch.empty(2, 10, 12, 32, device="cuda")
    ops.subsample(x, sub, 2)
    assert torch.equal(sub, x[:, ::2, ::2].contiguous())
    out = torch.zeros(2, 40, 48, 48, device="cuda")
    ops.resize_bilinear(x, out, 32)
    ref = R.resize_bilinear_tf1(x.cpu(), 40, 48)
    assert (out[..., :32].cpu() - ref).abs().max() < 1e-5 and (out[..., 32:] == 0).all()


def _sepconv_ref(x16, rate, wdw, dsc, dsh, wpw16, psc, psh):
    """fp32 CPU reference of one fused branch on the 16-bit-rounded inputs (oracle ops: TF 'same' atrous depthwise,
    folded BN + ReLU, pointwise, folded BN + ReLU) -- deeplabv3p.py:47-84 with depth_activation=True."""
    from oracle import ref_ops as R
    x = x16.float().cpu()
    if rate > 0:
        t = R.depthwise_same(x, wdw.cpu().unsqueeze(-1), 1, rate)
        x = (t * dsc.cpu() + dsh.cpu()).clamp_min(0)
    y = torch.einsum("bhwc,nc->bhwn", x, wpw16.float().cpu())
    return (y * psc.cpu() + psh.cpu()).clamp_min(0)


@pytest.mark.parametrize("dt,B,H,W,C,N,rates", [
    (torch.float16, 2, 16, 64, 128, 256, (0, 1, 2, 5)),       # TH=2: row groups overlap / touch / are disjoint
    (torch.float16, 1, 64, 64, 192, 256, (0, 12, 24, 36)),    # ASPP geometry at OS=8 (rates reach past the map)
    (torch.bfloat16, 2, 32, 32, 64, 128, (6, 12, 18)),        # ASPP geometry at OS=16, TH=4
    (torch.float16, 1, 9, 20, 304, 256, (1,)),                # decoder-like: ragged W (TH=6, 120 of 128 rows), K tail 304
    (torch.float16, 1, 5, 128, 72, 64, (1, 3)),               # W=128: one row per tile, K=72 (1 full + 8-channel tail)
])
def test_sepconv_fused_matches_oracle_ops(dt, B, H, W, C, N, rates):
    """dlb_sepconv_fused_fwd (one tcgen05 kernel for all branches) vs the oracle's separate ops."""
    from deeplab_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(H * W + C)
    x = torch.randn(B, H, W, C, device="cuda", generator=g).to(dt)
    nb = len(rates)
    wdw = [torch.randn(3, 3, C, device="cuda", generator=g) / 3 for _ in rates]
    dsc = [torch.rand(C, device="cuda", generator=g) + 0.5 for _ in rates]
    dsh = [torch.randn(C, device="cuda", generator=g) * 0.2 for _ in rates]
    wpw = [(torch.randn(N, C, device="cuda", generator=g) / C ** 0.5).to(dt) for _ in rates]
    psc = [torch.rand(N, device="cuda", generator=g) + 0.5 for _ in rates]
    psh = [torch.randn(N, device="cuda", generator=g) * 0.2 for _ in rates]
    cat = torch.full((B, H, W, nb * N + 8), -7.0, device="cuda", dtype=dt)
    outs = [cat[..., i * N:(i + 1) * N] for i in range(nb)]
    idx = [i for i, r in enumerate(rates) if r > 0]
    pack = ops.sepconv_pack_dw([wdw[i] for i in idx], [dsc[i] for i in idx], [dsh[i] for i in idx], dt)
    ops.sepconv_fused_fwd(x, rates, wpw, pack, psc, psh, outs)
    torch.cuda.synchronize()
    assert (cat[..., nb * N:] == -7.0).all()          # nothing written past the branch slices
    tol = 6e-3 if dt == torch.float16 else 4e-2
    for i, r in enumerate(rates):
        # the kernel rounds taps and the depthwise result to 16 bit: feed the reference the rounded taps
        ref = _sepconv_ref(x, r, wdw[i].to(dt).float(), dsc[i], dsh[i], wpw[i], psc[i], psh[i])
        err = (outs[i].float().cpu() - ref).abs().max() / ref.abs().max()
        assert err < tol, (i, r, float(err))


def test_xception_fused_aspp_equals_unfused():
    """Whole-network check: the fused ASPP + decoder kernels against the layer-wise path on the same weights."""
    from deeplab_b200.deeplabv3p import Deeplabv3
    from oracle import network as N
    W = N.random_xception_weights(seed=11)
    H, Wd, B = 256, 256, 2
    model = Deeplabv3(weights=None, input_shape=(H, Wd, 3), backbone='xception', OS=8, compute_dtype='float16')
    _push(model, W)
    x = np.random.RandomState(2).randint(0, 256, (B, H, Wd, 3)).astype(np.float32)
    e = model.engine
    e.fused_sepconv = True
    p1 = model.predict(x, batch_size=B).copy()
    l1 = e.workspace(B, False)["logits"][..., :21].float().cpu().clone()
    e.fused_sepconv = False
    p0 = model.predict(x, batch_size=B)
    l0 = e.workspace(B, False)["logits"][..., :21].float().cpu()
    assert (l1 - l0).abs().max() <= 2e-2 * l0.abs().max()
    assert (p1.argmax(-1) == p0.argmax(-1)).mean() > 0.995
