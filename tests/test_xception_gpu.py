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
