"""Dense-CRF CUDA path (dlb_crf_inference through utils.dense_crf / do_crf) vs the C restatement of densecrf.

Tolerance: north_star asks CRF within 1e-2 (abs on the marginals Q); the GPU lattice is the same lattice, so the
observed error is float-summation-order level and the tests hold 2e-3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _image(H, W, seed):
    rng = np.random.RandomState(seed)
    img = rng.rand(H, W, 3)
    import scipy.ndimage as ndi
    img = ndi.gaussian_filter(img, (6, 6, 0))
    img = (img - img.min()) / (img.max() - img.min())
    return (img * 255).astype(np.uint8)


@pytest.mark.parametrize("H,W,M,iters", [(64, 64, 3, 5), (96, 160, 21, 5), (128, 128, 21, 10), (50, 70, 2, 3)])
def test_dense_crf_matches_oracle(H, W, M, iters):
    from deeplab_b200.utils import dense_crf
    from oracle import crf as O
    rng = np.random.RandomState(H + W + M)
    logits = rng.randn(M, H * W).astype(np.float32) * 3
    un = -(logits - np.log(np.exp(logits).sum(0, keepdims=True)))
    img = _image(H, W, M)
    Qref = O.dense_crf(un, img, iters=iters)
    Q = dense_crf(un, img, iters=iters).cpu().numpy()
    assert np.abs(Q.sum(0) - 1).max() < 1e-4
    assert np.abs(Q - Qref).max() < 2e-3
    assert (Q.argmax(0) == Qref.argmax(0)).mean() > 0.999


def test_single_terms_and_batch():
    from deeplab_b200.utils import dense_crf
    from oracle import crf as O
    H, W, M = 48, 64, 5
    rng = np.random.RandomState(0)
    un = rng.rand(2, M, H * W).astype(np.float32) * 4
    img = np.stack([_image(H, W, 1), _image(H, W, 2)])
    for kw in (dict(compat_bilat=0.0), dict(compat_gauss=0.0)):
        Q, mp = dense_crf(un, img, iters=4, return_map=True, **kw)
        for b in range(2):
            ref = O.dense_crf(un[b], img[b], iters=4, compat_g=kw.get("compat_gauss", 3.0), compat_b=kw.get("compat_bilat", 10.0))
            assert np.abs(Q[b].cpu().numpy() - ref).max() < 2e-3
            assert (mp[b].cpu().numpy() == ref.argmax(0)).mean() > 0.999


def test_do_crf_matches_reference_semantics():
    """utils.do_crf (utils.py:74-91) incl. the zero_unsure=True label wrap, vs the literal oracle restatement."""
    from deeplab_b200.utils import do_crf
    from oracle import crf as O
    H, W = 96, 96
    img = _image(H, W, 3)
    mask = np.zeros((H, W), np.int32)
    mask[20:70, 30:80] = 15
    mask[5:15, 5:40] = 7
    for zu in (False, True):
        a = do_crf(img, mask, zero_unsure=zu)
        b = O.do_crf(img, mask, zero_unsure=zu)
        assert a.shape == (H, W) and (a == b).mean() > 0.999


def test_crf_from_device_probabilities_and_batch_of_eight():
    """soft unaries straight from pixel-major class probabilities on the device (-U = log p), no
    argmax -> unary_from_labels round trip; and the batched launch path (8 images, shared Gaussian lattice, per-image
    bilateral lattices) equals eight single-image runs of the oracle."""
    from deeplab_b200.utils import dense_crf
    from oracle import crf as O
    H, W, M, B = 64, 80, 21, 8
    rng = np.random.RandomState(3)
    logits = rng.randn(B, H * W, M).astype(np.float32) * 2
    probs = np.exp(logits - logits.max(-1, keepdims=True))
    probs /= probs.sum(-1, keepdims=True)
    imgs = np.stack([_image(H, W, 10 + b) for b in range(B)])
    Q, mp = dense_crf(None, torch.from_numpy(imgs).cuda(), iters=5, return_map=True, probs=torch.from_numpy(probs).cuda())
    assert Q.shape == (B, M, H * W) and mp.shape == (B, H * W)
    for b in range(B):
        ref = O.dense_crf(-np.log(probs[b].T), imgs[b], iters=5)
        assert np.abs(Q[b].cpu().numpy() - ref).max() < 2e-3
        assert (mp[b].cpu().numpy() == ref.argmax(0)).mean() > 0.999
