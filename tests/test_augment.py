"""SegmentationGenerator augmentations (reference utils.py:319-365): the device kernels against cv2 ITSELF -- OpenCV is
the one third-party dependency of the reference that exists in this image, so this parity is pinned to the true
dependency, not to a restatement.  The host-side restatements of cv2.getRotationMatrix2D / invertAffineTransform / the
gamma table are pinned on the CPU; the GPU test replays the reference's call sequence literally with cv2."""
import random

import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")


def _reference_getitem(image, label, p, n_classes):
    """utils.py:319-365 + 388-399 for one decoded image, executed with cv2 exactly as the reference does."""
    from oracle import ref_ops as R
    labels = np.unique(label)
    if p.blur_ksize:
        image = cv2.GaussianBlur(image, (p.blur_ksize, p.blur_ksize), 0)
    if p.hflip:
        image, label = cv2.flip(image, 1), cv2.flip(label, 1)
    if p.vflip:
        image, label = cv2.flip(image, 0), cv2.flip(label, 0)
    if p.gamma is not None:
        table = np.array([((i / 255.0) ** p.gamma) * 255 for i in np.arange(0, 256)]).astype(np.uint8)
        image = cv2.LUT(image, table)
    if p.warp:
        M = cv2.getRotationMatrix2D((image.shape[1] // 2, image.shape[0] // 2), p.angle, p.scale)
        image = cv2.warpAffine(image, M, (image.shape[1], image.shape[0]))
        label = cv2.warpAffine(label, M, (label.shape[1], label.shape[0]))
    label = label.astype("int32")
    for j in np.setxor1d(np.unique(label), labels):
        label[label == j] = n_classes
    y = label.flatten()
    y[y > (n_classes - 1)] = n_classes
    yr, sw = R.generator_labels_and_weights(y, n_classes)
    return image.astype(np.float32), yr.astype(np.float32), sw


def test_host_restatements_match_cv2():
    import deeplab_b200  # noqa: F401
    from deeplab_b200.utils import draw_augment_params, gamma_lut, invert_affine, rotation_matrix_2d
    rng = np.random.RandomState(0)
    for _ in range(200):
        c = (int(rng.randint(1, 700)), int(rng.randint(1, 700)))
        ang, sc = float(rng.randn() * 20), float(1 + rng.randn() * 0.2)
        M = cv2.getRotationMatrix2D(c, ang, sc)
        Mr = rotation_matrix_2d(c, ang, sc)
        assert np.array_equal(M, Mr), (M, Mr)
        assert np.array_equal(cv2.invertAffineTransform(M), invert_affine(Mr))
    for f in (0.5, 0.83, 1.0, 1.3, 2.2):
        ref = np.array([((i / 255.0) ** f) * 255 for i in np.arange(0, 256)]).astype(np.uint8)
        assert np.array_equal(ref, gamma_lut(f))
    # the draw order of the reference's __getitem__: same seed, same decisions
    r1, r2 = random.Random(5), random.Random(5)
    p = draw_augment_params(r1, blur=5, horizontal_flip=True, vertical_flip=False, brightness=0.3, rotation=False, zoom=0.1)
    blur = bool(5 and r2.randint(0, 1))
    hf = bool(r2.randint(0, 1))
    factor = 1.0 + r2.gauss(mu=0.0, sigma=0.3)
    if r2.randint(0, 1):
        factor = 1.0 / factor
    scale = r2.gauss(mu=1.0, sigma=0.1)
    assert (p.blur_ksize == 5) == blur and p.hflip == hf and p.gamma == factor and p.scale == scale and p.angle == 0.0 and p.warp


@pytest.mark.gpu
@pytest.mark.parametrize("H,W", [(64, 96), (320, 320), (512, 512)])
def test_augment_batch_bit_exact_with_cv2(H, W):
    import scipy.ndimage as ndi
    import deeplab_b200  # noqa: F401
    from deeplab_b200.utils import AugmentParams, augment_batch, draw_augment_params
    rng = np.random.RandomState(H + W)
    B, n_classes = 6, 21
    imgs = np.stack([(ndi.gaussian_filter(rng.rand(H, W, 3), (2, 2, 0)) * 255).astype(np.uint8) for _ in range(B)])
    labs = np.zeros((B, H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        for _ in range(4):
            cy, cx, r = rng.randint(0, H), rng.randint(0, W), rng.randint(H // 8, H // 2)
            m = (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
            labs[b][m] = rng.randint(1, n_classes)
        labs[b][rng.rand(H, W) < 0.01] = 255                      # VOC void pixels
    pr = random.Random(11)
    params = [draw_augment_params(pr, blur=5, horizontal_flip=True, vertical_flip=True, brightness=0.3, rotation=5.0, zoom=0.1)
              for _ in range(B - 2)]
    params.append(AugmentParams())                                  # identity
    params.append(AugmentParams(blur_ksize=7, hflip=True, gamma=0.7, angle=-17.0, scale=1.31, warp=True))
    X, Y, SWd = augment_batch(imgs, labs, params, n_classes)
    X, Y, SW = X.cpu().numpy(), Y.cpu().numpy(), SWd["pred_mask"].cpu().numpy()
    for b in range(B):
        xr, yr, swr = _reference_getitem(imgs[b].copy(), labs[b].copy(), params[b], n_classes)
        assert np.array_equal(X[b], xr), (b, np.abs(X[b] - xr).max())
        assert np.array_equal(Y[b, :, 0], yr), (b, (Y[b, :, 0] != yr).mean())
        assert np.array_equal(SW[b], swr), b
