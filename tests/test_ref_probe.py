"""Run-time upgrade of the oracle: when the reference's real third-party dependencies are importable in the
environment running the tests, pin the restated op semantics (oracle/ref_ops.py) and the densecrf
restatement against them.  In the build image none of them exists and every test here skips -- the probe itself is
always exercised."""
import numpy as np
import pytest
import torch


def test_probe_reports_environment():
    from oracle import ref_probe
    p = ref_probe.probe()
    assert set(p) == {"tensorflow", "keras", "pydensecrf", "h5py"}
    assert isinstance(ref_probe.summary(), str)


def test_tf_op_semantics_match_restatement():
    tf = pytest.importorskip("tensorflow")
    from oracle import ref_ops as R
    tf1 = tf.compat.v1
    rng = np.random.RandomState(0)
    x = rng.randn(2, 9, 11, 6).astype(np.float32)
    # legacy bilinear (deeplabv3p.py:382,418,439)
    y = tf1.image.resize_bilinear(tf.constant(x), (72, 88), align_corners=False).numpy()
    assert np.abs(y - R.resize_bilinear_tf1(torch.from_numpy(x), 72, 88).numpy()).max() < 1e-5
    # SAME padding, strided / dilated depthwise and dense convs (deeplabv3p.py:186-188,:317-321)
    wd = rng.randn(3, 3, 6, 1).astype(np.float32)
    for s, d in ((1, 1), (2, 1), (1, 2), (1, 4)):
        y = tf.nn.depthwise_conv2d(x, wd, [1, s, s, 1], "SAME", dilations=[d, d]).numpy()
        assert np.abs(y - R.depthwise_same(torch.from_numpy(x), torch.from_numpy(wd), s, d).numpy()).max() < 1e-4
    wc = rng.randn(3, 3, 6, 5).astype(np.float32)
    y = tf.nn.conv2d(x, wc, [1, 2, 2, 1], "SAME").numpy()
    assert np.abs(y - R.conv2d_same(torch.from_numpy(x), torch.from_numpy(wc), 2, 1).numpy()).max() < 1e-4
    # space_to_depth (ICNR, subpixel.py:36)
    z = rng.randn(1, 8, 8, 3).astype(np.float32)
    assert np.array_equal(tf.nn.space_to_depth(z, 2).numpy(), R.space_to_depth(torch.from_numpy(z), 2).numpy())


def test_pydensecrf_matches_c_restatement():
    pytest.importorskip("pydensecrf.densecrf")
    import scipy.ndimage as ndi
    from oracle import crf as O
    from oracle import ref_probe
    rng = np.random.RandomState(0)
    H, W, M = 64, 80, 5
    img = ndi.gaussian_filter(rng.rand(H, W, 3), (4, 4, 0))
    img = ((img - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
    un = rng.rand(M, H * W).astype(np.float32) * 4
    Q = ref_probe.pydensecrf_inference(un, img, iters=5)
    Qo = O.dense_crf(un, img, iters=5)
    assert np.abs(Q - Qo).max() < 1e-4
