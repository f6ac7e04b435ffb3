"""CPU tests that pin the ORACLE itself: against the committed golden vectors derived from the reference's own
artefacts (weights/*.h5, examples/*.JPG) and against closed-form / brute-force properties.  No GPU."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_weight_file_known_answers():
    """SURVEY 8c KAT (1): 165 layers, 272 datasets, 2 146 645 floats, head (1,1,256,21)+(21,), Keras 2.2.4/TF."""
    from oracle.hdf5_reader import load_keras_weights
    layers, attrs = load_keras_weights(os.path.join(GOLD, "mobilenetv2_original.h5"))
    assert len(layers) == 165
    assert sum(len(v) for v in layers.values()) == 272
    assert sum(a.size for v in layers.values() for _, a in v) == 2146645
    assert attrs["backend"] == b"tensorflow" and attrs["keras_version"] == b"2.2.4"
    names = list(layers)
    assert names[2:4] == ["Conv", "Conv_BN"] and names[-1] == "pred_mask" and names[-4] == "conv_upsample"
    k, b = layers["conv_upsample"]
    assert k[1].shape == (1, 1, 256, 21) and b[1].shape == (21,)
    assert layers["expanded_conv_depthwise"][0][1].shape == (3, 3, 32, 1)
    d = np.load(os.path.join(GOLD, "subpixel_delta.npz"))
    assert d["subpixel_1::0"].shape == (1, 1, 256, 1344) and d["subpixel_1::1"].shape == (1344,)


def test_oracle_reproduces_golden_logits():
    """The restatement, re-run here, reproduces the committed outputs (guards oracle drift)."""
    from oracle import network as N
    g = np.load(os.path.join(GOLD, "golden_mnv2.npz"))
    W = N.weights_from_h5(os.path.join(GOLD, "mobilenetv2_original.h5"))
    x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, (1, 512, 512, 3)).astype(np.float32))
    with torch.no_grad():
        logits, probs, _ = N.deeplabv3_forward(W, x)
    assert np.abs(logits.numpy() - g["logits"]).max() < 1e-4 * np.abs(g["logits"]).max()
    assert (probs.argmax(-1).numpy().reshape(512, 512) == g["argmax"]).mean() > 0.9999


def test_fp64_vs_fp32_self_consistency():
    from oracle import network as N
    W = N.random_mobilenetv2_weights(seed=0)
    x = torch.from_numpy(np.random.RandomState(1).randint(0, 256, (1, 96, 96, 3)).astype(np.float32))
    with torch.no_grad():
        l32, _, _ = N.deeplabv3_forward(W, x)
        W64 = {k: [t.double() for t in v] for k, v in W.items()}
        l64, _, _ = N.deeplabv3_forward(W64, x.double())
    assert (l32.double() - l64).abs().max() < 1e-4 * l64.abs().max()


def test_phase_shift_closed_form_vs_literal_and_not_depth_to_space():
    from oracle import ref_ops as R
    x = torch.arange(2 * 3 * 5 * 7 * 16, dtype=torch.float32).reshape(2, 3, 5, 7 * 16)
    a, b = R.phase_shift_literal(x, 4), R.phase_shift(x, 4)
    assert torch.equal(a, b)
    # out[n, a*r+j, b*r+i, k] = in[n, a, b, k*r*r + i*r + j]
    assert a[1, 2 * 4 + 3, 4 * 4 + 1, 5] == x[1, 2, 4, 5 * 16 + 1 * 4 + 3]
    ps = torch.pixel_shuffle(x.permute(0, 3, 1, 2), 4).permute(0, 2, 3, 1)
    assert not torch.equal(a, ps)
    perm = R.subpixel_column_perm(7, 4)
    assert sorted(perm.tolist()) == list(range(7 * 16))
    # permuted columns -> contiguous (i, k) runs per (row, jj)
    y = x[..., perm].reshape(2, 3, 5, 4, 4 * 7).permute(0, 1, 3, 2, 4).reshape(2, 12, 20, 7)
    assert torch.equal(y, a)


def test_tf_same_padding_and_legacy_bilinear():
    from oracle import ref_ops as R
    assert R.tf_same_pad(512, 3, 2, 1) == (256, 0, 1)        # stride-2 even input: extra pixel AFTER
    assert R.tf_same_pad(513, 3, 2, 1) == (257, 1, 1)
    assert R.tf_same_pad(64, 3, 1, 4) == (64, 4, 4)
    assert R.tf_same_pad(64, 3, 1, 36) == (64, 36, 36)
    x = torch.arange(4.0).reshape(1, 1, 4, 1)
    up = R.resize_bilinear_tf1(x, 1, 32)[0, 0, :, 0]
    # src = dst/8: weights k/8, last 7 outputs clamp to the edge value
    assert torch.allclose(up[:9], torch.arange(9.0) / 8)
    assert torch.all(up[24:] == 3.0)
    # differs from the half-pixel-centre resize torch implements
    tr = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2), size=(1, 32), mode="bilinear", align_corners=False)
    assert not torch.allclose(tr[0, 0, 0], up)


def test_loss_and_metrics_definitions():
    from oracle import ref_ops as R
    y = torch.tensor([[[0.], [1.], [2.], [3.]]])                 # 3 == void for 3 classes
    p = torch.tensor([[[.7, .2, .1], [.1, .8, .1], [.3, .3, .4], [.2, .5, .3]]])
    l = R.sparse_crossentropy_ignoring_last_label(y, p)
    assert torch.allclose(l[0], torch.tensor([-np.log(.7), -np.log(.8), -np.log(.4), 0.0]).float(), atol=1e-6)
    sw = torch.tensor([[1., 0., 2., 1.]])
    tot = R.keras_weighted_loss(y, p, sw)
    assert abs(tot.item() - ((-np.log(.7) - 2 * np.log(.4)) / 4 / 0.75)) < 1e-6
    assert abs(R.sparse_accuracy_ignoring_last_label(y, p).item() - 1.0) < 1e-6
    # the void pixel is predicted as class 1 and counts in that class's union (utils.py:146-147): (1 + 1/2 + 1) / 3
    assert abs(R.jaccard(y, p).item() - 2.5 / 3) < 1e-6
    assert abs(R.notebook_miou(np.array([0, 0, 1, 1]), np.array([0, 1, 1, 1])) - (0.5 + 2 / 3) / 2) < 1e-9


def test_keras_adam_first_steps():
    from oracle import ref_ops as R
    p, m, v = torch.ones(3), torch.zeros(3), torch.zeros(3)
    g = torch.tensor([1.0, -2.0, 0.5])
    p1, m, v = R.keras_adam(p, g, m, v, 0, lr=0.1, eps=1e-8, decay=0.0)
    assert torch.allclose(p1, torch.tensor([0.9, 1.1, 0.9]), atol=1e-6)      # first step = lr * sign(g)
    p2, _, _ = R.keras_adam(p1, g, m, v, 1, lr=0.1, eps=1e-8, decay=0.5)
    assert torch.allclose(p1 - p2, torch.tensor([1., -1., 1.]) * 0.1 / 1.5, atol=1e-6)


def test_permutohedral_oracle_properties():
    """barycentric weights >= 0 and sum to 1; K1 gain 0.882 of the exact Gaussian mass in the
    interior (d=2); normalised filter within ~1e-2 of the exact normalised Gaussian."""
    from oracle import crf
    H = W = 40
    ys, xs = np.mgrid[0:H, 0:W]
    f = np.stack([xs / 3.0, ys / 3.0], -1).reshape(-1, 2).astype(np.float32)
    v = np.random.RandomState(0).rand(H * W, 2).astype(np.float32)
    out, M, off, bary = crf.lattice_filter(f, np.concatenate([v, np.ones((H * W, 1), np.float32)], 1))
    assert bary.min() > -1e-6 and np.abs(bary.sum(1) - 1).max() < 1e-5
    assert off.min() >= 0 and off.max() < M
    d2 = ((f[:, None, :] - f[None, :, :]) ** 2).sum(-1)
    K = np.exp(-0.5 * d2)
    ratio = (out[:, 2] / K.sum(1)).reshape(H, W)[12:-12, 12:-12]
    assert abs(ratio.mean() - 0.882) < 0.01 and ratio.std() < 0.01
    assert np.abs(out[:, :2] / out[:, 2:3] - (K @ v) / K.sum(1, keepdims=True)).max() < 2e-2
    # d = 5 weights are a partition of unity too
    f5 = np.random.RandomState(1).rand(500, 5).astype(np.float32) * 10
    _, _, _, b5 = crf.lattice_filter(f5, np.ones((500, 1), np.float32))
    assert b5.min() > -1e-5 and np.abs(b5.sum(1) - 1).max() < 1e-5


def test_crf_oracle_semantics():
    from oracle import crf
    # one label: Q == 1 ; uniform image + uniform unary: Q stays uniform
    img = np.full((16, 16, 3), 128, np.uint8)
    Q = crf.dense_crf(np.zeros((3, 256), np.float32), img, iters=3)
    assert np.abs(Q - 1 / 3).max() < 1e-5
    U = crf.unary_from_labels(np.array([0, 1, 2, 1]), 3, 0.7, zero_unsure=False)
    assert abs(U[0, 0] + np.log(0.7)) < 1e-6 and abs(U[1, 0] + np.log(0.15)) < 1e-6
    Uz = crf.unary_from_labels(np.array([0, 1, 2, 1]), 3, 0.7, zero_unsure=True)
    assert np.allclose(Uz[:, 0], -np.log(1 / 3)) and abs(Uz[0, 1] + np.log(0.7)) < 1e-6   # label 1 -> row 0
    # smoothing: an isolated wrong pixel inside a uniform region gets corrected
    mask = np.zeros((24, 24), np.int32)
    mask[:, 12:] = 1
    mask[5, 3] = 1
    img2 = np.zeros((24, 24, 3), np.uint8)
    img2[:, 12:] = 255                                   # colour edge coincides with the label edge
    out = crf.do_crf(img2, mask, zero_unsure=False)
    assert out[5, 3] == 0 and out[5, 20] == 1 and (out[:, 12:] == 1).all()


def test_icnr_matches_reference_sequence():
    from oracle import ref_ops as R
    sub = torch.randn(1, 1, 8, 3)
    w = R.icnr(sub, 4)
    assert w.shape == (1, 1, 8, 48)
    # every group of scale^2 consecutive-by-C' channels repeats the sub-kernel: w[..., (i*s+j)*C' + k] == sub[..., k]
    for q in range(16):
        assert torch.equal(w[0, 0, :, q * 3:(q + 1) * 3], sub[0, 0])


def test_label_weights_restatement_matches_sklearn():
    """The oracle's class-weight arithmetic is pinned against the third-party function the reference calls
    (utils.py:393: sklearn.utils.class_weight.compute_class_weight('balanced', ...)), which is installed here."""
    from sklearn.utils import class_weight
    from oracle import ref_ops as R
    rng = np.random.RandomState(3)
    for n_classes, present in ((21, [0, 3, 7, 21, 255]), (2, [0, 1]), (21, [255]), (5, [2])):
        lab = rng.choice(present, size=(37, 29))
        y, sw = R.generator_labels_and_weights(lab, n_classes)
        yy = lab.flatten().astype(np.int64)
        yy[yy > n_classes - 1] = n_classes
        assert np.array_equal(y, yy)
        filt = yy[yy != n_classes]
        exp = np.zeros(yy.shape, dtype=np.float32)
        u = np.unique(filt)
        if len(u):
            cw = class_weight.compute_class_weight(class_weight="balanced", classes=u, y=filt)
            for c, w in zip(u, cw):
                np.putmask(exp, yy == c, w)
        assert np.array_equal(sw, exp)


def test_calculate_iou_conf_matches_literal_loop():
    from oracle import ref_ops as R
    rng = np.random.RandomState(4)
    C = 5
    pred = rng.randint(0, C, size=400)
    lab = rng.randint(0, C + 1, size=400)
    conf = np.zeros((C, C))
    for p, l in zip(pred, lab):          # the notebook's loop, verbatim semantics
        if l == C:
            continue
        if l < C and p < C:
            conf[l - 1, p - 1] += 1
    assert np.array_equal(R.calculate_iou_conf(pred, lab, C), conf)


# ------------------------------------------------------------------------------------------------------------------
# oracle independence (round-1 review): a second restatement written without anything shared with oracle/network.py,
# an oracle-owned HDF5 reader, and a gradient check that does not use autograd.
# ------------------------------------------------------------------------------------------------------------------
def _np_weights(W):
    return {k: [t.numpy() for t in v] for k, v in W.items()}


def _dbl(W):
    return {k: [t.double() for t in v] for k, v in W.items()}


@pytest.mark.parametrize("net,training", [("original", False), ("original", True), ("subpixel", False), ("subpixel", True)])
def test_second_restatement_agrees_mobilenetv2(net, training):
    """oracle/network.py (torch NHWC, F.conv2d) vs oracle/network_np.py (numpy fp64 NCHW, tap loops, direct-index
    resize / phase shift) on a non-square input: fp64 round-off level."""
    from oracle import network as N
    from oracle import network_np as NP
    head = "conv_upsample" if net == "original" else "subpixel_1"
    W = N.random_mobilenetv2_weights(seed=3, head=head, head_filters=21 * 64 if net == "subpixel" else 21)
    x = np.random.RandomState(1).randint(0, 256, (2, 64, 96, 3)).astype(np.float32)
    with torch.no_grad():
        lg, pr, ctx = N.deeplabv3_forward(_dbl(W), torch.from_numpy(x).double(), net=net, training=training)
    net2 = NP.Net(_np_weights(W), training=training)
    lg2, pr2 = net2.forward(x, net=net)
    assert np.abs(lg.numpy() - lg2).max() <= 1e-9 * np.abs(lg2).max()
    assert np.abs(pr.numpy() - pr2).max() <= 1e-10
    if training:      # batch statistics feeding the moving averages
        for name, (mu, var, cnt) in ctx.bn_batch_stats.items():
            mu2, var2, cnt2 = net2.batch_stats[name]
            assert cnt == cnt2
            assert np.abs(mu.numpy() - mu2).max() <= 1e-9 * max(1.0, np.abs(mu2).max()), name
            assert np.abs(var.numpy() - var2).max() <= 1e-9 * max(1.0, np.abs(var2).max()), name


@pytest.mark.parametrize("OS", [8, 16])
def test_second_restatement_agrees_xception(OS):
    """Xception path (explicit-padding stride-2 SepConvs, 1x1 stride-2 shortcuts, atrous ASPP, decoder)."""
    from oracle import network as N
    from oracle import network_np as NP
    W = N.random_xception_weights(seed=5)
    x = np.random.RandomState(2).randint(0, 256, (1, 64, 96, 3)).astype(np.float32)
    with torch.no_grad():
        lg, pr, _ = N.deeplabv3_forward(_dbl(W), torch.from_numpy(x).double(), backbone="xception", OS=OS)
    lg2, pr2 = NP.Net(_np_weights(W)).forward(x, backbone="xception", OS=OS)
    assert np.abs(lg.numpy() - lg2).max() <= 1e-9 * np.abs(lg2).max()
    assert np.abs(pr.numpy() - pr2).max() <= 1e-10


def test_second_restatement_reproduces_golden_logits_at_512():
    """config 1 at its own size through the second restatement and the oracle-owned reader: the committed golden logits
    (made by network.py + the product's reader) are reproduced by code that shares neither."""
    from oracle import network_np as NP
    from oracle.hdf5_reader import load_keras_weights
    g = np.load(os.path.join(GOLD, "golden_mnv2.npz"))
    layers, _ = load_keras_weights(os.path.join(GOLD, "mobilenetv2_original.h5"))
    W = {k: [a for _, a in v] for k, v in layers.items() if v}
    x = np.random.RandomState(0).randint(0, 256, (1, 512, 512, 3)).astype(np.float32)
    lg, pr = NP.Net(W).forward(x, head="conv_upsample")
    assert np.abs(lg - g["logits"]).max() < 1e-4 * np.abs(g["logits"]).max()      # goldens are fp32
    assert (pr.argmax(-1).reshape(512, 512) == g["argmax"]).mean() > 0.9999


def test_autograd_gradient_matches_finite_differences_of_second_restatement():
    """Backward oracle pinned without autograd: d loss / d eps along random parameter directions, by central finite
    differences of the fp64 numpy restatement's training loss, equals <autograd gradient of network.py, direction>."""
    from oracle import network as N
    from oracle import network_np as NP
    from oracle import train as T
    W = N.random_mobilenetv2_weights(seed=8, head="conv_upsample")
    rng = np.random.RandomState(0)
    B, H, Wd = 2, 32, 32
    x = rng.randint(0, 256, (B, H, Wd, 3)).astype(np.float32)
    y = rng.randint(0, 22, (B, H * Wd, 1)).astype(np.float32)           # label 21 = void
    sw = rng.uniform(0.5, 2.0, (B, H * Wd)).astype(np.float32)
    sw[rng.uniform(size=sw.shape) < 0.1] = 0.0
    _, grads, _, _ = T.loss_and_grads(W, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw),
                                      dtype=torch.float64)
    Wn = _np_weights(W)
    groups = [["Conv", "expanded_conv_depthwise", "expanded_conv_1_expand_BN"],
              ["expanded_conv_7_depthwise", "expanded_conv_7_project", "expanded_conv_13_project_BN"],
              ["aspp0", "image_pooling", "concat_projection", "conv_upsample", "image_pooling_BN"]]
    for names in groups:
        dirs, dot = {}, 0.0
        for n in names:
            for i, g in grads[n].items():
                d = rng.standard_normal(Wn[n][i].shape)
                d *= 1.0 / max(np.linalg.norm(d), 1e-12)
                dirs[(n, i)] = d
                dot += float((g.numpy().reshape(d.shape) * d).sum())
        # the loss is piecewise smooth (about 10^5 ReLU6 kinks): a step of 1e-4 crosses enough of them to be off by
        # tens of percent in the early layers; steps of 1e-8..1e-9 usually cross none and leave ~1e-7 of fp64
        # round-off.  The intervals are nested, so the best of three steps is taken.
        errs = []
        for eps in (1e-8, 3e-9, 1e-9):
            vals = []
            for sgn in (+1, -1):
                Wp = {k: [a.astype(np.float64).copy() for a in v] for k, v in Wn.items()}
                for (n, i), d in dirs.items():
                    Wp[n][i] += sgn * eps * d
                vals.append(NP.training_loss(Wp, x, y, sw))
            errs.append(abs((vals[0] - vals[1]) / (2 * eps) - dot))
        assert min(errs) <= 1e-5 * abs(dot) + 2e-6, (names, errs, dot)


def test_oracle_reader_is_independent_and_agrees_with_product_reader():
    """oracle/hdf5_reader.py shares no code with deeplab_b200.keras_h5; both parse the reference's file identically."""
    import inspect
    import deeplab_b200  # noqa: F401
    from deeplab_b200 import keras_h5 as P
    from oracle import hdf5_reader as O
    src = inspect.getsource(O)
    assert "importlib" not in src and "deeplab_b200 import" not in src
    path = os.path.join(GOLD, "mobilenetv2_original.h5")
    lo, ao = O.load_keras_weights(path)
    lp, ap = P.load_keras_weights(path)
    assert list(lo) == list(lp)
    for k in lo:
        assert [n for n, _ in lo[k]] == [n for n, _ in lp[k]]
        for (_, a), (_, b) in zip(lo[k], lp[k]):
            assert a.dtype == b.dtype == np.float32 and np.array_equal(a, b)
    assert ao["backend"] == ap["backend"] and ao["keras_version"] == ap["keras_version"]
