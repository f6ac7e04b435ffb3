"""CPU tests of the host-side logic that needs no GPU: HDF5 reader/writer, API argument validation mirrored from the
reference, metric finishing, ICNR, padding geometry, layer naming."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_h5_writer_roundtrip(tmp_path):
    import deeplab_b200  # noqa: F401
    from deeplab_b200 import keras_h5
    rng = np.random.RandomState(0)
    layers = OrderedDict()
    layers["input_1"] = []
    for i in range(40):       # enough layers to need several symbol nodes
        layers[f"conv_{i}"] = [(f"conv_{i}/kernel:0", rng.rand(1, 1, 8, 4).astype(np.float32)),
                               (f"conv_{i}/bias:0", rng.rand(4).astype(np.float32))]
        layers[f"bn_{i}"] = [(f"bn_{i}/{n}:0", rng.rand(4).astype(np.float32)) for n in ("gamma", "beta", "moving_mean", "moving_variance")]
    layers["pred_mask"] = []
    p = str(tmp_path / "w.h5")
    keras_h5.save_keras_weights(p, layers)
    back, attrs = keras_h5.load_keras_weights(p)
    assert list(back) == list(layers)
    assert attrs["backend"] == b"tensorflow" and attrs["keras_version"] == b"2.2.4"
    for k in layers:
        assert [n for n, _ in back[k]] == [n for n, _ in layers[k]]
        for (_, a), (_, b) in zip(back[k], layers[k]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("n_layers", [257, 300, 1100])
def test_h5_writer_multi_level_btree(tmp_path, n_layers):
    """More than 256 entries per group need a second B-tree level (round-1 advice: the 299-layer Xception model could
    not be saved); both the product's reader and the oracle's independent reader must get every layer back."""
    import deeplab_b200  # noqa: F401
    from deeplab_b200 import keras_h5 as P
    from oracle import hdf5_reader as O
    rng = np.random.RandomState(n_layers)
    layers = OrderedDict()
    for i in range(n_layers):
        nm = "unit_%d_%d" % (rng.randint(0, 10 ** 6), i)
        layers[nm] = [(nm + "/kernel:0", rng.randn(1, 1, 3, 2).astype(np.float32)),
                      (nm + "/bias:0", rng.randn(2).astype(np.float32))] if i % 4 else []
    path = str(tmp_path / "big.h5")
    P.save_keras_weights(path, layers)
    for reader in (P, O):
        got, attrs = reader.load_keras_weights(path)
        assert list(got) == list(layers)
        for k, ws in layers.items():
            assert [n for n, _ in got[k]] == [n for n, _ in ws]
            for (_, a), (_, b) in zip(got[k], ws):
                assert np.array_equal(a, b)


def test_reader_roundtrips_reference_file_through_writer(tmp_path):
    import deeplab_b200  # noqa: F401
    from deeplab_b200 import keras_h5
    layers, _ = keras_h5.load_keras_weights(os.path.join(GOLD, "mobilenetv2_original.h5"))
    p = str(tmp_path / "copy.h5")
    keras_h5.save_keras_weights(p, layers)
    back, _ = keras_h5.load_keras_weights(p)
    assert list(back) == list(layers)
    assert sum(a.size for v in back.values() for _, a in v) == 2146645
    for k in layers:
        for (_, a), (_, b) in zip(back[k], layers[k]):
            assert np.array_equal(a, b)


def test_deeplabv3_argument_validation_matches_reference():
    """deeplabv3p.py:247-258: ValueError for bad `weights` / `backbone` (checked before any device work)."""
    from deeplab_b200.deeplabv3p import Deeplabv3
    with pytest.raises(ValueError, match="`weights` argument"):
        Deeplabv3(weights="imagenet")
    with pytest.raises(ValueError, match="`backbone` argument"):
        Deeplabv3(weights=None, backbone="resnet")
    with pytest.raises(RuntimeError, match="no network"):
        Deeplabv3(weights="pascal_voc")


def test_metrics_from_confusion_match_oracle():
    from deeplab_b200.model import _metrics_from_confusion
    from oracle import ref_ops as R
    rng = np.random.RandomState(3)
    B, T, C = 3, 500, 5
    y = rng.randint(0, C + 1, (B, T, 1)).astype(np.float32)
    y[1][y[1] == 2] = 0                       # a class absent from one sample
    p = rng.rand(B, T, C).astype(np.float32)
    p /= p.sum(-1, keepdims=True)
    am = p.argmax(-1)
    conf = np.zeros((B, C + 1, C), np.int64)
    for b in range(B):
        for t in range(T):
            conf[b, int(y[b, t, 0]), am[b, t]] += 1
    jac, acc = _metrics_from_confusion(conf, C)
    assert abs(jac - R.jaccard(torch.from_numpy(y), torch.from_numpy(p)).item()) < 1e-9
    assert abs(acc - R.sparse_accuracy_ignoring_last_label(torch.from_numpy(y), torch.from_numpy(p)).item()) < 1e-6


def test_icnr_weights_structure_and_oracle_agreement():
    from deeplab_b200.subpixel import icnr_weights
    from oracle import ref_ops as R
    sub = {}

    def init(shape, rng):
        sub["w"] = rng.standard_normal(shape).astype(np.float32)
        return sub["w"]

    w = icnr_weights(init=init, scale=8, shape=[1, 1, 256, 1344], seed=1)
    assert w.shape == (1, 1, 256, 1344)
    ref = R.icnr(torch.from_numpy(sub["w"]), 8).numpy()
    assert np.array_equal(w, ref)
    w1 = icnr_weights(scale=1, shape=[3, 3, 4, 8], seed=0)
    assert w1.shape == (3, 3, 4, 8)


def test_subpixel_layer_config_quirks():
    from deeplab_b200.subpixel import Subpixel
    s = Subpixel(21, 1, 8, padding="same")
    assert s.filters == 21 * 64
    assert s.compute_output_shape((None, 64, 64, 256)) == (None, 512, 512, 21)
    assert s.get_config()["filters"] == 21 * 64          # reference precedence quirk, subpixel.py:101
    with pytest.raises(NotImplementedError):
        Subpixel(21, 3, 8)


def test_same_padding_geometry_and_unary():
    from deeplab_b200.ops import tf_same_pad
    from deeplab_b200.utils import unary_from_labels
    from oracle import crf, ref_ops as R
    for args in [(512, 3, 2, 1), (513, 3, 2, 1), (64, 3, 1, 2), (64, 3, 1, 36), (33, 3, 2, 1)]:
        assert tf_same_pad(*args) == R.tf_same_pad(*args)
    lab = np.array([0, 1, 2, 2, 0])
    for zu in (True, False):
        assert np.array_equal(unary_from_labels(lab, 3, 0.7, zu), crf.unary_from_labels(lab, 3, 0.7, zu))


def test_engine_requires_cuda_and_bad_input_shape():
    import deeplab_b200  # noqa: F401
    from deeplab_b200.engine import Engine
    with pytest.raises(ValueError, match="multiples of 8"):
        Engine(input_shape=(100, 100, 3), device="cpu")
    with pytest.raises(NotImplementedError):
        Engine(alpha=0.5, device="cpu")
    e = Engine(input_shape=(64, 64, 3), head="subpixel", device="cpu")
    n = sum(p.size for r in e.layers for p in r.params if p.trainable_kind)
    assert n == 2113557 - (256 * 21 + 21) + (256 * 1344 + 1344)      # SURVEY: 2 113 557 trainable ('original')
    # the Subpixel head hides its internal column permutation
    w = np.arange(256 * 1344, dtype=np.float32).reshape(1, 1, 256, 1344)
    b = np.arange(1344, dtype=np.float32)
    e.set_layer_weights(e.head_conv, [w, b])
    w2, b2 = e.get_layer_weights(e.head_conv)
    assert np.array_equal(w, w2) and np.array_equal(b, b2)
    assert not np.array_equal(e.head_conv.params[1].data.numpy(), b)       # stored permuted


def test_keras_callbacks_follow_keras_224_rules(tmp_path):
    """The notebook's callbacks (segmentation.ipynb cell 7: ModelCheckpoint(save_best_only, save_weights_only),
    ReduceLROnPlateau, EarlyStopping, TensorBoard) against the Keras 2.2.4 rules, driven with a stub model."""
    import json
    from deeplab_b200.model import EarlyStopping, ModelCheckpoint, ReduceLROnPlateau, TensorBoard

    class Opt:
        lr = 1e-3

    class Stub:
        def __init__(self):
            self.optimizer, self.saved, self.stop_training = Opt(), [], False

        def save_weights(self, path):
            self.saved.append(path)

        def set_lr(self, lr):
            self.optimizer.lr = lr

    # ModelCheckpoint: 'Jaccard' in the monitor name -> mode max (auto); saves only on improvement; formats the path
    m = Stub()
    ck = ModelCheckpoint(str(tmp_path / "w.{epoch:02d}-{val_Jaccard:.2f}.h5"), monitor="val_Jaccard", save_best_only=True,
                         save_weights_only=True)
    ck.set_model(m)
    for ep, v in enumerate([0.30, 0.25, 0.41, 0.41, float("nan"), 0.50]):
        ck.on_epoch_end(ep, {"val_Jaccard": v, "val_loss": 1.0})
    assert [os.path.basename(p) for p in m.saved] == ["w.01-0.30.h5", "w.03-0.41.h5", "w.06-0.50.h5"]
    # ... and mode min for a loss
    m = Stub()
    ck = ModelCheckpoint(str(tmp_path / "l.h5"), monitor="val_loss", save_best_only=True)
    ck.set_model(m)
    for ep, v in enumerate([1.0, 1.2, 0.9]):
        ck.on_epoch_end(ep, {"val_loss": v})
    assert len(m.saved) == 2
    # without save_best_only every epoch is written
    m = Stub()
    ck = ModelCheckpoint(str(tmp_path / "e{epoch}.h5"))
    ck.set_model(m)
    for ep in range(3):
        ck.on_epoch_end(ep, {})
    assert len(m.saved) == 3

    # ReduceLROnPlateau(monitor='val_loss', factor=0.2, patience=2, min_lr=1e-5): improvement = cur < best - min_delta
    m = Stub()
    rl = ReduceLROnPlateau(monitor="val_loss", factor=0.2, patience=2, min_lr=1e-5, min_delta=1e-4)
    rl.set_model(m)
    lrs = []
    for ep, v in enumerate([1.0, 0.99995, 1.1, 0.8, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9]):
        rl.on_epoch_end(ep, {"val_loss": v})
        lrs.append(m.optimizer.lr)
    # epochs 1, 2 do not improve by min_delta -> reduce after epoch 2; 0.8 improves; then every 2 stale epochs
    assert np.allclose(lrs[:4], [1e-3, 1e-3, 2e-4, 2e-4])
    assert np.isclose(lrs[5], 4e-5) and np.isclose(lrs[7], 1e-5) and np.isclose(lrs[-1], 1e-5)   # clamped at min_lr

    # EarlyStopping(patience=3): stops after 3 epochs without improvement
    m = Stub()
    es = EarlyStopping(monitor="val_loss", patience=3)
    es.set_model(m)
    stopped_at = None
    for ep, v in enumerate([1.0, 0.9, 0.95, 0.93, 0.91, 0.5]):
        es.on_epoch_end(ep, {"val_loss": v})
        if m.stop_training and stopped_at is None:
            stopped_at = ep
    assert stopped_at == 4

    tb = TensorBoard(log_dir=str(tmp_path / "logs"))
    tb.on_epoch_end(0, {"loss": 1.5, "val_Jaccard": 0.25})
    rec = json.loads(open(tmp_path / "logs" / "scalars.jsonl").read().strip())
    assert rec == {"epoch": 0, "loss": 1.5, "val_Jaccard": 0.25}
