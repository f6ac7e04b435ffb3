"""Network-level parity of the CUDA path (through the reference-facing API) against the oracle.

  * config 1 (BASELINE.json): Deeplabv3(backbone='mobilenetv2', input_shape=(512,512,3), classes=21, OS=16) with the
    reference's weights, fp32 mode vs the committed golden logits: 1e-3 relative (north_star tolerance).
  * 16-bit tensor-core mode vs the oracle (looser: storage rounding at every layer).
  * training step (fwd + bwd + Keras Adam + BN moving averages) vs the oracle's autograd step on seeded inputs.
Tolerances are written at each assert.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _push_weights(model, W):
    """oracle weight dict (Keras layer name -> list) into the model, by layer name."""
    for l in model.layers:
        if l.name in W:
            l.set_weights([w.detach().float().numpy() for w in W[l.name]])


def _pull_weights(model):
    return {l.name: [torch.from_numpy(a) for a in l.get_weights()] for l in model.layers if l._rec is not None}


# ----------------------------------------------------------------------------------------------- inference
def test_config1_fp32_matches_golden_logits():
    from deeplab_b200.deeplabv3p import Deeplabv3
    g = np.load(os.path.join(GOLD, "golden_mnv2.npz"))
    model = Deeplabv3(weights=None, input_shape=(512, 512, 3), classes=21, backbone='mobilenetv2', OS=16,
                      compute_dtype='float32')
    model.load_weights(os.path.join(GOLD, "mobilenetv2_original.h5"))     # topological load (head: conv_upsample)
    x = np.random.RandomState(0).randint(0, 256, (1, 512, 512, 3)).astype(np.float32)
    probs = model.predict(x)
    assert probs.shape == (1, 512 * 512, 21)
    ws = model.engine.workspace(1, False)
    logits = ws["logits"][..., :21].cpu().numpy()
    ref = g["logits"]
    # north_star: within 1e-3 rel fp32:  |a-b| <= 1e-3 * max|b|
    assert np.abs(logits - ref).max() <= 1e-3 * np.abs(ref).max()
    am = probs.argmax(-1).reshape(512, 512)
    assert (am == g["argmax"]).mean() > 0.9995
    assert abs(probs.max(-1).mean() - float(g["prob_max_mean"])) < 1e-4
    np.testing.assert_allclose(probs.sum(-1), 1.0, atol=1e-5)


@pytest.mark.parametrize("dtype,tol,agree", [("float16", 3e-2, 0.99), ("bfloat16", 2e-1, 0.95)])
def test_config1_16bit_tensor_core_path(dtype, tol, agree):
    from deeplab_b200.deeplabv3p import Deeplabv3
    g = np.load(os.path.join(GOLD, "golden_mnv2.npz"))
    model = Deeplabv3(weights=None, input_shape=(512, 512, 3), compute_dtype=dtype)
    model.load_weights(os.path.join(GOLD, "mobilenetv2_original.h5"))
    x = np.random.RandomState(0).randint(0, 256, (1, 512, 512, 3)).astype(np.float32)
    probs = model.predict(x)
    logits = model.engine.workspace(1, False)["logits"][..., :21].cpu().numpy()
    assert np.abs(logits - g["logits"]).max() <= tol * np.abs(g["logits"]).max()
    assert (probs.argmax(-1).reshape(512, 512) == g["argmax"]).mean() > agree


def test_example_figures_semantic_known_answer():
    """weak KAT on the reference's own example figures: dominant non-background classes."""
    from deeplab_b200.deeplabv3p import Deeplabv3
    ex = np.load(os.path.join(GOLD, "example_crops.npz"))
    model = Deeplabv3(weights=None, input_shape=(512, 512, 3), compute_dtype='float16')
    model.load_weights(os.path.join(GOLD, "mobilenetv2_original.h5"))
    expect = {"exp1": 1, "exp3": 15, "exp4": 12}          # airplane / person(+sheep) / dog
    for name, cls in expect.items():
        p = model.predict(ex[name][None].astype(np.float32))
        c, n = np.unique(p.argmax(-1), return_counts=True)
        top = [int(k) for k in c[np.argsort(-n)] if k != 0]
        assert top and top[0] == cls and top[0] == int(ex[f"dom_{name}_original"][0])


def test_batched_inference_matches_oracle_small():
    """B=3, 128x192 (non-square), fp32 mode, seeded random weights with non-trivial BN stats, vs the live oracle."""
    from deeplab_b200.deeplabv3p import Deeplabv3
    from oracle import network as N
    W = N.random_mobilenetv2_weights(seed=3)
    model = Deeplabv3(weights=None, input_shape=(128, 192, 3), compute_dtype='float32')
    _push_weights(model, W)
    x = np.random.RandomState(1).randint(0, 256, (3, 128, 192, 3)).astype(np.float32)
    probs = model.predict(x, batch_size=3)
    with torch.no_grad():
        _, pref, _ = N.deeplabv3_forward(W, torch.from_numpy(x))
    assert rel(probs, pref) < 1e-3


def test_subpixel_model_inference_matches_oracle():
    from deeplab_b200.utils import SegModel
    from oracle import network as N
    d = np.load(os.path.join(GOLD, "subpixel_delta.npz"))
    W = N.weights_from_h5(os.path.join(GOLD, "mobilenetv2_original.h5"))
    head = N.find_head_layer(W)
    Ws = type(W)()
    for k, v in W.items():
        if k == head:
            Ws["subpixel_1"] = [torch.from_numpy(d["subpixel_1::0"]), torch.from_numpy(d["subpixel_1::1"])]
        elif k == "concat_projection_BN":
            Ws[k] = [torch.from_numpy(d[f"{k}::{i}"]) for i in range(4)]
        else:
            Ws[k] = v
    x = np.random.RandomState(0).randint(0, 256, (1, 512, 512, 3)).astype(np.float32)
    with torch.no_grad():
        _, pref, _ = N.deeplabv3_forward(Ws, torch.from_numpy(x), net="subpixel")
    sm = SegModel(image_size=(512, 512), compute_dtype='float32')
    model = sm.create_seg_model('subpixel', n=21)
    assert model.layers[-3].name.startswith("subpixel") and model.layers[-1].name == "pred_mask"
    _push_weights(model, {**Ws, model.layers[-3].name: Ws["subpixel_1"]})
    probs = model.predict(x)
    assert rel(probs, pref) < 1e-3
    sm16 = SegModel(image_size=(512, 512), compute_dtype='float16')
    m16 = sm16.create_seg_model('subpixel', n=21)
    _push_weights(m16, {**Ws, m16.layers[-3].name: Ws["subpixel_1"]})
    p16 = m16.predict(x)
    assert (p16.argmax(-1) == pref.argmax(-1).numpy()).mean() > 0.98


# ----------------------------------------------------------------------------------------------- training
def _synthetic_batch(B, H, W, C=21, seed=0):
    rng = np.random.RandomState(seed)
    x = rng.randint(0, 256, (B, H, W, 3)).astype(np.float32)
    y = np.zeros((B, H, W), np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        for k in range(3):
            cy, cx, ry, rx = rng.randint(0, H), rng.randint(0, W), rng.randint(H // 8, H // 2), rng.randint(W // 8, W // 2)
            d = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
            lab = rng.randint(1, C)
            y[b][d < 1.0] = lab
            y[b][(d >= 1.0) & (d < 1.2)] = C          # void ring
    sw = rng.uniform(0.5, 2.0, (B, H * W)).astype(np.float32)
    sw[rng.uniform(size=sw.shape) < 0.1] = 0.0
    return x, y.reshape(B, H * W, 1), sw


@pytest.mark.parametrize("net", ["original", "subpixel"])
def test_train_step_fp32_matches_oracle(net):
    """one full step (all layers trainable, dropout off): loss, updated parameters, BN moving stats."""
    from deeplab_b200.model import Adam
    from deeplab_b200.utils import SegModel
    from oracle import network as N
    from oracle import train as T
    B, H, Wd = 2, 64, 96
    sm = SegModel(image_size=(H, Wd), compute_dtype='float32')
    model = sm.create_seg_model(net, n=21, seed=5)
    model.dropout_in_training = False
    head_name = model.engine.head_conv.name
    W = N.random_mobilenetv2_weights(seed=7, head=head_name, head_filters=21 * 64 if net == "subpixel" else 21)
    _push_weights(model, W)
    model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6), sample_weight_mode="temporal")
    x, y, sw = _synthetic_batch(B, H, Wd)
    vals = model.train_on_batch(x, y, {"pred_mask": sw})
    loss_ref, W1, _ = T.train_step(W, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw), net=net)
    _, grads, _, _ = T.loss_and_grads(W, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw), net=net)
    assert abs(vals[0] - loss_ref.item()) < 2e-4 * abs(loss_ref.item())
    got = _pull_weights(model)
    gmax = max(g.abs().max().item() for d in grads.values() for g in d.values())
    for name, ws in W1.items():
        for i, w in enumerate(ws):
            w0 = W[name][i].double()
            upd_ref = (w.double() - w0)
            upd_got = (got[name][i].double() - w0)
            if name in grads and i in grads[name] and grads[name][i].abs().max().item() < 1e-6 * gmax:
                continue      # structurally zero gradient: Adam's sign(noise) update is meaningless on both sides
            if name in grads and i in grads[name]:
                # Adam's first step is lr * sign(g) wherever |g| >> eps: an element whose gradient is at the fp32
                # noise level may legitimately flip sign, so compare where the oracle gradient is clearly non-zero
                # (strict) and the whole tensor in the mean (a few flips allowed)
                g = grads[name][i].reshape(upd_ref.shape).abs()
                mask = g > 0.2 * g.max()
                err = ((upd_got - upd_ref)[mask].abs().max() / upd_ref[mask].abs().max().clamp_min(1e-12)).item()
                assert err < 5e-2, (name, i, err)
                mean_err = ((upd_got - upd_ref).abs().mean() / upd_ref.abs().mean().clamp_min(1e-12)).item()
                assert mean_err < 0.1, (name, i, mean_err)
            else:                                                       # BN moving statistics
                err = ((upd_got - upd_ref).abs().max() / upd_ref.abs().max().clamp_min(1e-12)).item()
                assert err < 2e-3, (name, i, err)
    # second step continues from the device-side Adam state and iteration counter
    vals2 = model.train_on_batch(x, y, {"pred_mask": sw})
    assert np.isfinite(vals2[0]) and vals2[0] < vals[0] * 1.5


def test_gradients_fp32_match_oracle_autograd():
    """raw gradients of every parameter tensor vs torch autograd through the oracle (fp64), before the optimizer."""
    from deeplab_b200.utils import SegModel
    from oracle import network as N
    from oracle import train as T
    B, H, Wd = 2, 64, 64
    sm = SegModel(image_size=(H, Wd), compute_dtype='float32')
    model = sm.create_seg_model("original", n=21)
    W = N.random_mobilenetv2_weights(seed=11, head="conv_upsample")
    _push_weights(model, W)
    e = model.engine
    x, y, sw = _synthetic_batch(B, H, Wd, seed=4)
    ws = e.workspace(B, True)
    e.refresh_weight_copies()
    ws["img"].copy_(torch.from_numpy(x))
    ws["labels"].copy_(torch.from_numpy(y))
    ws["sample_w"].copy_(torch.from_numpy(sw))
    e.forward_train(ws, B, dropout=False)
    e.loss_and_head_grad(ws, B, True)
    e.backward(ws, B, dropout=False)
    torch.cuda.synchronize()
    tx, ty, tsw = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw)
    loss, grads, _, _ = T.loss_and_grads(W, tx, ty, tsw, dtype=torch.float64)
    _, g32, _, _ = T.loss_and_grads(W, tx, ty, tsw, dtype=torch.float32)
    assert abs(ws["loss_sum"].item() / ws["wcount"].item() - loss.item()) < 1e-4 * abs(loss.item())
    # The forward pass matches the oracle to ~1e-5, and every backward kernel matches an fp64
    # recomputation from its own inputs, but with 2 images at 64x64 the deep BatchNorms see only 128 samples per
    # channel: one ReLU6 mask that flips because z differs in the 6th digit moves a per-channel sum by ~1 %.  So:
    # strict max-norm check right behind the loss, direction check on the whole gradient, L2 check per tensor.
    gmax = max(g.abs().max().item() for d in grads.values() for g in d.values())
    flat, ref, l2 = [], [], []
    for rec in e.layers:
        for i, p in enumerate(rec.params):
            if not p.trainable_kind:
                continue
            gref = grads[rec.name][i].reshape(p.shape).double()
            got = p.grad.cpu().double() / e.loss_scale
            if gref.abs().max().item() < 1e-6 * gmax:
                # structurally zero gradient (the beta of a BN whose only consumers are 1x1 conv + BN)
                assert got.abs().max().item() < 1e-4 * gmax, rec.name
                continue
            flat.append(got.flatten()); ref.append(gref.flatten())
            err = ((got - gref).norm() / gref.norm()).item()
            l2.append(err)
            assert err < 0.2, (rec.name, i, err)               # a wrong index / missing term shows up as O(1)
            if rec.name == "conv_upsample":                    # directly behind the loss: no ReLU mask in between
                floor = rel(g32[rec.name][i].reshape(p.shape), gref)
                assert rel(got, gref) < max(5e-3, 5 * floor), (rec.name, i)
            elif rec.name in ("concat_projection", "concat_projection_BN", "aspp0"):
                assert err < 2e-2, (rec.name, i, err)          # one ReLU + BN deep: L2, robust to single mask flips
    flat, ref = torch.cat(flat), torch.cat(ref)
    cos = (torch.dot(flat, ref) / (flat.norm() * ref.norm())).item()
    assert cos > 0.9999, cos
    assert float(np.median(l2)) < 1e-2, float(np.median(l2))


def test_frozen_prefix_regime_and_keras_surface():
    """the notebook's fine-tuning regime (ipynb:147-155): everything before concat_projection frozen."""
    from deeplab_b200.model import Adam
    from deeplab_b200.utils import SegModel
    from oracle import network as N
    from oracle import train as T
    B, H, Wd = 2, 64, 64
    sm = SegModel(image_size=(H, Wd), compute_dtype='float32')
    model = sm.create_seg_model("original", n=21)
    model.dropout_in_training = False
    W = N.random_mobilenetv2_weights(seed=2, head="conv_upsample")
    _push_weights(model, W)
    frozen = []
    for layer in model.layers:
        if layer.name == 'concat_projection':
            break
        layer.trainable = False
        frozen.append(layer.name)
    model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6), sample_weight_mode="temporal")
    x, y, sw = _synthetic_batch(B, H, Wd, seed=9)
    vals = model.train_on_batch(x, y, {"pred_mask": sw})
    loss_ref, W1, _ = T.train_step(W, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw),
                                   frozen=set(frozen))
    assert abs(vals[0] - loss_ref.item()) < 2e-4 * abs(loss_ref.item())
    got = _pull_weights(model)
    for name, ws in W1.items():
        for i, w in enumerate(ws):
            if name in frozen:
                assert torch.equal(got[name][i], W[name][i]), name      # bit-identical: untouched
            else:
                upd_ref = w.double() - W[name][i].double()
                upd_got = got[name][i].double() - W[name][i].double()
                # mean error (sign flips of near-zero gradients under Adam's first step are legitimate outliers)
                assert ((upd_got - upd_ref).abs().mean() / upd_ref.abs().mean().clamp_min(1e-12)).item() < 2e-2, name


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_train_step_16bit_close_to_oracle(dtype):
    """tensor-core path: loss within 2%, gradient direction of the big conv kernels within cos > 0.98."""
    from deeplab_b200.utils import SegModel
    from oracle import network as N
    from oracle import train as T
    B, H, Wd = 4, 256, 256
    sm = SegModel(image_size=(H, Wd), compute_dtype=dtype)
    model = sm.create_seg_model("original", n=21)
    W = N.weights_from_h5(os.path.join(GOLD, "mobilenetv2_original.h5"))     # the reference's trained parameters
    _push_weights(model, W)
    e = model.engine
    x, y, sw = _synthetic_batch(B, H, Wd, seed=6)
    ws = e.workspace(B, True)
    e.refresh_weight_copies()
    ws["img"].copy_(torch.from_numpy(x)); ws["labels"].copy_(torch.from_numpy(y)); ws["sample_w"].copy_(torch.from_numpy(sw))
    e.forward_train(ws, B, dropout=False)
    e.loss_and_head_grad(ws, B, True)
    e.backward(ws, B, dropout=False)
    torch.cuda.synchronize()
    loss, grads, _, _ = T.loss_and_grads(W, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw),
                                         dtype=torch.float32)
    got = ws["loss_sum"].item() / ws["wcount"].item()
    assert abs(got - loss.item()) < (2e-2 if dtype == "float16" else 6e-2) * abs(loss.item())
    # Layers right behind the loss are well conditioned.  Deeper weight gradients are sums over pixels of
    # activation x (batch-norm-projected gradient): the projection removes the dominant (mean) component, so 16-bit
    # storage rounding of the operands is amplified there; they
    # are covered by the flat-gradient direction instead.
    for name in ["conv_upsample", "concat_projection", "aspp0"]:
        p = e._by_name[name].params[0]
        g = (p.grad.double().cpu() / e.loss_scale).flatten()
        r = grads[name][0].double().flatten()
        cos = torch.dot(g, r) / (g.norm() * r.norm())
        assert cos > (0.995 if dtype == "float16" else 0.97), (name, cos.item())
    flat, ref = [], []
    for rec in e.layers:
        for i, p in enumerate(rec.params):
            if p.trainable_kind:
                flat.append((p.grad.double().cpu() / e.loss_scale).flatten())
                ref.append(grads[rec.name][i].double().flatten())
    flat, ref = torch.cat(flat), torch.cat(ref)
    cos = torch.dot(flat, ref) / (flat.norm() * ref.norm())
    assert cos > (0.97 if dtype == "float16" else 0.75), cos.item()     # measured 0.99 / 0.83-0.92


def test_graph_replay_equals_eager_and_mious_match():
    """CUDA-graph replay reproduces the eager step from the same state; training reduces the loss; mIoU on a
    held-out synthetic mask set equals the oracle's +-0.1% (north_star)."""
    from deeplab_b200.model import Adam
    from deeplab_b200.utils import SegModel
    from oracle import network as N
    from oracle import ref_ops as R
    B, H, Wd = 2, 64, 64
    x, y, sw = _synthetic_batch(B, H, Wd, seed=21)
    xd, yd, swd = (torch.from_numpy(a).cuda() for a in (x, y, sw))
    sm = SegModel(image_size=(H, Wd), compute_dtype='float32')
    model = sm.create_seg_model("original", n=21)
    model.dropout_in_training = False
    _push_weights(model, N.random_mobilenetv2_weights(seed=5, head="conv_upsample"))
    model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6))
    e = model.engine
    ls, wc = e.train_step(xd, yd, swd, dropout=False)                     # warm-up + capture (performs step 1)
    first = ls.item() / wc.item()
    snap = [t.clone() for t in (e.params, e.adam_m, e.adam_v, e.adam_step, e.stats)]
    ls, wc = e.train_step(xd, yd, swd, dropout=False)                     # replayed step 2
    loss_g, grads_g, params_g = ls.item() / wc.item(), e.grads.clone(), e.params.clone()
    for dst, src in zip((e.params, e.adam_m, e.adam_v, e.adam_step, e.stats), snap):
        dst.copy_(src)
    e._weights_dirty = True
    ls, wc = e.train_step(xd, yd, swd, dropout=False, use_graph=False)    # the same step 2, eager
    loss_e = ls.item() / wc.item()
    assert abs(loss_g - loss_e) < 1e-5 * abs(loss_e)
    # float-atomic summation order is the only difference (the tiny batch through 50 batch-norms amplifies it in
    # the early layers, see test_gradients_fp32_match_oracle_autograd)
    cos = torch.dot(grads_g.double(), e.grads.double()) / (grads_g.double().norm() * e.grads.double().norm())
    assert cos > 0.9995, cos.item()
    assert (params_g - e.params).abs().max().item() <= 2.1 * 7e-4   # at most a sign flip of one Adam step
    assert (params_g - e.params).abs().mean().item() < 2e-5
    for _ in range(12):
        ls, wc = e.train_step(xd, yd, swd, dropout=False)
    assert ls.item() / wc.item() < 0.7 * first                            # it learns
    # held-out mIoU: model prediction vs oracle prediction under the notebook's mIOU (ipynb:203-210)
    W = _pull_weights(model)
    xv, yv, _ = _synthetic_batch(4, H, Wd, seed=99)
    p = model.predict(xv)
    with torch.no_grad():
        _, pref, _ = N.deeplabv3_forward({k: [t.float() for t in v] for k, v in W.items()}, torch.from_numpy(xv))
    for b in range(4):
        gt = yv[b, :, 0].astype(np.int64)
        m1 = R.notebook_miou(gt, p[b].argmax(-1))
        m2 = R.notebook_miou(gt, pref[b].argmax(-1).numpy())
        assert abs(m1 - m2) <= 1e-3


def test_generator_contract_and_calculate_iou():
    """The data formats either side of the hot path: device-side SegmentationGenerator labels / balanced
    sample weights feed train_on_batch, and calculate_iou's confusion matrix equals the notebook's counting loop
    applied to the model's own argmax."""
    from deeplab_b200.model import Adam
    from deeplab_b200.utils import SegModel, calculate_iou, generator_labels_and_weights
    from oracle import network as N
    from oracle import ref_ops as R
    B, H, Wd = 4, 64, 64
    x, y, _ = _synthetic_batch(B, H, Wd, seed=31)
    raw = y[:, :, 0].astype(np.int32).copy()
    raw[raw == 21] = 255                                   # VOC-style void value in the raw label files
    Y, SW = generator_labels_and_weights(raw.reshape(B, H, Wd), 21)
    assert Y.shape == (B, H * Wd, 1) and SW.shape == (B, H * Wd)
    for b in range(B):
        yr, swr = R.generator_labels_and_weights(raw[b], 21)
        assert np.array_equal(Y[b, :, 0].cpu().numpy(), yr.astype(np.float32))
        assert np.array_equal(SW[b].cpu().numpy(), swr)
    sm = SegModel(image_size=(H, Wd), compute_dtype='float32')
    model = sm.create_seg_model("original", n=21)
    _push_weights(model, N.random_mobilenetv2_weights(seed=6, head="conv_upsample"))
    model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6), sample_weight_mode="temporal")
    out = model.train_on_batch(x, Y, sample_weight={"pred_mask": SW})
    assert np.isfinite(np.asarray(out, dtype=np.float64)).all()
    conf = calculate_iou(model, nb_classes=21, data=(x, Y[:, :, 0].cpu().numpy()), batch_size=2)
    pred = model.predict(x).argmax(-1)
    assert np.array_equal(conf, R.calculate_iou_conf(pred, Y[:, :, 0].cpu().numpy(), 21))
    assert conf.sum() == (Y[:, :, 0].cpu().numpy() != 21).sum()


@pytest.mark.parametrize("alpha", [0.9, 1.1])
def test_width_multiplier_alpha_matches_oracle(alpha):
    """`alpha` != 1 (deeplabv3p.py:157-170): channel counts through _make_divisible, forward + one training step (fp32)
    against the oracle built with the same widths.  (The stem kernels are built for 32 filters: 0.9 <= alpha < 1.125.)"""
    from deeplab_b200.deeplabv3p import Deeplabv3
    from deeplab_b200.engine import _make_divisible
    from oracle import network as N
    W = N.random_mobilenetv2_weights(seed=13, alpha=alpha)
    model = Deeplabv3(weights=None, input_shape=(96, 128, 3), alpha=alpha, compute_dtype='float32')
    assert model.engine.c_last == _make_divisible(int(320 * alpha), 8) == W["aspp0"][0].shape[2]
    _push_weights(model, W)
    x = np.random.RandomState(2).randint(0, 256, (2, 96, 128, 3)).astype(np.float32)
    probs = model.predict(x, batch_size=2)
    with torch.no_grad():
        _, pref, _ = N.deeplabv3_forward(W, torch.from_numpy(x))
    assert rel(probs, pref) < 1e-3
    with pytest.raises(NotImplementedError):
        Deeplabv3(weights=None, input_shape=(96, 128, 3), alpha=0.5)


def test_input_tensor_fixes_geometry():
    from deeplab_b200.deeplabv3p import Deeplabv3
    t = torch.zeros(1, 64, 96, 3)
    model = Deeplabv3(weights=None, input_tensor=t, input_shape=(512, 512, 3), compute_dtype='float16')
    assert model.input is t and model.input_shape == (None, 64, 96, 3)
    assert model.predict(np.zeros((1, 64, 96, 3), np.float32)).shape == (1, 64 * 96, 21)


def test_bucketed_allreduce_schedule_single_gpu():
    """The data-parallel step body (gradient buckets handed to the all-reduce hook in completion order, inside the
    captured step) with a recording hook on one GPU: both buckets fire exactly once per step, in order, they tile the
    flat buffer, and the training trajectory equals the hook-free engine's."""
    from deeplab_b200.model import Adam
    from deeplab_b200.utils import SegModel
    from oracle import network as N
    B, H, Wd = 2, 64, 64
    x, y, sw = _synthetic_batch(B, H, Wd, seed=3)
    xd, yd, swd = (torch.from_numpy(a).cuda() for a in (x, y, sw))
    losses = []
    for hooked in (False, True):
        model = SegModel(image_size=(H, Wd), compute_dtype='float32').create_seg_model("original", n=21)
        model.dropout_in_training = False
        _push_weights(model, N.random_mobilenetv2_weights(seed=5, head="conv_upsample"))
        model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6))
        e = model.engine
        calls = []
        if hooked:
            def hook(t, async_op=False):
                calls.append((t.data_ptr() - e.grads.data_ptr()) // 4)
                calls.append(t.numel())
                return None
            e.grad_hook = hook
        out = []
        for _ in range(3):
            ls, wc = e.train_step(xd, yd, swd, dropout=False)
            out.append(ls.item() / wc.item())
        losses.append(out)
        if hooked:
            (lo0, hi0), (lo1, hi1) = e.grad_buckets()
            assert lo1 == 0 and hi1 == lo0 and hi0 == e.n_params and 0.7 < (hi0 - lo0) / e.n_params < 0.85
            # warm-up run + capture run: two body executions, each fires the suffix first, then the prefix
            assert calls == [lo0, hi0 - lo0, 0, hi1] * 2, calls
    # (same data, same start: the first loss is identical; later ones differ by the float-atomic summation order that
    # Adam's sign-like first steps amplify, see test_graph_replay_equals_eager_and_mious_match)
    # measured spread of the third loss between two runs of the SAME schedule on a B200: up to 9 %, so it only bounds
    # gross divergence; the second loss (one Adam step after identical gradients up to summation order) is held to 2 %
    a, b = losses
    assert abs(a[0] - b[0]) <= 1e-5 * abs(a[0]), losses
    assert abs(a[1] - b[1]) <= 2e-2 * abs(a[1]) and abs(a[2] - b[2]) <= 0.25 * abs(a[2]), losses


def test_test_on_batch_and_validation_match_oracle(tmp_path):
    """Keras `test_on_batch` / `evaluate_generator` / `fit_generator(validation_data=...)`: inference-phase forward
    (moving statistics, no dropout) + the fused loss / argmax / confusion kernels vs the oracle's loss
    (utils.py:127-130 with temporal weights), Jaccard (:139-157) and accuracy (:132-138)."""
    from deeplab_b200.model import Adam, ModelCheckpoint
    from deeplab_b200.utils import SegModel
    from oracle import network as N
    from oracle import ref_ops as R
    B, H, Wd = 3, 64, 96
    x, y, sw = _synthetic_batch(B, H, Wd, seed=17)
    W = N.random_mobilenetv2_weights(seed=9, head="conv_upsample")
    model = SegModel(image_size=(H, Wd), compute_dtype='float32').create_seg_model("original", n=21)
    _push_weights(model, W)
    model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6), sample_weight_mode="temporal")
    with torch.no_grad():
        _, pref, _ = N.deeplabv3_forward(W, torch.from_numpy(x))
    ty, tsw = torch.from_numpy(y), torch.from_numpy(sw)
    loss_ref = R.keras_weighted_loss(ty, pref, tsw).item()
    jac_ref, acc_ref = R.jaccard(ty, pref).item(), R.sparse_accuracy_ignoring_last_label(ty, pref).item()
    got = model.test_on_batch(x, y, {"pred_mask": sw})
    assert abs(got[0] - loss_ref) <= 1e-4 * abs(loss_ref), (got, loss_ref)
    assert abs(got[1] - jac_ref) <= 1e-6 and abs(got[2] - acc_ref) <= 1e-6, (got, jac_ref, acc_ref)
    # no sample weights: plain mean over all pixels (void pixels contribute 0)
    got2 = model.test_on_batch(x, y)
    assert abs(got2[0] - R.keras_weighted_loss(ty, pref).item()) <= 1e-4 * abs(got2[0])

    class Seq:
        def __len__(self):
            return 2

        def __getitem__(self, i):
            return x, y, {"pred_mask": sw}

    ev = model.evaluate_generator(Seq())
    assert abs(ev[0] - loss_ref) <= 1e-4 * abs(loss_ref)
    ck = str(tmp_path / "best.h5")
    hist = model.fit_generator(Seq(), steps_per_epoch=2, epochs=2, verbose=0, validation_data=Seq(), validation_steps=1,
                               callbacks=[ModelCheckpoint(ck, monitor="val_Jaccard", save_best_only=True, save_weights_only=True)])
    assert set(hist.history) >= {"loss", "Jaccard", "val_loss", "val_Jaccard"} and len(hist.history["val_loss"]) == 2
    assert all(np.isfinite(v) for v in hist.history["val_loss"]) and os.path.exists(ck)
    m2 = SegModel(image_size=(H, Wd), compute_dtype='float32').create_seg_model("original", n=21)
    m2.load_weights(ck)                                    # the checkpoint written by the callback loads back
    assert np.array_equal(m2.get_layer("aspp0").get_weights()[0].shape, (1, 1, 320, 256))
