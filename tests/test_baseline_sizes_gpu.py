"""Parity AT THE SIZES AND PRECISIONS BASELINE.json's configs are benchmarked in (round-1 review: every earlier parity
test ran at toy sizes or in fp32 only).

  config 2 / 4   MobileNetV2 512x512 training step in fp32, fp16 AND bf16 (batch 4 -- the statistics per BatchNorm are
                 4*64*64 = 16 384 samples, the same order as the benchmarked 16) on the config-2 inputs bench.py times
                 and on the reference's example photos: loss, per-tensor gradient cosine against the fp32 oracle, one
                 Adam update
  north_star     held-out mIoU in fp16 at 512x512 equal to the oracle's +-0.1 %
  config 3       Xception OS=8, 512x512, batch 4: fp32 within 1e-3 relative, fp16 argmax agreement
  config 5       dense CRF on 1024x1024x21, 10 iterations, within 1e-2 of the C restatement

Every tolerance is written at its assert and quoted in BASELINE.md section 4.  The measured values are also dumped to
gpurun_out/parity_baseline_sizes.json (when that directory exists) so the numbers in the docs can be traced.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _report(key, obj):
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    path = os.path.join(d, "parity_baseline_sizes.json")
    cur = {}
    if os.path.exists(path):
        try:
            cur = json.load(open(path))
        except Exception:
            cur = {}
    cur[key] = obj
    with open(path, "w") as f:
        json.dump(cur, f, indent=1, sort_keys=True)


def _push_weights(model, W):
    for l in model.layers:
        if l.name in W:
            l.set_weights([w.detach().float().numpy() for w in W[l.name]])


def _smooth_images(B, H, W, seed):
    """seeded blurred-noise images: natural-image-like statistics, 0..255"""
    import scipy.ndimage as ndi
    rng = np.random.RandomState(seed)
    x = ndi.gaussian_filter(rng.rand(B, H, W, 3), (0, 8, 8, 0))
    x = (x - x.min()) / (x.max() - x.min())
    return np.floor(x * 255.999).astype(np.float32)


def _ellipse_masks(B, H, W, C, seed):
    rng = np.random.RandomState(seed)
    y = np.zeros((B, H, W), np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        for _ in range(3):
            cy, cx, ry, rx = rng.randint(0, H), rng.randint(0, W), rng.randint(H // 8, H // 2), rng.randint(W // 8, W // 2)
            d = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
            y[b][d < 1.0] = rng.randint(1, C)
            y[b][(d >= 1.0) & (d < 1.1)] = C
    return y.reshape(B, H * W, 1)


def _balanced_weights(y, C):
    B = y.shape[0]
    sw = np.zeros(y.shape[:2], np.float32)
    for b in range(B):
        cls, cnt = np.unique(y[b, :, 0], return_counts=True)
        keep = cls != C
        n = cnt[keep].sum()
        for c, k in zip(cls[keep], cnt[keep]):
            sw[b][y[b, :, 0] == c] = n / (keep.sum() * k)
    return sw


# ------------------------------------------------------------------------------------------------- configs 2 / 4
_ORACLE_CACHE = {}


def _config2_inputs(kind, B, H, Wd, C):
    """'noise': SURVEY 8(d) config 2 -- X ~ U{0..255}, ellipse masks with a void ring, per-image balanced class weights
    (what bench.py times).  'photo': the three example-figure crops the reference ships + one noise image."""
    if kind == "photo":
        ex = np.load(os.path.join(GOLD, "example_crops.npz"))
        x = np.stack([ex[k].astype(np.float32) for k in ("exp1", "exp3", "exp4")] +
                     [np.random.RandomState(0).randint(0, 256, (H, Wd, 3)).astype(np.float32)])[:B]
    else:
        x = np.random.RandomState(0).randint(0, 256, (B, H, Wd, 3)).astype(np.float32)
    y = _ellipse_masks(B, H, Wd, C, seed=3)
    return x, y, _balanced_weights(y, C)


def _oracle_grads(kind, W, x, y, sw):
    from oracle import train as T
    if kind not in _ORACLE_CACHE:
        torch.set_num_threads(min(os.cpu_count() or 1, 32))
        loss_ref, grads, _, _ = T.loss_and_grads(W, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw),
                                                 dtype=torch.float32)
        _ORACLE_CACHE[kind] = (loss_ref.item(), grads)
    return _ORACLE_CACHE[kind]


# measured on a B200 (gpurun_out/parity_baseline_sizes.json, quoted in BASELINE.md section 4), batch 4, 512x512:
#   fp32  loss 1e-7,  flat cosine 0.999997, worst tensor 0.99999, Adam sign agreement 1.0
#   fp16  loss 8e-5 (noise) / 3.5e-3 (photo), flat cosine 0.989 / 0.994 (one of eight runs: 0.980 -- the fp32 atomics of
#         the BatchNorm statistics make the forward pass itself vary by ~4e-4 in the loss from run to run), worst tensor
#         0.973-0.983, sign agreement 0.994
#   bf16  loss 1.1e-2 / 0.9e-2, flat cosine 0.89 / 0.91-0.94, worst tensor 0.70-0.81 and sign agreement 0.93 over four runs
#         (run-to-run: the float atomics of the statistics; 8-bit mantissa of every stored activation
#         and gradient, amplified by ~50 BatchNorm backward projections -- inherent to bf16 storage, not a kernel fault:
#         the fp16 and fp32 instances of the same kernels meet the tighter bounds)
_TOL = {
    "float32": dict(loss=1e-5, flat=0.9999, median=0.9999, worst=0.9999, sign=0.999),
    "float16": dict(loss=1e-2, flat=0.97, median=0.97, worst=0.95, sign=0.98),
    "bfloat16": dict(loss=3e-2, flat=0.85, median=0.85, worst=0.6, sign=0.85),
}


@pytest.mark.parametrize("kind", ["noise", "photo"])
@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
def test_train_step_512_vs_oracle(dtype, kind):
    """One full 512x512 step in the benchmarked precisions (and fp32), from the reference's trained parameters:
    loss; cosine of every convolution-kernel gradient tensor (and of the whole flat gradient) to the fp32 oracle's;
    the first Adam update (= -lr*sign(g) wherever |g| >> eps) as sign agreement over the elements whose oracle gradient
    is above 10 % of the tensor's maximum.  Tolerances: _TOL above."""
    from deeplab_b200.model import Adam
    from deeplab_b200.utils import SegModel
    from oracle import network as N
    B, H, Wd, C = 4, 512, 512, 21
    tol = _TOL[dtype]
    sm = SegModel(image_size=(H, Wd), compute_dtype=dtype)
    model = sm.create_seg_model("original", n=C)
    model.dropout_in_training = False
    W = N.weights_from_h5(os.path.join(GOLD, "mobilenetv2_original.h5"))
    _push_weights(model, W)
    model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6), sample_weight_mode="temporal")
    e = model.engine
    x, y, sw = _config2_inputs(kind, B, H, Wd, C)
    tx, ty, tsw = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw)
    loss_ref, grads = _oracle_grads(kind, W, x, y, sw)

    ws = e.workspace(B, True)
    e.refresh_weight_copies()
    ws["img"].copy_(tx); ws["labels"].copy_(ty); ws["sample_w"].copy_(tsw)
    e.forward_train(ws, B, dropout=False)
    e.loss_and_head_grad(ws, B, True)
    e.backward(ws, B, dropout=False)
    torch.cuda.synchronize()
    loss = ws["loss_sum"].item() / ws["wcount"].item()
    ls = e.loss_scale
    cos, flat, ref = {}, [], []
    for rec in e.layers:
        for i, p in enumerate(rec.params):
            if not p.trainable_kind:
                continue
            g = (p.grad.double().cpu() / ls).flatten()
            r = grads[rec.name][i].double().flatten()
            flat.append(g); ref.append(r)
            if rec.kind != "bn" and r.norm() > 0:
                cos[f"{rec.name}:{i}"] = (torch.dot(g, r) / (g.norm() * r.norm()).clamp_min(1e-300)).item()
    flat, ref = torch.cat(flat), torch.cat(ref)
    cos_all = (torch.dot(flat, ref) / (flat.norm() * ref.norm())).item()
    vals = np.array(list(cos.values()))
    rep = {"loss": loss, "loss_ref": loss_ref, "loss_rel_err": abs(loss - loss_ref) / abs(loss_ref), "cos_flat": cos_all,
           "cos_min": float(vals.min()), "cos_median": float(np.median(vals)),
           "worst": sorted(cos.items(), key=lambda kv: kv[1])[:3]}

    # one optimizer step through the public API (graph-captured path) from the same state
    out = model.train_on_batch(x, y, {"pred_mask": sw})
    new = {l.name: [torch.from_numpy(a) for a in l.get_weights()] for l in model.layers if l._rec is not None}
    agree = {}
    for name in ("conv_upsample", "concat_projection", "aspp0", "expanded_conv_16_project", "expanded_conv_13_expand",
                 "expanded_conv_6_depthwise", "expanded_conv_3_expand", "Conv"):
        g = grads[name][0].double()
        upd = (new[name][0].double() - W[name][0].double()).reshape(g.shape)
        big = g.abs() > 0.1 * g.abs().max()
        agree[name] = (torch.sign(upd[big]) == -torch.sign(g[big])).double().mean().item()
    rep["adam_sign_agreement"] = agree
    rep["loss_api"] = out[0]
    _report(f"train_step_512_{dtype}_{kind}", rep)

    assert abs(loss - loss_ref) <= tol["loss"] * abs(loss_ref), rep
    assert abs(out[0] - loss_ref) <= tol["loss"] * abs(loss_ref), rep
    assert cos_all > tol["flat"] and np.median(vals) > tol["median"] and vals.min() > tol["worst"], rep
    assert min(agree.values()) > tol["sign"], rep


def test_miou_512_fp16_equals_oracle():
    """north_star: mIoU on a held-out synthetic mask set equal to the reference +-0.1 %, in the benchmarked fp16 mode at
    512x512.  Inputs: the three example-figure crops the reference ships plus five seeded smooth images; the notebook's
    mIOU (ipynb:203-210) of (mask, prediction) for the CUDA fp16 path and for the fp32 oracle must agree within 1e-3."""
    from deeplab_b200.deeplabv3p import Deeplabv3
    from oracle import network as N
    from oracle import ref_ops as R
    ex = np.load(os.path.join(GOLD, "example_crops.npz"))
    imgs = [ex[k].astype(np.float32) for k in ("exp1", "exp3", "exp4")]
    x = np.concatenate([np.stack(imgs), _smooth_images(5, 512, 512, seed=11)], 0)
    B = x.shape[0]
    masks = _ellipse_masks(B, 512, 512, 21, seed=12)[:, :, 0].astype(np.int64)
    W = N.weights_from_h5(os.path.join(GOLD, "mobilenetv2_original.h5"))
    model = Deeplabv3(weights=None, input_shape=(512, 512, 3), compute_dtype="float16")
    model.load_weights(os.path.join(GOLD, "mobilenetv2_original.h5"))
    p = model.predict(x, batch_size=4).argmax(-1)
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    diffs, agree, self_miou = [], [], []
    for b in range(B):
        with torch.no_grad():
            _, pref, _ = N.deeplabv3_forward(W, torch.from_numpy(x[b:b + 1]))
        pr = pref[0].argmax(-1).numpy()
        m1, m2 = R.notebook_miou(masks[b], p[b]), R.notebook_miou(masks[b], pr)
        diffs.append(abs(m1 - m2))
        agree.append(float((p[b] == pr).mean()))
        self_miou.append(R.notebook_miou(pr, p[b]))      # reported: the oracle's own segmentation as ground truth
    _report("miou_512_fp16", {"max_abs_miou_diff": max(diffs), "argmax_agreement": agree,
                              "miou_vs_oracle_segmentation": self_miou})
    assert max(diffs) <= 1e-3, diffs
    assert min(agree) > 0.99, agree


# ------------------------------------------------------------------------------------------------- config 3
def test_xception_512_bs4_fp32_and_fp16():
    """BASELINE config 3 at its own size: Xception OS=8, 512x512, batch 4.  fp32: logits within 1e-3 relative of the
    oracle (north_star tolerance); fp16 tensor-core path: max prob error 5e-2, argmax agreement > 0.97."""
    from deeplab_b200.deeplabv3p import Deeplabv3
    from oracle import network as N
    W = N.random_xception_weights(seed=8)
    H = Wd = 512
    B = 4
    x = np.random.RandomState(0).randint(0, 256, (B, H, Wd, 3)).astype(np.float32)
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    with torch.no_grad():
        logits_ref, pref, _ = N.deeplabv3_forward(W, torch.from_numpy(x), backbone="xception", OS=8)
    rep = {}
    for dtype in ("float32", "float16"):
        model = Deeplabv3(weights=None, input_shape=(H, Wd, 3), classes=21, backbone="xception", OS=8,
                          compute_dtype=dtype)
        _push_weights(model, W)
        probs = model.predict(x, batch_size=B)
        logits = model.engine.workspace(B, False)["logits"][..., :21].float().cpu()
        rel = ((logits - logits_ref).abs().max() / logits_ref.abs().max()).item()
        perr = float(np.abs(probs - pref.numpy()).max())
        agree = float((probs.argmax(-1) == pref.argmax(-1).numpy()).mean())
        rep[dtype] = {"logits_rel": rel, "probs_max_abs": perr, "argmax_agreement": agree}
        if dtype == "float32":
            assert rel <= 1e-3, rep
            assert agree > 0.999, rep
        else:
            assert perr < 5e-2 and agree > 0.97, rep
        del model
        torch.cuda.empty_cache()
    _report("xception_512_bs4", rep)


# ------------------------------------------------------------------------------------------------- config 5
def test_crf_1024x1024x21_10_iterations():
    """BASELINE config 5 at its own size (one image of the batch of 8): marginals within 1e-2 (north_star CRF
    tolerance) of the single-threaded C restatement of densecrf; MAP labels agree on > 99.9 % of the pixels."""
    import scipy.ndimage as ndi
    from deeplab_b200.utils import dense_crf
    from oracle import crf as O
    H = W = 1024
    M = 21
    rng = np.random.RandomState(5)
    logits = rng.randn(M, H * W).astype(np.float32) * 3
    un = -(logits - np.log(np.exp(logits).sum(0, keepdims=True)))
    img = ndi.gaussian_filter(rng.rand(H, W, 3), (8, 8, 0))
    img = ((img - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
    Qref = O.dense_crf(un, img, iters=10)
    Q = dense_crf(un, img, iters=10).cpu().numpy()
    err = float(np.abs(Q - Qref).max())
    agree = float((Q.argmax(0) == Qref.argmax(0)).mean())
    _report("crf_1024_21_10it", {"max_abs_Q": err, "map_agreement": agree})
    assert np.abs(Q.sum(0) - 1).max() < 1e-4
    assert err <= 1e-2, err
    assert agree > 0.999, agree
