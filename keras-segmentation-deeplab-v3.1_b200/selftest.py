"""smoke(): one small invocation of the hot path on cuda:0, checked against the oracle (called by __graft_entry__)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch


def smoke() -> None:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import network as N          # checker only
    from oracle import train as T
    from . import _lib
    from .model import Adam
    from .utils import SegModel, dense_crf

    if not torch.cuda.is_available() or not _lib.lib().dlb_device_ok():
        raise RuntimeError("smoke() needs a compute-capability 10.x CUDA device")
    torch.cuda.set_device(0)
    B, H, W = 2, 64, 64
    rng = np.random.RandomState(0)
    x = rng.randint(0, 256, (B, H, W, 3)).astype(np.float32)
    y = rng.randint(0, 22, (B, H * W, 1)).astype(np.float32)
    sw = rng.uniform(0.5, 1.5, (B, H * W)).astype(np.float32)
    Wt = N.random_mobilenetv2_weights(seed=1, head="conv_upsample")
    n0 = _lib.launch_count()
    for dtype, tol in (("float32", 2e-4), ("float16", 3e-2)):
        sm = SegModel(image_size=(H, W), compute_dtype=dtype)
        model = sm.create_seg_model("original", n=21)
        model.dropout_in_training = False
        for l in model.layers:
            if l.name in Wt:
                l.set_weights([w.numpy() for w in Wt[l.name]])
        model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6), sample_weight_mode="temporal")
        with torch.no_grad():
            _, pref, _ = N.deeplabv3_forward(Wt, torch.from_numpy(x))
        p = model.predict(x)
        err = np.abs(p - pref.numpy()).max()
        assert err < (1e-3 if dtype == "float32" else 5e-2), f"inference parity {dtype}: {err}"
        loss = model.train_on_batch(x, y, {"pred_mask": sw})[0]
        ref, _, _ = T.train_step(Wt, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw))
        assert abs(loss - ref.item()) < tol * abs(ref.item()), f"training loss parity {dtype}: {loss} vs {ref.item()}"
    # dense-CRF kernel against the C restatement
    from oracle import crf as O
    un = rng.rand(3, 32 * 32).astype(np.float32) * 3
    img = (rng.rand(32, 32, 3) * 255).astype(np.uint8)
    q = dense_crf(un, img, iters=3).cpu().numpy()
    qr = O.dense_crf(un, img, iters=3)
    assert np.abs(q - qr).max() < 5e-3, f"crf parity: {np.abs(q - qr).max()}"
    print(f"smoke ok: {_lib.launch_count() - n0} kernel launches from {_lib.LIB_PATH}")
