"""Gradient precision of the 16-bit training path at 512x512 vs the fp32 oracle, as a function of the fp16 loss scale
(is the gap to fp32 rounding noise or underflow of the scaled gradients?).  GPU box only; prints one JSON line per run.
    python tools/grad_precision_probe.py [batch]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deeplab_b200  # noqa: E402,F401
from deeplab_b200.utils import SegModel  # noqa: E402
from oracle import network as N  # noqa: E402
from oracle import train as T  # noqa: E402
import test_baseline_sizes_gpu as TB  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = Wd = 512
W = N.weights_from_h5(os.path.join(ROOT, "tests", "golden", "mobilenetv2_original.h5"))
KIND = sys.argv[2] if len(sys.argv) > 2 else "noise"
if KIND == "smooth":
    x = TB._smooth_images(B, H, Wd, seed=2)
elif KIND == "photo":      # the reference's three example-figure crops + one noise image
    ex = np.load(os.path.join(ROOT, "tests", "golden", "example_crops.npz"))
    x = np.stack([ex[k].astype(np.float32) for k in ("exp1", "exp3", "exp4")] +
                 [np.random.RandomState(0).randint(0, 256, (H, Wd, 3)).astype(np.float32)])[:B]
else:                      # X ~ U{0..255}
    x = np.random.RandomState(0).randint(0, 256, (B, H, Wd, 3)).astype(np.float32)
y = TB._ellipse_masks(B, H, Wd, 21, seed=3)
sw = TB._balanced_weights(y, 21)
tx, ty, tsw = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw)
torch.set_num_threads(min(os.cpu_count() or 1, 32))
loss_ref, grads, _, _ = T.loss_and_grads(W, tx, ty, tsw, dtype=torch.float32)
for dtype, scales in (("float16", [1024.0]),):
    model = SegModel(image_size=(H, Wd), compute_dtype=dtype).create_seg_model("original", n=21)
    TB._push_weights(model, W)
    e = model.engine
    ws = e.workspace(B, True)
    e.refresh_weight_copies()
    for s in scales:
        if e.ls_state is not None:
            e.loss_scale = s
        ws["img"].copy_(tx); ws["labels"].copy_(ty); ws["sample_w"].copy_(tsw)
        e.forward_train(ws, B, dropout=False)
        e.loss_and_head_grad(ws, B, True)
        e.backward(ws, B, dropout=False)
        torch.cuda.synchronize()
        cos, flat, ref = {}, [], []
        for rec in e.layers:
            for i, p in enumerate(rec.params):
                if not p.trainable_kind:
                    continue
                g = (p.grad.double().cpu() / s).flatten()
                r = grads[rec.name][i].double().flatten()
                flat.append(g); ref.append(r)
                if rec.kind != "bn":
                    cos[rec.name] = (torch.dot(g, r) / (g.norm() * r.norm()).clamp_min(1e-300)).item()
        flat, ref = torch.cat(flat), torch.cat(ref)
        v = np.array(list(cos.values()))
        print(json.dumps({"dtype": dtype, "batch": B, "loss_scale": s, "finite": bool(torch.isfinite(flat).all()),
                          "loss": ws["loss_sum"].item() / ws["wcount"].item(), "loss_ref": loss_ref.item(),
                          "cos_flat": (torch.dot(flat, ref) / (flat.norm() * ref.norm())).item(),
                          "cos_median": float(np.median(v)), "cos_min": float(v.min()),
                          "cos_head": cos.get("conv_upsample"), "cos_stem": cos.get("Conv"), "kind": KIND,
                          "by_depth": [round(cos[k], 4) for k in ("conv_upsample", "concat_projection", "aspp0",
                                       "expanded_conv_16_project", "expanded_conv_16_depthwise", "expanded_conv_16_expand",
                                       "expanded_conv_13_project", "expanded_conv_10_expand", "expanded_conv_6_project",
                                       "expanded_conv_3_expand", "expanded_conv_1_expand", "expanded_conv_depthwise", "Conv")]}),
              flush=True)
    del model, e, ws
    torch.cuda.empty_cache()
