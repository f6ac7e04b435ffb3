"""Diagnostic (GPU): per-layer cosine / norm ratio of 16-bit-path gradients vs the fp32-path gradients of the same
engine code, real weights, for several loss scales.  Usage: python tools/diag_grads.py [H]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import deeplab_b200
from deeplab_b200.utils import SegModel
from test_model_gpu import _synthetic_batch

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = 4
x, y, sw = _synthetic_batch(B, H, H, seed=6)


def grads(dtype, loss_scale=None):
    sm = SegModel(image_size=(H, H), compute_dtype=dtype)
    m = sm.create_seg_model("original", n=21)
    m.load_weights(os.path.join(ROOT, "tests", "golden", "mobilenetv2_original.h5"))
    e = m.engine
    if loss_scale is not None:
        e.loss_scale = loss_scale
    ws = e.workspace(B, True)
    e.refresh_weight_copies()
    ws["img"].copy_(torch.from_numpy(x)); ws["labels"].copy_(torch.from_numpy(y)); ws["sample_w"].copy_(torch.from_numpy(sw))
    e.forward_train(ws, B, False); e.loss_and_head_grad(ws, B, True); e.backward(ws, B, False)
    torch.cuda.synchronize()
    return e, {r.name: (r.params[0].grad.double().flatten() / e.loss_scale).clone() for r in e.layers if r.kind != "bn"}, \
        (e.grads.double() / e.loss_scale).clone()


_, g32, f32 = grads("float32")
for dt, ls in (("float16", 1024.0), ("float16", 65536.0), ("bfloat16", 1.0)):
    _, g, f = grads(dt, ls)
    print(f"== {dt} loss_scale {ls}: flat cos {torch.dot(f, f32) / (f.norm() * f32.norm()):.4f}")
    for name in ["conv_upsample", "concat_projection", "aspp0", "expanded_conv_16_project", "expanded_conv_16_depthwise",
                 "expanded_conv_16_expand", "expanded_conv_13_project", "expanded_conv_10_expand", "expanded_conv_6_project",
                 "expanded_conv_3_expand", "expanded_conv_1_expand", "expanded_conv_project", "Conv"]:
        a, b = g[name], g32[name]
        print(f"   {name:28s} cos {torch.dot(a, b) / (a.norm() * b.norm() + 1e-30):7.4f}  |g16|/|g32| {a.norm() / (b.norm() + 1e-30):7.4f}  |g32| {b.norm():.3e}")
