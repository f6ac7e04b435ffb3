"""Debug (GPU): run one fp32 training fwd+bwd and check every bn_bwd / dw_conv_bwd / pw_wgrad call against a torch
fp64 recomputation from the very tensors the call received.  Prints the calls whose outputs disagree."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import torch.nn.functional as F
import deeplab_b200
from deeplab_b200 import ops
from deeplab_b200.utils import SegModel
from oracle import network as N
from test_model_gpu import _synthetic_batch, _push_weights

B, H, Wd = 2, 64, 64
dtype = sys.argv[1] if len(sys.argv) > 1 else "float32"
sm = SegModel(image_size=(H, Wd), compute_dtype=dtype)
model = sm.create_seg_model("original", n=21)
_push_weights(model, N.random_mobilenetv2_weights(seed=11, head="conv_upsample"))
e = model.engine
x, y, sw = _synthetic_batch(B, H, Wd, seed=4)
ws = e.workspace(B, True)
e.refresh_weight_copies()
ws["img"].copy_(torch.from_numpy(x)); ws["labels"].copy_(torch.from_numpy(y)); ws["sample_w"].copy_(torch.from_numpy(sw))
e.forward_train(ws, B, dropout=False)
e.loss_and_head_grad(ws, B, True)

def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()

o_bn, o_dw, o_wg = ops.bn_bwd, ops.dw_conv_bwd, ops.pw_wgrad
cnt = [0]

def bn_bwd(x, da, dx, *, scale, shift, mean, rstd, act, red, dgamma=None, dbeta=None, **k):
    xd, dad = x.double().reshape(-1, x.shape[-1]), da.double().reshape(-1, x.shape[-1])
    z = xd * scale.double() + shift.double()
    m = torch.ones_like(z) if act == 0 else ((z > 0) & ((z < 6) if act == 2 else True)).double()
    dz = dad * m
    xh = (xd - mean.double()) * rstd.double()
    db, dg = dz.sum(0), (dz * xh).sum(0)
    M = xd.shape[0]
    ref = scale.double() * (dz - db / M - xh * dg / M)
    r = o_bn(x, da, dx, scale=scale, shift=shift, mean=mean, rstd=rstd, act=act, red=red, dgamma=dgamma, dbeta=dbeta, **k)
    torch.cuda.synchronize()
    e1, e2, e3 = rel(dx.reshape(-1, x.shape[-1]), ref), rel(red[:x.shape[-1]], db), rel(red[x.shape[-1]:], dg)
    cnt[0] += 1
    if max(e1, e2, e3) > 2e-3:
        print(f"bn_bwd#{cnt[0]} shape {tuple(x.shape)} act {act}: dx {e1:.2e} dbeta {e2:.2e} dgamma {e3:.2e}")
    return r

def dw_bwd(x, dy, w, *, dx=None, dw=None, in_shape=None, stride, dilation, pad_top, pad_left, in_scale=None, in_shift=None, in_act=0):
    shp = in_shape
    a = x.double()
    if in_scale is not None:
        a = a * in_scale.double() + in_shift.double()
        a = a.clamp(0, 6) if in_act == 2 else a
    a = a.detach().requires_grad_(True)
    wd = w.double().detach().requires_grad_(True)
    C = shp[3]
    Ho, Wo = dy.shape[1], dy.shape[2]
    pb = (Ho - 1) * stride + 2 * dilation + 1 - shp[1] - pad_top
    pr = (Wo - 1) * stride + 2 * dilation + 1 - shp[2] - pad_left
    ap = F.pad(a.permute(0, 3, 1, 2), (pad_left, max(pr, 0), pad_top, max(pb, 0)))
    out = F.conv2d(ap, wd.view(3, 3, C).permute(2, 0, 1).unsqueeze(1), stride=stride, dilation=dilation, groups=C).permute(0, 2, 3, 1)
    ga, gw = torch.autograd.grad(out, [a, wd], dy.double())
    o_dw(x, dy, w, dx=dx, dw=dw, in_shape=in_shape, stride=stride, dilation=dilation, pad_top=pad_top, pad_left=pad_left,
         in_scale=in_scale, in_shift=in_shift, in_act=in_act)
    torch.cuda.synchronize()
    e1 = rel(dx, ga) if dx is not None else 0
    e2 = rel(dw, gw.view(dw.shape)) if dw is not None else 0
    if max(e1, e2) > 2e-3:
        print(f"dw_bwd shape {tuple(shp)} s{stride} d{dilation}: dx {e1:.2e} dw {e2:.2e}")

def pw_wgrad(A, dY, dW, **k):
    o_wg(A, dY, dW, **k)
    torch.cuda.synchronize()
    Kk, Nn = k.get("K") or A.shape[-1], k.get("N") or dY.shape[-1]
    ref = A.double().reshape(-1, A.shape[-1])[:, :Kk].t() @ dY.double().reshape(-1, dY.shape[-1])[:, :Nn]
    er = rel(dW.reshape(Kk, -1)[:, :Nn], ref)
    if er > 2e-3:
        print(f"pw_wgrad K{Kk} N{Nn}: {er:.2e}")
    return dW

ops.bn_bwd, ops.dw_conv_bwd, ops.pw_wgrad = bn_bwd, dw_bwd, pw_wgrad
e.backward(ws, B, dropout=False)
torch.cuda.synchronize()
print("debug_bwd done:", cnt[0], "bn_bwd calls checked")

# ---- forward comparison against the oracle (training-mode taps) and final gradient comparison
from oracle import train as T
W = N.random_mobilenetv2_weights(seed=11, head="conv_upsample")
tap = {}
Wd_ = {k: [t.double() for t in v] for k, v in W.items()}
_, _, ctx = N.deeplabv3_forward(Wd_, torch.from_numpy(x).double(), training=True, tap=tap)
for i in range(17):
    name = f"expanded_conv_{i}_out" if i else "expanded_conv_out"
    print(f"fwd x{i + 1} vs oracle: {rel(ws[f'x{i + 1}'].cpu(), tap[name]):.2e}")
print("feat:", rel(ws["feat"].cpu(), tap["features"]), "logits:", rel(ws["logits"][..., :21].cpu(), tap["logits"]))
loss, grads, _, _ = T.loss_and_grads(W, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(sw))
for rec in e.layers:
    for i, p in enumerate(rec.params):
        if p.trainable_kind:
            gref = grads[rec.name][i].reshape(p.shape)
            er = rel(p.grad.cpu() / e.loss_scale, gref)
            if er > 2e-3 and gref.abs().max() > 1e-9:
                print(f"grad {rec.name}[{i}] vs oracle: {er:.2e} (|g| {gref.abs().max():.2e})")
