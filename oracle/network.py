"""ORACLE (test infrastructure only): CPU restatement of the reference's DeepLabV3+ graph in torch-CPU.

Follows /root/reference line by line:
  deeplabv3p.py:47-84   SepConv_BN            -> sepconv_bn
  deeplabv3p.py:87-116  _conv2d_same          -> conv2d_same_fixed
  deeplabv3p.py:119-155 _xception_block       -> xception_block
  deeplabv3p.py:167-206 _inverted_res_block   -> inverted_res_block
  deeplabv3p.py:209-466 Deeplabv3             -> deeplabv3_forward
  utils.py:169-198      create_seg_model heads ('original' = conv_upsample + bilinear, 'subpixel' = Subpixel(n,1,8))
Third-party op semantics (TF SAME padding, Keras BN, legacy bilinear, phase shift) come from oracle/ref_ops.py.

Parity status: UNPINNED by the reference (no tests, no stored activations; Keras/TF are not installable here).
Pins we do have: exact parameters + layer names/shapes from weights/*.h5 (tests/test_oracle.py), the semantic
smoke test on the reference's example figures (tests/golden/make_golden.py), fp64-vs-fp32 self consistency.

`W` is a dict: Keras layer name -> list of torch tensors in Keras `weight_names` order
(Conv2D [kernel(, bias)], DepthwiseConv2D [depthwise_kernel], BatchNormalization [gamma, beta, mean, var]).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ref_ops as R


class Ctx:
    """Forward context: training-mode BN uses batch statistics (Keras 2.2.4 does so even for frozen layers)."""

    def __init__(self, W: Dict[str, List[torch.Tensor]], training: bool = False, dropout_mask=None, tap=None):
        self.W = W
        self.training = training
        self.dropout_mask = dropout_mask     # explicit keep-mask (already scaled) or None = no dropout
        self.tap = tap                       # optional dict collecting named intermediates
        self.bn_batch_stats = OrderedDict()  # layer -> (mean, biased var, count) seen in training mode

    def bn(self, x, name, eps):
        g, b, m, v = self.W[name]
        if self.training:
            y, mean, var = R.batchnorm_train(x, g, b, eps)
            self.bn_batch_stats[name] = (mean.detach(), var.detach(), x.numel() // x.shape[-1])
            return y
        return R.batchnorm_infer(x, g, b, m, v, eps)

    def save(self, name, t):
        if self.tap is not None:
            self.tap[name] = t


def _make_divisible(v, divisor, min_value=None):
    """deeplabv3p.py:157-164."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def inverted_res_block(c: Ctx, x, expansion, stride, block_id, skip_connection, rate=1):
    """deeplabv3p.py:167-206."""
    inputs = x
    prefix = "expanded_conv_{}_".format(block_id)
    if block_id:
        x = R.pointwise(x, c.W[prefix + "expand"][0])
        x = c.bn(x, prefix + "expand_BN", 1e-3)
        x = R.relu6(x)
    else:
        prefix = "expanded_conv_"
    x = R.depthwise_same(x, c.W[prefix + "depthwise"][0], stride=stride, dilation=rate)
    x = c.bn(x, prefix + "depthwise_BN", 1e-3)
    x = R.relu6(x)
    x = R.pointwise(x, c.W[prefix + "project"][0])
    x = c.bn(x, prefix + "project_BN", 1e-3)
    if skip_connection:
        x = inputs + x
    c.save(prefix + "out", x)
    return x


MNV2_BLOCKS = [
    # (expansion, stride, block_id, skip, rate)   deeplabv3p.py:327-367
    (1, 1, 0, False, 1), (6, 2, 1, False, 1), (6, 1, 2, True, 1), (6, 2, 3, False, 1), (6, 1, 4, True, 1),
    (6, 1, 5, True, 1), (6, 1, 6, False, 1), (6, 1, 7, True, 2), (6, 1, 8, True, 2), (6, 1, 9, True, 2),
    (6, 1, 10, False, 2), (6, 1, 11, True, 2), (6, 1, 12, True, 2), (6, 1, 13, False, 2), (6, 1, 14, True, 4),
    (6, 1, 15, True, 4), (6, 1, 16, False, 4),
]


def mobilenetv2_backbone(c: Ctx, x):
    """deeplabv3p.py:315-367 (OS is forced to 8, :316).  x: preprocessed NHWC."""
    x = R.conv2d_same(x, c.W["Conv"][0], stride=2)
    x = c.bn(x, "Conv_BN", 1e-3)
    x = R.relu6(x)
    c.save("stem", x)
    for (t, s, bid, skip, rate) in MNV2_BLOCKS:
        x = inverted_res_block(c, x, t, s, bid, skip, rate)
    return x


def sepconv_bn(c: Ctx, x, prefix, stride=1, rate=1, depth_activation=False, epsilon=1e-3):
    """deeplabv3p.py:47-84."""
    if not depth_activation:
        x = torch.relu(x)
    wdw = c.W[prefix + "_depthwise"][0]
    if stride == 1:
        x = R.depthwise_same(x, wdw, 1, rate)
    else:
        x = R.depthwise_explicit(x, wdw, stride, rate)
    x = c.bn(x, prefix + "_depthwise_BN", epsilon)
    if depth_activation:
        x = torch.relu(x)
    x = R.pointwise(x, c.W[prefix + "_pointwise"][0])
    x = c.bn(x, prefix + "_pointwise_BN", epsilon)
    if depth_activation:
        x = torch.relu(x)
    return x


def conv2d_same_fixed(c: Ctx, x, prefix, stride=1, rate=1):
    """deeplabv3p.py:87-116."""
    w = c.W[prefix][0]
    if stride == 1:
        return R.conv2d_same(x, w, 1, rate)
    return R.conv2d_explicit(x, w, stride, rate)


def xception_block(c: Ctx, inputs, prefix, skip_connection_type, stride, rate=1, depth_activation=False,
                   return_skip=False):
    """deeplabv3p.py:119-155 (with the `layers.add` NameError read as its evident intent, Add)."""
    residual = inputs
    skip = None
    for i in range(3):
        residual = sepconv_bn(c, residual, prefix + "_separable_conv{}".format(i + 1),
                              stride=stride if i == 2 else 1, rate=rate, depth_activation=depth_activation)
        if i == 1:
            skip = residual
    if skip_connection_type == "conv":
        shortcut = conv2d_same_fixed(c, inputs, prefix + "_shortcut", stride=stride)
        shortcut = c.bn(shortcut, prefix + "_shortcut_BN", 1e-3)
        outputs = residual + shortcut
    elif skip_connection_type == "sum":
        outputs = residual + inputs
    else:
        outputs = residual
    return (outputs, skip) if return_skip else outputs


XCEPTION_SPEC = {
    8: dict(entry_block3_stride=1, middle_block_rate=2, exit_block_rates=(2, 4), atrous_rates=(12, 24, 36)),
    16: dict(entry_block3_stride=2, middle_block_rate=1, exit_block_rates=(1, 2), atrous_rates=(6, 12, 18)),
}


def xception_backbone(c: Ctx, x, OS):
    """deeplabv3p.py:272-313."""
    s = XCEPTION_SPEC[8 if OS == 8 else 16]
    x = R.conv2d_same(x, c.W["entry_flow_conv1_1"][0], stride=2)
    x = torch.relu(c.bn(x, "entry_flow_conv1_1_BN", 1e-3))
    x = conv2d_same_fixed(c, x, "entry_flow_conv1_2", stride=1)
    x = torch.relu(c.bn(x, "entry_flow_conv1_2_BN", 1e-3))
    x = xception_block(c, x, "entry_flow_block1", "conv", 2)
    x, skip1 = xception_block(c, x, "entry_flow_block2", "conv", 2, return_skip=True)
    x = xception_block(c, x, "entry_flow_block3", "conv", s["entry_block3_stride"])
    for i in range(16):
        x = xception_block(c, x, "middle_flow_unit_{}".format(i + 1), "sum", 1, rate=s["middle_block_rate"])
    x = xception_block(c, x, "exit_flow_block1", "conv", 1, rate=s["exit_block_rates"][0])
    x = xception_block(c, x, "exit_flow_block2", "none", 1, rate=s["exit_block_rates"][1], depth_activation=True)
    return x, skip1, s["atrous_rates"]


def deeplabv3_features(c: Ctx, img, backbone="mobilenetv2", OS=16):
    """Everything up to and including Dropout (= model.layers[-5].output for MobileNetV2, utils.py:181).
    img: NHWC float in 0..255.  Returns the 256-channel feature map (stride 8 MobileNetV2, stride 4 Xception)."""
    H, W_ = img.shape[1], img.shape[2]
    x = img / 127.5 - 1.0                                    # deeplabv3p.py:270
    if backbone == "xception":
        x, skip1, atrous_rates = xception_backbone(c, x, OS)
    else:
        OS = 8                                               # deeplabv3p.py:316
        x = mobilenetv2_backbone(c, x)
    c.save("backbone", x)
    fh, fw = int(math.ceil(H / OS)), int(math.ceil(W_ / OS))
    # image pooling branch, deeplabv3p.py:375-382 (AveragePooling2D(pool=(fh,fw)), strides=pool, VALID)
    b4 = x[:, : (x.shape[1] // fh) * fh, : (x.shape[2] // fw) * fw]
    b4 = b4.reshape(b4.shape[0], x.shape[1] // fh, fh, x.shape[2] // fw, fw, b4.shape[-1]).mean((2, 4))
    b4 = R.pointwise(b4, c.W["image_pooling"][0])
    b4 = torch.relu(c.bn(b4, "image_pooling_BN", 1e-5))
    b4 = R.resize_bilinear_tf1(b4, fh, fw)
    b0 = R.pointwise(x, c.W["aspp0"][0])
    b0 = torch.relu(c.bn(b0, "aspp0_BN", 1e-5))
    if backbone == "xception":
        b1 = sepconv_bn(c, x, "aspp1", rate=atrous_rates[0], depth_activation=True, epsilon=1e-5)
        b2 = sepconv_bn(c, x, "aspp2", rate=atrous_rates[1], depth_activation=True, epsilon=1e-5)
        b3 = sepconv_bn(c, x, "aspp3", rate=atrous_rates[2], depth_activation=True, epsilon=1e-5)
        x = torch.cat([b4, b0, b1, b2, b3], -1)
    else:
        x = torch.cat([b4, b0], -1)
    x = R.pointwise(x, c.W["concat_projection"][0])
    x = torch.relu(c.bn(x, "concat_projection_BN", 1e-5))
    if c.training and c.dropout_mask is not None:            # Dropout(0.1), deeplabv3p.py:410
        x = x * c.dropout_mask
    if backbone == "xception":                               # decoder, deeplabv3p.py:414-429
        x = R.resize_bilinear_tf1(x, int(math.ceil(H / 4)), int(math.ceil(W_ / 4)))
        d = R.pointwise(skip1, c.W["feature_projection0"][0])
        d = torch.relu(c.bn(d, "feature_projection0_BN", 1e-5))
        x = torch.cat([x, d], -1)
        x = sepconv_bn(c, x, "decoder_conv0", depth_activation=True, epsilon=1e-5)
        x = sepconv_bn(c, x, "decoder_conv1", depth_activation=True, epsilon=1e-5)
    c.save("features", x)
    return x


def head_forward(c: Ctx, feat, H, W_, head_layer: str, net: str = "original", scale: int = 8):
    """utils.py:188-198 ('original' / 'subpixel') and deeplabv3p.py:438-444 (bare Deeplabv3 == 'original' form).
    Returns (logits_lowres_or_full, probs [B, H*W, C])."""
    k, b = c.W[head_layer]
    y = R.pointwise(feat, k, b)
    if net == "subpixel":
        up = R.phase_shift_literal(y, scale)
    else:
        up = R.resize_bilinear_tf1(y, H, W_)
    c.save("logits", y)
    c.save("logits_up", up)
    probs = torch.softmax(up.reshape(up.shape[0], H * W_, -1), -1)
    return y, probs


def find_head_layer(W) -> str:
    """The last Conv2D-with-bias layer (logits_semantic / custom_logits_semantic / conv_upsample / subpixel_N)."""
    for name in reversed(list(W.keys())):
        ws = W[name]
        if len(ws) == 2 and ws[0].dim() == 4 and ws[1].dim() == 1:
            return name
    raise KeyError("no head layer")


def deeplabv3_forward(W, img, backbone="mobilenetv2", OS=16, net="original", training=False, dropout_mask=None,
                      tap=None):
    c = Ctx(W, training, dropout_mask, tap)
    feat = deeplabv3_features(c, img, backbone, OS)
    scale = 4 if backbone == "xception" else 8
    logits, probs = head_forward(c, feat, img.shape[1], img.shape[2], find_head_layer(W), net, scale)
    return logits, probs, c


def weights_from_h5(path: str, dtype=torch.float32) -> "OrderedDict[str, List[torch.Tensor]]":
    from .hdf5_reader import load_keras_weights
    layers, _ = load_keras_weights(path)
    out = OrderedDict()
    for name, ws in layers.items():
        if ws:
            out[name] = [torch.from_numpy(a).to(dtype) for _, a in ws]
    return out


# ---------------------------------------------------------------------------------------------------------
# seeded Keras-default initialisation for configurations without shipped weights
# ---------------------------------------------------------------------------------------------------------
def _glorot_uniform(rng, shape, fan_in, fan_out):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return torch.from_numpy(rng.uniform(-limit, limit, size=shape).astype(np.float32))


def _conv(rng, kh, cin, cout, bias=False):
    w = _glorot_uniform(rng, (kh, kh, cin, cout), kh * kh * cin, kh * kh * cout)
    return [w, torch.zeros(cout)] if bias else [w]


def _dw(rng, c):
    # Keras DepthwiseConv2D glorot_uniform on shape (3,3,C,1): fan_in = 9*C, fan_out = 9*1 (Keras _compute_fans)
    return [_glorot_uniform(rng, (3, 3, c, 1), 9 * c, 9)]


def _bn(rng, c, perturb=True):
    if perturb:   # non-trivial statistics so BN is exercised
        return [torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32)),
                torch.from_numpy(rng.uniform(-0.5, 0.5, c).astype(np.float32)),
                torch.from_numpy(rng.uniform(-0.5, 0.5, c).astype(np.float32)),
                torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32))]
    return [torch.ones(c), torch.zeros(c), torch.zeros(c), torch.ones(c)]


def random_mobilenetv2_weights(seed=0, classes=21, head="logits_semantic", head_filters=None, perturb_bn=True, alpha=1.0):
    """alpha: MobileNetV2 width multiplier -- channel counts as deeplabv3p.py:168-170 / :317 compute them."""
    rng = np.random.RandomState(seed)
    W = OrderedDict()
    c0 = _make_divisible(32 * alpha, 8)
    W["Conv"] = _conv(rng, 3, 3, c0)
    W["Conv_BN"] = _bn(rng, c0, perturb_bn)
    cin = c0
    outs = [16, 24, 24, 32, 32, 32, 64, 64, 64, 64, 96, 96, 96, 160, 160, 160, 320]
    outs = [_make_divisible(int(f * alpha), 8) for f in outs]
    for (t, s, bid, skip, rate), cout in zip(MNV2_BLOCKS, outs):
        prefix = "expanded_conv_{}_".format(bid) if bid else "expanded_conv_"
        mid = cin * t
        if bid:
            W[prefix + "expand"] = _conv(rng, 1, cin, mid)
            W[prefix + "expand_BN"] = _bn(rng, mid, perturb_bn)
        W[prefix + "depthwise"] = _dw(rng, mid)
        W[prefix + "depthwise_BN"] = _bn(rng, mid, perturb_bn)
        W[prefix + "project"] = _conv(rng, 1, mid, cout)
        W[prefix + "project_BN"] = _bn(rng, cout, perturb_bn)
        cin = cout
    W["image_pooling"] = _conv(rng, 1, cin, 256)
    W["image_pooling_BN"] = _bn(rng, 256, perturb_bn)
    W["aspp0"] = _conv(rng, 1, cin, 256)
    W["aspp0_BN"] = _bn(rng, 256, perturb_bn)
    W["concat_projection"] = _conv(rng, 1, 512, 256)
    W["concat_projection_BN"] = _bn(rng, 256, perturb_bn)
    hf = head_filters if head_filters is not None else classes
    W[head] = _conv(rng, 1, 256, hf, bias=True)
    W[head][1] = torch.from_numpy(rng.uniform(-0.1, 0.1, hf).astype(np.float32))
    return W


def random_xception_weights(seed=0, classes=21, head="logits_semantic", perturb_bn=True):
    rng = np.random.RandomState(seed)
    W = OrderedDict()

    def sep(prefix, cin, cout):
        W[prefix + "_depthwise"] = _dw(rng, cin)
        W[prefix + "_depthwise_BN"] = _bn(rng, cin, perturb_bn)
        W[prefix + "_pointwise"] = _conv(rng, 1, cin, cout)
        W[prefix + "_pointwise_BN"] = _bn(rng, cout, perturb_bn)

    def block(prefix, cin, depths, skip):
        c = cin
        for i, d in enumerate(depths):
            sep(prefix + "_separable_conv{}".format(i + 1), c, d)
            c = d
        if skip == "conv":
            W[prefix + "_shortcut"] = _conv(rng, 1, cin, depths[-1])
            W[prefix + "_shortcut_BN"] = _bn(rng, depths[-1], perturb_bn)
        return c

    W["entry_flow_conv1_1"] = _conv(rng, 3, 3, 32)
    W["entry_flow_conv1_1_BN"] = _bn(rng, 32, perturb_bn)
    W["entry_flow_conv1_2"] = _conv(rng, 3, 32, 64)
    W["entry_flow_conv1_2_BN"] = _bn(rng, 64, perturb_bn)
    c = block("entry_flow_block1", 64, [128, 128, 128], "conv")
    c = block("entry_flow_block2", c, [256, 256, 256], "conv")
    c = block("entry_flow_block3", c, [728, 728, 728], "conv")
    for i in range(16):
        c = block("middle_flow_unit_{}".format(i + 1), c, [728, 728, 728], "sum")
    c = block("exit_flow_block1", c, [728, 1024, 1024], "conv")
    c = block("exit_flow_block2", c, [1536, 1536, 2048], "none")
    W["image_pooling"] = _conv(rng, 1, 2048, 256)
    W["image_pooling_BN"] = _bn(rng, 256, perturb_bn)
    W["aspp0"] = _conv(rng, 1, 2048, 256)
    W["aspp0_BN"] = _bn(rng, 256, perturb_bn)
    for i in (1, 2, 3):
        sep("aspp{}".format(i), 2048, 256)
    W["concat_projection"] = _conv(rng, 1, 1280, 256)
    W["concat_projection_BN"] = _bn(rng, 256, perturb_bn)
    W["feature_projection0"] = _conv(rng, 1, 256, 48)
    W["feature_projection0_BN"] = _bn(rng, 48, perturb_bn)
    sep("decoder_conv0", 304, 256)
    sep("decoder_conv1", 256, 256)
    W[head] = _conv(rng, 1, 256, classes, bias=True)
    return W
