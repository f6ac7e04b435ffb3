"""ORACLE (test infrastructure only): SECOND, independently written restatement of the reference graph.

Purpose (SURVEY 8c (4), round-1 review): `oracle/network.py` (torch, NHWC, F.conv2d, autograd) is what the CUDA path
is compared against; this file restates the same reference code a second time with nothing shared -- numpy float64,
NCHW, convolutions as explicit tap loops over strided slices, TF-SAME / explicit padding computed from first
principles per axis, the legacy bilinear resize and the phase shift as direct index loops, BatchNorm from its
definition -- so that a mistake in one restatement shows up as a disagreement between the two
(`tests/test_oracle.py::test_second_restatement_*`).  It also gives a gradient check that does not use autograd:
central finite differences of the fp64 loss along random directions vs the autograd gradient of network.py.

Followed line by line from /root/reference (NOT from network.py):
  deeplabv3p.py:47-84 SepConv_BN, :87-116 _conv2d_same, :119-155 _xception_block, :167-206 _inverted_res_block,
  :260-444 Deeplabv3 body, subpixel.py:77-88 _phase_shift, utils.py:127-130 loss, utils.py:188-197 heads.
Third-party op semantics: SURVEY Appendix B.  Parity status: unpinned by the reference (no tests / goldens exist).
"""
from __future__ import annotations

import math

import numpy as np

F64 = np.float64


# ------------------------------------------------------------------------------------------------- primitives
def _same_pads(n, k, s, d):
    """TF 'SAME' along one axis: output ceil(n/s); total padding so that the last window fits; the odd pixel goes
    to the END (Appendix B.1)."""
    out = (n + s - 1) // s
    span = (k - 1) * d + 1
    need = (out - 1) * s + span - n
    if need < 0:
        need = 0
    lo = need // 2
    return out, lo, need - lo


def _fixed_pads(k, d):
    """deeplabv3p.py:61-66 / :106-111: pad_total = k_eff - 1 split beg = total//2, end = rest; then VALID."""
    k_eff = k + (k - 1) * (d - 1)
    total = k_eff - 1
    beg = total // 2
    return beg, total - beg


def _pad_hw(x, top, bottom, left, right):
    n, c, h, w = x.shape
    y = np.zeros((n, c, h + top + bottom, w + left + right), F64)
    y[:, :, top:top + h, left:left + w] = x
    return y


def _windows(xp, i, j, d, s, oh, ow):
    """the input samples tap (i, j) sees for every output position: xp[.., y*s + i*d, x*s + j*d]"""
    return xp[:, :, i * d: i * d + (oh - 1) * s + 1: s, j * d: j * d + (ow - 1) * s + 1: s]


def conv2d(x, w_hwio, stride=1, rate=1, mode="same"):
    """Conv2D, no bias.  x [N,C,H,W]; Keras kernel [kh,kw,Cin,Cout].  mode: 'same' (TF) or 'fixed' (ZeroPadding2D
    ((beg,end)) + VALID, the stride>1 branch of _conv2d_same)."""
    k = w_hwio.shape[0]
    n, c, h, w = x.shape
    if mode == "same":
        oh, pt, pb = _same_pads(h, k, stride, rate)
        ow, pl, pr = _same_pads(w, k, stride, rate)
    else:
        pt, pb = _fixed_pads(k, rate)
        pl, pr = pt, pb
        span = (k - 1) * rate + 1
        oh = (h + pt + pb - span) // stride + 1
        ow = (w + pl + pr - span) // stride + 1
    xp = _pad_hw(x, pt, pb, pl, pr)
    out = np.zeros((n, w_hwio.shape[3], oh, ow), F64)
    for i in range(k):
        for j in range(k):
            win = _windows(xp, i, j, rate, stride, oh, ow)                  # [N,C,oh,ow]
            out += np.tensordot(w_hwio[i, j].astype(F64), win, axes=([0], [1])).transpose(1, 0, 2, 3)
    return out


def depthwise(x, w_hwc1, stride=1, rate=1, mode="same"):
    """DepthwiseConv2D 3x3, no bias; Keras depthwise_kernel [3,3,C,1]."""
    n, c, h, w = x.shape
    k = w_hwc1.shape[0]
    if mode == "same":
        oh, pt, pb = _same_pads(h, k, stride, rate)
        ow, pl, pr = _same_pads(w, k, stride, rate)
    else:
        pt, pb = _fixed_pads(k, rate)
        pl, pr = pt, pb
        span = (k - 1) * rate + 1
        oh = (h + pt + pb - span) // stride + 1
        ow = (w + pl + pr - span) // stride + 1
    xp = _pad_hw(x, pt, pb, pl, pr)
    out = np.zeros((n, c, oh, ow), F64)
    for i in range(k):
        for j in range(k):
            out += _windows(xp, i, j, rate, stride, oh, ow) * w_hwc1[i, j, :, 0].astype(F64)[None, :, None, None]
    return out


def conv1x1(x, w_hwio, bias=None):
    out = np.einsum("nchw,co->nohw", x, w_hwio[0, 0].astype(F64))
    if bias is not None:
        out = out + bias.astype(F64)[None, :, None, None]
    return out


def resize_bilinear_legacy(x, oh, ow):
    """TF1 tf.image.resize_bilinear defaults: src = dst * (in/out) (scale in float32 like TF's kernel), lower index
    floor(src), upper min(lower+1, in-1), weight src - lower; rows then columns.  Direct per-output-line loops."""
    n, c, h, w = x.shape
    sy = np.float32(h) / np.float32(oh)
    sx = np.float32(w) / np.float32(ow)
    rows = np.zeros((n, c, oh, w), F64)
    for y in range(oh):
        src = float(np.float32(y) * sy)
        lo = int(math.floor(src))
        hi = lo + 1 if lo + 1 < h else h - 1
        t = src - lo
        rows[:, :, y, :] = x[:, :, lo, :] + (x[:, :, hi, :] - x[:, :, lo, :]) * t
    out = np.zeros((n, c, oh, ow), F64)
    for xx in range(ow):
        src = float(np.float32(xx) * sx)
        lo = int(math.floor(src))
        hi = lo + 1 if lo + 1 < w else w - 1
        t = src - lo
        out[:, :, :, xx] = rows[:, :, :, lo] + (rows[:, :, :, hi] - rows[:, :, :, lo]) * t
    return out


def phase_shift_loops(x_nhwc, r):
    """subpixel.py:77-88 executed literally on index tuples: reshape to (b, a, bb, c/r^2, r, r), permute
    (0,1,2,5,4,3), then the two rounds of `[X[:, i] for i in range(..)]` + concatenate on axis 2."""
    bsz, a, b, c = x_nhwc.shape
    cs = c // (r * r)
    X = x_nhwc.reshape(bsz, a, b, cs, r, r)
    X = np.transpose(X, (0, 1, 2, 5, 4, 3))                         # bsz, a, b, r, r, cs
    X = np.concatenate([X[:, i] for i in range(a)], axis=2)         # bsz, b, a*r, r, cs
    X = np.concatenate([X[:, i] for i in range(b)], axis=2)         # bsz, a*r, b*r, cs
    return X


class Net:
    """Graph walk with the weights of one model; training=True uses batch statistics in every BatchNorm."""

    def __init__(self, W, training=False, dropout_mask=None):
        self.W = {k: [np.asarray(t, dtype=F64) for t in v] for k, v in W.items()}
        self.training = training
        self.dropout_mask = dropout_mask
        self.batch_stats = {}

    def bn(self, x, name, eps):
        g, b, mu, var = self.W[name]
        if self.training:
            mu = x.mean(axis=(0, 2, 3))
            var = ((x - mu[None, :, None, None]) ** 2).mean(axis=(0, 2, 3))
            self.batch_stats[name] = (mu, var, x.shape[0] * x.shape[2] * x.shape[3])
        inv = 1.0 / np.sqrt(var + eps)
        return (x - mu[None, :, None, None]) * (g * inv)[None, :, None, None] + b[None, :, None, None]

    # deeplabv3p.py:167-206
    def inverted_res_block(self, x, expansion, stride, block_id, skip, rate=1):
        inp = x
        p = "expanded_conv_%d_" % block_id
        if block_id:
            x = conv1x1(x, self.W[p + "expand"][0])
            x = np.clip(self.bn(x, p + "expand_BN", 1e-3), 0.0, 6.0)
        else:
            p = "expanded_conv_"
        x = depthwise(x, self.W[p + "depthwise"][0], stride, rate, "same")
        x = np.clip(self.bn(x, p + "depthwise_BN", 1e-3), 0.0, 6.0)
        x = conv1x1(x, self.W[p + "project"][0])
        x = self.bn(x, p + "project_BN", 1e-3)
        return inp + x if skip else x

    # deeplabv3p.py:47-84
    def sepconv_bn(self, x, prefix, stride=1, rate=1, depth_activation=False, eps=1e-3):
        if not depth_activation:
            x = np.maximum(x, 0.0)
        x = depthwise(x, self.W[prefix + "_depthwise"][0], stride, rate, "same" if stride == 1 else "fixed")
        x = self.bn(x, prefix + "_depthwise_BN", eps)
        if depth_activation:
            x = np.maximum(x, 0.0)
        x = conv1x1(x, self.W[prefix + "_pointwise"][0])
        x = self.bn(x, prefix + "_pointwise_BN", eps)
        if depth_activation:
            x = np.maximum(x, 0.0)
        return x

    # deeplabv3p.py:119-155
    def xception_block(self, inputs, prefix, skip_type, stride, rate=1, depth_activation=False, return_skip=False):
        res = inputs
        skip = None
        for i in range(3):
            res = self.sepconv_bn(res, "%s_separable_conv%d" % (prefix, i + 1), stride if i == 2 else 1, rate,
                                  depth_activation)
            if i == 1:
                skip = res
        if skip_type == "conv":
            w = self.W[prefix + "_shortcut"][0]
            sc = conv2d(inputs, w, stride, 1, "same" if stride == 1 else "fixed")
            sc = self.bn(sc, prefix + "_shortcut_BN", 1e-3)
            out = res + sc
        elif skip_type == "sum":
            out = res + inputs
        else:
            out = res
        return (out, skip) if return_skip else out

    def forward(self, img_nhwc, backbone="mobilenetv2", OS=16, net="original", head=None):
        """img [N,H,W,3] in 0..255 -> (low-res logits NHWC, probabilities [N, H*W, C])."""
        H, Wd = img_nhwc.shape[1], img_nhwc.shape[2]
        x = np.transpose(np.asarray(img_nhwc, F64), (0, 3, 1, 2)) / 127.5 - 1.0
        skip1 = None
        if backbone == "xception":
            if OS == 8:
                b3s, mid_rate, exit_rates, atrous = 1, 2, (2, 4), (12, 24, 36)
            else:
                b3s, mid_rate, exit_rates, atrous = 2, 1, (1, 2), (6, 12, 18)
            x = np.maximum(self.bn(conv2d(x, self.W["entry_flow_conv1_1"][0], 2, 1, "same"), "entry_flow_conv1_1_BN", 1e-3), 0)
            x = np.maximum(self.bn(conv2d(x, self.W["entry_flow_conv1_2"][0], 1, 1, "same"), "entry_flow_conv1_2_BN", 1e-3), 0)
            x = self.xception_block(x, "entry_flow_block1", "conv", 2)
            x, skip1 = self.xception_block(x, "entry_flow_block2", "conv", 2, return_skip=True)
            x = self.xception_block(x, "entry_flow_block3", "conv", b3s)
            for i in range(16):
                x = self.xception_block(x, "middle_flow_unit_%d" % (i + 1), "sum", 1, mid_rate)
            x = self.xception_block(x, "exit_flow_block1", "conv", 1, exit_rates[0])
            x = self.xception_block(x, "exit_flow_block2", "none", 1, exit_rates[1], depth_activation=True)
        else:
            OS = 8
            x = np.clip(self.bn(conv2d(x, self.W["Conv"][0], 2, 1, "same"), "Conv_BN", 1e-3), 0.0, 6.0)
            plan = [(1, 1, 0, False, 1), (6, 2, 1, False, 1), (6, 1, 2, True, 1), (6, 2, 3, False, 1), (6, 1, 4, True, 1),
                    (6, 1, 5, True, 1), (6, 1, 6, False, 1), (6, 1, 7, True, 2), (6, 1, 8, True, 2), (6, 1, 9, True, 2),
                    (6, 1, 10, False, 2), (6, 1, 11, True, 2), (6, 1, 12, True, 2), (6, 1, 13, False, 2),
                    (6, 1, 14, True, 4), (6, 1, 15, True, 4), (6, 1, 16, False, 4)]
            for (t, s, bid, skip, rate) in plan:
                x = self.inverted_res_block(x, t, s, bid, skip, rate)
        ph, pw = int(np.ceil(H / OS)), int(np.ceil(Wd / OS))
        # AveragePooling2D(pool_size=(ph,pw)) -> strides = pool, VALID
        n, c, h, w = x.shape
        qh, qw = h // ph, w // pw
        b4 = np.zeros((n, c, qh, qw), F64)
        for i in range(qh):
            for j in range(qw):
                b4[:, :, i, j] = x[:, :, i * ph:(i + 1) * ph, j * pw:(j + 1) * pw].mean(axis=(2, 3))
        b4 = np.maximum(self.bn(conv1x1(b4, self.W["image_pooling"][0]), "image_pooling_BN", 1e-5), 0)
        b4 = resize_bilinear_legacy(b4, ph, pw)
        b0 = np.maximum(self.bn(conv1x1(x, self.W["aspp0"][0]), "aspp0_BN", 1e-5), 0)
        if backbone == "xception":
            bs = [self.sepconv_bn(x, "aspp%d" % (i + 1), 1, atrous[i], True, 1e-5) for i in range(3)]
            x = np.concatenate([b4, b0] + bs, axis=1)
        else:
            x = np.concatenate([b4, b0], axis=1)
        x = np.maximum(self.bn(conv1x1(x, self.W["concat_projection"][0]), "concat_projection_BN", 1e-5), 0)
        if self.training and self.dropout_mask is not None:
            x = x * np.transpose(np.asarray(self.dropout_mask, F64), (0, 3, 1, 2))
        if backbone == "xception":
            x = resize_bilinear_legacy(x, int(np.ceil(H / 4)), int(np.ceil(Wd / 4)))
            d = np.maximum(self.bn(conv1x1(skip1, self.W["feature_projection0"][0]), "feature_projection0_BN", 1e-5), 0)
            x = np.concatenate([x, d], axis=1)
            x = self.sepconv_bn(x, "decoder_conv0", 1, 1, True, 1e-5)
            x = self.sepconv_bn(x, "decoder_conv1", 1, 1, True, 1e-5)
        if head is None:
            head = [k for k, v in self.W.items() if len(v) == 2 and v[0].ndim == 4 and v[1].ndim == 1][-1]
        k, b = self.W[head]
        y = conv1x1(x, k, b)
        y_nhwc = np.transpose(y, (0, 2, 3, 1))
        if net == "subpixel":
            scale = 4 if backbone == "xception" else 8
            up = phase_shift_loops(y_nhwc, scale)
        else:
            up = np.transpose(resize_bilinear_legacy(y, H, Wd), (0, 2, 3, 1))
        z = up.reshape(up.shape[0], H * Wd, -1)
        z = z - z.max(-1, keepdims=True)
        e = np.exp(z)
        return y_nhwc, e / e.sum(-1, keepdims=True)


def keras_loss(y_true, probs, sample_w=None):
    """utils.py:127-130 through Keras' categorical_crossentropy + weighted_masked_objective (Appendix B.5)."""
    nb = probs.shape[-1]
    lab = np.asarray(y_true)[:, :, 0].astype(np.int64)
    p = probs / probs.sum(-1, keepdims=True)
    p = np.clip(p, 1e-7, 1 - 1e-7)
    score = np.zeros(lab.shape, F64)
    for c in range(nb):                       # one_hot(.., nb+1)[:, :, :-1]: label nb (void) selects nothing
        m = lab == c
        score[m] = -np.log(p[..., c][m])
    if sample_w is not None:
        sw = np.asarray(sample_w, F64)
        score = score * sw
        score = score / (sw != 0).mean()
    return float(score.mean())


def training_loss(W, img, y, sw=None, net="original", backbone="mobilenetv2"):
    probs = Net(W, training=True).forward(img, backbone=backbone, net=net)[1]
    return keras_loss(y, probs, sw)
