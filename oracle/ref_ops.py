"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement, in torch-CPU / numpy, of the third-party op semantics the reference's graph relies on
(Keras 2.2.4 / TF 1.x; not vendored under /root/reference).  Each function cites the
reference call site it stands in for.  Parity status: UNPINNED by the reference (it ships no tests and no golden
activations); pinned instead by (a) closed-form vs literal-emulation cross checks in tests/test_oracle.py,
(b) the weight-file known-answer tests, (c) the semantic smoke test on the reference's example figures.

Everything is differentiable torch so autograd provides the backward oracle.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------
# padding / convolution  (Conv2D / DepthwiseConv2D padding='same', ZeroPadding2D; deeplabv3p.py:61-74,:99-116,:186,:318)
# ---------------------------------------------------------------------------------------------------------
def tf_same_pad(in_size: int, k: int, stride: int, dilation: int):
    """TF SAME: out = ceil(in/s); pad_total = max((out-1)*s + (k-1)*d + 1 - in, 0); extra pixel goes AFTER."""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + (k - 1) * dilation + 1 - in_size, 0)
    return out, total // 2, total - total // 2


def _pad_same(x_nchw, k, stride, dilation):
    H, W = x_nchw.shape[2], x_nchw.shape[3]
    _, pt, pb = tf_same_pad(H, k, stride, dilation)
    _, pl, pr = tf_same_pad(W, k, stride, dilation)
    return F.pad(x_nchw, (pl, pr, pt, pb))


def conv2d_same(x, w_hwio, stride=1, dilation=1):
    """x NHWC, w HWIO (Keras kernel layout), padding='same' -> NHWC."""
    k = w_hwio.shape[0]
    xp = _pad_same(x.permute(0, 3, 1, 2), k, stride, dilation)
    y = F.conv2d(xp, w_hwio.permute(3, 2, 0, 1), stride=stride, dilation=dilation)
    return y.permute(0, 2, 3, 1)


def conv2d_explicit(x, w_hwio, stride, dilation):
    """_conv2d_same / SepConv_BN stride>1 branch (deeplabv3p.py:61-69,:106-116): explicit symmetric-ish pad + VALID."""
    k = w_hwio.shape[0]
    k_eff = k + (k - 1) * (dilation - 1)
    pad_total = k_eff - 1
    pb, pe = pad_total // 2, pad_total - pad_total // 2
    xp = F.pad(x.permute(0, 3, 1, 2), (pb, pe, pb, pe))
    y = F.conv2d(xp, w_hwio.permute(3, 2, 0, 1), stride=stride, dilation=dilation)
    return y.permute(0, 2, 3, 1)


def depthwise_same(x, w_hwc1, stride=1, dilation=1):
    """DepthwiseConv2D(3, padding='same') with Keras depthwise_kernel (3,3,C,1) (deeplabv3p.py:186-188)."""
    C = x.shape[-1]
    w = w_hwc1.reshape(3, 3, C)
    xp = _pad_same(x.permute(0, 3, 1, 2), 3, stride, dilation)
    y = F.conv2d(xp, w.permute(2, 0, 1).unsqueeze(1), stride=stride, dilation=dilation, groups=C)
    return y.permute(0, 2, 3, 1)


def depthwise_explicit(x, w_hwc1, stride, dilation):
    C = x.shape[-1]
    w = w_hwc1.reshape(3, 3, C)
    k_eff = 3 + 2 * (dilation - 1)
    pad_total = k_eff - 1
    pb, pe = pad_total // 2, pad_total - pad_total // 2
    xp = F.pad(x.permute(0, 3, 1, 2), (pb, pe, pb, pe))
    y = F.conv2d(xp, w.permute(2, 0, 1).unsqueeze(1), stride=stride, dilation=dilation, groups=C)
    return y.permute(0, 2, 3, 1)


def pointwise(x, w_hwio, bias=None):
    """Conv2D(filters, (1,1)) (deeplabv3p.py:78,:175,:194,:385,:406,:438)."""
    y = x @ w_hwio.reshape(w_hwio.shape[2], w_hwio.shape[3])
    return y if bias is None else y + bias


# ---------------------------------------------------------------------------------------------------------
# BatchNormalization (Keras 2.2.4, TF backend)   deeplabv3p.py:76,:80,:178,:189,:197,:322,:379,:386,:408
# ---------------------------------------------------------------------------------------------------------
def batchnorm_infer(x, gamma, beta, mean, var, eps):
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def batchnorm_train(x, gamma, beta, eps):
    """Normalise with the biased batch statistics over all axes but the last; returns (y, mean, biased var)."""
    dims = tuple(range(x.dim() - 1))
    mean = x.mean(dims)
    var = ((x - mean) ** 2).mean(dims)
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta, mean, var


def relu6(x):
    """Lambda(lambda x: relu(x, max_value=6.)) (deeplabv3p.py:181,:192,:325)."""
    return x.clamp(0.0, 6.0)


# ---------------------------------------------------------------------------------------------------------
# legacy TF1 bilinear resize   K.tf.image.resize_bilinear(x, size)  deeplabv3p.py:382,:418,:439 ; utils.py:190
# ---------------------------------------------------------------------------------------------------------
def _axis_coeffs(out_size, in_size, dtype):
    scale = torch.tensor(in_size / out_size, dtype=torch.float32)       # TF computes the scale in float32
    src = torch.arange(out_size, dtype=torch.float32) * scale
    lo = torch.floor(src).long()
    hi = torch.clamp(lo + 1, max=in_size - 1)
    frac = (src - lo.float()).to(dtype)
    return lo, hi, frac


def resize_bilinear_tf1(x, H, W):
    """align_corners=False, half_pixel_centers=False: src = dst * in/out, hi = min(lo+1, in-1). x NHWC."""
    h, w = x.shape[1], x.shape[2]
    y0, y1, fy = _axis_coeffs(H, h, x.dtype)
    x0, x1, fx = _axis_coeffs(W, w, x.dtype)
    fy = fy.view(1, H, 1, 1)
    fx = fx.view(1, 1, W, 1)
    top_rows, bot_rows = x[:, y0], x[:, y1]
    tl, tr = top_rows[:, :, x0], top_rows[:, :, x1]
    bl, br = bot_rows[:, :, x0], bot_rows[:, :, x1]
    top = tl + (tr - tl) * fx
    bot = bl + (br - bl) * fx
    return top + (bot - top) * fy


# ---------------------------------------------------------------------------------------------------------
# Subpixel   subpixel.py:77-88 (_phase_shift), :13-39 (ICNR)
# ---------------------------------------------------------------------------------------------------------
def phase_shift_literal(I, r):
    """Literal emulation of the reference op sequence (reshape / permute / per-row slice + concat)."""
    bsize, a, b, c = I.shape
    X = I.reshape(bsize, a, b, c // (r * r), r, r)
    X = X.permute(0, 1, 2, 5, 4, 3)                       # bsize, a, b, r, r, c/(r*r)
    X = [X[:, i] for i in range(a)]                       # a x [bsize, b, r, r, c']
    X = torch.cat(X, 2)                                   # bsize, b, a*r, r, c'
    X = [X[:, i] for i in range(b)]                       # b x [bsize, a*r, r, c']
    X = torch.cat(X, 2)                                   # bsize, a*r, b*r, c'
    return X


def phase_shift(I, r):
    """Closed form: out[n, a*r+j, b*r+i, k] = in[n, a, b, k*r*r + i*r + j]."""
    n, a, b, c = I.shape
    cs = c // (r * r)
    X = I.reshape(n, a, b, cs, r, r)            # [n, a, b, k, i, j]
    X = X.permute(0, 1, 5, 2, 4, 3)             # [n, a, j, b, i, k]
    return X.reshape(n, a * r, b * r, cs)


def subpixel_column_perm(cs: int, r: int) -> np.ndarray:
    """perm[j'] = Keras column feeding internal column j' = (jj*r + i)*cs + k  (Keras column k*r*r + i*r + jj).

    With the weight columns stored in this order, one GEMM row writes, for each jj, a contiguous run of r*cs
    output elements out[n, a*r+jj, b*r : (b+1)*r, :] -- the fused phase-shift store of dlb_pw_gemm."""
    perm = np.empty(cs * r * r, dtype=np.int64)
    for jj in range(r):
        for i in range(r):
            for k in range(cs):
                perm[(jj * r + i) * cs + k] = k * r * r + i * r + jj
    return perm


def space_to_depth(x, bs):
    """tf.space_to_depth NHWC: out[n,h,w,(dy*bs+dx)*C + c] = in[n,h*bs+dy,w*bs+dx,c]."""
    n, H, W, c = x.shape
    x = x.reshape(n, H // bs, bs, W // bs, bs, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(n, H // bs, W // bs, bs * bs * c)


def icnr(sub_kernel, scale):
    """ICNR.__call__ (subpixel.py:27-39) given the already-sampled sub-kernel [kh,kw,Cin,Cout/s^2]."""
    x = sub_kernel.permute(2, 0, 1, 3)                                   # [Cin, kh, kw, C']
    kh, kw = x.shape[1], x.shape[2]
    x = x.repeat_interleave(scale, 1).repeat_interleave(scale, 2)        # resize_nearest_neighbor x scale
    assert x.shape[1] == kh * scale and x.shape[2] == kw * scale
    x = space_to_depth(x, scale)
    return x.permute(1, 2, 0, 3)


# ---------------------------------------------------------------------------------------------------------
# loss / metrics   utils.py:127-157 ; ipynb:203-210 ; Keras weighted_masked_objective (temporal sample weights)
# ---------------------------------------------------------------------------------------------------------
def sparse_crossentropy_ignoring_last_label(y_true, y_pred):
    """utils.py:127-130 + keras.backend.categorical_crossentropy on probabilities. -> [B, T]"""
    nb = y_pred.shape[-1]
    lab = y_true[:, :, 0].long()
    onehot = F.one_hot(lab.clamp(0, nb), nb + 1)[:, :, :-1].to(y_pred.dtype)
    p = y_pred / y_pred.sum(-1, keepdim=True)
    p = p.clamp(1e-7, 1 - 1e-7)
    return -(onehot * torch.log(p)).sum(-1)


def keras_weighted_loss(y_true, y_pred, sample_w=None):
    """score *= w ; score /= mean(w != 0) ; mean(score)   (Keras 2.2.4 training_utils.weighted_masked_objective)."""
    score = sparse_crossentropy_ignoring_last_label(y_true, y_pred)
    if sample_w is not None:
        score = score * sample_w
        score = score / (sample_w != 0).to(score.dtype).mean()
    return score.mean()


def sparse_accuracy_ignoring_last_label(y_true, y_pred):
    """utils.py:132-138."""
    nb = y_pred.shape[-1]
    pred = y_pred.reshape(-1, nb).argmax(-1)
    t = y_true.reshape(-1).long()
    legal = t != nb
    return ((t == pred) & legal).float().sum() / legal.float().sum()


def jaccard(y_true, y_pred):
    """utils.py:139-157: per class, mean IoU over the samples that contain the class; NaN classes dropped."""
    nb = y_pred.shape[-1]
    pred = y_pred.argmax(-1)
    t = y_true[:, :, 0].long()
    ious = []
    for i in range(nb):
        tl, pl = t == i, pred == i
        inter = (tl & pl).sum(1).double()
        union = (tl | pl).sum(1).double()
        legal = tl.sum(1) > 0
        if legal.any():
            ious.append((inter[legal] / union[legal]).mean())
    return torch.stack(ious).mean() if ious else torch.tensor(float("nan"))


def notebook_miou(gt, pred):
    """segmentation.ipynb mIOU (ipynb:203-210): mean IoU over the labels present in gt."""
    ious = []
    for l in np.unique(gt):
        inter = np.logical_and(gt == l, pred == l).sum()
        union = np.logical_or(gt == l, pred == l).sum()
        ious.append(inter / union)
    return float(np.mean(ious))


# ---------------------------------------------------------------------------------------------------------
# Keras 2.2.4 Adam (ipynb:107: Adam(lr=7e-4, epsilon=1e-8, decay=1e-6))
# ---------------------------------------------------------------------------------------------------------
def keras_adam(p, g, m, v, iterations, lr=7e-4, beta1=0.9, beta2=0.999, eps=1e-8, decay=0.0):
    lr_t = lr
    if decay > 0:
        lr_t = lr_t * (1.0 / (1.0 + decay * iterations))
    t = iterations + 1
    lr_t = lr_t * (math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t))
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    p = p - lr_t * m / (torch.sqrt(v) + eps)
    return p, m, v

# ---------------------------------------------------------------------------------------------------------
# data formats either side of the hot path (test infrastructure, like everything in oracle/)
# ---------------------------------------------------------------------------------------------------------
def generator_labels_and_weights(label, n_classes):
    """Restatement of SegmentationGenerator.__getitem__'s label handling for ONE image (reference utils.py:360-399):
    void remap (:360-365), then adaptive per-pixel weights from sklearn's compute_class_weight('balanced', classes,
    y) = n_samples / (n_classes_present * bincount(y)) over the non-void pixels (:388-397), void weight 0 (:399).
    label: int array of any shape; returns (y [P] int32, sw [P] float32)."""
    y = np.asarray(label).astype(np.int64).flatten()
    y = np.where((y < 0) | (y > n_classes - 1), n_classes, y)         # setxor1d remap + `y[y>(n_classes-1)] = n_classes`
    filt = y[y != n_classes]
    sw = np.zeros(y.shape, dtype=np.float32)
    u = np.unique(filt)
    if len(u):
        counts = np.bincount(filt, minlength=n_classes)[u]
        w = len(filt) / (len(u) * counts.astype(np.float64))          # float64, as sklearn
        for cls, wc in zip(u, w):
            np.putmask(sw, y == cls, wc)                              # stored into the float32 SW buffer
    return y.astype(np.int32), sw


def calculate_iou_conf(pred_argmax, label, nb_classes):
    """The counting loop of segmentation.ipynb cell 10 (`calculate_iou`), vectorised: conf_m[l-1, p-1] += 1 for every
    pixel whose label is not void."""
    p = np.ravel(pred_argmax).astype(int)
    l = np.ravel(label).astype(int)
    conf = np.zeros((nb_classes, nb_classes), dtype=float)
    keep = (l != nb_classes) & (l < nb_classes) & (p < nb_classes)
    np.add.at(conf, ((l[keep] - 1) % nb_classes, (p[keep] - 1) % nb_classes), 1)
    return conf
