"""The hot-path part of the reference's utils.py on the B200-native engine.

  SegModel.create_seg_model / train_generator / train / load_weights / set_*   utils.py:160-254
  sparse_crossentropy_ignoring_last_label, sparse_accuracy_ignoring_last_label, Jaccard   utils.py:127-157
  do_crf                                                                      utils.py:74-91
  get_VOC2012_classes                                                         utils.py:99-124
SegmentationGenerator's per-image augmentations + label / weight contract run on the device (augment_batch);
its file handling (glob / cv2.imread / resize of VOC files) and plot_confusion_matrix stay out of scope.

The loss / metric callables are accepted by `model.compile(...)` for API compatibility; the training step always
uses the fused CUDA implementation of exactly these functions (dlb_resize_softmax_ce / dlb_confusion).  Calling
them directly works on CUDA tensors and runs the same kernels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L
from . import ops
from .engine import Engine
from .model import Model, _metrics_from_confusion, keras_layer_names
from .subpixel import Subpixel, icnr_weights  # noqa: F401


def get_VOC2012_classes():
    names = ['background', 'airplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow',
             'table', 'dog', 'horse', 'motorbike', 'person', 'potted_plant', 'sheep', 'sofa', 'train', 'tv', 'void']
    return dict(enumerate(names))


# ---------------------------------------------------------------------------------------------------------
# loss / metrics on CUDA tensors (y_true [B, T, 1] float labels with `classes` = void, y_pred [B, T, C] probs)
# ---------------------------------------------------------------------------------------------------------
def _as_cuda(t):
    t = torch.as_tensor(t, dtype=torch.float32)
    return t.cuda().contiguous() if not t.is_cuda else t.contiguous()


def sparse_crossentropy_ignoring_last_label(y_true, y_pred):
    """utils.py:127-130 -> per-pixel loss [B, T] on CUDA tensors.

    Convenience entry point for user code that evaluates the loss on its own probabilities (a gather + log, a few
    torch device ops).  The training step never calls it: there the same definition is fused with the bilinear
    resize, the softmax and the backward pass in dlb_resize_softmax_ce (head_ops.cu)."""
    y_true, y_pred = _as_cuda(y_true), _as_cuda(y_pred)
    Cn = y_pred.shape[-1]
    lab = y_true[:, :, 0].long()
    valid = (lab >= 0) & (lab < Cn)
    p = (y_pred / y_pred.sum(-1, keepdim=True)).gather(2, lab.clamp(0, Cn - 1).unsqueeze(-1)).squeeze(-1)
    return torch.where(valid, -torch.log(p.clamp(1e-7, 1 - 1e-7)), torch.zeros_like(p))


def _confusion(y_true, y_pred):
    y_true, y_pred = _as_cuda(y_true), _as_cuda(y_pred)
    B, T, Cn = y_pred.shape
    am = torch.empty(B, T, device="cuda", dtype=torch.uint8)
    # argmax through the head kernel (scale-1 path): probs -> log -> softmax is monotone, argmax identical
    ops.resize_softmax_fwd(torch.log(y_pred.clamp_min(1e-30)).view(B, T, 1, Cn).contiguous(), Cn, T, 1, None, am)
    conf = torch.zeros(B, Cn + 1, Cn, device="cuda", dtype=torch.int64)
    ops.confusion(y_true.view(B, T, 1), am, Cn, conf)
    return conf.cpu().numpy(), Cn


def sparse_accuracy_ignoring_last_label(y_true, y_pred):
    """utils.py:132-138."""
    conf, Cn = _confusion(y_true, y_pred)
    return _metrics_from_confusion(conf, Cn)[1]


def Jaccard(y_true, y_pred):
    """utils.py:139-157."""
    conf, Cn = _confusion(y_true, y_pred)
    return _metrics_from_confusion(conf, Cn)[0]


# ---------------------------------------------------------------------------------------------------------
# dense CRF (utils.py:74-91): pydensecrf replaced by the permutohedral-lattice CUDA kernels
# ---------------------------------------------------------------------------------------------------------
def unary_from_labels(labels, n_labels, gt_prob, zero_unsure=True):
    """pydensecrf.utils.unary_from_labels (host-side table fill, replicated literally incl. the label-0 wrap)."""
    assert 0 < gt_prob < 1
    labels = np.asarray(labels).flatten()
    with np.errstate(divide="ignore"):
        n_energy = -np.log((1.0 - gt_prob) / (n_labels - 1)) if n_labels > 1 else np.inf
    p_energy = -np.log(gt_prob)
    U = np.full((n_labels, len(labels)), n_energy, dtype='float32')
    U[labels - 1 if zero_unsure else labels, np.arange(U.shape[1])] = p_energy
    if zero_unsure:
        U[:, labels == 0] = -np.log(1.0 / n_labels)
    return U


_crf_ws = {}


def dense_crf(unary, image, iters=5, sxy_gauss=3.0, compat_gauss=3.0, sxy_bilat=80.0, srgb_bilat=13.0,
              compat_bilat=10.0, return_map=False, probs=None):
    """Batched dense-CRF mean field on the GPU.
    unary [B, M, H*W] (or [M, H*W]) float32 energies, image [B, H, W, 3] (or [H, W, 3]) uint8 -> Q [B, M, H*W].
    Alternatively `probs` [B, H*W, M] (or [H*W, M]): class probabilities pixel-major, exactly what `model.predict` /
    the engine's softmax leaves in device memory -- soft unaries -log p without leaving the GPU."""
    if probs is not None:
        un = torch.as_tensor(probs, dtype=torch.float32)
        single = un.dim() == 2
    else:
        un = torch.as_tensor(unary, dtype=torch.float32)
        single = un.dim() == 2
    im = torch.as_tensor(image, dtype=torch.uint8)
    if single:
        un, im = un[None], im[None]
    un, im = un.cuda().contiguous(), im.cuda().contiguous()
    if probs is not None:
        B, N, M = un.shape
    else:
        B, M, N = un.shape
    H, W = im.shape[1], im.shape[2]
    assert N == H * W
    cfg = L.CrfConfig(H, W, M, iters, sxy_gauss, compat_gauss, sxy_bilat, srgb_bilat, compat_bilat,
                      1 if probs is not None else 0)
    # the whole batch goes through ONE set of launches (batch = grid.y); very large batches are chunked so that the
    # worst-case lattice workspace (~2.4 KB per pixel and image at 21 labels) stays below ~24 GB
    per_img = int(L.lib().dlb_crf_workspace_bytes_batched(C.byref(cfg), 1))
    chunk = max(1, min(B, int((24 << 30) // max(per_img, 1))))
    nbytes = int(L.lib().dlb_crf_workspace_bytes_batched(C.byref(cfg), chunk))
    key = (H, W, M)
    ws = _crf_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _crf_ws.pop(key, None)
        ws = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
        _crf_ws[key] = ws
    Q = torch.empty(B, M, N, device="cuda")
    mp = torch.empty(B, N, device="cuda", dtype=torch.uint8) if return_map else None
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        L.check(L.lib().dlb_crf_inference_batched(C.byref(cfg), nb, un[b0:b0 + nb].data_ptr(), im[b0:b0 + nb].data_ptr(),
                                                  Q[b0:b0 + nb].data_ptr(), mp[b0:b0 + nb].data_ptr() if return_map else None,
                                                  ws.data_ptr(), ws.numel(), L.stream_ptr()), "crf_inference")
    if return_map:
        return (Q[0], mp[0]) if single else (Q, mp)
    return Q[0] if single else Q


def do_crf(im, mask, zero_unsure=True):
    """utils.py:74-91, same arguments and return value (int label map [H, W])."""
    colors, labels = np.unique(mask, return_inverse=True)
    labels = labels.reshape(-1)
    image_size = mask.shape[:2]
    n_labels = len(set(labels.flat))
    U = unary_from_labels(labels, n_labels, gt_prob=.7, zero_unsure=zero_unsure)
    Q = dense_crf(U, np.ascontiguousarray(im).astype('uint8'), iters=5)
    MAP = Q.argmax(0).cpu().numpy().reshape(image_size)
    unique_map = np.unique(MAP)
    for u in unique_map:
        np.putmask(MAP, MAP == u, colors[u])
    return MAP


# ---------------------------------------------------------------------------------------------------------
# SegModel (utils.py:160-254)
# ---------------------------------------------------------------------------------------------------------
class SegModel:
    epochs = 20
    batch_size = 16

    def __init__(self, dataset='VOCdevkit/VOC2012', image_size=(320, 320), compute_dtype='float16'):
        self.sz = image_size
        self.mainpath = dataset
        self.crop = False
        self.compute_dtype = compute_dtype

    def create_seg_model(self, net, n=21, backbone='mobilenetv2', load_weights=False, multi_gpu=False, seed=0):
        """utils.py:169-214: Deeplabv3(classes=21, OS=16) cut at the Dropout output + a new head:
        net == 'original': Conv2D(n, 1, name='conv_upsample') + bilinear + softmax 'pred_mask'
        net == 'subpixel': Subpixel(n, 1, scale) (ICNR-initialised) + softmax 'pred_mask'."""
        from .deeplabv3p import _resolve_dtype
        if backbone not in ('mobilenetv2', 'xception'):
            raise ValueError('The `backbone` argument should be either `xception`  or `mobilenetv2` ')
        if net not in ('original', 'subpixel'):
            raise ValueError("net must be 'original' or 'subpixel'")
        if backbone == 'xception':
            raise NotImplementedError("training heads on the Xception backbone are not built (the reference's own "
                                      "Xception path raises NameError, deeplabv3p.py:147)")
        self.net = net
        self.modelpath = 'weights/{}_{}.h5'.format(backbone, net)
        scale = 8
        dt = _resolve_dtype(self.compute_dtype)
        engine = Engine(input_shape=tuple(self.sz) + (3,), classes=21, head=net, n_out=n, compute_dtype=dt, seed=seed,
                        head_layer_name="conv_upsample" if net == 'original' else None)
        names = keras_layer_names(backbone, net, engine.head_conv.name)
        model = Model(engine, 'deeplabv3p' if net == 'original' else 'deeplabv3p_subpixel', names)
        if net == 'subpixel':      # "Do ICNR" (utils.py:200-204)
            layer = model.get_layer(engine.head_conv.name)
            c, b = layer.get_weights()
            w = icnr_weights(scale=scale, shape=c.shape, seed=seed)
            layer.set_weights([w, b])
        if load_weights:
            model.load_weights('weights/{}_{}.h5'.format(backbone, net))
        if multi_gpu:
            from .parallel import make_data_parallel
            make_data_parallel(model)
        self.model = model
        return model

    def create_generators(self, *args, **kwargs):
        raise NotImplementedError("SegmentationGenerator (cv2 / VOC2012 data pipeline, utils.py:257-423) is out of "
                                  "scope of the hot path; feed any keras.utils.Sequence-like object yielding "
                                  "(X, Y, {'pred_mask': SW}) to train_generator")

    def load_weights(self, model):
        model.load_weights(self.modelpath)

    def train_generator(self, model, train_generator, valid_generator, callbacks, mp=True):
        steps = len(train_generator)
        h = model.fit_generator(train_generator, steps_per_epoch=steps, epochs=self.epochs, verbose=1,
                                callbacks=callbacks, validation_data=valid_generator,
                                validation_steps=len(valid_generator) if valid_generator is not None else None,
                                max_queue_size=10, workers=max(1, (os.cpu_count() or 2) // 2), use_multiprocessing=mp)
        return h

    def train(self, model, X, y, val_data, tf_board=False, plot_train_process=True, callbacks=None):
        # the reference calls an undefined self.build_callbacks here (utils.py:246); callbacks are passed in instead
        h = model.fit(X, y, validation_data=val_data, verbose=1, batch_size=self.batch_size, epochs=self.epochs,
                      callbacks=callbacks or [])
        return h

    @classmethod
    def set_num_epochs(cls, new_epochs):
        cls.epochs = new_epochs

    @classmethod
    def set_batch_size(cls, new_batch_size):
        cls.batch_size = new_batch_size

# ---------------------------------------------------------------------------------------------------------
# the data formats either side of the hot path
# ---------------------------------------------------------------------------------------------------------
def generator_labels_and_weights(labels, n_classes):
    """What `SegmentationGenerator.__getitem__` hands to `fit_generator` besides the image (reference utils.py:360-399),
    computed on the device: labels [B, H, W] or [B, P] (uint8 / int32 / float32, raw dataset values) ->
      Y  [B, P, 1] float32 with every value outside 0..n_classes-1 mapped to the void label n_classes (:360-365)
      SW [B, P]    float32 per-image balanced class weights, 0 on void pixels (:388-399; sklearn 'balanced')
    Returns CUDA tensors (`{'pred_mask': SW}` is the sample-weight dict the notebook passes)."""
    t = torch.as_tensor(labels)
    if t.dtype not in (torch.uint8, torch.int32, torch.float32):
        t = t.to(torch.int32)
    t = t.cuda().contiguous()
    B = t.shape[0]
    t = t.view(B, -1)
    y = torch.empty(B, t.shape[1], device="cuda", dtype=torch.float32)
    sw = torch.empty_like(y)
    ops.label_weights(t, int(n_classes), y, sw)
    return y.unsqueeze(-1), sw


class AugmentParams:
    """The random decisions SegmentationGenerator.__getitem__ takes for ONE image (reference utils.py:319-357)."""

    def __init__(self, blur_ksize=0, hflip=False, vflip=False, gamma=None, angle=0.0, scale=1.0, warp=False):
        self.blur_ksize, self.hflip, self.vflip = int(blur_ksize), bool(hflip), bool(vflip)
        self.gamma, self.angle, self.scale, self.warp = gamma, float(angle), float(scale), bool(warp)


def draw_augment_params(rng, blur=0, horizontal_flip=True, vertical_flip=0, brightness=0.1, rotation=5.0, zoom=0.1):
    """Draws from `rng` (Python's `random` module or a random.Random) in exactly the order of the reference's
    __getitem__ (utils.py:323-352), so the same seed gives the same augmentation stream."""
    p = AugmentParams()
    if blur and rng.randint(0, 1):
        p.blur_ksize = int(blur)
    if horizontal_flip and rng.randint(0, 1):
        p.hflip = True
    if vertical_flip and rng.randint(0, 1):
        p.vflip = True
    if brightness:
        factor = 1.0 + rng.gauss(mu=0.0, sigma=brightness)
        if rng.randint(0, 1):
            factor = 1.0 / factor
        p.gamma = factor
    p.angle = rng.gauss(mu=0.0, sigma=rotation) if rotation else 0.0
    p.scale = rng.gauss(mu=1.0, sigma=zoom) if zoom else 1.0
    p.warp = bool(rotation or zoom)
    return p


def rotation_matrix_2d(center, angle, scale):
    """cv2.getRotationMatrix2D restated (float64): [[a, b, (1-a)cx - b cy], [-b, a, b cx + (1-a)cy]]."""
    a_ = np.float64(angle) * (np.pi / 180.0)          # (OpenCV folds CV_PI / 180 first)
    alpha, beta = np.cos(a_) * scale, np.sin(a_) * scale
    cx, cy = np.float64(center[0]), np.float64(center[1])
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]], np.float64)


def invert_affine(M):
    """cv2.invertAffineTransform restated (float64)."""
    M = np.asarray(M, np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22, A12, A21 = M[1, 1] * D, M[0, 0] * D, -M[0, 1] * D, -M[1, 0] * D
    b1 = -A11 * M[0, 2] - A12 * M[1, 2]
    b2 = -A21 * M[0, 2] - A22 * M[1, 2]
    return np.array([[A11, A12, b1], [A21, A22, b2]], np.float64)


def gamma_lut(factor):
    """the brightness table of utils.py:344 (float64 power, truncated to uint8)."""
    return np.array([((i / 255.0) ** factor) * 255 for i in np.arange(0, 256)]).astype(np.uint8)


def augment_batch(images, labels, params, n_classes=21):
    """SegmentationGenerator.__getitem__ from the decoded (already resized) arrays on, on the device:
    images [B,H,W,3] uint8 BGR, labels [B,H,W] uint8, params: list of AugmentParams ->
    (X [B,H,W,3] float32, Y [B,H*W,1] float32, {'pred_mask': SW [B,H*W] float32}) CUDA tensors -- what fit_generator
    takes.  Bit-exact with the cv2 calls of the reference (tests/test_augment.py)."""
    img = torch.as_tensor(images, dtype=torch.uint8).cuda().contiguous()
    lab = torch.as_tensor(labels, dtype=torch.uint8).cuda().contiguous()
    B, H, W = lab.shape
    arr = (L.AugParams * B)()
    luts = np.tile(np.arange(256, dtype=np.uint8), (B, 1))
    any_lut = False
    for b, p in enumerate(params):
        arr[b].hflip, arr[b].vflip, arr[b].blur_ksize, arr[b].warp = int(p.hflip), int(p.vflip), int(p.blur_ksize), int(p.warp)
        if p.blur_ksize not in (0, 3, 5, 7):
            raise ValueError("blur kernel size must be 0, 3, 5 or 7")
        if p.warp:
            minv = invert_affine(rotation_matrix_2d((W // 2, H // 2), p.angle, p.scale))
            for i in range(6):
                arr[b].minv[i] = float(minv.flat[i])
        if p.gamma is not None:
            luts[b] = gamma_lut(p.gamma)
            any_lut = True
    prm = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    lut_t = torch.from_numpy(luts).cuda() if any_lut else None
    tmp_img, tmp_lab = torch.empty_like(img), torch.empty_like(lab)
    present = torch.empty(B, 8, device="cuda", dtype=torch.int32)
    X = torch.empty(B, H, W, 3, device="cuda", dtype=torch.float32)
    lab_out = torch.empty_like(lab)
    L.check(L.lib().dlb_augment_batch(B, H, W, int(n_classes), img.data_ptr(), lab.data_ptr(), prm.data_ptr(), L.ptr(lut_t),
                                      tmp_img.data_ptr(), tmp_lab.data_ptr(), present.data_ptr(), X.data_ptr(),
                                      lab_out.data_ptr(), L.stream_ptr()), "augment_batch")
    Y, SW = generator_labels_and_weights(lab_out, n_classes)
    return X, Y, {"pred_mask": SW}


def calculate_iou(model, nb_classes=21, data=None, batch_size=16):
    """Dataset-level confusion matrix of the notebook (segmentation.ipynb cell 10, `calculate_iou`): predict, argmax,
    count (label, prediction) pairs over all non-void pixels.  The reference walks 262 144 x N pixels in a Python loop;
    here the counts come from the device (dlb_confusion).  `data` = (X [N,H,W,3], label [N,P]) replaces the notebook's
    global validation generator.  Returns conf_m [nb_classes, nb_classes] float64 with the reference's indexing
    `conf_m[l-1, p-1] += 1` (a cyclic shift of both axes; IoU / mean IoU are invariant to it)."""
    if data is None:
        raise ValueError("calculate_iou needs data=(X, label): the notebook's global SegClass generator is not part of this package")
    X, label = data
    e = model.engine
    if nb_classes != e.n_out:
        raise ValueError(f"nb_classes={nb_classes} but the model predicts {e.n_out} classes")
    X = torch.as_tensor(X)
    label = torch.as_tensor(label).reshape(X.shape[0], -1)
    conf = torch.zeros(nb_classes + 1, nb_classes, device="cuda", dtype=torch.int64)
    for i in range(0, X.shape[0], batch_size):
        xb = X[i:i + batch_size].to(torch.float32)
        Bb = xb.shape[0]
        ws = e.workspace(Bb, False)
        ws["img"].copy_(xb, non_blocking=True)
        am = e.forward_infer(ws["img"], want_probs=False)          # argmax straight from the head kernel, on the device
        cb = torch.zeros(Bb, nb_classes + 1, nb_classes, device="cuda", dtype=torch.int64)
        ops.confusion(label[i:i + batch_size].float().cuda().contiguous().view(Bb, -1, 1), am, nb_classes, cb)
        conf += cb.sum(0)
    c = conf[:nb_classes].cpu().numpy().astype(np.float64)     # void row dropped (reference: `if l == nb_classes: continue`)
    return np.roll(np.roll(c, -1, axis=0), -1, axis=1)          # conf_m[l-1, p-1]
