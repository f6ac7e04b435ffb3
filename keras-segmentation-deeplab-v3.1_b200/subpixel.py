"""`Subpixel` layer and `icnr_weights` -- the reference's subpixel.py on sm_100a kernels.

  Subpixel(filters, kernel_size, r, padding='valid', ...)   subpixel.py:41-103
      = Conv2D(filters * r * r, kernel_size) followed by _phase_shift:
        out[n, a*r+j, b*r+i, k] = conv[n, a, b, k*r*r + i*r + j]          (subpixel.py:77-88)
    Inside a model (SegModel.create_seg_model(net='subpixel'), utils.py:194-198) the phase shift is fused into
    the 1x1 GEMM's epilogue (dlb_pw_gemm shuffle store).  As a standalone layer object it is callable on NHWC CUDA
    tensors and runs the same kernels.
  icnr_weights(init, scale, shape, dtype)                    subpixel.py:9-39
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops


def _glorot_normal(shape, rng):
    kh, kw, cin, cout = shape
    fan_in, fan_out = kh * kw * cin, kh * kw * cout
    std = math.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978     # TF truncated-normal correction
    a = rng.standard_normal(shape)
    # truncated normal at 2 sigma, by resampling (tf.glorot_normal_initializer)
    bad = np.abs(a) > 2
    while bad.any():
        a[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(a) > 2
    return (a * std).astype(np.float32)


def icnr_weights(init=None, scale=2, shape=[3, 3, 32, 4], dtype=np.float32, seed=None):
    """ICNR (subpixel.py:13-39): sample a [kh,kw,Cin,Cout/scale^2] sub-kernel, nearest-neighbour upsample it by
    `scale`, space_to_depth it back.  `init`: callable(shape, rng) -> ndarray (default glorot normal, as the
    reference's tf.glorot_normal_initializer()).  NB the resulting channel order (i*s+j)*C'+k does not match
    _phase_shift's k*r*r+i*r+j (reference quirk) -- replicated, not fixed."""
    shape = list(shape)
    rng = np.random.RandomState(seed)
    init = init or _glorot_normal
    if scale == 1:
        return init(shape, rng).astype(dtype)
    new_shape = shape[:3] + [shape[3] // (scale ** 2)]
    x = init(new_shape, rng)                                   # [kh, kw, Cin, C']
    x = np.transpose(x, (2, 0, 1, 3))                          # [Cin, kh, kw, C']
    x = np.repeat(np.repeat(x, scale, axis=1), scale, axis=2)  # resize_nearest_neighbor
    n, H, W, c = x.shape
    x = x.reshape(n, H // scale, scale, W // scale, scale, c).transpose(0, 1, 3, 2, 4, 5)
    x = x.reshape(n, H // scale, W // scale, scale * scale * c)   # space_to_depth
    x = np.transpose(x, (1, 2, 0, 3))
    return x.astype(dtype)


class Subpixel:
    def __init__(self, filters, kernel_size, r, padding='valid', data_format=None, strides=(1, 1), activation=None,
                 use_bias=True, kernel_initializer='glorot_uniform', bias_initializer='zeros',
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, name=None, **kwargs):
        ks = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        if tuple(ks) != (1, 1) or tuple(strides) != (1, 1):
            raise NotImplementedError("Subpixel: only the 1x1 / stride-1 convolution the reference uses is built")
        if activation is not None:
            raise NotImplementedError("Subpixel: activation is not used by the reference")
        self.filters = r * r * filters        # Conv2D sees the expanded count (subpixel.py:59-60)
        self.kernel_size, self.strides, self.padding, self.use_bias = tuple(ks), tuple(strides), padding, use_bias
        self.r = r
        self.name = name
        self.kernel = None
        self.bias = None
        self._seed = kwargs.get("seed", 0)

    # Keras-like weight API: [kernel HWIO (1,1,Cin,r*r*filters), bias]
    def build(self, cin, device="cuda"):
        rng = np.random.RandomState(self._seed)
        lim = math.sqrt(6.0 / (cin + self.filters))
        self.kernel = torch.from_numpy(rng.uniform(-lim, lim, (1, 1, cin, self.filters)).astype(np.float32)).to(device)
        self.bias = torch.zeros(self.filters, device=device)

    def get_weights(self):
        return [self.kernel.cpu().numpy(), self.bias.cpu().numpy()] if self.use_bias else [self.kernel.cpu().numpy()]

    def set_weights(self, ws):
        self.kernel = torch.as_tensor(ws[0], dtype=torch.float32).cuda()
        if self.use_bias:
            self.bias = torch.as_tensor(ws[1], dtype=torch.float32).cuda()

    def _phase_shift(self, I: torch.Tensor) -> torch.Tensor:
        """subpixel.py:77-88 on an NHWC CUDA tensor."""
        B, a, b, c = I.shape
        out = torch.empty(B, a * self.r, b * self.r, c // (self.r * self.r), device=I.device, dtype=I.dtype)
        return ops.phase_shift(I.contiguous(), out, self.r)

    def call(self, inputs: torch.Tensor) -> torch.Tensor:
        """conv (tcgen05 GEMM for 16-bit inputs, exact SIMT for fp32) + phase shift, NHWC CUDA tensor in/out."""
        B, h, w, cin = inputs.shape
        if self.kernel is None:
            self.build(cin, inputs.device)
        wt = self.kernel.view(cin, self.filters).t().contiguous().to(inputs.dtype)
        conv = torch.empty(B, h, w, self.filters, device=inputs.device, dtype=inputs.dtype)
        ops.pw_gemm(inputs.contiguous(), wt, conv, col_shift=self.bias if self.use_bias else None)
        return self._phase_shift(conv)

    __call__ = call

    def compute_output_shape(self, input_shape):
        n, h, w, _ = input_shape
        return (n, self.r * h, self.r * w, int(self.filters / (self.r * self.r)))

    def get_config(self):
        # reference quirk (subpixel.py:101): `filters / r*r` == filters, i.e. the *expanded* count is returned
        return dict(name=self.name, filters=int(self.filters / self.r * self.r), kernel_size=self.kernel_size,
                    strides=self.strides, padding=self.padding, use_bias=self.use_bias, r=self.r)
