"""Execution engine for the MobileNetV2 DeepLabV3+ graph (deeplabv3p.py:315-444, utils.py:169-198).

Host-side plumbing only: parameter storage (flat fp32 buffers), activation / gradient workspaces, the explicit
forward + backward schedule (no torch.autograd, no torch ops on the hot path), CUDA-graph capture of the whole
training step.  Every arithmetic step is one C-ABI call into libdeeplab_b200.so (ops.py).

Numerics (Keras 2.2.4 semantics):
  * training forward = BatchNorm with batch statistics for every BN layer, frozen or not; moving statistics and
    gamma/beta are only updated for trainable BN layers;
  * storage dtype of activations: fp16 / bf16 (tcgen05 GEMMs, fp32 accumulate) or fp32 (exact SIMT parity mode);
  * parameters, BN statistics, gradients of parameters and Adam state are fp32.

Data flow of one inverted-residual block in training (deeplabv3p.py:167-206), raw = pre-BN conv output:
    x_in --pw_gemm(+stats)--> y_e --dw_conv(prologue BN_e+ReLU6, +stats)--> y_d
         --pw_gemm(A-tile transform BN_d+ReLU6, +stats)--> y_p --bn_act_apply(+x_in)--> x_out
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from ._lib import ACT_NONE, ACT_RELU, ACT_RELU6
from ._lib import _DT as L_DT

# (expansion, stride, block_id, skip, rate, out_channels)   deeplabv3p.py:327-367 with alpha = 1
MNV2_BLOCKS = [
    (1, 1, 0, False, 1, 16), (6, 2, 1, False, 1, 24), (6, 1, 2, True, 1, 24), (6, 2, 3, False, 1, 32),
    (6, 1, 4, True, 1, 32), (6, 1, 5, True, 1, 32), (6, 1, 6, False, 1, 64), (6, 1, 7, True, 2, 64),
    (6, 1, 8, True, 2, 64), (6, 1, 9, True, 2, 64), (6, 1, 10, False, 2, 96), (6, 1, 11, True, 2, 96),
    (6, 1, 12, True, 2, 96), (6, 1, 13, False, 2, 160), (6, 1, 14, True, 4, 160), (6, 1, 15, True, 4, 160),
    (6, 1, 16, False, 4, 320),
]


def _make_divisible(v, divisor, min_value=None):
    """deeplabv3p.py:157-164."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


class Param:
    """One Keras weight tensor living in a flat buffer."""

    def __init__(self, name: str, shape: Tuple[int, ...], trainable: bool, init: str):
        self.name, self.shape, self.trainable_kind, self.init = name, tuple(shape), trainable, init
        self.size = int(np.prod(shape))
        self.offset = -1          # in the params buffer (trainable kinds) or the stats buffer (moving stats)
        self.data: torch.Tensor = None
        self.grad: Optional[torch.Tensor] = None


class LayerRec:
    """A weighted layer of the Keras graph (Conv2D / DepthwiseConv2D / BatchNormalization)."""

    def __init__(self, name: str, kind: str, params: List[Param], **attrs):
        self.name, self.kind, self.params = name, kind, params
        self.trainable = True
        self.__dict__.update(attrs)


class BN:
    def __init__(self, layer: LayerRec, C: int, eps: float, momentum: float):
        self.layer, self.C, self.eps, self.momentum = layer, C, eps, momentum
        self.gamma, self.beta, self.mm, self.mv = layer.params
        # views assigned by Engine._alloc_bn_work
        self.sum = self.sqs = self.scale = self.shift = self.mean = self.rstd = self.red = None


class Engine:
    """MobileNetV2 DeepLabV3+ with the 'bare' (Deeplabv3), 'original' or 'subpixel' head."""

    def __init__(self, input_shape=(512, 512, 3), classes=21, head="bare", n_out: Optional[int] = None, alpha=1.0,
                 compute_dtype=torch.float16, device="cuda", seed: int = 0, head_layer_name: Optional[str] = None):
        self.alpha = float(alpha)        # MobileNetV2 width multiplier (deeplabv3p.py:157-164, :168-170, :317)
        self.H, self.W = int(input_shape[0]), int(input_shape[1])
        if self.H % 8 or self.W % 8:
            raise ValueError("input height/width must be multiples of 8 (output stride 8, deeplabv3p.py:316)")
        self.classes = classes
        self.head = head                       # 'bare' | 'original' | 'subpixel'
        self.n_out = n_out if n_out is not None else classes
        self.dtype = compute_dtype
        self.device = torch.device(device)
        self.scale = 8
        self.fh, self.fw = self.H // 8, self.W // 8
        self.ldl = (self.n_out + 7) // 8 * 8     # channel pitch of the low-resolution logits buffers
        self.layers: List[LayerRec] = []
        self._by_name: Dict[str, LayerRec] = {}
        self._build_spec(head_layer_name)
        self._alloc_params(seed)
        self._ws: Dict[Tuple, dict] = {}
        self._graphs: Dict[Tuple, object] = {}
        # fp16: dynamic loss scaling on the device (dlb_adam_step): {scale, good steps, found_inf, growth interval}.
        # Starts at 1024 (never overflowed on this network), halves on an inf/NaN gradient (that update is skipped),
        # doubles every 2000 clean steps.  bf16 / fp32 do not scale.
        self.ls_state = (torch.tensor([1024.0, 0.0, 0.0, 2000.0], device=self.device)
                         if compute_dtype == torch.float16 else None)
        self.dropout_rate = 0.1
        self.dropout_seed = 0x5EED + seed
        self.adam_cfg = dict(lr=7e-4, beta1=0.9, beta2=0.999, eps=1e-8, decay=1e-6)
        self.world_size = 1
        self.grad_hook = None                  # called with the flat grad buffer before the optimizer step
        self._weights_dirty = True
        self._any_frozen = False
        # depthwise_BN + relu6 applied to the A tiles inside the project GEMM / its weight gradient (no a_d tensor).
        # DLB_FUSE_DW_BN=0 restores the materialised activation (A/B measurements only).
        import os
        self.fuse_dw_bn = os.environ.get("DLB_FUSE_DW_BN", "1") != "0"
        # inference: depthwise -> project of an inverted-residual block in one kernel (DLB_FUSE_MBCONV=0: separate)
        self.fuse_dw_project = os.environ.get("DLB_FUSE_MBCONV", "1") != "0"
        # BatchNorm statistics are finalised in the prologue of the kernel that consumes them (dlb_bn_fin) instead of by
        # 54 stand-alone launches per step; DLB_FUSE_BN_FIN=0 restores the stand-alone kernel
        self.fuse_bn_fin = os.environ.get("DLB_FUSE_BN_FIN", "1") != "0"
        # data parallel: gradient buckets are all-reduced while the rest of the backward pass runs, inside the captured
        # step (DLB_AR_IN_GRAPH=0: one all-reduce of the whole buffer between two graphs, the round-1 schedule)
        self.ar_in_graph = os.environ.get("DLB_AR_IN_GRAPH", "1") != "0"
        self._on_bucket = None
        self._mb_packs = {}

    @property
    def loss_scale(self) -> float:
        """the factor currently folded into d loss / d logits (reads the device state: a host sync)."""
        return float(self.ls_state[0].item()) if self.ls_state is not None else 1.0

    @loss_scale.setter
    def loss_scale(self, v: float):
        if self.ls_state is None:
            if float(v) != 1.0:
                raise ValueError("loss scaling is an fp16 feature")
            return
        self.ls_state[0] = float(v)

    # ------------------------------------------------------------------------------------------------
    # graph spec (weighted layers in Keras model order)
    # ------------------------------------------------------------------------------------------------
    def _add(self, rec: LayerRec):
        self.layers.append(rec)
        self._by_name[rec.name] = rec
        return rec

    def _conv(self, name, kh, cin, cout, bias=False):
        ps = [Param(name + "/kernel:0", (kh, kh, cin, cout), True, "glorot")]
        if bias:
            ps.append(Param(name + "/bias:0", (cout,), True, "zeros"))
        return self._add(LayerRec(name, "conv", ps, cin=cin, cout=cout, kh=kh, use_bias=bias))

    def _dw(self, name, c):
        return self._add(LayerRec(name, "dw", [Param(name + "/depthwise_kernel:0", (3, 3, c, 1), True, "glorot_dw")], C=c))

    def _bn(self, name, c, eps, momentum):
        ps = [Param(name + "/gamma:0", (c,), True, "ones"), Param(name + "/beta:0", (c,), True, "zeros"),
              Param(name + "/moving_mean:0", (c,), False, "zeros"), Param(name + "/moving_variance:0", (c,), False, "ones")]
        rec = self._add(LayerRec(name, "bn", ps, C=c))
        return BN(rec, c, eps, momentum)

    def _build_spec(self, head_layer_name):
        # deeplabv3p.py:317: first_block_filters = _make_divisible(32 * alpha, 8); :168-170: every block's output is
        # _make_divisible(int(filters * alpha), 8)
        c0 = _make_divisible(32 * self.alpha, 8)
        if c0 != 32:
            raise NotImplementedError(f"alpha={self.alpha}: the stem kernels are built for 32 output channels "
                                      f"(_make_divisible(32 * alpha, 8) = {c0}); 0.9 <= alpha < 1.125 keeps it at 32")
        self.stem = self._conv("Conv", 3, 3, c0)
        self.stem_bn = self._bn("Conv_BN", c0, 1e-3, 0.999)
        self.blocks = []
        cin = c0
        for (t, s, bid, skip, rate, filters) in MNV2_BLOCKS:
            cout = _make_divisible(int(filters * self.alpha), 8)
            prefix = "expanded_conv_{}_".format(bid) if bid else "expanded_conv_"
            mid = cin * t
            blk = dict(bid=bid, stride=s, skip=skip, rate=rate, cin=cin, mid=mid, cout=cout, prefix=prefix)
            if bid:
                blk["expand"] = self._conv(prefix + "expand", 1, cin, mid)
                blk["expand_bn"] = self._bn(prefix + "expand_BN", mid, 1e-3, 0.999)
            blk["dw"] = self._dw(prefix + "depthwise", mid)
            blk["dw_bn"] = self._bn(prefix + "depthwise_BN", mid, 1e-3, 0.999)
            blk["project"] = self._conv(prefix + "project", 1, mid, cout)
            blk["project_bn"] = self._bn(prefix + "project_BN", cout, 1e-3, 0.999)
            self.blocks.append(blk)
            cin = cout
        self.c_last = cin
        self.image_pooling = self._conv("image_pooling", 1, cin, 256)
        self.image_pooling_bn = self._bn("image_pooling_BN", 256, 1e-5, 0.99)
        self.aspp0 = self._conv("aspp0", 1, cin, 256)
        self.aspp0_bn = self._bn("aspp0_BN", 256, 1e-5, 0.99)
        self.concat_projection = self._conv("concat_projection", 1, 512, 256)
        self.concat_projection_bn = self._bn("concat_projection_BN", 256, 1e-5, 0.99)
        if self.head == "subpixel":
            name = head_layer_name or "subpixel_1"
            self.head_conv = self._conv(name, 1, 256, self.n_out * self.scale * self.scale, bias=True)
        else:
            if head_layer_name is None:
                head_layer_name = ("logits_semantic" if self.classes == 21 else "custom_logits_semantic") \
                    if self.head == "bare" else "conv_upsample"
            self.head_conv = self._conv(head_layer_name, 1, 256, self.n_out, bias=True)
        self.bns: List[BN] = [self.stem_bn]
        for b in self.blocks:
            if b["bid"]:
                self.bns.append(b["expand_bn"])
            self.bns += [b["dw_bn"], b["project_bn"]]
        self.bns += [self.image_pooling_bn, self.aspp0_bn, self.concat_projection_bn]

    # ------------------------------------------------------------------------------------------------
    # parameter storage
    # ------------------------------------------------------------------------------------------------
    def _alloc_params(self, seed):
        dev = self.device
        n_train = n_stat = 0
        for rec in self.layers:
            for p in rec.params:
                if p.trainable_kind:
                    p.offset = n_train
                    n_train += (p.size + 3) // 4 * 4
                else:
                    p.offset = n_stat
                    n_stat += (p.size + 3) // 4 * 4
        self.n_params = n_train
        self.params = torch.zeros(n_train, device=dev)
        self.grads = torch.zeros(n_train, device=dev)
        self.adam_m = torch.zeros(n_train, device=dev)
        self.adam_v = torch.zeros(n_train, device=dev)
        self.adam_step = torch.zeros(1, device=dev, dtype=torch.int64)
        self.stats = torch.zeros(n_stat, device=dev)
        self.train_mask = torch.ones(n_train, device=dev)     # 0 for frozen (trainable=False) parameters
        for rec in self.layers:
            for p in rec.params:
                buf = self.params if p.trainable_kind else self.stats
                p.data = buf[p.offset:p.offset + p.size].view(p.shape)
                if p.trainable_kind:
                    p.grad = self.grads[p.offset:p.offset + p.size].view(p.shape)
        # BN work buffers (flat, so one memset / one launch can cover all layers)
        ctot = sum(b.C for b in self.bns)
        self.bn_sums = torch.zeros(2 * ctot, device=dev, dtype=torch.float64)
        self.bn_red = torch.zeros(2 * ctot, device=dev, dtype=torch.float64)
        self.bn_work = torch.zeros(4 * ctot, device=dev)
        o = 0
        for b in self.bns:
            C = b.C
            b.sum, b.sqs = self.bn_sums[2 * o:2 * o + C], self.bn_sums[2 * o + C:2 * o + 2 * C]
            b.red = self.bn_red[2 * o:2 * o + 2 * C]
            b.scale, b.shift = self.bn_work[4 * o:4 * o + C], self.bn_work[4 * o + C:4 * o + 2 * C]
            b.mean, b.rstd = self.bn_work[4 * o + 2 * C:4 * o + 3 * C], self.bn_work[4 * o + 3 * C:4 * o + 4 * C]
            o += C
        self._init_weights(seed)
        # low-precision / transposed weight copies for the GEMMs
        self.wcopies: Dict[str, Dict[str, torch.Tensor]] = {}
        for rec in self.layers:
            if rec.kind == "conv" and rec.kh == 1:
                K, N = rec.cin, rec.cout
                d = dict(kn=torch.empty(K, N, device=dev, dtype=self.dtype), nk=torch.empty(N, K, device=dev, dtype=self.dtype))
                if rec is self.concat_projection or rec is self.image_pooling:
                    d["nk32"] = torch.empty(N, K, device=dev, dtype=torch.float32)
                if rec is self.head_conv and N % 8:
                    # dgrad reads W[K, N] as a [K, ld] matrix through TMA: pad the pitch to 32 (16-byte rule)
                    d["kn"] = torch.zeros(K, self.ldl, device=dev, dtype=self.dtype)
                self.wcopies[rec.name] = d
        # descriptor table of the batched master -> copy cast (one launch after every optimizer step)
        rows, start = [], 0
        for rec in self.layers:
            if rec.name in self.wcopies:
                d, K, N = self.wcopies[rec.name], rec.cin, rec.cout
                wp = rec.params[0].data.data_ptr()
                rows.append([wp, d["kn"].data_ptr(), d["nk"].data_ptr(), K, N, d["kn"].shape[1], L_DT[self.dtype], start])
                start += K * N
                if "nk32" in d:
                    rows.append([wp, 0, d["nk32"].data_ptr(), K, N, N, L_DT[torch.float32], start])
                    start += K * N
        self._cast_table = torch.tensor(rows, dtype=torch.int64, device=dev)
        self._cast_total = start
        # Subpixel: GEMM columns are stored permuted (jj, i, k) so the fused phase-shift store is contiguous
        self.sub_perm = self.sub_perm_inv = None
        if self.head == "subpixel":
            r, cs = self.scale, self.n_out
            perm = np.empty(cs * r * r, dtype=np.int64)
            for jj in range(r):
                for i in range(r):
                    for k in range(cs):
                        perm[(jj * r + i) * cs + k] = k * r * r + i * r + jj
            self.sub_perm = torch.from_numpy(perm)
            self.sub_perm_inv = torch.empty_like(self.sub_perm)
            self.sub_perm_inv[self.sub_perm] = torch.arange(perm.size)

    def _init_weights(self, seed):
        """Keras defaults: glorot_uniform kernels, zero biases, BN gamma 1 / beta 0 / mean 0 / var 1 (host RNG)."""
        rng = np.random.RandomState(seed)
        for rec in self.layers:
            for p in rec.params:
                if p.init == "glorot":
                    kh, _, cin, cout = p.shape
                    lim = math.sqrt(6.0 / (kh * kh * cin + kh * kh * cout))
                    a = rng.uniform(-lim, lim, p.shape).astype(np.float32)
                elif p.init == "glorot_dw":
                    c = p.shape[2]
                    lim = math.sqrt(6.0 / (9 * c + 9))
                    a = rng.uniform(-lim, lim, p.shape).astype(np.float32)
                elif p.init == "ones":
                    a = np.ones(p.shape, np.float32)
                else:
                    a = np.zeros(p.shape, np.float32)
                p.data.copy_(torch.from_numpy(a))

    # Keras-facing weight access (internal Subpixel column permutation hidden here)
    def get_layer_weights(self, rec: LayerRec) -> List[np.ndarray]:
        out = []
        for p in rec.params:
            a = p.data.detach().cpu()
            if rec is self.head_conv and self.sub_perm is not None:
                a = a.index_select(a.dim() - 1, self.sub_perm_inv)
            out.append(a.numpy().copy())
        return out

    def set_layer_weights(self, rec: LayerRec, arrays: List[np.ndarray]):
        if len(arrays) != len(rec.params):
            raise ValueError(f"layer {rec.name}: expected {len(rec.params)} weight arrays, got {len(arrays)}")
        for p, a in zip(rec.params, arrays):
            a = np.asarray(a, dtype=np.float32)
            if tuple(a.shape) != p.shape:
                raise ValueError(f"layer {rec.name}: weight {p.name} expects shape {p.shape}, got {tuple(a.shape)}")
            t = torch.from_numpy(a)
            if rec is self.head_conv and self.sub_perm is not None:
                t = t.index_select(t.dim() - 1, self.sub_perm)
            p.data.copy_(t)
        self._weights_dirty = True

    def set_trainable(self, rec: LayerRec, flag: bool):
        rec.trainable = bool(flag)
        for p in rec.params:
            if p.trainable_kind:
                self.train_mask[p.offset:p.offset + p.size] = 1.0 if flag else 0.0
        self._any_frozen = any(not r.trainable for r in self.layers)
        self._graphs.clear()

    def refresh_weight_copies(self):
        """fp32 master -> 16-bit [K,N] / [N,K] GEMM operands (after load / set_weights / optimizer step)."""
        ops.cast_weights_batched(self._cast_table, self._cast_total)
        self._weights_dirty = False

    # ------------------------------------------------------------------------------------------------
    # workspaces
    # ------------------------------------------------------------------------------------------------
    def _geometry(self):
        """spatial size at the input of each block + TF-SAME padding of its depthwise conv."""
        H, W = self.H, self.W
        h, pt, _ = ops.tf_same_pad(H, 3, 2, 1)
        w, pl, _ = ops.tf_same_pad(W, 3, 2, 1)
        geo = []
        for b in self.blocks:
            ho, bpt, _ = ops.tf_same_pad(h, 3, b["stride"], b["rate"])
            wo, bpl, _ = ops.tf_same_pad(w, 3, b["stride"], b["rate"])
            geo.append(dict(h=h, w=w, ho=ho, wo=wo, pt=bpt, pl=bpl))
            h, w = ho, wo
        return geo

    def workspace(self, B: int, training: bool):
        key = (B, training)
        if key in self._ws:
            return self._ws[key]
        dev, dt = self.device, self.dtype
        geo = self._geometry()
        ws = dict(geo=geo)
        E = lambda *shape, dtype=dt: torch.empty(*shape, device=dev, dtype=dtype)
        h0, w0 = geo[0]["h"], geo[0]["w"]
        ws["img"] = E(B, self.H, self.W, 3, dtype=torch.float32)
        fh, fw = self.fh, self.fw
        if training:
            ws["y_stem"] = E(B, h0, w0, 32)
            ws["x0"] = E(B, h0, w0, 32)
            for i, (b, g) in enumerate(zip(self.blocks, geo)):
                if b["bid"]:
                    ws[f"y_e{i}"] = E(B, g["h"], g["w"], b["mid"])
                ws[f"y_d{i}"] = E(B, g["ho"], g["wo"], b["mid"])
                if not self.fuse_dw_bn:
                    ws[f"a_d{i}"] = E(B, g["ho"], g["wo"], b["mid"])
                ws[f"y_p{i}"] = E(B, g["ho"], g["wo"], b["cout"])
                ws[f"x{i + 1}"] = E(B, g["ho"], g["wo"], b["cout"])
            ws["y_a0"] = E(B, fh, fw, 256)
            ws["a_a0"] = E(B, fh, fw, 256)
            ws["y_cp"] = E(B, fh, fw, 256)
            ws["feat"] = E(B, fh, fw, 256)
            ws["y_ip"] = E(B, 256, dtype=torch.float32)
            ws["dy_ip"] = E(B, 256, dtype=torch.float32)
            ws["d_b4"] = E(B, 256, dtype=torch.float32)
            ws["d_rowbias"] = E(B, 256, dtype=torch.float32)
            ws["d_pooled"] = E(B, self.c_last, dtype=torch.float32)
            # gradient scratch
            wide = max(B * g["h"] * g["w"] * b["mid"] for b, g in zip(self.blocks, geo))
            narrow = max(B * h0 * w0 * 32, max(B * g["ho"] * g["wo"] * b["cout"] for b, g in zip(self.blocks, geo)),
                         B * fh * fw * self.c_last)
            ws["g_wide"] = [E(wide), E(wide)]
            ws["g_narrow"] = [E(narrow), E(narrow), E(narrow)]
            ws["g256"] = [E(B, fh, fw, 256), E(B, fh, fw, 256)]
            ws["labels"] = E(B, self.H * self.W, 1, dtype=torch.float32)
            ws["sample_w"] = E(B, self.H * self.W, dtype=torch.float32)
            ws["grad_scale"] = torch.zeros(1, device=dev)
            ws["wcount"] = torch.zeros(1, device=dev, dtype=torch.float64)
            ws["loss_sum"] = torch.zeros(1, device=dev, dtype=torch.float64)
            ws["argmax"] = torch.empty(B, self.H * self.W, device=dev, dtype=torch.uint8)
            if self.head == "subpixel":
                ws["logits"] = E(B, self.H, self.W, self.n_out, dtype=torch.float32)
                ws["dlogits"] = E(B, self.H, self.W, self.n_out, dtype=torch.float32)
                ws["dlogits_lo"] = E(B, fh, fw, self.n_out * 64)
            else:
                ws["logits"] = torch.zeros(B, fh, fw, self.ldl, device=dev)
                ws["dlogits"] = torch.zeros(B, fh, fw, self.ldl, device=dev)
                ws["dlogits_lo"] = E(B, fh, fw, self.ldl)
        else:
            act = max(B * g["h"] * g["w"] * max(b["mid"], b["cin"]) for b, g in zip(self.blocks, geo))
            ws["t"] = [E(act), E(act), E(act), E(act)]
            ws["a_a0"] = E(B, fh, fw, 256)
            ws["feat"] = E(B, fh, fw, 256)
            if self.head == "subpixel":
                ws["logits"] = E(B, self.H, self.W, self.n_out, dtype=torch.float32)
            else:
                ws["logits"] = torch.zeros(B, fh, fw, self.ldl, device=dev)
            ws["probs"] = E(B, self.H * self.W, self.n_out, dtype=torch.float32)
            ws["argmax"] = torch.empty(B, self.H * self.W, device=dev, dtype=torch.uint8)
            ctot = sum(b.C for b in self.bns)
            ws["fold"] = torch.empty(2 * ctot, device=dev)
        ws["pooled"] = E(B, self.c_last, dtype=torch.float32)
        ws["b4"] = E(B, 256, dtype=torch.float32)
        ws["rowbias"] = E(B, 256, dtype=torch.float32)
        ws["y_ip"] = E(B, 256, dtype=torch.float32)
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------------
    # inference forward (BatchNorm folded into the producing kernel's epilogue)
    # ------------------------------------------------------------------------------------------------
    def _wi(self, name, cols=None):
        """inference-time [N, K] weight operand of a 1x1 conv: the 16-bit copy, or in fp32 mode the pre-split (hi, lo)
        pair of the 3xTF32 GEMM (weights are constant between optimizer steps; re-split in _fold_all)."""
        if self.dtype == torch.float32:
            hi, lo = self._wsplit[name]
            return (hi, lo) if cols is None else (hi[:, cols], lo[:, cols])
        w = self.wcopies[name]["nk"]
        return w if cols is None else w[:, cols]

    def _fold_all(self, ws):
        self._mb_packs = {}           # packed depthwise taps + folded BN of the fused depthwise -> project kernel
        if self.dtype == torch.float32:
            self._wsplit = {name: ops.f32_split(d["nk"]) for name, d in self.wcopies.items()}
        o = 0
        for b in self.bns:
            sc, sh = ws["fold"][2 * o:2 * o + b.C], ws["fold"][2 * o + b.C:2 * o + 2 * b.C]
            ops.bn_fold(b.gamma.data, b.beta.data, b.mm.data, b.mv.data, b.eps, sc, sh)
            b.fscale, b.fshift = sc, sh
            o += b.C

    def forward_infer(self, img: torch.Tensor, want_probs=True, want_argmax=False):
        """img: [B,H,W,3] fp32 (0..255) on the device -> probs [B, H*W, n_out] fp32 (and/or argmax uint8)."""
        B = img.shape[0]
        ws = self.workspace(B, False)
        if self._weights_dirty:
            self.refresh_weight_copies()
            self._fold_dirty = True
        if getattr(self, "_fold_dirty", True) or getattr(self, "_fold_ws", None) is not ws:
            self._fold_all(ws)
            self._fold_dirty, self._fold_ws = False, ws
        geo = ws["geo"]
        T = ws["t"]
        dt = self.dtype

        def view(buf, *shape):
            n = int(np.prod(shape))
            return buf[:n].view(*shape)

        h0, w0 = geo[0]["h"], geo[0]["w"]
        x = view(T[0], B, h0, w0, 32)
        ops.stem_conv_fwd(img, self.stem.params[0].data, x, out_scale=self.stem_bn.fscale,
                          out_shift=self.stem_bn.fshift, out_act=ACT_RELU6)
        cur = 0
        for b, g in zip(self.blocks, geo):
            free = [i for i in range(4) if i != cur]
            xin = x
            if b["bid"]:
                a_e = view(T[free[0]], B, g["h"], g["w"], b["mid"])
                bn = b["expand_bn"]
                ops.pw_gemm(xin, self._wi(b["expand"].name), a_e, col_scale=bn.fscale, col_shift=bn.fshift,
                            act=ACT_RELU6)
            else:
                a_e = xin
            xo = view(T[free[2]], B, g["ho"], g["wo"], b["cout"])
            dbn, pbn = b["dw_bn"], b["project_bn"]
            if self.fuse_dw_project and self.dtype != torch.float32 and b["stride"] == 1 and g["w"] <= 128 \
                    and b["cout"] <= 256 and b["mid"] >= 16:
                # depthwise + BN + relu6 -> project + BN (+ add) in one kernel: the expanded activation is read once,
                # its depthwise result is produced in shared memory as the tensor-core A operand (sepconv_fused.cu)
                pk = self._mb_packs.get(b["bid"])
                if pk is None:
                    pk = self._mb_packs[b["bid"]] = ops.sepconv_pack_dw([b["dw"].params[0].data], [dbn.fscale],
                                                                        [dbn.fshift], self.dtype)
                ops.sepconv_fused_fwd(a_e, [b["rate"]], [self.wcopies[b["project"].name]["nk"]], pk, [pbn.fscale],
                                      [pbn.fshift], [xo], dw_act=ACT_RELU6, pw_act=ACT_NONE,
                                      residuals=[xin if b["skip"] else None])
            else:
                a_d = view(T[free[1]], B, g["ho"], g["wo"], b["mid"])
                ops.dw_conv_fwd(a_e, b["dw"].params[0].data, a_d, stride=b["stride"], dilation=b["rate"], pad_top=g["pt"],
                                pad_left=g["pl"], out_scale=dbn.fscale, out_shift=dbn.fshift, out_act=ACT_RELU6)
                ops.pw_gemm(a_d, self._wi(b["project"].name), xo, col_scale=pbn.fscale, col_shift=pbn.fshift,
                            residual=xin if b["skip"] else None)
            x, cur = xo, free[2]
        self._aspp_head_infer(ws, x, B)
        C_ = self.n_out
        if self.head == "subpixel":
            lg = ws["logits"]
            ops.resize_softmax_fwd(lg, C_, self.H, self.W, ws["probs"] if want_probs else None,
                                   ws["argmax"] if want_argmax or not want_probs else None)
        else:
            ops.resize_softmax_fwd(ws["logits"], C_, self.H, self.W, ws["probs"] if want_probs else None,
                                   ws["argmax"] if want_argmax or not want_probs else None)
        return ws["probs"] if want_probs else ws["argmax"]

    def eval_batch(self, img: torch.Tensor, labels: torch.Tensor, sample_w: Optional[torch.Tensor] = None):
        """Inference-mode forward (moving BatchNorm statistics, no dropout) + the training step's fused loss kernel:
        -> device scalars (loss_sum, wcount) and the uint8 argmax [B, H*W] (Keras `test_on_batch` / validation)."""
        B = img.shape[0]
        self.forward_infer(img, want_probs=False)
        ws = self.workspace(B, False)
        if "ev_labels" not in ws:
            dev = self.device
            ws["ev_labels"] = torch.empty(B, self.H * self.W, 1, device=dev)
            ws["ev_sw"] = torch.empty(B, self.H * self.W, device=dev)
            ws["ev_dlogits"] = torch.zeros_like(ws["logits"])
            ws["ev_scale"] = torch.zeros(1, device=dev)
            ws["ev_wcount"] = torch.zeros(1, device=dev, dtype=torch.float64)
            ws["ev_loss"] = torch.zeros(1, device=dev, dtype=torch.float64)
        ws["ev_labels"].copy_(labels.view(B, -1, 1), non_blocking=True)
        sw = None
        if sample_w is not None:
            ws["ev_sw"].copy_(sample_w.view(B, -1), non_blocking=True)
            sw = ws["ev_sw"]
        ops.ce_grad_scale(B * self.H * self.W, sw, ws["ev_scale"], ws["ev_wcount"])
        ops.fill_zero(ws["ev_loss"])
        ops.resize_softmax_ce(ws["logits"], self.n_out, self.H, self.W, ws["ev_labels"], sw, ws["ev_scale"],
                              ws["ev_dlogits"], ws["ev_loss"], ws["ev_wcount"], ws["argmax"])
        return ws["ev_loss"], ws["ev_wcount"], ws["argmax"]

    def _aspp_head_infer(self, ws, x16, B):
        fh, fw = self.fh, self.fw
        # image pooling branch (deeplabv3p.py:375-382); the 1x1 -> HxW bilinear resize is a broadcast
        ops.global_avgpool_fwd(x16, ws["pooled"])
        bn = self.image_pooling_bn
        ops.pw_gemm(ws["pooled"], self.wcopies["image_pooling"]["nk32"], ws["b4"], col_scale=bn.fscale,
                    col_shift=bn.fshift, act=ACT_RELU)
        # concat([b4, b0]) @ W_cp = b4 @ W_cp[:256] (per-image bias) + b0 @ W_cp[256:]
        bn = self.concat_projection_bn
        wcp = self.wcopies["concat_projection"]
        ops.pw_gemm(ws["b4"], wcp["nk32"], ws["rowbias"], K=256, col_scale=bn.fscale)
        bn0 = self.aspp0_bn
        ops.pw_gemm(x16, self._wi("aspp0"), ws["a_a0"], col_scale=bn0.fscale, col_shift=bn0.fshift,
                    act=ACT_RELU)
        ops.pw_gemm(ws["a_a0"], self._wi("concat_projection", slice(256, None)), ws["feat"], col_scale=bn.fscale,
                    col_shift=bn.fshift, row_bias=ws["rowbias"], rows_per_img=fh * fw, act=ACT_RELU)
        self._head_fwd(ws, ws["feat"], infer=True)

    def _head_fwd(self, ws, feat, infer=False):
        w = self._wi(self.head_conv.name) if infer else self.wcopies[self.head_conv.name]["nk"]
        bias = self.head_conv.params[1].data
        if self.head == "subpixel":
            ops.pw_gemm(feat, w, ws["logits"], col_shift=bias, shuffle=(self.scale, self.fh, self.fw))
        else:
            ops.pw_gemm(feat, w, ws["logits"], col_shift=bias, n_store=self.ldl)

    # ------------------------------------------------------------------------------------------------
    # training forward / backward
    # ------------------------------------------------------------------------------------------------
    def _finalize(self, bn: BN, count):
        upd = bn.layer.trainable
        ops.bn_finalize(count, bn.sum, bn.sqs, bn.gamma.data, bn.beta.data, bn.eps, bn.momentum,
                        bn.mm.data if upd else None, bn.mv.data if upd else None, bn.scale, bn.shift, bn.mean, bn.rstd,
                        reset=False)

    def _fin(self, bn: BN, count):
        """scale / shift of `bn` for its consumer: a dlb_bn_fin (the consumer finalises in its prologue), or the finished
        tables after a stand-alone finalize launch (DLB_FUSE_BN_FIN=0).  Returns (fin, scale, shift)."""
        if self.fuse_bn_fin:
            upd = bn.layer.trainable
            return ops.bn_fin(count, bn.sum, bn.sqs, bn.gamma.data, bn.beta.data, bn.eps, bn.momentum,
                              bn.mm.data if upd else None, bn.mv.data if upd else None, bn.scale, bn.shift, bn.mean,
                              bn.rstd), None, None
        self._finalize(bn, count)
        return None, bn.scale, bn.shift

    def forward_train(self, ws, B, dropout: bool):
        geo = ws["geo"]
        img = ws["img"]
        h0, w0 = geo[0]["h"], geo[0]["w"]
        bn = self.stem_bn
        ops.fill_zero(self.bn_sums)      # every BatchNorm's fp64 accumulators, once per step (consumers do not clear them)
        ops.stem_conv_fwd(img, self.stem.params[0].data, ws["y_stem"], stat_sum=bn.sum, stat_sqs=bn.sqs)
        fin, sc, sh = self._fin(bn, B * h0 * w0)
        ops.bn_act_apply(ws["y_stem"], ws["x0"], scale=sc, shift=sh, act=ACT_RELU6, fin=fin)
        for i, (b, g) in enumerate(zip(self.blocks, geo)):
            xin = ws[f"x{i}"]
            Min, Mout = B * g["h"] * g["w"], B * g["ho"] * g["wo"]
            dbn = b["dw_bn"]
            if b["bid"]:
                ebn = b["expand_bn"]
                ops.pw_gemm(xin, self.wcopies[b["expand"].name]["nk"], ws[f"y_e{i}"], stat_sum=ebn.sum, stat_sqs=ebn.sqs)
                fin, sc, sh = self._fin(ebn, Min)
                ops.dw_conv_fwd(ws[f"y_e{i}"], b["dw"].params[0].data, ws[f"y_d{i}"], stride=b["stride"],
                                dilation=b["rate"], pad_top=g["pt"], pad_left=g["pl"], in_scale=sc, in_shift=sh,
                                in_fin=fin, in_act=ACT_RELU6, stat_sum=dbn.sum, stat_sqs=dbn.sqs)
            else:
                ops.dw_conv_fwd(xin, b["dw"].params[0].data, ws[f"y_d{i}"], stride=b["stride"], dilation=b["rate"],
                                pad_top=g["pt"], pad_left=g["pl"], stat_sum=dbn.sum, stat_sqs=dbn.sqs)
            fin, sc, sh = self._fin(dbn, Mout)
            # depthwise_BN + relu6 are applied to the A tiles inside the project GEMM: the normalised activation is
            # never written (one read + one write of the 6C-wide tensor less per block, and nothing to save for backward)
            pbn = b["project_bn"]
            if self.fuse_dw_bn:
                ops.pw_gemm(ws[f"y_d{i}"], self.wcopies[b["project"].name]["nk"], ws[f"y_p{i}"], stat_sum=pbn.sum,
                            stat_sqs=pbn.sqs, a_scale=sc, a_shift=sh, a_fin=fin, a_act=ACT_RELU6)
            else:
                ops.bn_act_apply(ws[f"y_d{i}"], ws[f"a_d{i}"], scale=sc, shift=sh, fin=fin, act=ACT_RELU6)
                ops.pw_gemm(ws[f"a_d{i}"], self.wcopies[b["project"].name]["nk"], ws[f"y_p{i}"], stat_sum=pbn.sum,
                            stat_sqs=pbn.sqs)
            fin, sc, sh = self._fin(pbn, Mout)
            ops.bn_act_apply(ws[f"y_p{i}"], ws[f"x{i + 1}"], scale=sc, shift=sh, fin=fin, act=ACT_NONE,
                             res=xin if b["skip"] else None)
        x16 = ws["x17"]
        fh, fw = self.fh, self.fw
        M = B * fh * fw
        # image pooling branch: BN over the batch of B pooled vectors
        ops.global_avgpool_fwd(x16, ws["pooled"])
        ibn = self.image_pooling_bn
        ops.pw_gemm(ws["pooled"], self.wcopies["image_pooling"]["nk32"], ws["y_ip"], stat_sum=ibn.sum, stat_sqs=ibn.sqs)
        fin, sc, sh = self._fin(ibn, B)
        ops.bn_act_apply(ws["y_ip"], ws["b4"], scale=sc, shift=sh, fin=fin, act=ACT_RELU)
        wcp = self.wcopies["concat_projection"]
        ops.pw_gemm(ws["b4"], wcp["nk32"], ws["rowbias"], K=256)
        abn = self.aspp0_bn
        ops.pw_gemm(x16, self.wcopies["aspp0"]["nk"], ws["y_a0"], stat_sum=abn.sum, stat_sqs=abn.sqs)
        fin, sc, sh = self._fin(abn, M)
        ops.bn_act_apply(ws["y_a0"], ws["a_a0"], scale=sc, shift=sh, fin=fin, act=ACT_RELU)
        cbn = self.concat_projection_bn
        ops.pw_gemm(ws["a_a0"], wcp["nk"][:, 256:], ws["y_cp"], row_bias=ws["rowbias"], rows_per_img=fh * fw,
                    stat_sum=cbn.sum, stat_sqs=cbn.sqs)
        fin, sc, sh = self._fin(cbn, M)
        # the dropout mask is a counter-based hash of (seed, optimizer iteration on the device, element index):
        # forward and backward regenerate the same mask, and a replayed CUDA graph still gets a fresh one per step
        ops.bn_act_apply(ws["y_cp"], ws["feat"], scale=sc, shift=sh, fin=fin, act=ACT_RELU,
                         drop_rate=self.dropout_rate if dropout else 0.0, drop_seed=self.dropout_seed,
                         drop_seed_dev=self.adam_step)
        self._head_fwd(ws, ws["feat"])

    def loss_and_head_grad(self, ws, B, use_sample_w: bool):
        """softmax + void-ignoring weighted CE (utils.py:127-130) and d loss / d logits, loss-scaled."""
        npix = B * self.H * self.W
        sw = ws["sample_w"] if use_sample_w else None
        ops.ce_grad_scale(npix, sw, ws["grad_scale"], ws["wcount"], 1.0, self.ls_state)
        ops.fill_zero(ws["loss_sum"])
        if self.head != "subpixel":
            ops.fill_zero(ws["dlogits"])
        ops.resize_softmax_ce(ws["logits"], self.n_out, self.H, self.W, ws["labels"], sw, ws["grad_scale"],
                              ws["dlogits"], ws["loss_sum"], ws["wcount"], ws["argmax"])

    def backward(self, ws, B, dropout: bool):
        geo = ws["geo"]
        fh, fw = self.fh, self.fw
        HW = fh * fw
        ops.fill_zero(self.bn_red)
        ops.fill_zero(self.grads)          # one memset: depthwise / stem gradients are accumulated with atomics
        first = self._first_trainable_index()
        # ---- head
        hc = self.head_conv
        hw = self.wcopies[hc.name]
        g256a, g256b = ws["g256"]
        if self.head == "subpixel":
            # d logits [B,H,W,n] -> GEMM-column order [B,fh,fw,(jj,i,k)] (inverse of the fused store)
            self._unshuffle_dlogits(ws)
            dl = ws["dlogits_lo"]
        else:
            dl = ops.cast(ws["dlogits"], ws["dlogits_lo"])
        if hc.trainable:
            ops.pw_wgrad(ws["feat"], dl, hc.params[0].grad.view(256, -1), N=hc.cout, dbias=hc.params[1].grad,
                         beta=1.0)
        if first > self._order("concat_projection_BN"):
            return
        d_feat = ops.pw_gemm(dl, hw["kn"], g256a, K=hc.cout if self.head == "subpixel" else self.ldl, N=256)
        # ---- concat_projection (+BN, ReLU, Dropout)
        cbn = self.concat_projection_bn
        dy_cp = ops.bn_bwd(ws["y_cp"], d_feat, g256b, scale=cbn.scale, shift=cbn.shift, mean=cbn.mean, rstd=cbn.rstd,
                           act=ACT_RELU, red=cbn.red, dgamma=cbn.gamma.grad, dbeta=cbn.beta.grad,
                           drop_rate=self.dropout_rate if dropout else 0.0, drop_seed=self.dropout_seed,
                           drop_seed_dev=self.adam_step)
        cp = self.concat_projection
        wcp = self.wcopies["concat_projection"]
        gW = cp.params[0].grad.view(512, 256)
        if cp.trainable:
            ops.pw_wgrad(ws["a_a0"], dy_cp, gW[256:], beta=1.0)
        # per-image bias gradient = column sums of dy_cp per image
        ops.global_avgpool_fwd(dy_cp, ws["d_rowbias"])
        if cp.trainable:
            ops.small_gemm(ws["b4"], ws["d_rowbias"], gW[:256], M=256, N=256, K=B, transA=True, alpha=float(HW))
        if first > self._order("aspp0_BN"):
            return
        da_a0 = ops.pw_gemm(dy_cp, wcp["kn"][256:], g256a)
        abn = self.aspp0_bn
        dy_a0 = ops.bn_bwd(ws["y_a0"], da_a0, g256b, scale=abn.scale, shift=abn.shift, mean=abn.mean, rstd=abn.rstd,
                           act=ACT_RELU, red=abn.red, dgamma=abn.gamma.grad, dbeta=abn.beta.grad)
        x16 = ws["x17"]
        if self.aspp0.trainable:
            ops.pw_wgrad(x16, dy_a0, self.aspp0.params[0].grad.view(self.c_last, 256), beta=1.0)
        # ---- image pooling branch
        ops.small_gemm(ws["d_rowbias"], cp.params[0].data.view(512, 256), ws["d_b4"], M=B, N=256, K=256, transB=True,
                       alpha=float(HW))
        ibn = self.image_pooling_bn
        ops.bn_bwd(ws["y_ip"], ws["d_b4"], ws["dy_ip"], scale=ibn.scale, shift=ibn.shift, mean=ibn.mean, rstd=ibn.rstd,
                   act=ACT_RELU, red=ibn.red, dgamma=ibn.gamma.grad, dbeta=ibn.beta.grad)
        ip = self.image_pooling
        if ip.trainable:
            ops.small_gemm(ws["pooled"], ws["dy_ip"], ip.params[0].grad.view(self.c_last, 256), M=self.c_last, N=256, K=B, transA=True)
        if first > self._order("expanded_conv_16_project_BN"):
            return
        ops.small_gemm(ws["dy_ip"], ip.params[0].data.view(self.c_last, 256), ws["d_pooled"], M=B, N=self.c_last, K=256, transB=True)
        gn = ws["g_narrow"]

        def nview(buf, *shape):
            return buf[:int(np.prod(shape))].view(*shape)

        dx = nview(gn[0], B, fh, fw, self.c_last)
        ops.pw_gemm(dy_a0, self.wcopies["aspp0"]["kn"], dx)
        ops.global_avgpool_bwd(ws["d_pooled"], dx, True)
        # ---- backbone blocks in reverse
        cur = 0
        gw = ws["g_wide"]
        for i in range(len(self.blocks) - 1, -1, -1):
            b, g = self.blocks[i], geo[i]
            if i == 12 and self._on_bucket is not None:
                self._on_bucket(0)          # gradients of blocks 13..16, ASPP and the head are complete
            others = [k for k in range(3) if k != cur]
            pbn, dbn = b["project_bn"], b["dw_bn"]
            xin = ws[f"x{i}"]
            # project BN (no activation); residual passes dx straight through
            dy_p = nview(gn[others[0]], B, g["ho"], g["wo"], b["cout"])
            ops.bn_bwd(ws[f"y_p{i}"], dx, dy_p, scale=pbn.scale, shift=pbn.shift, mean=pbn.mean, rstd=pbn.rstd,
                       act=ACT_NONE, red=pbn.red, dgamma=pbn.gamma.grad, dbeta=pbn.beta.grad)
            pj = b["project"]
            if pj.trainable:
                if self.fuse_dw_bn:
                    ops.pw_wgrad(ws[f"y_d{i}"], dy_p, pj.params[0].grad.view(b["mid"], b["cout"]), beta=1.0,
                                 a_scale=dbn.scale, a_shift=dbn.shift, a_act=ACT_RELU6)
                else:
                    ops.pw_wgrad(ws[f"a_d{i}"], dy_p, pj.params[0].grad.view(b["mid"], b["cout"]), beta=1.0)
            if first > self._order(b["dw_bn"].layer.name):
                return
            da_d = nview(gw[0], B, g["ho"], g["wo"], b["mid"])
            ops.pw_gemm(dy_p, self.wcopies[pj.name]["kn"], da_d)
            dy_d = nview(gw[1], B, g["ho"], g["wo"], b["mid"])
            ops.bn_bwd(ws[f"y_d{i}"], da_d, dy_d, scale=dbn.scale, shift=dbn.shift, mean=dbn.mean, rstd=dbn.rstd,
                       act=ACT_RELU6, red=dbn.red, dgamma=dbn.gamma.grad, dbeta=dbn.beta.grad)
            dwl = b["dw"]
            dw_grad = dwl.params[0].grad.view(3, 3, b["mid"]) if dwl.trainable else None
            if b["bid"]:
                ebn = b["expand_bn"]
                stop_here = first > self._order(ebn.layer.name)
                da_e = None if stop_here else nview(gw[0], B, g["h"], g["w"], b["mid"])
                ops.dw_conv_bwd(ws[f"y_e{i}"], dy_d, dwl.params[0].data, dx=da_e, dw=dw_grad,
                                in_shape=(B, g["h"], g["w"], b["mid"]), stride=b["stride"], dilation=b["rate"],
                                pad_top=g["pt"], pad_left=g["pl"], in_scale=ebn.scale, in_shift=ebn.shift, in_act=ACT_RELU6)
                if stop_here:
                    return
                dy_e = nview(gw[1], B, g["h"], g["w"], b["mid"])
                ops.bn_bwd(ws[f"y_e{i}"], da_e, dy_e, scale=ebn.scale, shift=ebn.shift, mean=ebn.mean, rstd=ebn.rstd,
                           act=ACT_RELU6, red=ebn.red, dgamma=ebn.gamma.grad, dbeta=ebn.beta.grad)
                ex = b["expand"]
                if ex.trainable:
                    ops.pw_wgrad(xin, dy_e, ex.params[0].grad.view(b["cin"], b["mid"]), beta=1.0)
                prev_bn = self.blocks[i - 1]["project_bn"].layer.name if i > 0 else "Conv_BN"
                if first > self._order(prev_bn):
                    return
                dx_in = nview(gn[others[1]], B, g["h"], g["w"], b["cin"])
                ops.pw_gemm(dy_e, self.wcopies[ex.name]["kn"], dx_in, residual=dx if b["skip"] else None)
                dx, cur = dx_in, others[1]
            else:
                # block 0: depthwise acts directly on x0 = relu6(BN(stem))
                stop_here = first > self._order("Conv_BN")
                dx_in = None if stop_here else nview(gn[others[1]], B, g["h"], g["w"], b["cin"])
                ops.dw_conv_bwd(xin, dy_d, dwl.params[0].data, dx=dx_in, dw=dw_grad,
                                in_shape=(B, g["h"], g["w"], b["mid"]), stride=b["stride"], dilation=b["rate"],
                                pad_top=g["pt"], pad_left=g["pl"])
                if stop_here:
                    return
                dx, cur = dx_in, others[1]
        # ---- stem
        sbn = self.stem_bn
        others = [k for k in range(3) if k != cur]
        h0, w0 = geo[0]["h"], geo[0]["w"]
        dy_s = nview(gn[others[0]], B, h0, w0, 32)
        ops.bn_bwd(ws["y_stem"], dx, dy_s, scale=sbn.scale, shift=sbn.shift, mean=sbn.mean, rstd=sbn.rstd,
                   act=ACT_RELU6, red=sbn.red, dgamma=sbn.gamma.grad, dbeta=sbn.beta.grad)
        if self.stem.trainable:
            ops.stem_conv_wgrad(ws["img"], dy_s, self.stem.params[0].grad)

    def grad_buckets(self):
        """[(lo, hi)] slices of the flat gradient buffer in the order the backward pass completes them.  Parameters
        lie in forward order, so everything from block 13 on (76 % of the 2.11 M parameters) is final after the first
        ~30 % of the backward pass: that suffix is reduced while blocks 12..0 and the stem are still running, and only
        the 2 MB prefix is left for the end."""
        off = self._by_name["expanded_conv_13_expand"].params[0].offset
        return [(off, self.n_params), (0, off)]

    def _unshuffle_dlogits(self, ws):
        """[B,H,W,n] fp32 gradient -> [B,fh,fw,(jj,i,k)] in the compute dtype (transpose of the fused store)."""
        ops.subpixel_grad_gather(ws["dlogits"], ws["dlogits_lo"], self.fh, self.fw, self.scale)

    def _order(self, layer_name: str) -> int:
        return self._order_map[layer_name]

    def _first_trainable_index(self) -> int:
        self._order_map = {rec.name: i for i, rec in enumerate(self.layers)}
        for i, rec in enumerate(self.layers):
            if rec.trainable:
                return i
        return len(self.layers)

    # ------------------------------------------------------------------------------------------------
    # one optimizer step (graph-captured)
    # ------------------------------------------------------------------------------------------------
    def _fwd_bwd_body(self, ws, B, dropout, use_sample_w):
        self.forward_train(ws, B, dropout)
        self.loss_and_head_grad(ws, B, use_sample_w)
        self.backward(ws, B, dropout)

    def _step_body_bucketed(self, ws, B, dropout, use_sample_w):
        """forward + backward with the gradient buckets all-reduced asynchronously as they complete, then Adam: the
        whole data-parallel step in one capturable body (the collectives run on the process group's own stream;
        `wait()` makes the optimizer depend on them)."""
        buckets, works, fired = self.grad_buckets(), [], set()

        def fire(k):
            lo, hi = buckets[k]
            works.append(self.grad_hook(self.grads[lo:hi], True))
            fired.add(k)

        self._on_bucket = fire
        try:
            self._fwd_bwd_body(ws, B, dropout, use_sample_w)
        finally:
            self._on_bucket = None
        for k in range(len(buckets)):          # (a frozen prefix ends the backward pass early)
            if k not in fired:
                fire(k)
        for w in works:
            if w is not None:
                w.wait()
        self._update_body()

    def _update_body(self):
        c = self.adam_cfg
        if self.ls_state is not None:
            ops.grad_finite_check(self.grads, self.ls_state)
        # frozen layers (trainable=False) are left out of the update inside the kernel: p, m and v untouched
        ops.adam_step(self.params, self.grads, self.adam_m, self.adam_v, self.adam_step, lr=c["lr"], beta1=c["beta1"],
                      beta2=c["beta2"], eps=c["eps"], decay=c["decay"], grad_mult=1.0 / self.world_size,
                      train_mask=self.train_mask if self._any_frozen else None, loss_scale_state=self.ls_state)
        self.refresh_weight_copies()

    def _capture(self, fn):
        """warm-up on a side stream (first-use attribute calls, allocations), then capture into a CUDA graph.
        The warm-up run performs the work once for real; the capture only records."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.n_launches = _lib.launch_count() - n0
        return g

    def train_step(self, img: torch.Tensor, labels: torch.Tensor, sample_w: Optional[torch.Tensor] = None,
                   dropout: bool = True, use_graph: bool = True):
        """img [B,H,W,3] fp32, labels [B,H*W,1] fp32, sample_w [B,H*W] fp32 (device or pinned host tensors).
        Returns the device scalars (loss_sum, wcount): loss = loss_sum / wcount (Keras weighted mean).

        The step is two CUDA graphs -- (forward + loss + backward) and (Adam + weight re-cast) -- with the optional
        gradient hook (the NCCL all-reduce of parallel.py) between them; a single graph when there is no hook."""
        B = img.shape[0]
        ws = self.workspace(B, True)
        if self._weights_dirty:
            self.refresh_weight_copies()
        ws["img"].copy_(img, non_blocking=True)
        ws["labels"].copy_(labels.view(B, -1, 1), non_blocking=True)
        use_sw = sample_w is not None
        if use_sw:
            ws["sample_w"].copy_(sample_w.view(B, -1), non_blocking=True)
        self._fold_dirty = True
        bucketed = self.grad_hook is not None and self.ar_in_graph
        if not use_graph:
            if bucketed:
                self._step_body_bucketed(ws, B, dropout, use_sw)
                return ws["loss_sum"], ws["wcount"]
            self._fwd_bwd_body(ws, B, dropout, use_sw)
            if self.grad_hook is not None:
                self.grad_hook(self.grads)
            self._update_body()
            return ws["loss_sum"], ws["wcount"]
        key = (B, dropout, use_sw)
        graphs = self._graphs.get(key)
        if graphs is None:
            if self.grad_hook is None:
                def whole():
                    self._fwd_bwd_body(ws, B, dropout, use_sw)
                    self._update_body()
                self._graphs[key] = (self._capture(whole), None)
            elif bucketed:
                self._graphs[key] = (self._capture(lambda: self._step_body_bucketed(ws, B, dropout, use_sw)), None)
            else:
                ga = self._capture(lambda: self._fwd_bwd_body(ws, B, dropout, use_sw))
                self.grad_hook(self.grads)
                gb = self._capture(self._update_body)
                self._graphs[key] = (ga, gb)
            return ws["loss_sum"], ws["wcount"]      # the warm-up runs already performed this step
        ga, gb = graphs
        ga.replay()
        if gb is not None:
            self.grad_hook(self.grads)
            gb.replay()
        return ws["loss_sum"], ws["wcount"]

    def graph_launches_per_step(self) -> int:
        """number of libdeeplab_b200 kernel launches recorded in the captured step graph(s)."""
        n = 0
        for ga, gb in self._graphs.values():
            n = max(n, ga.n_launches + (gb.n_launches if gb is not None else 0))
        return n
