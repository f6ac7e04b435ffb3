"""Modified aligned Xception-65 DeepLabV3+ (backbone='xception', deeplabv3p.py:272-313, ASPP :375-410, decoder
:414-429) -- inference executor on the sm_100a kernels (BASELINE config 3: OS=8, bs=4, fp32).

Built from the same kernels as the MobileNetV2 engine plus three small ones (dense 3x3 conv, pixel subsample,
feature-map bilinear resize).  BatchNorm is folded into the producing kernel's epilogue:
    SepConv_BN (deeplabv3p.py:47-84)  =  dw_conv(prologue ReLU if not depth_activation; epilogue BN_dw [+ReLU])
                                         -> pw_gemm(epilogue BN_pw [+ReLU] [+ residual / shortcut])
The ASPP concat (5 branches) is never built: the image-pooling branch enters `concat_projection` as a per-image
bias and the four spatial branches are written straight into channel slices of one [M, 1024] buffer.

The reference's own Xception path is broken as shipped (`layers.add` NameError, deeplabv3p.py:147,149) and has no
weights; it is implemented by its evident intent (bonlime/keras-deeplab-v3-plus v1.1).  Training on this backbone
is not built.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops
from ._lib import ACT_NONE, ACT_RELU
from .engine import BN, Engine, LayerRec
from .model import Model, _auto_name


class XceptionEngine(Engine):
    def __init__(self, input_shape=(512, 512, 3), classes=21, OS=16, compute_dtype=torch.float32, device="cuda", seed=0):
        self.OS = 8 if OS == 8 else 16
        if self.OS == 8:
            self.entry_block3_stride, self.middle_rate, self.exit_rates, self.atrous = 1, 2, (2, 4), (12, 24, 36)
        else:
            self.entry_block3_stride, self.middle_rate, self.exit_rates, self.atrous = 2, 1, (1, 2), (6, 12, 18)
        H, W = int(input_shape[0]), int(input_shape[1])
        if H % 16 or W % 16:
            raise ValueError("xception: input height/width must be multiples of 16")
        super().__init__(input_shape=(H, W, 3), classes=classes, head="bare", compute_dtype=compute_dtype, device=device,
                         seed=seed)
        self.scale = 4
        self.fused_aspp = True          # (layer-wise path) one pass over x for the three atrous depthwise convs
        self.fused_sepconv = True       # 16-bit: ASPP branches / decoder SepConvs as ONE tcgen05 kernel each
        self._packs: Dict = {}
        self._bufs: Dict = {}
        self._ones = torch.ones(2048, device=self.device)
        self._zeros = torch.zeros(2048, device=self.device)

    # ------------------------------------------------------------------ spec
    def _sep(self, prefix, cin, cout, eps):
        return dict(dw=self._dw(prefix + "_depthwise", cin), dw_bn=self._bn(prefix + "_depthwise_BN", cin, eps, 0.99),
                    pw=self._conv(prefix + "_pointwise", 1, cin, cout), pw_bn=self._bn(prefix + "_pointwise_BN", cout, eps, 0.99))

    def _xblock(self, prefix, cin, depths, skip, stride, rate, depth_activation=False, return_skip=False):
        seps, c = [], cin
        for i, d in enumerate(depths):
            seps.append(self._sep(prefix + "_separable_conv{}".format(i + 1), c, d, 1e-3))
            c = d
        blk = dict(prefix=prefix, seps=seps, skip=skip, stride=stride, rate=rate, depth_activation=depth_activation,
                   return_skip=return_skip, cin=cin, cout=c)
        if skip == "conv":
            blk["shortcut"] = self._conv(prefix + "_shortcut", 1, cin, depths[-1])
            blk["shortcut_bn"] = self._bn(prefix + "_shortcut_BN", depths[-1], 1e-3, 0.99)
        return blk

    def _build_spec(self, head_layer_name):
        self.stem = self._conv("entry_flow_conv1_1", 3, 3, 32)
        self.stem_bn = self._bn("entry_flow_conv1_1_BN", 32, 1e-3, 0.99)
        self.conv1_2 = self._conv("entry_flow_conv1_2", 3, 32, 64)
        self.conv1_2_bn = self._bn("entry_flow_conv1_2_BN", 64, 1e-3, 0.99)
        xb = []
        xb.append(self._xblock("entry_flow_block1", 64, [128, 128, 128], "conv", 2, 1))
        xb.append(self._xblock("entry_flow_block2", 128, [256, 256, 256], "conv", 2, 1, return_skip=True))
        xb.append(self._xblock("entry_flow_block3", 256, [728, 728, 728], "conv", self.entry_block3_stride, 1))
        for i in range(16):
            xb.append(self._xblock("middle_flow_unit_{}".format(i + 1), 728, [728, 728, 728], "sum", 1, self.middle_rate))
        xb.append(self._xblock("exit_flow_block1", 728, [728, 1024, 1024], "conv", 1, self.exit_rates[0]))
        xb.append(self._xblock("exit_flow_block2", 1024, [1536, 1536, 2048], "none", 1, self.exit_rates[1],
                               depth_activation=True))
        self.xblocks = xb
        self.image_pooling = self._conv("image_pooling", 1, 2048, 256)
        self.image_pooling_bn = self._bn("image_pooling_BN", 256, 1e-5, 0.99)
        self.aspp0 = self._conv("aspp0", 1, 2048, 256)
        self.aspp0_bn = self._bn("aspp0_BN", 256, 1e-5, 0.99)
        self.aspp = [self._sep("aspp{}".format(i), 2048, 256, 1e-5) for i in (1, 2, 3)]
        self.concat_projection = self._conv("concat_projection", 1, 1280, 256)
        self.concat_projection_bn = self._bn("concat_projection_BN", 256, 1e-5, 0.99)
        self.feature_projection0 = self._conv("feature_projection0", 1, 256, 48)
        self.feature_projection0_bn = self._bn("feature_projection0_BN", 48, 1e-5, 0.99)
        self.decoder = [self._sep("decoder_conv0", 304, 256, 1e-5), self._sep("decoder_conv1", 256, 256, 1e-5)]
        name = head_layer_name or ("logits_semantic" if self.classes == 21 else "custom_logits_semantic")
        self.head_conv = self._conv(name, 1, 256, self.n_out, bias=True)
        self.bns: List[BN] = [self.stem_bn, self.conv1_2_bn]
        for b in xb:
            for s in b["seps"]:
                self.bns += [s["dw_bn"], s["pw_bn"]]
            if b["skip"] == "conv":
                self.bns.append(b["shortcut_bn"])
        self.bns += [self.image_pooling_bn, self.aspp0_bn]
        for s in self.aspp:
            self.bns += [s["dw_bn"], s["pw_bn"]]
        self.bns += [self.concat_projection_bn, self.feature_projection0_bn]
        for s in self.decoder:
            self.bns += [s["dw_bn"], s["pw_bn"]]
        self.blocks = []        # MobileNetV2-only structures of the base class stay empty

    # ------------------------------------------------------------------ buffers
    def _buf(self, tag, *shape, dtype=None):
        dtype = dtype or self.dtype
        key = (tag, shape, dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(*shape, device=self.device, dtype=dtype)
            self._bufs[key] = t
        return t

    def workspace(self, B: int, training: bool):
        if training:
            raise NotImplementedError("training on the Xception backbone is not built")
        key = (B, False)
        if key not in self._ws:
            ctot = sum(b.C for b in self.bns)
            self._ws[key] = dict(img=torch.empty(B, self.H, self.W, 3, device=self.device),
                                 fold=torch.empty(2 * ctot, device=self.device),
                                 logits=torch.zeros(B, self.H // 4, self.W // 4, self.ldl, device=self.device),
                                 probs=torch.empty(B, self.H * self.W, self.n_out, device=self.device),
                                 argmax=torch.empty(B, self.H * self.W, device=self.device, dtype=torch.uint8))
        return self._ws[key]

    def train_step(self, *a, **k):
        raise NotImplementedError("training on the Xception backbone is not built")

    # ------------------------------------------------------------------ forward
    def _sepconv(self, tag, x, sep, stride, rate, depth_activation, residual=None, out=None):
        """SepConv_BN (deeplabv3p.py:47-84) with folded BatchNorms."""
        B, H, W, C = x.shape
        dwbn, pwbn = sep["dw_bn"], sep["pw_bn"]
        if stride == 1:
            Ho, pt, _ = ops.tf_same_pad(H, 3, 1, rate)
            Wo, pl, _ = ops.tf_same_pad(W, 3, 1, rate)
        else:   # explicit ZeroPadding2D + 'valid' (deeplabv3p.py:61-69)
            k_eff = 3 + 2 * (rate - 1)
            pt = pl = (k_eff - 1) // 2
            Ho = (H + (k_eff - 1) - k_eff) // stride + 1
            Wo = (W + (k_eff - 1) - k_eff) // stride + 1
        y = self._buf(tag + "/dw", B, Ho, Wo, C)
        pre = not depth_activation       # Activation('relu') before the depthwise conv
        ops.dw_conv_fwd(x, sep["dw"].params[0].data, y, stride=stride, dilation=rate, pad_top=pt, pad_left=pl,
                        in_scale=self._ones if pre else None, in_shift=self._zeros if pre else None,
                        in_act=ACT_RELU if pre else ACT_NONE, out_scale=dwbn.fscale, out_shift=dwbn.fshift,
                        out_act=ACT_RELU if depth_activation else ACT_NONE)
        cout = sep["pw"].cout
        z = out if out is not None else self._buf(tag + "/pw", B, Ho, Wo, cout)
        ops.pw_gemm(y, self._wi(sep["pw"].name), z, N=cout, n_store=cout, col_scale=pwbn.fscale,
                    col_shift=pwbn.fshift, act=ACT_RELU if depth_activation else ACT_NONE, residual=residual)
        return z

    def _dw_pack(self, tag, seps):
        """Packed depthwise taps + folded BN of `seps` for the fused kernel (rebuilt whenever the BNs are re-folded)."""
        pk = self._packs.get(tag)
        if pk is None:
            pk = ops.sepconv_pack_dw([s_["dw"].params[0].data for s_ in seps], [s_["dw_bn"].fscale for s_ in seps],
                                     [s_["dw_bn"].fshift for s_ in seps], self.dtype)
            self._packs[tag] = pk
        return pk

    def _xception_block(self, x, blk):
        """_xception_block (deeplabv3p.py:119-155)."""
        B, H, W, C = x.shape
        tag = blk["prefix"]
        shortcut = None
        if blk["skip"] == "conv":
            s = blk["stride"]
            xin = x
            if s > 1:
                xin = ops.subsample(x, self._buf(tag + "/sub", B, (H + s - 1) // s, (W + s - 1) // s, C), s)
            bn = blk["shortcut_bn"]
            shortcut = self._buf(tag + "/short", xin.shape[0], xin.shape[1], xin.shape[2], blk["cout"])
            ops.pw_gemm(xin, self._wi(blk["shortcut"].name), shortcut, col_scale=bn.fscale, col_shift=bn.fshift)
        elif blk["skip"] == "sum":
            shortcut = x
        r, skip_t = x, None
        for i, sep in enumerate(blk["seps"]):
            last = i == 2
            r = self._sepconv(f"{tag}/{i}", r, sep, blk["stride"] if last else 1, blk["rate"], blk["depth_activation"],
                              residual=shortcut if last else None)
            if i == 1:
                skip_t = r
        return (r, skip_t) if blk["return_skip"] else (r, None)

    def forward_infer(self, img: torch.Tensor, want_probs=True, want_argmax=False):
        B = img.shape[0]
        ws = self.workspace(B, False)
        if self._weights_dirty:
            self.refresh_weight_copies()
            self._fold_dirty = True
        if getattr(self, "_fold_dirty", True) or getattr(self, "_fold_ws", None) is not ws:
            self._fold_all(ws)
            self._fold_dirty, self._fold_ws = False, ws
            self._packs = {}
        H, W = self.H, self.W
        x = self._buf("stem", B, H // 2, W // 2, 32)
        ops.stem_conv_fwd(img, self.stem.params[0].data, x, out_scale=self.stem_bn.fscale, out_shift=self.stem_bn.fshift,
                          out_act=ACT_RELU)
        y = self._buf("conv1_2", B, H // 2, W // 2, 64)
        ops.conv3x3_fwd(x, self.conv1_2.params[0].data, y, out_scale=self.conv1_2_bn.fscale,
                        out_shift=self.conv1_2_bn.fshift, out_act=ACT_RELU)
        x, skip1 = y, None
        for blk in self.xblocks:
            x, sk = self._xception_block(x, blk)
            if sk is not None:
                skip1 = sk
        fh, fw = x.shape[1], x.shape[2]
        M = B * fh * fw
        # ---- ASPP (deeplabv3p.py:375-410)
        pooled = self._buf("pooled", B, 2048, dtype=torch.float32)
        ops.global_avgpool_fwd(x, pooled)
        bn = self.image_pooling_bn
        b4 = self._buf("b4", B, 256, dtype=torch.float32)
        ops.pw_gemm(pooled, self.wcopies["image_pooling"]["nk32"], b4, col_scale=bn.fscale, col_shift=bn.fshift, act=ACT_RELU)
        cbn = self.concat_projection_bn
        wcp = self.wcopies["concat_projection"]
        rowbias = self._buf("rowbias", B, 256, dtype=torch.float32)
        ops.pw_gemm(b4, wcp["nk32"], rowbias, K=256, col_scale=cbn.fscale)
        cat = self._buf("aspp_cat", B, fh, fw, 1024)
        bn0 = self.aspp0_bn
        fused = self.fused_sepconv and self.dtype != torch.float32
        if fused and fw <= 128:
            # aspp0 + aspp1..3 in one launch: x is read once, the depthwise results stay in shared memory
            # (dlb_sepconv_fused_fwd; deeplabv3p.py:385-399)
            pw_bns = [bn0] + [s_["pw_bn"] for s_ in self.aspp]
            ops.sepconv_fused_fwd(x, [0] + list(self.atrous),
                                  [self._wi("aspp0")] + [self.wcopies[s_["pw"].name]["nk"] for s_ in self.aspp],
                                  self._dw_pack("aspp", self.aspp), [b_.fscale for b_ in pw_bns],
                                  [b_.fshift for b_ in pw_bns], [cat[..., 256 * i:256 * (i + 1)] for i in range(4)])
        else:
            ops.pw_gemm(x, self._wi("aspp0"), cat[..., 0:256], N=256, n_store=256, col_scale=bn0.fscale,
                        col_shift=bn0.fshift, act=ACT_RELU)
        if fused and fw <= 128:
            pass
        elif self.fused_aspp and fh * fw * 32 <= 200 * 1024:
            # fused atrous depthwise stage: x is read once for the three rates (dlb_aspp_dw3_fwd)
            dws = [self._buf(f"aspp{i + 1}/dw", B, fh, fw, 2048) for i in range(3)]
            ops.aspp_dw3_fwd(x, [s_["dw"].params[0].data for s_ in self.aspp], list(self.atrous),
                             [s_["dw_bn"].fscale for s_ in self.aspp], [s_["dw_bn"].fshift for s_ in self.aspp], dws)
            for i, sep in enumerate(self.aspp):
                pwbn = sep["pw_bn"]
                ops.pw_gemm(dws[i], self._wi(sep["pw"].name), cat[..., 256 * (i + 1):256 * (i + 2)], N=256,
                            n_store=256, col_scale=pwbn.fscale, col_shift=pwbn.fshift, act=ACT_RELU)
        else:
            for i, sep in enumerate(self.aspp):
                self._sepconv(f"aspp{i + 1}", x, sep, 1, self.atrous[i], True, out=cat[..., 256 * (i + 1):256 * (i + 2)])
        feat = self._buf("aspp_out", B, fh, fw, 256)
        ops.pw_gemm(cat, self._wi("concat_projection", slice(256, None)), feat, col_scale=cbn.fscale, col_shift=cbn.fshift, row_bias=rowbias,
                    rows_per_img=fh * fw, act=ACT_RELU)
        # ---- decoder (deeplabv3p.py:414-429)
        dh, dw_ = H // 4, W // 4
        dcat = self._buf("dec_cat", B, dh, dw_, 304)
        ops.resize_bilinear(feat, dcat, 256)
        fbn = self.feature_projection0_bn
        ops.pw_gemm(skip1, self._wi("feature_projection0"), dcat[..., 256:304], N=48, n_store=48,
                    col_scale=fbn.fscale, col_shift=fbn.fshift, act=ACT_RELU)
        if fused and dw_ <= 128:
            # decoder_conv0/1 (deeplabv3p.py:426-429): depthwise 3x3 + BN + ReLU + 1x1 + BN + ReLU, one kernel each
            d = dcat
            for i, sep in enumerate(self.decoder):
                o = self._buf(f"decoder_conv{i}/pw", B, dh, dw_, 256)
                ops.sepconv_fused_fwd(d, [1], [self.wcopies[sep["pw"].name]["nk"]], self._dw_pack(f"dec{i}", [sep]),
                                      [sep["pw_bn"].fscale], [sep["pw_bn"].fshift], [o])
                d = o
        else:
            d = self._sepconv("decoder_conv0", dcat, self.decoder[0], 1, 1, True)
            d = self._sepconv("decoder_conv1", d, self.decoder[1], 1, 1, True)
        # ---- head
        hw = self.wcopies[self.head_conv.name]
        ops.pw_gemm(d, self._wi(self.head_conv.name), ws["logits"], col_shift=self.head_conv.params[1].data, n_store=self.ldl)
        ops.resize_softmax_fwd(ws["logits"], self.n_out, H, W, ws["probs"] if want_probs else None,
                               ws["argmax"] if want_argmax or not want_probs else None)
        return ws["probs"] if want_probs else ws["argmax"]


def build_xception_model(input_shape, classes, OS, infer, dtype, seed):
    e = XceptionEngine(input_shape=input_shape, classes=classes, OS=OS, compute_dtype=dtype, seed=seed)
    names = [(_auto_name("input"), "input"), (_auto_name("lambda"), "lambda")]
    names += [(rec.name, rec.kind) for rec in e.layers[:-1]]
    names += [(_auto_name("dropout"), "dropout"), (e.head_conv.name, "conv"), (_auto_name("lambda"), "lambda"),
              (_auto_name("reshape"), "reshape"), (_auto_name("activation"), "activation")]
    return Model(e, "deeplabv3p", names, infer=infer)
