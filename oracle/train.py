"""ORACLE (test infrastructure only): one Keras-semantics training step of the restated graph under torch autograd.

Restates what `model.fit_generator` executes per batch in the reference (utils.py:231-241 + ipynb:107):
forward in training phase (BatchNorm batch statistics), `sparse_crossentropy_ignoring_last_label` with temporal
sample weights (Keras weighted_masked_objective), backward, Keras-2.2.4 Adam, BN moving-average update
(momentum 0.999 backbone / 0.99 ASPP, Bessel-corrected variance).  Parity status: unpinned (see network.py).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import network as N
from . import ref_ops as R

BN_MOMENTUM_DEFAULT = 0.99


def bn_momentum(layer_name: str) -> float:
    """deeplabv3p.py:178,189,197,322: momentum=0.999 for the MobileNetV2 backbone BNs; Keras default 0.99 elsewhere."""
    if layer_name == "Conv_BN" or layer_name.startswith("expanded_conv"):
        return 0.999
    return BN_MOMENTUM_DEFAULT


def is_bn(ws):
    return len(ws) == 4 and all(w.dim() == 1 for w in ws)


def loss_and_grads(W, x, y, sw=None, net="original", backbone="mobilenetv2", dropout_mask=None, dtype=torch.float64):
    """-> (loss, grads: name -> list of grads for trainable tensors, ctx with batch stats, probs)."""
    Wd = OrderedDict()
    leaves = []
    for name, ws in W.items():
        new = []
        for i, w in enumerate(ws):
            t = w.detach().to(dtype)
            if not (is_bn(ws) and i >= 2):
                t.requires_grad_(True)
                leaves.append((name, i, t))
            new.append(t)
        Wd[name] = new
    logits, probs, ctx = N.deeplabv3_forward(Wd, x.to(dtype), backbone=backbone, net=net, training=True,
                                             dropout_mask=dropout_mask)
    loss = R.keras_weighted_loss(y, probs, None if sw is None else sw.to(dtype))
    gs = torch.autograd.grad(loss, [t for _, _, t in leaves])
    grads = OrderedDict()
    for (name, i, _), g in zip(leaves, gs):
        grads.setdefault(name, {})[i] = g
    return loss.detach(), grads, ctx, probs.detach()


def train_step(W, x, y, sw=None, net="original", iterations=0, lr=7e-4, eps=1e-8, decay=1e-6, frozen=(),
               adam_state=None, dtype=torch.float64):
    """One optimizer step.  Returns (loss, new W, adam_state).  `frozen`: layer names with trainable=False
    (they still normalise with batch statistics, but neither their parameters nor moving stats change)."""
    loss, grads, ctx, _ = loss_and_grads(W, x, y, sw, net, dtype=dtype)
    adam_state = adam_state if adam_state is not None else {}
    newW = OrderedDict()
    for name, ws in W.items():
        out = []
        for i, w in enumerate(ws):
            w = w.to(dtype)
            if is_bn(ws) and i >= 2:
                if name in frozen:
                    out.append(w)
                    continue
                mean, var, cnt = ctx.bn_batch_stats[name]
                mom = bn_momentum(name)
                if i == 2:
                    out.append(w * mom + (1 - mom) * mean)
                else:
                    unbiased = var * cnt / max(cnt - 1, 1)
                    out.append(w * mom + (1 - mom) * unbiased)
                continue
            if name in frozen:
                out.append(w)
                continue
            g = grads[name][i]
            m, v = adam_state.get((name, i), (torch.zeros_like(w), torch.zeros_like(w)))
            p, m, v = R.keras_adam(w, g, m, v, iterations, lr=lr, eps=eps, decay=decay)
            adam_state[(name, i)] = (m, v)
            out.append(p)
        newW[name] = out
    return loss, newW, adam_state
