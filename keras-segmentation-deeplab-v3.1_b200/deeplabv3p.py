"""`Deeplabv3(...)` -- the reference's model constructor (deeplabv3p.py:209-466) on the B200-native engine.

Same signature, defaults, argument validation and exceptions; returns a `model.Model` exposing the Keras surface
the notebook uses.  Nothing of the Keras graph-building code is reused: the graph is the Engine's layer spec and
every layer executes as a hand-written sm_100a kernel.

Deliberate differences (documented in INTEGRATION.md):
  * weights='pascal_voc' cannot download (no network; the reference itself raises NameError there,
    deeplabv3p.py:458): raises RuntimeError with that explanation.  Use load_weights(path).
  * extra keyword `compute_dtype` ('float16' default | 'bfloat16' | 'float32' exact-parity mode).
"""
from __future__ import annotations

import torch

from .engine import Engine
from .model import Model, keras_layer_names

_DTYPES = {"float16": torch.float16, "fp16": torch.float16, "bfloat16": torch.bfloat16, "bf16": torch.bfloat16,
           "float32": torch.float32, "fp32": torch.float32}


def _resolve_dtype(compute_dtype):
    if isinstance(compute_dtype, torch.dtype):
        return compute_dtype
    try:
        return _DTYPES[str(compute_dtype)]
    except KeyError:
        raise ValueError(f"compute_dtype must be one of {sorted(_DTYPES)}")


def Deeplabv3(weights='pascal_voc', input_tensor=None, infer=False, input_shape=(512, 512, 3), classes=21,
              backbone='mobilenetv2', OS=16, alpha=1., compute_dtype='float16', seed=0):
    """Instantiates the DeepLabV3+ architecture (see deeplabv3p.py:209-246 for the argument semantics).

    backbone='mobilenetv2' silently runs at output stride 8 whatever `OS` says (deeplabv3p.py:316)."""
    if not (weights in {'pascal_voc', None}):
        raise ValueError('The `weights` argument should be either '
                         '`None` (random initialization) or `pascal_voc` '
                         '(pre-trained on PASCAL VOC)')
    if not (backbone in {'xception', 'mobilenetv2'}):
        raise ValueError('The `backbone` argument should be either '
                         '`xception`  or `mobilenetv2` ')
    if input_tensor is not None:
        # Keras builds the graph on top of an existing tensor (deeplabv3p.py:260-266); here the tensor only fixes the
        # input geometry -- its (H, W, 3) overrides input_shape -- and is returned as model.input
        shp = tuple(getattr(input_tensor, "shape", ()))
        if len(shp) not in (3, 4) or shp[-1] != 3:
            raise ValueError("input_tensor must have shape (H, W, 3) or (batch, H, W, 3)")
        input_shape = tuple(int(v) for v in shp[-3:])
    if weights == 'pascal_voc':
        raise RuntimeError("weights='pascal_voc' needs a download (WEIGHTS_PATH_X / WEIGHTS_PATH_MOBILE, "
                           "deeplabv3p.py:42-43); this box has no network and the reference itself fails there "
                           "(get_file is never imported, deeplabv3p.py:458). Build with weights=None and call "
                           "model.load_weights(path).")
    dt = _resolve_dtype(compute_dtype)
    if backbone == 'xception':
        from .xception import build_xception_model
        return build_xception_model(input_shape, classes, OS, infer, dt, seed)
    engine = Engine(input_shape=input_shape, classes=classes, head="bare", alpha=alpha, compute_dtype=dt, seed=seed)
    names = keras_layer_names(backbone, "bare", engine.head_conv.name)
    model = Model(engine, "deeplabv3p", names, infer=infer)
    if input_tensor is not None:
        model.input = input_tensor
    return model
