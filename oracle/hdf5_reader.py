"""Minimal pure-Python reader for the Keras-2.2.4 HDF5 weight files the reference ships (weights/*.h5).

No h5py / libhdf5 exists in this image.  The files are: superblock v0, 8-byte offsets/lengths, v1 object headers,
symbol-table groups (B-tree v1 `TREE` + `SNOD` + local `HEAP`), contiguous uncompressed little-endian float32
datasets, fixed-length string attributes (`layer_names`, `weight_names`) and two vlen-string attributes in one
global heap (`backend`, `keras_version`).  SURVEY.md Appendix D records the byte layout this follows.

Used by the oracle AND by the product's `load_weights` (it is host-side file parsing, not arithmetic), so it
lives in oracle/ only as a thin re-export: the implementation is `deeplab_b200.keras_h5`.
"""
from __future__ import annotations

import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_src = os.path.join(os.path.dirname(_here), "keras-segmentation-deeplab-v3.1_b200", "keras_h5.py")
_spec = importlib.util.spec_from_file_location("_dlb_keras_h5", _src)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["_dlb_keras_h5"] = _mod
_spec.loader.exec_module(_mod)

H5File = _mod.H5File
load_keras_weights = _mod.load_keras_weights
save_keras_weights = getattr(_mod, "save_keras_weights", None)
