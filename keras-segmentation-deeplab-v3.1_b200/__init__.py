"""deeplab_b200 -- B200-native DeepLabV3+ hot path behind the reference's Python API.

Mirrors Golbstein/Keras-segmentation-deeplab-v3.1: `deeplabv3p.Deeplabv3`, `subpixel.Subpixel / icnr_weights`,
`utils.SegModel / do_crf / sparse_crossentropy_ignoring_last_label / Jaccard ...`; the arithmetic runs in the
hand-written sm_100a kernels of libdeeplab_b200.so (include/deeplab_b200.h), reached through ctypes.
"""
__version__ = "0.1.0"

from . import _lib, ops  # noqa: F401
