"""ORACLE (test infrastructure only): run-time probe for the TRUE reference dependencies (SURVEY 8c "Run-time upgrade").

The reference delegates its arithmetic to Keras 2.2.4 / TensorFlow 1.x / pydensecrf, none of which exist in the
build image.  A GPU box (or a developer machine) may have them; this module finds out at run time and, when they are
there, exposes the real third-party ops so `tests/test_ref_probe.py` can pin the restatements in oracle/ref_ops.py and
oracle/densecrf_oracle.c against them, and `bench.py` can say which CPU baseline it timed.  Nothing is installed and
nothing under /root/reference is read.
"""
from __future__ import annotations

import importlib


def _version(mod_name: str):
    try:
        m = importlib.import_module(mod_name)
    except Exception:
        return None
    return getattr(m, "__version__", "present")


def probe() -> dict:
    """-> {'tensorflow': version | None, 'keras': ..., 'pydensecrf': ..., 'h5py': ...}"""
    return {name: _version(mod) for name, mod in (("tensorflow", "tensorflow"), ("keras", "keras"),
                                                  ("pydensecrf", "pydensecrf.densecrf"), ("h5py", "h5py"))}


def summary() -> str:
    p = probe()
    have = [f"{k} {v}" for k, v in p.items() if v]
    return "true-reference dependencies present: " + ", ".join(have) if have else \
        "no true-reference dependency (tensorflow / keras / pydensecrf / h5py) importable: CPU restatement used"


def pydensecrf_inference(unary, image, iters=5, sxy_gauss=3, compat_gauss=3, sxy_bilat=80, srgb_bilat=13, compat_bilat=10):
    """The reference's CRF call sequence (utils.py:78-86) on the real pydensecrf.  unary [M, H*W] f32, image [H,W,3] u8."""
    import numpy as np
    import pydensecrf.densecrf as dcrf
    H, W = image.shape[:2]
    d = dcrf.DenseCRF2D(W, H, unary.shape[0])
    d.setUnaryEnergy(np.ascontiguousarray(unary, dtype=np.float32))
    d.addPairwiseGaussian(sxy=sxy_gauss, compat=compat_gauss)
    d.addPairwiseBilateral(sxy=sxy_bilat, srgb=srgb_bilat, rgbim=np.ascontiguousarray(image), compat=compat_bilat)
    return np.array(d.inference(iters), dtype=np.float32)
