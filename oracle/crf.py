"""ORACLE (test infrastructure only): dense-CRF post-process, restating utils.py:74-91 + pydensecrf.utils.

`unary_from_labels` restates pydensecrf/utils.py (not vendored); the mean-field inference is the
C restatement in densecrf_oracle.c, compiled by __graft_entry__.build_oracle() into oracle/_build/liboracle_crf.so.
Parity status: unpinned by the reference (no golden CRF outputs), see densecrf_oracle.c header.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_crf.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            import subprocess
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-std=c99", "-o", _SO,
                                   os.path.join(_HERE, "densecrf_oracle.c"), "-lm"])
        L = C.CDLL(_SO)
        L.oracle_lattice_build.restype = C.c_void_p
        L.oracle_lattice_build.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_lattice_free.argtypes = [C.c_void_p]
        L.oracle_lattice_size.argtypes = [C.c_void_p]
        L.oracle_lattice_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_lattice_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_crf_inference.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                           C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib = L
    return _lib


def unary_from_labels(labels, n_labels, gt_prob, zero_unsure=True):
    """pydensecrf.utils.unary_from_labels."""
    assert 0 < gt_prob < 1
    labels = labels.flatten()
    with np.errstate(divide="ignore"):
        n_energy = -np.log((1.0 - gt_prob) / (n_labels - 1)) if n_labels > 1 else np.inf
    p_energy = -np.log(gt_prob)
    U = np.full((n_labels, len(labels)), n_energy, dtype="float32")
    U[labels - 1 if zero_unsure else labels, np.arange(U.shape[1])] = p_energy
    if zero_unsure:
        U[:, labels == 0] = -np.log(1.0 / n_labels)
    return U


def lattice_filter(features: np.ndarray, values: np.ndarray):
    """K(values) for arbitrary features [N, d] and values [N, vs] (unnormalised splat-blur-slice)."""
    L = lib()
    f = np.ascontiguousarray(features, dtype=np.float32)
    v = np.ascontiguousarray(values, dtype=np.float32)
    N, d = f.shape
    h = L.oracle_lattice_build(f.ctypes.data, d, N)
    out = np.empty_like(v)
    L.oracle_lattice_compute(h, out.ctypes.data, v.ctypes.data, v.shape[1])
    M = L.oracle_lattice_size(h)
    off = np.empty((N, d + 1), dtype=np.int32)
    bary = np.empty((N, d + 1), dtype=np.float32)
    L.oracle_lattice_get(h, off.ctypes.data, bary.ctypes.data)
    L.oracle_lattice_free(h)
    return out, M, off, bary


def dense_crf(unary: np.ndarray, image: np.ndarray, iters=5, sxy_g=3.0, compat_g=3.0, sxy_b=80.0, srgb_b=13.0,
              compat_b=10.0) -> np.ndarray:
    """unary [M, N] float32 energies, image [H, W, 3] uint8 -> Q [M, N]."""
    H, W = image.shape[:2]
    M = unary.shape[0]
    u = np.ascontiguousarray(unary, dtype=np.float32)
    im = np.ascontiguousarray(image, dtype=np.uint8)
    Q = np.empty_like(u)
    lib().oracle_crf_inference(H, W, M, iters, u.ctypes.data, im.ctypes.data, sxy_g, compat_g, sxy_b, srgb_b,
                               compat_b, Q.ctypes.data)
    return Q


def do_crf(im, mask, zero_unsure=True):
    """utils.py:74-91, literal."""
    colors, labels = np.unique(mask, return_inverse=True)
    labels = labels.reshape(-1)
    image_size = mask.shape[:2]
    n_labels = len(set(labels.flat))
    U = unary_from_labels(labels, n_labels, gt_prob=.7, zero_unsure=zero_unsure)
    Q = dense_crf(U, im.astype("uint8"), iters=5)
    MAP = np.argmax(Q, axis=0).reshape(image_size)
    unique_map = np.unique(MAP)
    for u in unique_map:
        np.putmask(MAP, MAP == u, colors[u])
    return MAP
