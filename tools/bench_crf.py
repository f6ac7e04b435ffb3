"""Dense-CRF micro-benchmark (BASELINE config 5): 10 mean-field iterations on 1024x1024x21 unaries.
  python tools/bench_crf.py [batch] [--once]      # --once: a single call after warm-up (for ncu launch lists)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deeplab_b200  # noqa: E402,F401
from deeplab_b200.utils import dense_crf  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = int(args[0]) if args else 1
    H = W = int(args[1]) if len(args) > 1 else 1024
    M, iters = 21, 10
    import scipy.ndimage as ndi
    rng = np.random.RandomState(7)
    base = ndi.gaussian_filter(rng.rand(H, W, 3), (8, 8, 0))
    base = ((base - base.min()) / (base.max() - base.min()) * 255).astype(np.uint8)
    img = torch.from_numpy(np.stack([np.roll(base, 61 * b, axis=(0, 1)) for b in range(B)])).cuda()
    g = torch.Generator(device="cuda").manual_seed(1234)
    un = -torch.log_softmax(torch.randn(B, M, H * W, device="cuda", generator=g) * 3.0, dim=1)
    dense_crf(un, img, iters=iters)
    torch.cuda.synchronize()
    if "--once" in sys.argv:
        dense_crf(un, img, iters=iters)
        torch.cuda.synchronize()
        return
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dense_crf(un, img, iters=iters)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2] / B
    N = H * W
    algo = iters * (12 * N * M + 72 * N) + 92 * N
    print(json.dumps({"batch": B, "H": H, "W": W, "M": M, "iters": iters, "ms_per_img": round(ms, 3),
                      "algorithmic_GBps": round(algo / ms / 1e6, 1)}))


if __name__ == "__main__":
    main()
