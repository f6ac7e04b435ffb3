"""Pointwise-GEMM micro-benchmark: the (M, K, N) shapes of the MobileNetV2 training step measured alone, L2 flushed
between iterations, CUDA events on the launch stream.  One JSON line per shape with the algorithmic HBM GB/s
(A + weights + output) and its fraction of the measured copy peak.

  python tools/bench_gemm.py            # all hot shapes
  python tools/bench_gemm.py --ncu      # one un-timed launch per shape (for `ncu -k regex:pw_gemm_tc`)
  python tools/bench_gemm.py 160x960    # only shapes whose KxN matches (suffix r = the residual variant)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import deeplab_b200  # noqa: E402,F401
from deeplab_b200 import ops  # noqa: E402

pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
HBM = float(json.load(open(pk)).get("hbm_gbs", 6650.0)) if os.path.exists(pk) else 6650.0

# (pixels per image side, K, N, stats, residual)  -- bs 16
SHAPES = [
    (256, 16, 96, True, False), (256, 96, 16, False, False), (256, 32, 16, True, False), (128, 24, 144, True, False),
    (128, 144, 24, True, False), (64, 32, 192, True, False), (64, 64, 384, True, False), (64, 96, 576, True, False),
    (64, 576, 96, True, False), (64, 160, 960, True, False), (64, 960, 160, True, False), (64, 160, 960, False, True),
    (64, 960, 320, True, False), (64, 320, 960, False, False), (64, 320, 256, True, False), (64, 256, 256, True, False),
]


def main():
    ncu = "--ncu" in sys.argv
    sel = [a for a in sys.argv[1:] if not a.startswith("--")]
    dt = torch.bfloat16 if "--bf16" in sys.argv else torch.float16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for side, K, N, stats, res in SHAPES:
        if sel and f"{K}x{N}" + ("r" if res else "") not in sel:
            continue
        M = 16 * side * side
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(M, K, device="cuda", generator=g).to(dt)
        Bt = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(dt)
        out = torch.empty(M, N, device="cuda", dtype=dt)
        R = torch.randn(M, N, device="cuda", generator=g).to(dt) if res else None
        ssum = torch.zeros(N, device="cuda", dtype=torch.float64) if stats else None
        ssqs = torch.zeros(N, device="cuda", dtype=torch.float64) if stats else None

        xf = "--xform" in sys.argv
        asc = torch.rand(K, device="cuda", generator=g) + 0.5 if xf else None
        ash = torch.randn(K, device="cuda", generator=g) if xf else None

        def fn():
            ops.pw_gemm(A, Bt, out, residual=R, stat_sum=ssum, stat_sqs=ssqs, a_scale=asc, a_shift=ash,
                        a_act=ops.ACT_RELU6 if xf else ops.ACT_NONE)

        fn()
        torch.cuda.synchronize()
        if ncu:
            flush.zero_()
            fn()          # the launch to capture: ncu -k regex:pw_gemm_tc -s 1 -c 1 with ONE shape selected
            torch.cuda.synchronize()
            continue
        ts = []
        for _ in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        byt = (M * K + N * K + M * N * (2 if res else 1)) * 2
        print(json.dumps({"M": M, "K": K, "N": N, "stats": stats, "residual": res, "us": round(ms * 1e3, 1),
                          "gbs": round(byt / ms / 1e6, 1), "frac_hbm": round(byt / ms / 1e6 / HBM, 3)}))


if __name__ == "__main__":
    main()
