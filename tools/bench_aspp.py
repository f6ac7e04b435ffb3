"""The fused ASPP block (north_star's named roofline target) measured alone: BASELINE config 3 geometry, Xception OS=8,
bs 4 -> x [4,64,64,2048], branches aspp0 (1x1) + aspp1..3 (atrous depthwise 3x3 rates 12/24/36 + BN + ReLU + 1x1 +
BN + ReLU), outputs into the [M,1024] concat buffer.

  fused    : ONE dlb_sepconv_fused_fwd launch (tcgen05; depthwise results never leave the SM)
  layerwise: aspp0 GEMM + dlb_aspp_dw3_fwd + 3 GEMMs (what the fused kernel replaces)

Algorithmic bytes: x once + the four [M,256] outputs + the weights once.  FLOPs: 2*M*C*N per
branch + 2*9*M*C per depthwise.  L2 is flushed (256 MB write) between timed iterations; CUDA events on the launch
stream.  One JSON line per variant.  `--ncu` runs one un-timed launch of each (for ncu captures)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import deeplab_b200  # noqa: E402,F401
from deeplab_b200 import ops  # noqa: E402
from deeplab_b200._lib import ACT_RELU  # noqa: E402

pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAKS = json.load(open(pk)) if os.path.exists(pk) else {}
HBM = float(PEAKS.get("hbm_gbs", 6650.0))
TF = float(PEAKS.get("bf16_tflops", 1650.0))


def timed(fn, flush, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dt = torch.bfloat16 if "--bf16" in sys.argv else torch.float16
    B, H, W, C, N = 4, 64, 64, 2048, 256
    rates = [0, 12, 24, 36]
    if "--os16" in sys.argv:
        H = W = 32
        rates = [0, 6, 12, 18]
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, W, C, device="cuda", generator=g).to(dt)
    wdw = [torch.randn(3, 3, C, device="cuda", generator=g) / 3 for _ in range(3)]
    dsc = [torch.rand(C, device="cuda", generator=g) + 0.5 for _ in range(3)]
    dsh = [torch.randn(C, device="cuda", generator=g) * 0.1 for _ in range(3)]
    wpw = [(torch.randn(N, C, device="cuda", generator=g) / C ** 0.5).to(dt) for _ in range(4)]
    psc = [torch.rand(N, device="cuda", generator=g) + 0.5 for _ in range(4)]
    psh = [torch.randn(N, device="cuda", generator=g) * 0.1 for _ in range(4)]
    cat_f = torch.zeros(B, H, W, 4 * N, device="cuda", dtype=dt)
    cat_l = torch.zeros_like(cat_f)
    outs_f = [cat_f[..., i * N:(i + 1) * N] for i in range(4)]
    outs_l = [cat_l[..., i * N:(i + 1) * N] for i in range(4)]
    pack = ops.sepconv_pack_dw(wdw, dsc, dsh, dt)
    dws = [torch.empty_like(x) for _ in range(3)]
    flush = torch.empty(64 * 1024 * 1024, device="cuda")

    def fused():
        ops.sepconv_fused_fwd(x, rates, wpw, pack, psc, psh, outs_f)

    def layerwise():
        ops.pw_gemm(x, wpw[0], outs_l[0], N=N, n_store=N, col_scale=psc[0], col_shift=psh[0], act=ACT_RELU)
        ops.aspp_dw3_fwd(x, wdw, rates[1:], dsc, dsh, dws)
        for i in range(3):
            ops.pw_gemm(dws[i], wpw[i + 1], outs_l[i + 1], N=N, n_store=N, col_scale=psc[i + 1], col_shift=psh[i + 1],
                        act=ACT_RELU)

    fused(); layerwise()
    torch.cuda.synchronize()
    diff = (cat_f.float() - cat_l.float()).abs().max().item() / cat_l.float().abs().max().item()
    if "--ncu" in sys.argv:
        print(json.dumps({"what": "ncu pass (one launch of each variant)", "rel_diff_fused_vs_layerwise": diff}))
        return
    M = B * H * W
    es = x.element_size()
    algo = M * C * es + 4 * M * N * es + 4 * N * C * es + 3 * 9 * C * 4
    flops = 4 * 2 * M * C * N + 3 * 2 * 9 * M * C
    layer_bytes = (M * C * es) * (1 + 1 + 3 + 3) + 4 * M * N * es + 4 * N * C * es
    for name, fn, nl in (("fused", fused, 1), ("layerwise", layerwise, 5)):
        ms = timed(fn, flush)
        print(json.dumps({
            "what": f"ASPP block aspp0+aspp1..3, x [{B},{H},{W},{C}] -> [M,{4 * N}], {name}", "dtype": str(dt).split(".")[-1],
            "launches": nl, "ms": ms, "img_per_s": B / ms * 1e3,
            "algorithmic_bytes": algo, "hbm_gbs_on_algorithmic_bytes": algo / ms / 1e6, "hbm_peak_gbs": HBM,
            "hbm_frac": algo / ms / 1e6 / HBM, "flops": flops, "tflops": flops / ms / 1e9, "tensor_peak_tflops": TF,
            "tensor_frac": flops / ms / 1e9 / TF, "layerwise_bytes_if_materialised": layer_bytes,
            "rel_diff_fused_vs_layerwise": diff, "l2": "flushed between iterations"}))


if __name__ == "__main__":
    main()
