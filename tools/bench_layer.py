"""Single-kernel micro-benchmarks of the streaming families at the largest MobileNetV2 layer ([16,64,64,960] fp16,
the block-14..16 expanded tensor): BatchNorm apply / backward (reduce + apply), depthwise forward / backward, and the
1x1 weight gradient.  L2 flushed between iterations, CUDA events on the launch stream, one JSON line per kernel with
algorithmic GB/s and its fraction of the measured HBM copy peak.

  python tools/bench_layer.py                # all
  python tools/bench_layer.py --ncu bn_bwd   # warm launch + ONE more launch of the named family (for ncu -s / -c)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import deeplab_b200  # noqa: E402,F401
from deeplab_b200 import ops  # noqa: E402
from deeplab_b200._lib import ACT_RELU6  # noqa: E402

pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
HBM = float(json.load(open(pk)).get("hbm_gbs", 6650.0)) if os.path.exists(pk) else 6650.0


def main():
    ncu = "--ncu" in sys.argv
    sel = [a for a in sys.argv[1:] if not a.startswith("--")]
    dt = torch.bfloat16 if "--bf16" in sys.argv else torch.float16
    B, H, W, C, Cn = 16, 64, 64, 960, 160
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, H, W, C, device=dev, generator=g).to(dt)
    da = torch.randn(B, H, W, C, device=dev, generator=g).to(dt)
    dx = torch.empty_like(x)
    y = torch.empty_like(x)
    dyn = torch.randn(B, H, W, Cn, device=dev, generator=g).to(dt)
    f32 = dict(device=dev, dtype=torch.float32)
    scale, shift = torch.rand(C, **f32) + 0.5, torch.randn(C, **f32) * 0.1
    mean, rstd = torch.randn(C, **f32) * 0.1, torch.rand(C, **f32) + 0.5
    red = torch.zeros(2 * C, device=dev, dtype=torch.float64)
    dgamma, dbeta = torch.empty(C, **f32), torch.empty(C, **f32)
    wdw = torch.randn(3, 3, C, 1, **f32) * 0.3
    ddw = torch.zeros(3, 3, C, 1, **f32)
    ssum, ssqs = torch.zeros(C, device=dev, dtype=torch.float64), torch.zeros(C, device=dev, dtype=torch.float64)
    dW = torch.empty(C, Cn, **f32)
    es = 2
    n = x.numel()

    def bn_bwd():
        red.zero_()
        ops.bn_bwd(x, da, dx, scale=scale, shift=shift, mean=mean, rstd=rstd, act=ACT_RELU6, red=red, dgamma=dgamma,
                   dbeta=dbeta)

    cases = {
        "bn_act_apply": (lambda: ops.bn_act_apply(x, y, scale=scale, shift=shift, act=ACT_RELU6), 2 * n * es),
        "bn_bwd": (bn_bwd, 5 * n * es),
        "dw_conv_fwd": (lambda: ops.dw_conv_fwd(x, wdw, y, stride=1, dilation=4, pad_top=4, pad_left=4, in_scale=scale,
                                                in_shift=shift, in_act=ACT_RELU6, stat_sum=ssum, stat_sqs=ssqs), 2 * n * es),
        "dw_conv_bwd": (lambda: ops.dw_conv_bwd(x, da, wdw, dx=dx, dw=ddw, stride=1, dilation=4, pad_top=4, pad_left=4,
                                                in_scale=scale, in_shift=shift, in_act=ACT_RELU6), 4 * n * es),
        # the project conv's weight gradient as the training step runs it: A-operand transform (depthwise_BN + relu6) on
        "pw_wgrad": (lambda: ops.pw_wgrad(x.view(-1, C), dyn.view(-1, Cn), dW, a_scale=scale, a_shift=shift, a_act=ACT_RELU6),
                     (n + dyn.numel()) * es + dW.numel() * 4),
    }
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, (fn, byt) in cases.items():
        if sel and name not in sel:
            continue
        fn()
        torch.cuda.synchronize()
        if ncu:
            flush.zero_()
            fn()
            torch.cuda.synchronize()
            continue
        ts = []
        for _ in range(12):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        print(json.dumps({"kernel": name, "shape": [B, H, W, C], "us": round(ms * 1e3, 1), "algorithmic_MB": round(byt / 1e6, 1),
                          "gbs": round(byt / ms / 1e6, 1), "frac_hbm": round(byt / ms / 1e6 / HBM, 3)}))


if __name__ == "__main__":
    main()
