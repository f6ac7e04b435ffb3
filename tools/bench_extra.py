"""Secondary measurements for BASELINE.md section 4 (not the driver's bench contract -- that is bench.py):
  config 1  MobileNetV2 inference img/s (fp16 / fp32, bs 1 and 16), next to the oracle's CPU inference
  config 3  Xception OS=8 bs=4 inference (fp32 / fp16) + the fused ASPP atrous depthwise stage vs the HBM roofline
  config 5  dense CRF 1024x1024x21, 10 mean-field iterations, bs 8: ms/img (+ the C oracle on one image)
Each result is one JSON line.  CUDA-event timing, warm-up 3, inputs resident on the device unless stated."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import deeplab_b200
from deeplab_b200 import ops
from deeplab_b200.deeplabv3p import Deeplabv3
from deeplab_b200.utils import dense_crf

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
which = set(sys.argv[1:]) or {"1", "3", "5"}


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if "1" in which:
    for dt in ("float16", "float32"):
        m = Deeplabv3(weights=None, input_shape=(512, 512, 3), compute_dtype=dt)
        m.load_weights(os.path.join(ROOT, "tests", "golden", "mobilenetv2_original.h5"))
        for B in (1, 16):
            x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, (B, 512, 512, 3)).astype(np.float32)).cuda()
            ws = m.engine.workspace(B, False)
            ws["img"].copy_(x)
            ms = timed(lambda: m.engine.forward_infer(ws["img"]))
            ms_am = timed(lambda: m.engine.forward_infer(ws["img"], want_probs=False))
            print(json.dumps({"config": 1, "what": "MobileNetV2 'original' inference 512x512 -> probs [B,262144,21] fp32",
                              "dtype": dt, "batch": B, "ms": ms, "img_per_s": B / ms * 1e3,
                              "img_per_s_argmax_only": B / ms_am * 1e3}))
    from oracle import network as N
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    W = N.weights_from_h5(os.path.join(ROOT, "tests", "golden", "mobilenetv2_original.h5"))
    xc = torch.from_numpy(np.random.RandomState(0).randint(0, 256, (1, 512, 512, 3)).astype(np.float32))
    with torch.no_grad():
        N.deeplabv3_forward(W, xc)
        t0 = time.perf_counter()
        for _ in range(3):
            N.deeplabv3_forward(W, xc)
        t = (time.perf_counter() - t0) / 3
    print(json.dumps({"config": 1, "what": "oracle torch-CPU fp32 inference bs 1", "cores": torch.get_num_threads(),
                      "ms": t * 1e3, "img_per_s": 1 / t}))

if "3" in which:
    for dt in ("float32", "float16"):
        m = Deeplabv3(weights=None, input_shape=(512, 512, 3), backbone="xception", OS=8, compute_dtype=dt)
        B = 4
        x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, (B, 512, 512, 3)).astype(np.float32)).cuda()
        e = m.engine
        ws = e.workspace(B, False)
        ws["img"].copy_(x)
        ms = timed(lambda: e.forward_infer(ws["img"]), iters=5)
        e.fused_aspp = False
        ms_unf = timed(lambda: e.forward_infer(ws["img"]), iters=5)
        e.fused_aspp = True
        print(json.dumps({"config": 3, "what": "Xception OS=8 inference 512x512 bs 4", "dtype": dt, "ms": ms,
                          "img_per_s": B / ms * 1e3, "ms_unfused_aspp": ms_unf}))
        # the fused atrous depthwise stage alone
        tdt = torch.float32 if dt == "float32" else torch.float16
        feat = torch.randn(B, 64, 64, 2048, device="cuda").to(tdt)
        wts = [torch.randn(3, 3, 2048, device="cuda") for _ in range(3)]
        sc = [torch.rand(2048, device="cuda") + 0.5 for _ in range(3)]
        sh = [torch.randn(2048, device="cuda") for _ in range(3)]
        ys = [torch.empty_like(feat) for _ in range(3)]
        flush = torch.empty(64 * 1024 * 1024, device="cuda")      # 256 MB > L2: written between timed iterations

        def one():
            flush.zero_()
            ops.aspp_dw3_fwd(feat, wts, [12, 24, 36], sc, sh, ys)

        def only_flush():
            flush.zero_()

        t_all, t_fl = timed(one, iters=10), timed(only_flush, iters=10)
        ms_k = t_all - t_fl
        nbytes = 4 * feat.numel() * feat.element_size()
        print(json.dumps({"config": 3, "what": "fused ASPP atrous depthwise stage (rates 12/24/36), 4x64x64x2048",
                          "dtype": dt, "ms": ms_k, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms_k / 1e6,
                          "peak_gbs": PEAK, "frac": nbytes / ms_k / 1e6 / PEAK, "l2": "flushed between iterations"}))

if "5" in which:
    B, H, W, M = 8, 1024, 1024, 21
    rng = np.random.RandomState(0)
    logits = torch.from_numpy(rng.randn(B, M, H * W).astype(np.float32) * 3).cuda()
    un = -torch.log_softmax(logits, 1)
    import scipy.ndimage as ndi
    imgs = np.stack([(ndi.gaussian_filter(rng.rand(H, W, 3), (8, 8, 0)) * 4 % 1 * 255).astype(np.uint8) for _ in range(B)])
    im = torch.from_numpy(imgs).cuda()
    ms = timed(lambda: dense_crf(un, im, iters=10), iters=2, warm=1)
    ms0 = timed(lambda: dense_crf(un, im, iters=0), iters=2, warm=1)
    print(json.dumps({"config": 5, "what": "dense CRF 1024x1024x21, 10 iters, bs 8 (sxy 3/w 3; sxy 80, srgb 13/w 10)",
                      "ms_per_img": ms / B, "ms_per_img_lattice_build_only": ms0 / B,
                      "ms_per_img_meanfield": (ms - ms0) / B}))
    from oracle import crf as O
    hs = 256
    u_s = un[0].view(M, H, W)[:, :hs, :hs].reshape(M, -1).cpu().numpy()
    t0 = time.perf_counter()
    O.dense_crf(u_s, imgs[0][:hs, :hs], iters=10)
    t = time.perf_counter() - t0
    print(json.dumps({"config": 5, "what": f"C oracle (densecrf restatement, 1 thread) on a {hs}x{hs} crop, 10 iters",
                      "ms": t * 1e3, "ms_per_img_extrapolated_1024": t * 1e3 * (H * W) / (hs * hs)}))
