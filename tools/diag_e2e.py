"""Where does the end-to-end arm lose time against the device-resident arm?  (diagnostic, not a bench)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import deeplab_b200
from deeplab_b200.model import Adam
from deeplab_b200.utils import SegModel
from deeplab_b200 import ops

B, H, W = 16, 512, 512
sm = SegModel(image_size=(H, W), compute_dtype="float16")
model = sm.create_seg_model("original", n=21, seed=0)
model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6), sample_weight_mode="temporal")
e = model.engine
x, y, sw = bench.synthetic_batch(B, seed=0)
xp, yp, swp = (torch.from_numpy(a).pin_memory() for a in (x, y, sw))
xd, yd, swd = xp.cuda(), yp.cuda(), swp.cuda()


class Seq:
    def __init__(s, n, items): s.n, s.items = n, items
    def __len__(s): return s.n
    def __getitem__(s, i): return s.items[0], s.items[1], {"pred_mask": s.items[2]}


def t_fit(items, steps=20):
    model.fit_generator(Seq(3, items), steps_per_epoch=3, epochs=1, verbose=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.fit_generator(Seq(steps, items), steps_per_epoch=steps, epochs=1, verbose=0)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


def t_steps(steps=20):
    for _ in range(3): e.train_step(xd, yd, swd)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): e.train_step(xd, yd, swd)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


def t_h2d(steps=10):
    d = [torch.empty_like(xd), torch.empty_like(yd), torch.empty_like(swd)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps):
        d[0].copy_(xp, non_blocking=True); d[1].copy_(yp, non_blocking=True); d[2].copy_(swp, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


print(json.dumps({"train_step_device_ms": t_steps(), "fit_generator_pinned_host_ms": t_fit((xp, yp, swp)),
                  "fit_generator_device_tensors_ms": t_fit((xd, yd, swd)), "h2d_84MB_alone_ms": t_h2d()}))
