"""Timeline of the pipelined training loop: when do the H2D copies run relative to the steps?  (diagnostic)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import deeplab_b200
from deeplab_b200.model import Adam
from deeplab_b200.utils import SegModel

B, H, W = 16, 512, 512
sm = SegModel(image_size=(H, W), compute_dtype="float16")
model = sm.create_seg_model("original", n=21, seed=0)
model.compile(optimizer=Adam(lr=7e-4, epsilon=1e-8, decay=1e-6), sample_weight_mode="temporal")
e = model.engine
x, y, sw = bench.synthetic_batch(B, seed=0)
xp, yp, swp = (torch.from_numpy(a).pin_memory() for a in (x, y, sw))
dev = e.device
main = torch.cuda.current_stream()
cs = torch.cuda.Stream()
slots = [dict(img=torch.empty(B, H, W, 3, device=dev), labels=torch.empty(B, H * W, 1, device=dev),
              sw=torch.empty(B, H * W, device=dev), ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
for s in slots: s["free"].record(main)
for _ in range(3): e.train_step(slots[0]["img"].zero_(), slots[0]["labels"].zero_(), slots[0]["sw"].fill_(1))
torch.cuda.synchronize()
N = 8
E = lambda: torch.cuda.Event(enable_timing=True)
t0 = E(); t0.record(main)
rec = []
pend = None
for i in range(N):
    sl = slots[i % 2]
    c0, c1, s0, s1 = E(), E(), E(), E()
    with torch.cuda.stream(cs):
        cs.wait_event(sl["free"])
        c0.record(cs)
        sl["img"].copy_(xp, non_blocking=True); sl["labels"].copy_(yp, non_blocking=True); sl["sw"].copy_(swp, non_blocking=True)
        c1.record(cs)
        sl["ready"].record(cs)
    main.wait_event(sl["ready"])
    s0.record(main)
    ls, wc = e.train_step(sl["img"], sl["labels"], sl["sw"])
    s1.record(main)
    sl["free"].record(main)
    cur = ls.clone()
    if pend is not None: pend.cpu()
    pend = cur
    rec.append((c0, c1, s0, s1))
torch.cuda.synchronize()
for i, (c0, c1, s0, s1) in enumerate(rec):
    print(i, "copy %.2f-%.2f  step %.2f-%.2f" % (t0.elapsed_time(c0), t0.elapsed_time(c1), t0.elapsed_time(s0), t0.elapsed_time(s1)))
