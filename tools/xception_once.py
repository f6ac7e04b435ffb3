"""One warm + one measured Xception OS=8 inference (bs 4, 512x512) for ncu launch lists:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/xception_once.py float32"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import deeplab_b200  # noqa: F401
from deeplab_b200.deeplabv3p import Deeplabv3

dt = sys.argv[1] if len(sys.argv) > 1 else "float32"
bb = sys.argv[2] if len(sys.argv) > 2 else "xception"
m = Deeplabv3(weights=None, input_shape=(512, 512, 3), backbone=bb, OS=8, compute_dtype=dt)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, (B, 512, 512, 3)).astype(np.float32)).cuda()
ws = m.engine.workspace(B, False)
ws["img"].copy_(x)
for _ in range(2):
    m.engine.forward_infer(ws["img"])
    torch.cuda.synchronize()
