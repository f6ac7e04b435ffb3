"""Generates the committed golden fixtures from the reference's own artefacts.  Run in the build container, where
/root/reference exists (it does not exist on the GPU box; tests only read the files written here).

  mobilenetv2_original.h5   verbatim copy of the reference's weight FILE (a data artefact, not source): the exact
                            parameters are the only hard pin the reference offers
  subpixel_delta.npz        the layers of weights/mobilenetv2_subpixel.h5 that differ from the original file
                            (concat_projection_BN statistics + the Subpixel head)
  golden_mnv2.npz           oracle outputs (torch-CPU fp32 restatement) on the config-1 input
                            RandomState(0).randint(0,256,(1,512,512,3)): low-res logits, argmax map
  example_crops.npz         3 crops of the reference's example figures (examples/exp{1,3,4}.JPG panel 3) resized to
                            512x512 BGR + the dominant classes the restatement predicts (weak known-answer test)
"""
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import network as N  # noqa: E402
from oracle.hdf5_reader import load_keras_weights  # noqa: E402


def main():
    shutil.copyfile(f"{REF}/weights/mobilenetv2_original.h5", f"{HERE}/mobilenetv2_original.h5")
    a, _ = load_keras_weights(f"{REF}/weights/mobilenetv2_original.h5")
    b, _ = load_keras_weights(f"{REF}/weights/mobilenetv2_subpixel.h5")
    wa = [(k, v) for k, v in a.items() if v]
    wb = [(k, v) for k, v in b.items() if v]
    delta = {}
    for (ka, va), (kb, vb) in zip(wa, wb):
        if not all(np.array_equal(x[1], y[1]) for x, y in zip(va, vb)):
            for i, (_, arr) in enumerate(vb):
                delta[f"{kb}::{i}"] = arr
    np.savez_compressed(f"{HERE}/subpixel_delta.npz", **delta)
    print("subpixel delta layers:", sorted({k.split('::')[0] for k in delta}))

    W = N.weights_from_h5(f"{REF}/weights/mobilenetv2_original.h5")
    x = torch.from_numpy(np.random.RandomState(0).randint(0, 256, (1, 512, 512, 3)).astype(np.float32))
    with torch.no_grad():
        logits, probs, _ = N.deeplabv3_forward(W, x)
    np.savez_compressed(f"{HERE}/golden_mnv2.npz", logits=logits.numpy(),
                        argmax=probs.argmax(-1).numpy().astype(np.uint8).reshape(512, 512),
                        prob_max_mean=np.float64(probs.max(-1).values.double().mean().item()))

    import cv2
    boxes = {"exp1": (551, 768, 53, 260), "exp3": (571, 783, 40, 250), "exp4": (558, 771, 47, 258)}
    crops, dom = {}, {}
    Ws = N.weights_from_h5(f"{REF}/weights/mobilenetv2_subpixel.h5")
    for name, (x0, x1, y0, y1) in boxes.items():
        img = cv2.imread(f"{REF}/examples/{name}.JPG")          # BGR, as the generator feeds the network
        crop = cv2.resize(img[y0:y1, x0:x1], (512, 512), interpolation=cv2.INTER_LINEAR)
        crops[name] = crop
        xin = torch.from_numpy(crop.astype(np.float32))[None]
        with torch.no_grad():
            _, p, _ = N.deeplabv3_forward(W, xin)
            _, ps, _ = N.deeplabv3_forward(Ws, xin, net="subpixel")
        for tag, pp in (("original", p), ("subpixel", ps)):
            cls, cnt = np.unique(pp.argmax(-1).numpy(), return_counts=True)
            order = [int(c) for c in cls[np.argsort(-cnt)] if c != 0][:2]
            dom[f"{name}_{tag}"] = np.array(order, dtype=np.int64)
            print(name, tag, dict(zip(cls.tolist(), cnt.tolist())))
    np.savez_compressed(f"{HERE}/example_crops.npz", **crops, **{"dom_" + k: v for k, v in dom.items()})


if __name__ == "__main__":
    main()
