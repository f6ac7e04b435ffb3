"""Keras-surface shim over the Engine: exactly the part of keras.Model / Layer / callbacks / optimizers that the
reference's notebook and utils.py touch:

  model.input, model.layers[i].name/.trainable/.get_weights()/.set_weights(), model.compile(optimizer, loss,
  metrics, sample_weight_mode), model.fit_generator(...), model.fit(...), model.predict(x, batch_size),
  model.train_on_batch, model.evaluate_generator, model.load_weights(path, by_name), model.save_weights(path),
  callbacks ModelCheckpoint / ReduceLROnPlateau / EarlyStopping / TensorBoard (ipynb:157-169), Adam (ipynb:107).

Host-side orchestration only -- the arithmetic is in the engine's CUDA kernels.
"""
from __future__ import annotations

import itertools
import math
import time
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch

from . import keras_h5, ops
from .engine import Engine, LayerRec

_uid: Dict[str, itertools.count] = {}


def _auto_name(prefix: str) -> str:
    c = _uid.setdefault(prefix, itertools.count(1))
    return f"{prefix}_{next(c)}"


class Adam:
    """keras.optimizers.Adam(lr, beta_1, beta_2, epsilon, decay) -- parameters only; the update is dlb_adam_step."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0.0, amsgrad=False, **kwargs):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by the reference")
        self.lr, self.beta_1, self.beta_2 = float(lr), float(beta_1), float(beta_2)
        self.epsilon = 1e-7 if epsilon is None else float(epsilon)     # K.epsilon() default
        self.decay = float(decay)


class Layer:
    """One entry of model.layers."""

    def __init__(self, model, name: str, rec: Optional[LayerRec] = None, kind: str = ""):
        self._model, self.name, self._rec, self.kind = model, name, rec, kind

    @property
    def trainable(self):
        return self._rec.trainable if self._rec is not None else True

    @trainable.setter
    def trainable(self, v):
        if self._rec is not None:
            self._model.engine.set_trainable(self._rec, v)

    def get_weights(self):
        return self._model.engine.get_layer_weights(self._rec) if self._rec is not None else []

    def set_weights(self, ws):
        if self._rec is None:
            if ws:
                raise ValueError(f"layer {self.name} has no weights")
            return
        self._model.engine.set_layer_weights(self._rec, ws)

    @property
    def weights(self):
        return [p.name for p in self._rec.params] if self._rec is not None else []

    @property
    def output(self):
        return SymbolicTensor(self._model, self.name)

    def __repr__(self):
        return f"<Layer {self.name}>"


class SymbolicTensor:
    def __init__(self, model, layer_name):
        self.model, self.layer_name = model, layer_name


class History:
    def __init__(self):
        self.history: Dict[str, list] = {}
        self.epoch: List[int] = []


class Callback:
    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None): ...
    def on_epoch_end(self, epoch, logs=None): ...
    def on_train_end(self, logs=None): ...


class ModelCheckpoint(Callback):
    """ipynb:160-161: ModelCheckpoint(filepath, save_best_only=True, save_weights_only=True, monitor, mode)."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", period=1):
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only = save_best_only
        if mode == "auto":
            mode = "max" if ("acc" in monitor or "Jaccard" in monitor or monitor.startswith("fmeasure")) else "min"
        self.sign = 1.0 if mode == "max" else -1.0
        self.best = -math.inf

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None or not np.isfinite(cur):
                return
            if self.sign * cur > self.best:
                self.best = self.sign * cur
                self.model.save_weights(path)
        else:
            self.model.save_weights(path)


class ReduceLROnPlateau(Callback):
    """ipynb:163-164."""

    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, mode="auto", min_delta=1e-4,
                 cooldown=0, min_lr=0, **kw):
        self.monitor, self.factor, self.patience, self.min_lr, self.min_delta = monitor, factor, patience, min_lr, min_delta
        if mode == "auto":
            mode = "max" if ("acc" in monitor or "Jaccard" in monitor) else "min"
        self.sign = 1.0 if mode == "max" else -1.0
        self.best, self.wait = -math.inf, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.sign * cur > self.best + self.min_delta:
            self.best, self.wait = self.sign * cur, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                new_lr = max(self.model.optimizer.lr * self.factor, self.min_lr)
                self.model.set_lr(new_lr)
                self.wait = 0


class EarlyStopping(Callback):
    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto", **kw):
        self.monitor, self.patience, self.min_delta = monitor, patience, min_delta
        if mode == "auto":
            mode = "max" if ("acc" in monitor or "Jaccard" in monitor) else "min"
        self.sign = 1.0 if mode == "max" else -1.0
        self.best, self.wait = -math.inf, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.sign * cur > self.best + self.min_delta:
            self.best, self.wait = self.sign * cur, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.model.stop_training = True


class TensorBoard(Callback):
    """Scalar logging only (ipynb:158-159 uses histogram_freq=0, write_graph=False): JSON lines in log_dir."""

    def __init__(self, log_dir="./logs", **kw):
        self.log_dir = log_dir

    def on_epoch_end(self, epoch, logs=None):
        import json
        import os
        os.makedirs(self.log_dir, exist_ok=True)
        with open(os.path.join(self.log_dir, "scalars.jsonl"), "a") as f:
            f.write(json.dumps(dict(epoch=epoch, **{k: float(v) for k, v in (logs or {}).items()})) + "\n")


def _metrics_from_confusion(conf: np.ndarray, C: int):
    """conf [B, C+1, C] (row = true label incl. void, col = prediction) -> (Jaccard utils.py:139-157,
    sparse_accuracy_ignoring_last_label utils.py:132-138)."""
    conf = conf.astype(np.float64)
    tp = np.stack([conf[:, i, i] for i in range(C)], 1)                  # [B, C]
    true_cnt = conf[:, :C, :].sum(2)                                      # [B, C] pixels of class i
    pred_cnt = conf.sum(1)                                                # [B, C] predictions of class i (incl. on void)
    union = true_cnt + pred_cnt - tp
    ious = []
    for i in range(C):
        legal = true_cnt[:, i] > 0
        if legal.any():
            ious.append((tp[legal, i] / union[legal, i]).mean())
    jac = float(np.mean(ious)) if ious else float("nan")
    legal_px = conf[:, :C, :].sum()
    acc = float(tp.sum() / legal_px) if legal_px > 0 else float("nan")
    return jac, acc


class Model:
    """The object Deeplabv3(...) / SegModel.create_seg_model(...) return."""

    def __init__(self, engine: Engine, name: str, layer_names: List[str], infer: bool = False):
        self.engine, self.name, self.infer = engine, name, infer
        self.layers: List[Layer] = []
        by = engine._by_name
        for n, kind in layer_names:
            self.layers.append(Layer(self, n, by.get(n), kind))
        self.input = SymbolicTensor(self, self.layers[0].name)
        self.optimizer: Optional[Adam] = None
        self.loss = None
        self.metrics_names = ["loss"]
        self.sample_weight_mode = None
        self.stop_training = False
        self._pinned: Dict[tuple, torch.Tensor] = {}
        self.dropout_in_training = True

    # ---------------------------------------------------------------- introspection
    @property
    def output(self):
        return SymbolicTensor(self, self.layers[-1].name)

    @property
    def input_shape(self):
        return (None, self.engine.H, self.engine.W, 3)

    @property
    def output_shape(self):
        e = self.engine
        return (None, e.H, e.W, e.n_out) if self.infer else (None, e.H * e.W, e.n_out)

    def get_layer(self, name=None, index=None):
        if index is not None:
            return self.layers[index]
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(f"No such layer: {name}")

    def count_params(self):
        return sum(p.size for rec in self.engine.layers for p in rec.params)

    def summary(self, print_fn=print):
        print_fn(f'Model "{self.name}": {len(self.layers)} layers, {self.count_params():,} parameters')

    # ---------------------------------------------------------------- weights
    def _weighted(self):
        return [l for l in self.layers if l._rec is not None]

    def load_weights(self, filepath, by_name=False):
        """Keras topological load (i-th weighted layer <- i-th weighted group) or by_name (deeplabv3p.py:465)."""
        layers, _ = keras_h5.load_keras_weights(filepath)
        file_layers = [(n, [a for _, a in ws]) for n, ws in layers.items() if ws]
        mine = self._weighted()
        if by_name:
            idx = {n: ws for n, ws in file_layers}
            for l in mine:
                if l.name in idx:
                    l.set_weights(idx[l.name])
            return
        if len(file_layers) != len(mine):
            raise ValueError(f"You are trying to load a weight file containing {len(file_layers)} layers into a "
                             f"model with {len(mine)} layers.")
        for l, (_, ws) in zip(mine, file_layers):
            l.set_weights(ws)

    def save_weights(self, filepath, overwrite=True):
        out = OrderedDict()
        for l in self.layers:
            if l._rec is None:
                out[l.name] = []
            else:
                out[l.name] = [(p.name, a) for p, a in zip(l._rec.params, l.get_weights())]
        keras_h5.save_keras_weights(filepath, out)

    def get_weights(self):
        return [a for l in self._weighted() for a in l.get_weights()]

    def set_weights(self, ws):
        it = iter(ws)
        for l in self._weighted():
            l.set_weights([next(it) for _ in l._rec.params])

    # ---------------------------------------------------------------- compile / optimizer
    def compile(self, optimizer, loss=None, metrics=None, sample_weight_mode=None, **kwargs):
        if isinstance(optimizer, str):
            if optimizer.lower() != "adam":
                raise ValueError("only Adam is built (the reference trains with Adam, ipynb:107)")
            optimizer = Adam()
        self.optimizer = optimizer
        self.loss, self.metrics, self.sample_weight_mode = loss, metrics, sample_weight_mode
        e = self.engine
        e.adam_cfg = dict(lr=optimizer.lr, beta1=optimizer.beta_1, beta2=optimizer.beta_2, eps=optimizer.epsilon,
                          decay=optimizer.decay)
        e._graphs.clear()
        self.metrics_names = ["loss", "Jaccard", "sparse_accuracy_ignoring_last_label"]

    def set_lr(self, lr):
        self.optimizer.lr = float(lr)
        self.engine.adam_cfg["lr"] = float(lr)
        self.engine._graphs.clear()        # lr is baked into the captured step

    # ---------------------------------------------------------------- host <-> device staging
    def _stage(self, key, arr: np.ndarray, dtype=torch.float32) -> torch.Tensor:
        """numpy -> pinned host buffer (reused) -> returned as a pinned tensor; the engine issues the async H2D."""
        a = np.asarray(arr)
        buf = self._pinned.get((key, a.shape))
        if buf is None:
            buf = torch.empty(a.shape, dtype=dtype, pin_memory=True)
            self._pinned[(key, a.shape)] = buf
        np.copyto(buf.numpy(), a, casting="unsafe")          # converts (e.g. uint8 images) while copying
        return buf

    def _registered(self, arr):
        """A generator that refills the SAME preallocated arrays for every batch (the reference's SegmentationGenerator
        does: self.X / self.Y / self.SW, utils.py:293-307) needs no staging copy at all: the second time an array
        (same address, size, dtype float32, C-contiguous) shows up it is page-locked in place (cudaHostRegister) and
        from then on the host->device copy reads it directly.  Returns the pinned tensor view or None."""
        a = arr if isinstance(arr, np.ndarray) else None
        if a is None or a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"] or a.nbytes < (1 << 20):
            return None
        reg = getattr(self, "_reg", None)
        if reg is None:
            reg = self._reg = {"seen": {}, "pinned": OrderedDict()}
        ident = (a.ctypes.data, a.nbytes)
        hit = reg["pinned"].get(ident)
        if hit is not None:
            return hit[1]
        n = reg["seen"].get(ident, 0) + 1
        reg["seen"][ident] = n
        if n < 2:
            return None
        if len(reg["seen"]) > 64:
            reg["seen"].clear()
        try:
            rc = torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
            ok = int(rc) == 0 if not isinstance(rc, tuple) else int(rc[0]) == 0
        except Exception:
            ok = False
        if not ok:
            return None
        t = torch.from_numpy(a)
        reg["pinned"][ident] = (a, t)                    # keeps the array alive while it is registered
        while len(reg["pinned"]) > 12:
            _, (old, _) = reg["pinned"].popitem(last=False)
            try:
                torch.cuda.cudart().cudaHostUnregister(old.ctypes.data)
            except Exception:
                pass
        return t

    def _stage_pool(self):
        pool = getattr(self, "_pool", None)
        if pool is None:
            from concurrent.futures import ThreadPoolExecutor
            pool = self._pool = ThreadPoolExecutor(max_workers=4, thread_name_prefix="dlb-stage")
        return pool

    def _stage_parallel(self, key, arr, guard: Optional[torch.cuda.Event]) -> torch.Tensor:
        """Like _stage, but the copy into the pinned buffer is split over the pool's threads (numpy releases the GIL
        inside copyto; one thread moves ~8 GB/s, an 84 MB batch would otherwise cost most of a 9 ms step).  `guard` is
        the event recorded after the previous host->device copy out of this buffer."""
        a = np.asarray(arr)
        buf = self._pinned.get((key, a.shape))
        if buf is None:
            buf = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
            self._pinned[(key, a.shape)] = buf
        if guard is not None:
            guard.synchronize()
        dst = buf.numpy()
        n = a.shape[0]
        parts = min(4, n) if a.nbytes > (1 << 22) else 1
        if parts <= 1:
            np.copyto(dst, a, casting="unsafe")
            return buf
        step = (n + parts - 1) // parts
        futs = [self._stage_pool().submit(np.copyto, dst[i:i + step], a[i:i + step], "unsafe") for i in range(0, n, step)]
        for f in futs:
            f.result()
        return buf

    # ---------------------------------------------------------------- training
    def train_on_batch(self, x, y, sample_weight=None, class_weight=None, return_device=False):
        if self.optimizer is None:
            raise RuntimeError("You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.")
        e = self.engine
        if isinstance(sample_weight, dict):
            sample_weight = sample_weight.get("pred_mask", next(iter(sample_weight.values())))
        xt = x if torch.is_tensor(x) else self._stage("x", x)
        yt = y if torch.is_tensor(y) else self._stage("y", y)
        swt = None
        if sample_weight is not None:
            swt = sample_weight if torch.is_tensor(sample_weight) else self._stage("sw", sample_weight)
        loss_sum, wcount = e.train_step(xt, yt, swt, dropout=self.dropout_in_training)
        ws = e.workspace(xt.shape[0], True)
        conf = torch.zeros(xt.shape[0], e.n_out + 1, e.n_out, device=e.device, dtype=torch.int64)
        ops.confusion(ws["labels"], ws["argmax"], e.n_out, conf)
        if return_device:
            return loss_sum, wcount, conf
        vals = torch.stack([loss_sum[0], wcount[0]]).cpu().numpy()       # D2H read of the step's result
        loss = float(vals[0] / vals[1]) if vals[1] > 0 else float("nan")
        jac, acc = _metrics_from_confusion(conf.cpu().numpy(), e.n_out)
        return [loss, jac, acc]

    def test_on_batch(self, x, y, sample_weight=None):
        """Keras test_on_batch: inference-phase forward, then the same fused loss / argmax / confusion kernels the
        training step uses (dlb_resize_softmax_ce, dlb_confusion)."""
        e = self.engine
        if isinstance(sample_weight, dict):
            sample_weight = sample_weight.get("pred_mask", next(iter(sample_weight.values())))
        xt = x if torch.is_tensor(x) else self._stage("tx", x)
        yt = y if torch.is_tensor(y) else self._stage("ty", y)
        swt = None
        if sample_weight is not None:
            swt = sample_weight if torch.is_tensor(sample_weight) else self._stage("tsw", sample_weight)
        B = xt.shape[0]
        ws = e.workspace(B, False)
        ws["img"].copy_(xt, non_blocking=True)
        loss_sum, wcount, am = e.eval_batch(ws["img"], yt.to(e.device, non_blocking=True) if not yt.is_cuda else yt,
                                            None if swt is None else (swt.to(e.device, non_blocking=True) if not swt.is_cuda else swt))
        conf = torch.zeros(B, e.n_out + 1, e.n_out, device=e.device, dtype=torch.int64)
        ops.confusion(ws["ev_labels"], am, e.n_out, conf)
        vals = torch.stack([loss_sum[0], wcount[0]]).cpu().numpy()
        jac, acc = _metrics_from_confusion(conf.cpu().numpy(), e.n_out)
        return [float(vals[0] / vals[1]) if vals[1] > 0 else float("nan"), jac, acc]

    def _train_pipelined(self, batches, max_steps):
        """Software-pipelined training loop: the host->device copy of batch i+1 (copy stream, double-buffered
        device slots) overlaps the captured step of batch i, and the loss / confusion read-back of step i happens
        after step i+1 has been enqueued.  Yields [loss, Jaccard, accuracy] per step, in order."""
        if self.optimizer is None:
            raise RuntimeError("You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.")
        e = self.engine
        dev = e.device
        main = torch.cuda.current_stream()
        copy_stream = getattr(self, "_copy_stream", None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream()
        slots = getattr(self, "_slots", None)
        pending = None

        # read-backs go through their own stream into pinned host buffers: a `.cpu()` on the main stream would be
        # ordered behind the step that was just enqueued and make the host wait for step i+1 instead of step i (the
        # next batch's host->device copy would then start too late to overlap anything).
        d2h = getattr(self, "_d2h_stream", None)
        if d2h is None:
            d2h = self._d2h_stream = torch.cuda.Stream()

        def finish(p):
            p["done"].synchronize()
            vals = p["h_vals"].numpy()
            loss = float(vals[0] / vals[1]) if vals[1] > 0 else float("nan")
            jac, acc = _metrics_from_confusion(p["h_conf"].numpy().copy(), e.n_out)
            return [loss, jac, acc]

        NSLOT = 3          # pinned staging buffers: one being filled, one waiting, one being read by the copy engine
        guards = getattr(self, "_stage_guards", None)
        if guards is None:
            guards = self._stage_guards = [None] * NSLOT

        def stage(i, batch):
            """numpy (or tensor) batch -> pinned tensors"""
            x, y, sw = self._unpack(batch)
            if isinstance(sw, dict):
                sw = sw.get("pred_mask", next(iter(sw.values())))
            k, g = i % NSLOT, guards[i % NSLOT]

            def one(tag, a):
                if torch.is_tensor(a):
                    return a
                t = self._registered(a)
                if t is not None:
                    aliased[0] = True        # the copy engine reads the caller's own array: see `staged`
                    return t
                return self._stage_parallel((tag, k), a, g)

            xt, yt = one("x", x), one("y", y)
            swt = None if sw is None else one("sw", sw)
            return xt, yt, swt

        aliased = [False]
        h2d_done = [None]

        def staged(batches):
            """Pull -> stage -> yield, one batch at a time on the calling thread.  The captured step of batch i-1 is
            already running on the GPU while batch i is pulled from the generator and copied (4 threads) into pinned
            memory, so the host work is hidden as long as it is shorter than a step; nothing is pulled ahead of time
            because a generator may reuse its arrays for the next batch (the reference's does)."""
            it = iter(batches)
            i = -1
            while True:
                # arrays registered in place are read by the copy engine itself: the previous batch's host->device copy
                # (issued one iteration ago, a few ms) must have finished before the generator may overwrite them
                if aliased[0] and h2d_done[0] is not None:
                    h2d_done[0].synchronize()
                try:
                    batch = next(it)
                except StopIteration:
                    break
                i += 1
                if i >= max_steps:
                    break
                yield stage(i, batch)

        for i, (xt, yt, swt) in enumerate(staged(batches)):
            B = xt.shape[0]
            if slots is None or slots[0]["img"].shape[0] != B:
                slots = self._slots = [dict(img=torch.empty(B, e.H, e.W, 3, device=dev),
                                            labels=torch.empty(B, e.H * e.W, 1, device=dev),
                                            sw=torch.empty(B, e.H * e.W, device=dev),
                                            ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
                for sl in slots:
                    sl["free"].record(main)
            sl = slots[i % 2]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(sl["free"])
                sl["img"].copy_(xt, non_blocking=True)
                sl["labels"].copy_(yt.view(B, -1, 1), non_blocking=True)
                if swt is not None:
                    sl["sw"].copy_(swt.view(B, -1), non_blocking=True)
                sl["ready"].record(copy_stream)
                if not xt.is_cuda:
                    guards[i % NSLOT] = torch.cuda.Event()
                    guards[i % NSLOT].record(copy_stream)
                    h2d_done[0] = guards[i % NSLOT]
            main.wait_event(sl["ready"])
            loss_sum, wcount = e.train_step(sl["img"], sl["labels"], sl["sw"] if swt is not None else None,
                                            dropout=self.dropout_in_training)
            sl["free"].record(main)
            ws = e.workspace(B, True)
            conf = torch.zeros(B, e.n_out + 1, e.n_out, device=dev, dtype=torch.int64)
            ops.confusion(ws["labels"], ws["argmax"], e.n_out, conf)
            vals = torch.stack([loss_sum[0].double(), wcount[0].double()])
            hb = getattr(self, "_host_bufs", None)
            if hb is None or hb[0]["h_conf"].shape != conf.shape:
                hb = self._host_bufs = [dict(h_vals=torch.empty(2, dtype=torch.float64).pin_memory(),
                                             h_conf=torch.empty(conf.shape, dtype=torch.int64).pin_memory()) for _ in range(2)]
            cur = dict(hb[i % 2], done=torch.cuda.Event())
            produced = torch.cuda.Event()
            produced.record(main)
            with torch.cuda.stream(d2h):
                d2h.wait_event(produced)
                cur["h_vals"].copy_(vals, non_blocking=True)
                cur["h_conf"].copy_(conf, non_blocking=True)
                cur["done"].record(d2h)
            vals.record_stream(d2h)
            conf.record_stream(d2h)
            if pending is not None:
                yield finish(pending)
            pending = cur
        if pending is not None:
            yield finish(pending)

    @staticmethod
    def _unpack(batch):
        if len(batch) == 3:
            return batch
        return batch[0], batch[1], None

    def fit_generator(self, generator, steps_per_epoch=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
                      validation_steps=None, max_queue_size=10, workers=1, use_multiprocessing=False, shuffle=True,
                      initial_epoch=0, **kwargs):
        callbacks = list(callbacks or [])
        hist = History()
        for cb in callbacks:
            cb.set_model(self)
            cb.on_train_begin()
        self.stop_training = False
        steps = steps_per_epoch if steps_per_epoch is not None else len(generator)
        for epoch in range(initial_epoch, epochs):
            t0 = time.time()
            agg = np.zeros(3)
            cnt = 0
            it = (generator[i] for i in range(steps)) if hasattr(generator, "__getitem__") else generator
            for vals in self._train_pipelined(it, steps):
                agg += np.nan_to_num(np.array(vals))
                cnt += 1
            logs = {n: v for n, v in zip(self.metrics_names, agg / max(cnt, 1))}
            if validation_data is not None:
                vsteps = validation_steps if validation_steps is not None else len(validation_data)
                vals = self._evaluate_iter(validation_data, vsteps)
                logs.update({"val_" + n: v for n, v in zip(self.metrics_names, vals)})
            if hasattr(generator, "on_epoch_end"):
                generator.on_epoch_end()
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(float(v))
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - {time.time() - t0:.1f}s - " +
                      " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            cb.on_train_end()
        return hist

    def _evaluate_iter(self, data, steps):
        agg = np.zeros(3)
        cnt = 0
        it = (data[i] for i in range(steps)) if hasattr(data, "__getitem__") and not isinstance(data, tuple) else [data]
        for batch in it:
            x, y, sw = self._unpack(batch)
            agg += np.nan_to_num(np.array(self.test_on_batch(x, y, sw)))
            cnt += 1
        return agg / max(cnt, 1)

    def evaluate_generator(self, generator, steps=None, **kw):
        return list(self._evaluate_iter(generator, steps if steps is not None else len(generator)))

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
            sample_weight=None, shuffle=True, **kwargs):
        bs = batch_size or 32
        n = len(x)

        class _Seq:
            def __len__(s):
                return (n + bs - 1) // bs

            def __getitem__(s, i):
                sl = slice(i * bs, min(n, (i + 1) * bs))
                return (x[sl], y[sl], None if sample_weight is None else sample_weight[sl])

        vd = None
        if validation_data is not None:
            vd = tuple(validation_data)
        return self.fit_generator(_Seq(), epochs=epochs, verbose=verbose, callbacks=callbacks, validation_data=vd,
                                  validation_steps=1 if vd is not None else None)

    # ---------------------------------------------------------------- inference
    def _predict_device(self, x) -> torch.Tensor:
        e = self.engine
        xt = x if torch.is_tensor(x) else self._stage("px", x)
        ws = e.workspace(xt.shape[0], False)
        ws["img"].copy_(xt, non_blocking=True)
        return e.forward_infer(ws["img"])

    def predict(self, x, batch_size=32, verbose=0, steps=None):
        """-> numpy [N, H*W, classes] (or [N, H, W, classes] when built with infer=True, deeplabv3p.py:440-444)."""
        e = self.engine
        x = np.asarray(x) if not torch.is_tensor(x) else x
        outs = []
        for i in range(0, len(x), batch_size):
            p = self._predict_device(x[i:i + batch_size])
            outs.append(p.cpu().numpy())
        out = np.concatenate(outs, 0)
        if self.infer:
            out = out.reshape(-1, e.H, e.W, e.n_out)
        return out

    def predict_on_batch(self, x):
        return self.predict(x, batch_size=len(x))


def keras_layer_names(backbone: str, head: str, head_layer_name: str):
    """model.layers in Keras order for the MobileNetV2 graph (cf. the `layer_names` attr of weights/*.h5)."""
    names = [(_auto_name("input"), "input"), (_auto_name("lambda"), "lambda"), ("Conv", "conv"), ("Conv_BN", "bn"),
             (_auto_name("lambda"), "lambda")]
    from .engine import MNV2_BLOCKS
    for (t, s, bid, skip, rate, cout) in MNV2_BLOCKS:
        prefix = "expanded_conv_{}_".format(bid) if bid else "expanded_conv_"
        if bid:
            names += [(prefix + "expand", "conv"), (prefix + "expand_BN", "bn"), (prefix + "expand_relu", "lambda")]
        names += [(prefix + "depthwise", "dw"), (prefix + "depthwise_BN", "bn"), (prefix + "depthwise_relu", "lambda"),
                  (prefix + "project", "conv"), (prefix + "project_BN", "bn")]
        if skip:
            names.append((prefix + "add", "add"))
    a1 = _auto_name("activation")
    names += [(_auto_name("average_pooling2d"), "pool"), ("image_pooling", "conv"), ("image_pooling_BN", "bn"),
              ("aspp0", "conv"), (a1, "activation"), ("aspp0_BN", "bn"), (_auto_name("lambda"), "lambda"),
              ("aspp0_activation", "activation"), (_auto_name("concatenate"), "concat"), ("concat_projection", "conv"),
              ("concat_projection_BN", "bn"), (_auto_name("activation"), "activation"), (_auto_name("dropout"), "dropout")]
    if head == "subpixel":
        names += [(head_layer_name, "subpixel"), (_auto_name("reshape"), "reshape"), ("pred_mask", "activation")]
    elif head == "original":
        names += [(head_layer_name, "conv"), (_auto_name("lambda"), "lambda"), (_auto_name("reshape"), "reshape"),
                  ("pred_mask", "activation")]
    else:
        names += [(head_layer_name, "conv"), (_auto_name("lambda"), "lambda"), (_auto_name("reshape"), "reshape"),
                  (_auto_name("activation"), "activation")]
    return names
