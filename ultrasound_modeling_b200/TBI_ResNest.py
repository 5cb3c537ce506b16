"""Drop-in for the reference's ``TBI_ResNest.ResNest`` (TBI_ResNest.py:15-55): same constructor,
same attributes (``resModel``, ``optimizer``, ``loss``, ``learning_rate``, ``class_factor``), same
``step(x, y, train) -> (loss[H,W], accuracy, probs)``; the arithmetic runs in libtbi_sm100.so.

    from ultrasound_modeling_b200.TBI_ResNest import ResNest
    net = ResNest(256, 256, 1, 3, ksize=3, radix=2, kpaths=1, learning_rate=5e-3)
    loss, acc, probs = net.step(x, y, train=True)          # x,y: numpy or torch, NHWC

Extra keyword arguments (all optional) choose the storage dtype, the device, CUDA-graph replay and
data-parallel gradient averaging; defaults reproduce the reference semantics, including its quirks
(always-on dropout, inference-mode BatchNorm, channel-axis softmax in split-attention, [H,W] loss).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .engine import Engine


class _Adam:
    """holder mirroring tf.optimizers.Adam's public knobs (TBI_ResNest.py:28); the update itself is tbi_adam_multi"""

    def __init__(self, learning_rate):
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = 0.9, 0.999, 1e-7


class _ResModel:
    """what the reference gets back from tf.keras.Model(img_input, result): callable x -> probs"""

    def __init__(self, owner: "ResNest"):
        self._o = owner

    def __call__(self, x, training=None):
        return self._o.predict(x)

    @property
    def trainable_variables(self):
        named = self._o.engine._named(self._o.engine.params, self._o.engine.stats)
        return [v for k, v in named.items() if not (k.endswith("/moving_mean") or k.endswith("/moving_variance"))]

    def save(self, path):
        os.makedirs(path, exist_ok=True)
        torch.save({k: v.cpu() for k, v in self._o.engine.state_dict().items()}, os.path.join(path, "variables.pt"))


class ResNest:
    def __init__(self, height, width, channel, num_class, ksize, radix=4, kpaths=4, learning_rate=1e-3,
                 ckpt_dir='./Checkpoint', *, dtype="bf16", device=None, impl=_lib.IMPL_AUTO, use_cuda_graph=True,
                 seed=0, grad_sync=None):
        self.height, self.width, self.channel, self.num_class = height, width, channel, num_class
        self.ksize, self.learning_rate = ksize, learning_rate
        self.radix, self.kpaths = radix, kpaths
        self.ckpt_dir = ckpt_dir
        if device is None:
            device = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cuda"
        self.engine = Engine(height, width, channel, num_class, ksize, radix, kpaths, dtype=dtype, device=device,
                             impl=impl, seed=seed)
        self.resModel = _ResModel(self)
        self.optimizer = _Adam(learning_rate)
        self.class_factor = [0.06329, 0.027567, 0.90914]     # unused by the reference too (TBI_ResNest.py:30)
        self.loss = self.my_loss_cat
        self.use_cuda_graph = use_cuda_graph
        self.grad_sync = grad_sync                           # parallel.GradSync or None
        self._graphs = {}
        self._pin = {}
        self._warm = set()
        self._copy_stream = None
        self._y_event = None

    # ------------------------------------------------------------------ inputs
    def _stage(self, name, arr, dst: torch.Tensor):
        """host (numpy/torch, fp64/fp32) or device tensor -> the engine's static fp32 input buffer"""
        if isinstance(arr, np.ndarray):
            arr = torch.from_numpy(np.ascontiguousarray(arr))
        if arr.device.type == "cpu" and arr.dtype == torch.float32 and arr.is_contiguous() and arr.is_pinned():
            dst.copy_(arr.reshape(dst.shape), non_blocking=True)      # caller's pinned fp32 buffer: DMA straight from it
        elif arr.device.type == "cpu":
            pin = self._pin.get(name)
            if pin is None or pin.shape != dst.shape:
                pin = torch.empty(dst.shape, dtype=torch.float32, pin_memory=True)
                self._pin[name] = pin
            pin.copy_(arr.reshape(dst.shape))                # casts fp64 -> fp32 on the host, like Keras' autocast
            dst.copy_(pin, non_blocking=True)
        else:
            dst.copy_(arr.reshape(dst.shape))

    def _stage_labels(self, y, dst: torch.Tensor):
        """the labels are first needed by the loss, a whole forward pass after the step starts: a pinned fp32 host tensor is
        copied on a side stream while the forward runs (3/4 of the step's host->device bytes are labels); anything else goes
        through _stage on the main stream."""
        if torch.is_tensor(y) and y.device.type == "cpu" and y.dtype == torch.float32 and y.is_contiguous() and y.is_pinned():
            dev = self.engine.device
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(dev)
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))      # the previous step's loss has consumed y_in
            with torch.cuda.stream(self._copy_stream):
                dst.copy_(y.reshape(dst.shape), non_blocking=True)
            self._y_event = self._copy_stream.record_event()
        else:
            self._stage("y", y, dst)
            self._y_event = None

    def _wait_labels(self):
        if self._y_event is not None:
            torch.cuda.current_stream(self.engine.device).wait_event(self._y_event)
            self._y_event = None

    # ------------------------------------------------------------------ device step
    def _device_step(self, train: bool, draw: bool):
        e = self.engine
        e.prepare()
        if draw:
            e.draw_dropout()
        e.forward()
        self._wait_labels()
        e.loss()
        if train:
            if self.grad_sync is not None:
                self.grad_sync.backward_and_sync(e)
                e.adam(self.optimizer.learning_rate, 1.0 / self.grad_sync.world_size)
            else:
                e.backward()
                e.adam(self.optimizer.learning_rate, 1.0)

    def _run(self, train: bool, draw: bool):
        key = (self.engine.N, train, draw)
        if not self.use_cuda_graph:
            self._device_step(train, draw)
            return
        g = self._graphs.get(key)
        if g is None:
            if key not in self._warm:                        # first call eager: lazy CUDA/NCCL init must not be captured
                self._warm.add(key)
                self._device_step(train, draw)
                return
        if train and self.grad_sync is not None:
            # data parallel: graph segments + eager NCCL between them (collectives are never captured)
            e = self.engine

            def pre():
                e.prepare()
                if draw:
                    e.draw_dropout()
                e.forward()

            self.grad_sync.step_graphed(e, pre, lambda: e.adam(self.optimizer.learning_rate, 1.0 / self.grad_sync.world_size), key,
                                        loss_fn=e.loss, before_loss=self._wait_labels)
            return
        if g is None:
            # two graphs: [prepare, dropout, forward] | [loss, backward, Adam]; the label copy (side stream) joins between them
            e = self.engine
            self._wait_labels()
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                e.prepare()
                if draw:
                    e.draw_dropout()
                e.forward()
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=g1.pool()):
                e.loss()
                if train:
                    e.backward()
                    e.adam(self.optimizer.learning_rate, 1.0)
            g = self._graphs[key] = (g1, g2)
        g[0].replay()
        self._wait_labels()
        g[1].replay()

    def step(self, x, y, train=False, dropout_masks=None):
        """reference: TBI_ResNest.py:35-55.  dropout_masks: None = draw fresh masks (reference behaviour,
        dropout is on even when train=False); a list of three 0/1 keep-masks = use exactly these (parity);
        False = dropout disabled."""
        e = self.engine
        n = int(x.shape[0])
        e.build(n)
        self._stage("x", x, e.x_in)
        self._stage_labels(y, e.y_in)
        draw = dropout_masks is None
        if not draw:
            e.set_dropout(None if dropout_masks is False else dropout_masks)
        self._run(bool(train), draw)
        acc = e.correct.to(torch.float32) / float(n * self.height * self.width)
        return e.loss_map, acc.reshape(()), e.probs

    def predict(self, x, dropout_masks=None):
        """forward only (the reference's resModel(x)); returns probabilities NHWC"""
        e = self.engine
        e.build(int(x.shape[0]))
        self._stage("x", x, e.x_in)
        if dropout_masks is None:
            e.draw_dropout()
        else:
            e.set_dropout(None if dropout_masks is False else dropout_masks)
        e.prepare()
        e.forward()
        e.y_in.zero_()
        e.loss()
        return e.probs

    def my_loss_cat(self, y_true, y_pred):
        """reference: TBI_ResNest.py:234-248 on probabilities; evaluated by the fused softmax+loss kernel
        (softmax(log p) == p for normalised p)."""
        e = self.engine
        y_true = torch.as_tensor(y_true).to(device=e.device, dtype=torch.float32).contiguous()
        y_pred = torch.as_tensor(y_pred).to(device=e.device, dtype=torch.float32).contiguous()
        n, h, w, c = y_pred.shape
        logits = torch.log(y_pred.clamp_min(1e-30))
        probs = torch.empty_like(y_pred)
        out = torch.empty(h, w, dtype=torch.float32, device=e.device)
        correct = torch.zeros(1, dtype=torch.int32, device=e.device)
        _lib.check(e.L.tbi_softmax_loss_fwd_bwd(_lib.F32, n, h, w, c, logits.data_ptr(), y_true.data_ptr(), probs.data_ptr(),
                                                out.data_ptr(), correct.data_ptr(), None, 0, e.stream()), "softmax_loss")
        # the kernel normalises by the model's H*W like the reference does (self.height*self.width)
        return out * (float(h * w) / float(self.height * self.width))

    # ------------------------------------------------------------------ persistence (reference :57-78 is broken; this works)
    def save_params(self):
        os.makedirs(self.ckpt_dir, exist_ok=True)
        e = self.engine
        torch.save({"variables": {k: v.cpu() for k, v in e.state_dict().items()}, "adam_m": e.adam_m.cpu(), "adam_v": e.adam_v.cpu(),
                    "step": int(e.step_count.item())}, os.path.join(self.ckpt_dir, "tbi_resnest.pt"))

    def load_params(self):
        e = self.engine
        ck = torch.load(os.path.join(self.ckpt_dir, "tbi_resnest.pt"), map_location="cpu")
        e.load_state_dict(ck["variables"])
        e.adam_m.copy_(ck["adam_m"]); e.adam_v.copy_(ck["adam_v"]); e.step_count.fill_(ck["step"])

    # convenience for tests / interchange
    def state_dict(self):
        return self.engine.state_dict()

    def load_state_dict(self, sd):
        self.engine.load_state_dict(sd)
