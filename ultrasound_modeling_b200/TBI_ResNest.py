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
        """reference: self.resModel.save(path) (TBI_ResNest.py:472).  Writes <path>/variables.npz: Keras variable name ->
        array in Keras layout (HWIO / HWOI); `load(path)` reads it back."""
        os.makedirs(path, exist_ok=True)
        np.savez(os.path.join(path, "variables.npz"), **{k: v.cpu().numpy() for k, v in self._o.engine.state_dict().items()})

    def load(self, path):
        with np.load(os.path.join(path, "variables.npz")) as z:
            self._o.engine.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files})



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
        if grad_sync is not None:
            grad_sync.attach(self.engine)
        self._graphs = {}
        self._pin = {}
        self._warm = set()
        self._copy_stream = None
        self._y_event = None

    # ------------------------------------------------------------------ inputs
    PIN_RING = 2

    def _pinned(self, name, shape):
        """a ring of pinned staging buffers per input; each slot remembers the event of the async copy that last read it, and
        the host waits for that event before overwriting the slot (a step never synchronises, so without this the host could
        rewrite the buffer while the previous step's DMA is still queued behind ~10 ms of GPU work)."""
        ring = self._pin.get(name)
        if ring is None or ring["shape"] != tuple(shape):
            ring = self._pin[name] = dict(shape=tuple(shape), i=0,
                                          bufs=[torch.empty(shape, dtype=torch.float32, pin_memory=True) for _ in range(self.PIN_RING)],
                                          events=[None] * self.PIN_RING)
        i = ring["i"]
        ring["i"] = (i + 1) % self.PIN_RING
        if ring["events"][i] is not None:
            ring["events"][i].synchronize()
        return ring, i

    def _stage(self, name, arr, dst: torch.Tensor, stream=None):
        """host (numpy/torch, fp64/fp32) or device tensor -> the engine's static fp32 input buffer"""
        if isinstance(arr, np.ndarray):
            arr = torch.from_numpy(np.ascontiguousarray(arr))
        if arr.device.type == "cpu" and arr.dtype == torch.float32 and arr.is_contiguous() and arr.is_pinned():
            dst.copy_(arr.reshape(dst.shape), non_blocking=True)      # caller's pinned fp32 buffer: DMA straight from it
        elif arr.device.type == "cpu":
            ring, i = self._pinned(name, dst.shape)
            pin = ring["bufs"][i]
            pin.copy_(arr.reshape(dst.shape))                # casts fp64 -> fp32 on the host, like Keras' autocast
            dst.copy_(pin, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.engine.device))
            ring["events"][i] = ev
        else:
            dst.copy_(arr.reshape(dst.shape))

    def _stage_labels(self, y, dst: torch.Tensor):
        """the labels are first needed by the loss, a whole forward pass after the step starts: host labels are copied on a
        side stream while the forward runs (3/4 of the step's host->device bytes are labels); a device tensor goes through
        _stage on the main stream."""
        if isinstance(y, np.ndarray) or (torch.is_tensor(y) and y.device.type == "cpu"):
            dev = self.engine.device
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(dev)
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))      # the previous step's loss has consumed y_in
            with torch.cuda.stream(self._copy_stream):
                self._stage("y", y, dst)
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
            else:
                e.backward()
            e.adam()                                        # hyper-parameters come from the device buffer (set_hyper in _run)

    def _purge_graphs(self):
        """drop CUDA graphs captured against builds whose buffers the engine has released"""
        e = self.engine
        if e.evicted:
            dead = set(e.evicted)
            e.evicted.clear()
            for cache in (self._graphs, self.grad_sync._seg_graphs if self.grad_sync is not None else {}):
                for k in [k for k in cache if k[0] in dead]:
                    del cache[k]
            self._warm = {k for k in self._warm if k[0] not in dead}

    def _run(self, train: bool, draw: bool):
        e = self.engine
        # the learning rate the caller sees (self.optimizer.learning_rate, TBI_ResNest.py:28) reaches the device before the step,
        # eager or replayed: a schedule assigned between steps is honoured by a captured graph too
        world = self.grad_sync.world_size if self.grad_sync is not None else 1
        scale = 1.0 / world if (self.grad_sync is None or self.grad_sync.average) else 1.0
        e.set_hyper(self.optimizer.learning_rate, scale)
        self._purge_graphs()
        key = (e.gen, train, draw)                           # e.gen: a graph is only valid for the build (buffers) it captured
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
            def pre():
                e.prepare()
                if draw:
                    e.draw_dropout()
                e.forward()

            self.grad_sync.step_graphed(e, pre, e.adam, key, loss_fn=e.loss, before_loss=self._wait_labels)
            return
        if g is None:
            # two graphs: [prepare, dropout, forward] | [loss, backward, Adam]; the label copy (side stream) joins between them
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
                    e.adam()
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
    # Format: a directory of .npz archives whose keys are the Keras variable names of the reference graph
    # ("Conv1/kernel", "conv2_1_car_k0_att2_r1/bias", "batch_normalization_3/moving_variance", ...) and whose arrays are
    # in Keras layouts (Conv2D HWIO, Conv2DTranspose HWOI): what `model.get_weights()` / `layer.set_weights()` on the
    # reference side produce and accept, so weights trained on either stack interchange (INTEGRATION.md has the TF-side loop).
    def save_params(self, path: Optional[str] = None):
        path = self.ckpt_dir if path is None else path
        os.makedirs(path, exist_ok=True)
        e = self.engine
        np.savez(os.path.join(path, "variables.npz"), **{k: v.cpu().numpy() for k, v in e.state_dict().items()})
        opt = {"step": np.asarray(int(e.step_count.item()), dtype=np.int64)}
        for tag, buf in (("m", e.adam_m), ("v", e.adam_v)):
            for k, v in e._named(buf, None, trainable_only=True).items():
                opt[f"{tag}/{k}"] = v.detach().cpu().numpy()
        np.savez(os.path.join(path, "optimizer.npz"), **opt)
        return path

    def load_params(self, path: Optional[str] = None):
        path = self.ckpt_dir if path is None else path
        e = self.engine
        with np.load(os.path.join(path, "variables.npz")) as z:
            e.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files})
        opt_path = os.path.join(path, "optimizer.npz")
        if os.path.exists(opt_path):
            with np.load(opt_path) as z:
                for tag, buf in (("m", e.adam_m), ("v", e.adam_v)):
                    for k, dst in e._named(buf, None, trainable_only=True).items():
                        dst.copy_(torch.from_numpy(z[f"{tag}/{k}"]).to(e.device).reshape(dst.shape))
                e.step_count.fill_(int(z["step"]))

    # convenience for tests / interchange
    def state_dict(self):
        return self.engine.state_dict()

    def load_state_dict(self, sd):
        self.engine.load_state_dict(sd)

    def tensor_core_report(self):
        """which tap-GEMM launches of the last steps left the tcgen05 path (see Engine.fallback_report)"""
        return self.engine.fallback_report()
