"""Drop-in for the reference's ``VisionTransformer.py`` (the model ``MainParallel.py`` / ``MainNumpy.py`` train): the Variant B
ResNeSt encoder, the 8-block ViT bridge on its 80 tokens, and the DecoderCup, with the reference's training step
(label-smoothed categorical cross-entropy / global batch, global-norm clip 1.0, Keras Adam) on the B200 path.

    from ultrasound_modeling_b200.VisionTransformer import VisionTransformer
    net = VisionTransformer(batch_size=64)                       # img_size=(256, 80), num_classes=3, learning_rate=1e-3
    loss, probs = net.train_step(x, y)                           # x [N,256,80,10], y [N,256,80,3] (numpy or torch, NHWC)
    loss, probs = net.step(x, y)                                 # evaluation: forward + loss
    probs, attn_weights = net.forward(x)                         # == net.visionModel(x)

Same constructor, attributes (``visionModel``, ``optimizer``, ``loss``, ``learning_rate``, ``batch_size``, ``transformer``,
``decoder``) and return values as VisionTransformer.py:192-257.  Every arithmetic step is a libtbi_sm100.so entry point:
the Dense layers are 1x1 tap-GEMMs over the token grid (tcgen05 in bf16), LayerNorm = tbi_layernorm_c_*, the attention core
one fused kernel per block (tbi_attention_*), GELU and the loss their own kernels, Adam + clip one launch over a flat buffer
(tbi_sumsq + tbi_adam_multi_dev).  torch allocates, reshapes and owns the variables; the backward pass is the tape of
ResNest.py in this package, shared by the three modules.

Reference quirks kept: attention scores are divided by sqrt(num_heads) (:42); ``position_embeddings`` is the constant
``tf.zeros`` (:108) -- not a variable, so nothing is added; every dropout rate is 0; the model returns probabilities and the
loss is applied to them; ``compute_loss`` divides the SUM over all pixels by the constructor's ``batch_size``.
Variable names = attribute paths of the reference objects (oracle/vit_oracle.py lists them); Dense kernels are stored as
[1,1,in,out].
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from ._lib import ACT_NONE
from .Decoder import DecoderCup
from .ResNest import ResNest, VariableStore, _Layer

VIT_LN_EPS = 1e-6
ENC = "transformer/embeddings/hybrid_model/"
DEC = "decoder/"
TR = "transformer/encoder/"


class _Adam:
    """tf.optimizers.Adam's public knobs (VisionTransformer.py:204)"""

    def __init__(self, learning_rate):
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = 0.9, 0.999, 1e-7


class Transformer(_Layer):
    """VisionTransformer.py:81-190 (Embeddings, Encoder of 8 Blocks, Transformer): tokens live as [1, N*T/8, 8, C] NHWC tensors
    so that every Dense is a 1x1 convolution over full 16 x 8 pixel tiles."""

    def __init__(self, img_size, *, store: VariableStore, dtype, device, hidden=512, mlp_dim=2048, num_heads=4, num_layers=8):
        super().__init__(store, "")
        self.hidden, self.mlp_dim, self.num_heads, self.num_layers = hidden, mlp_dim, num_heads, num_layers
        self.hybrid_model = ResNest(img_size[0], img_size[1], 10, radix=3, ksize=3, kpaths=3, dtype=dtype, device=device, _store=store, _prefix=ENC)
        self.attn_scale = 1.0 / math.sqrt(float(num_heads))          # the reference's choice (:42), not 1/sqrt(head_dim)

    # -- variables: Dense / patch-embedding kernels get Keras' default glorot_uniform ----------------------------------
    def _dense_var(self, name, cin, cout):
        s = self._s
        if name + "/kernel" not in s.vars:
            lim = math.sqrt(6.0 / (cin + cout))
            s.vars[name + "/kernel"] = ((torch.rand(1, 1, cin, cout, generator=s.gen) * 2 - 1) * lim).to(s.device)

    def _dense(self, x, name, cout, residual=None):
        self._dense_var(name, x.shape[3], cout)
        return self._conv(x, name, 1, cout, residual=residual)

    # -- tape ops that ResNest.py's _Layer does not have -----------------------------------------------------------------
    def _reshape(self, x, shape):
        y = x.reshape(shape)
        if self._s.recording:
            s = self._s

            def bwd():
                g = s.gget(y)
                if g is not None:
                    s.gacc(x, g.reshape(x.shape))
            s.tape.append(bwd)
        return y

    def _norm(self, x, name):
        """LayerNormalization(epsilon=1e-6) out of place (the input stays alive as the residual)"""
        c = x.shape[3]
        g = self._s.vector(name + "/gamma", c, 1.0); b = self._s.vector(name + "/beta", c, 0.0)
        y = ops.layernorm_c(x, g, b, act=ACT_NONE, eps=VIT_LN_EPS)
        if self._s.recording:
            s = self._s

            def bwd():
                dy = s.gget(y)
                if dy is None:
                    return
                dx, dg, db = ops.layernorm_c_bwd(x, y, dy, g, act=ACT_NONE, eps=VIT_LN_EPS)
                s.pacc(name + "/gamma", dg); s.pacc(name + "/beta", db)
                s.gacc(x, dx)
            s.tape.append(bwd)
        return y

    def _gelu(self, x):
        y = ops.gelu(x)
        if self._s.recording:
            s = self._s

            def bwd():
                dy = s.gget(y)
                if dy is not None:
                    s.gacc(x, ops.gelu_bwd(x, dy))
            s.tape.append(bwd)
        return y

    def _attention(self, q, k, v, n, t):
        """q, k, v: [1, n*t/8, 8, C] -> context in the same shape, probabilities fp32 [n, heads, t, t]"""
        c = q.shape[3]
        tok = lambda z: z.reshape(n, t, c)
        ctx3, probs = ops.attention(tok(q), tok(k), tok(v), self.num_heads, self.attn_scale)
        ctx = ctx3.reshape(q.shape)
        if self._s.recording:
            s = self._s

            def bwd():
                d = s.gget(ctx)
                if d is None:
                    return
                dq, dk, dv = ops.attention_bwd(tok(q), tok(k), tok(v), probs, d.reshape(n, t, c), self.num_heads, self.attn_scale)
                s.gacc(q, dq.reshape(q.shape)); s.gacc(k, dk.reshape(k.shape)); s.gacc(v, dv.reshape(v.shape))
            s.tape.append(bwd)
        return ctx, probs

    def block(self, h, i, n, t):
        """Block.forward :136-147"""
        p = f"{TR}layer_{i}/"
        a = self._norm(h, p + "attention_norm")
        q = self._dense(a, p + "attn/query", self.hidden); k = self._dense(a, p + "attn/key", self.hidden); v = self._dense(a, p + "attn/value", self.hidden)
        ctx, probs = self._attention(q, k, v, n, t)
        h = self._dense(ctx, p + "attn/out", self.hidden, residual=h)
        m = self._norm(h, p + "ffn_norm")
        f = self._gelu(self._dense(m, p + "ffn/fc1", self.mlp_dim))
        return self._dense(f, p + "ffn/fc2", self.hidden, residual=h), probs

    def forward(self, x):
        """-> (encoded tokens [N, T, hidden], [attention probabilities per layer], features [x_3, x_2, x_1])"""
        x4, feats = self.hybrid_model.forward(x, record=False)          # the caller decides about recording (shared tape)
        n, gh, gw, _ = x4.shape
        t = gh * gw
        if (n * t) % 8 != 0:
            raise _lib.TbiError(f"VisionTransformer: batch*tokens = {n}*{t} must be a multiple of 8 (token tiles of 8)")
        e = self._dense(x4, "transformer/embeddings/patch_embeddings", self.hidden)       # 1x1 Conv2D; + position_embeddings == 0
        h = self._reshape(e, (1, n * t // 8, 8, self.hidden))
        weights = []
        for i in range(self.num_layers):
            h, w = self.block(h, i, n, t)
            weights.append(w)
        h = self._norm(h, TR + "encoder_norm")
        return self._reshape(h, (n, t, self.hidden)), weights, feats


class _VisionModel:
    """what the reference gets from tf.keras.Model(inputs, outputs) (:213-218): callable, trainable_variables, save"""

    def __init__(self, owner: "VisionTransformer"):
        self._o = owner

    def __call__(self, x, training=None):
        return self._o.forward(x)

    @property
    def trainable_variables(self):
        return [v for k, v in self._o.variables().items() if not (k.endswith("/moving_mean") or k.endswith("/moving_variance"))]

    def save(self, path):
        os.makedirs(path, exist_ok=True)
        np.savez(os.path.join(path, "variables.npz"), **{k: v.detach().cpu().numpy() for k, v in self._o.variables().items()})

    def load(self, path):
        with np.load(os.path.join(path, "variables.npz")) as z:
            self._o.load_variables({k: z[k] for k in z.files})


class VisionTransformer:
    def __init__(self, batch_size, img_size=(256, 80), num_classes=3, learning_rate=1e-3, weight_decay=1e-4, *, dtype="bf16",
                 device=None, seed=0, grad_sync=None, clip_norm=1.0, label_smoothing=0.1, num_layers=8, use_cuda_graph=True):
        if not torch.cuda.is_available():
            raise _lib.TbiError("ultrasound_modeling_b200 needs a CUDA device (sm_100); there is no CPU path")
        self.num_classes, self.batch_size, self.weight_decay = num_classes, batch_size, weight_decay
        self.input_shape = [img_size[0], img_size[1], 10]
        self.learning_rate = learning_rate
        self.optimizer = _Adam(learning_rate)
        self.alpha = 2
        self.class_factor = [0.06329, 0.027567, 0.90914]
        self.clip_norm, self.label_smoothing = float(clip_norm), float(label_smoothing)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.tdtype = torch.bfloat16 if dtype in ("bf16", torch.bfloat16) else torch.float32
        self.store = VariableStore(self.device, seed)
        self.transformer = Transformer(img_size, store=self.store, dtype=dtype, device=self.device, num_layers=num_layers)
        self.decoder = DecoderCup(num_classes, grid=(img_size[0] // 16, img_size[1] // 16), dtype=dtype, device=self.device, _store=self.store, _prefix=DEC)
        self.loss = self.compute_loss
        self.visionModel = _VisionModel(self)
        self.grad_sync = grad_sync
        # CUDA-graph replay of forward / loss+backward (the define-by-run tape issues ~2 600 launches per training step and is
        # host-bound without it).  Like TBI_ResNest.ResNest: results live in static buffers that the next call overwrites.
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graphs = {}
        self._hyper = torch.tensor([learning_rate, 1.0, self.clip_norm], dtype=torch.float32, device=self.device)
        self._hyper_host = None
        self._gnorm_sq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.L = _lib.lib()

    # ------------------------------------------------------------------ variables
    def variables(self):
        return OrderedDict(self.store.vars)

    def load_variables(self, variables):
        self.store.load(variables)

    def gradients(self):
        """variable name -> gradient of the last train_step / backward (before clipping)"""
        return OrderedDict(self.store.grads)

    # ------------------------------------------------------------------ forward / loss
    def _x(self, x):
        return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(device=self.device, dtype=self.tdtype).contiguous()

    def _y(self, y):
        return torch.as_tensor(np.asarray(y) if not torch.is_tensor(y) else y).to(device=self.device, dtype=torch.float32).contiguous()

    def _graphed(self, tag, fn, inputs):
        """fn(*device tensors) -> tensors, through a CUDA graph: per (entry point, input shapes, variable-storage generation)
        two eager calls (they create / freeze the variables), then one capture, then replays over static input buffers."""
        if not self.use_cuda_graph:
            return fn(*inputs)
        key = (tag, tuple(tuple(t.shape) for t in inputs), self.store.generation)
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs = {k: v for k, v in self._graphs.items() if k[2] == self.store.generation}     # stale pointers: drop
            ent = self._graphs[key] = dict(warm=0, static=[torch.empty_like(t) for t in inputs], graph=None, out=None)
        for st, t in zip(ent["static"], inputs):
            st.copy_(t, non_blocking=True)
        if ent["graph"] is None:
            if ent["warm"] < 2:
                ent["warm"] += 1
                return fn(*ent["static"])
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(g):
                out = fn(*ent["static"])
            ent["graph"], ent["out"] = g, out
        ent["graph"].replay()
        return ent["out"]

    def _forward_logits(self, x):
        tokens, weights, feats = self.transformer.forward(self._x(x))
        z = self.decoder.forward(tokens, feats, logits=True, record=False)
        return z, weights

    def forward_logits(self, x):
        """pre-softmax class scores [N,H,W,num_classes] and the attention weights (graph-replayed; the evaluator's entry)"""
        return self._graphed("logits", self._forward_logits, (self._x(x),))

    def _forward_dev(self, xd):
        z, weights = self._forward_logits(xd)
        probs, _, _ = ops.softmax_cce(z, torch.zeros_like(z), 0.0, 1.0, need_grad=False)
        return probs, weights

    def forward(self, x):
        """VisionTransformer.forward :220-223 -> (probabilities [N,H,W,num_classes], [attention weights per layer])"""
        return self._graphed("forward", self._forward_dev, (self._x(x),))

    def compute_loss(self, y_true, y_pred):
        """:225-227 on PROBABILITIES (as the reference calls it); evaluated by the fused kernel on log(p) (softmax(log p) == p)"""
        y_true = self._y(y_true)
        y_pred = self._y(y_pred)
        _, loss, _ = ops.softmax_cce(torch.log(y_pred.clamp_min(1e-30)), y_true, self.label_smoothing, float(self.batch_size), need_grad=False)
        return loss.reshape(())

    def _step_dev(self, xd, yd):
        z, _ = self._forward_logits(xd)
        probs, loss, _ = ops.softmax_cce(z, yd, self.label_smoothing, float(self.batch_size), need_grad=False)
        return loss.reshape(()), probs

    def step(self, x, y):
        """:248-254 -> (loss, probabilities)"""
        return self._graphed("step", self._step_dev, (self._x(x), self._y(y)))

    # ------------------------------------------------------------------ training step
    def _backward_dev(self, xd, yd):
        s = self.store
        s.start_recording()
        z, _ = self._forward_logits(xd)
        if s.flat is None:                                      # first step: every variable exists now
            s.freeze()
            s.flat["grads"].zero_()
        probs, loss, dz = ops.softmax_cce(z, yd, self.label_smoothing, float(self.batch_size))
        s.gacc(z, dz)
        s.run_backward()
        return loss.reshape(()), probs

    def backward(self, x, y):
        """forward + loss + backward (no optimizer): -> (loss, probabilities); gradients() holds d loss / d variable"""
        return self._graphed("backward", self._backward_dev, (self._x(x), self._y(y)))

    def train_step(self, x, y):
        """:235-246: gradients of the averaged loss -> clip_by_global_norm(1.0) -> Adam; returns (loss, probabilities).
        Under data parallelism (grad_sync) each replica clips ITS gradients, the clipped gradients are summed over the replicas
        (what apply_gradients does under MirroredStrategy, MainParallel.py:130) and Adam applies the sum: the loss is already
        divided by the GLOBAL batch, so no further scaling."""
        loss, probs = self.backward(x, y)
        f = self.store.flat
        st = torch.cuda.current_stream(self.device).cuda_stream
        lr = float(self.optimizer.learning_rate)
        n = f["params"].numel()
        if self.grad_sync is None:
            want = (lr, 1.0, self.clip_norm)
            if want != self._hyper_host:
                self._hyper.copy_(torch.tensor(want, dtype=torch.float32)); self._hyper_host = want
            self._gnorm_sq.zero_()
            _lib.check(self.L.tbi_sumsq(n, f["grads"].data_ptr(), self._gnorm_sq.data_ptr(), st), "sumsq")
            gn = self._gnorm_sq.data_ptr()
        else:
            self._gnorm_sq.zero_()
            _lib.check(self.L.tbi_sumsq(n, f["grads"].data_ptr(), self._gnorm_sq.data_ptr(), st), "sumsq")
            if self.clip_norm > 0:
                f["grads"].mul_(self.clip_norm / torch.clamp(self._gnorm_sq.sqrt(), min=self.clip_norm))
            self.grad_sync.allreduce(f["grads"])
            want = (lr, 1.0, 0.0)
            if want != self._hyper_host:
                self._hyper.copy_(torch.tensor(want, dtype=torch.float32)); self._hyper_host = want
            gn = None
        _lib.check(self.L.tbi_adam_multi_dev(n, f["params"].data_ptr(), f["grads"].data_ptr(), f["m"].data_ptr(), f["v"].data_ptr(),
                                             f["step"].data_ptr(), self._hyper.data_ptr(), gn, 0.9, 0.999, 1e-7, st), "adam")
        _lib.check(self.L.tbi_adam_advance(f["step"].data_ptr(), st), "adam_advance")
        return loss, probs

    def global_grad_norm(self):
        """||g|| of the last train_step before clipping (device sync)"""
        return float(self._gnorm_sq.sqrt().item())

    def __call__(self, x, *args, **kwargs):
        return self.forward(x)
