"""CPU oracle for the ViT bridge and the Variant B training step: ``VisionTransformer.py`` of the reference.

TEST INFRASTRUCTURE ONLY (same rule as the other oracle files): only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs may import it.  PINNED AGAINST THE REFERENCE'S OWN CODE (not TensorFlow's binaries): the unmodified ``/root/reference/VisionTransformer.py`` (with
``ResNest.py`` and ``Decoder.py``) runs under the stand-in ``tensorflow`` package ``oracle/tfshim``; ``tests/golden/make_golden_ref.py``
recorded its ``train_step`` x 2, ``step`` and forward at [1,256,80,10] (``tests/golden/ref_vit_256x80.npz``) and
``tests/test_oracle_pinned.py`` holds this file to them in float64 at 1e-9: the 494-variable inventory by attribute path, loss,
probabilities, every gradient, every variable after clip_by_global_norm + Adam, the attention weights.  The shim's primitives are
restated from TF/Keras' documented definitions; that layer of arithmetic is what no TensorFlow binary has confirmed.

What it follows (reference file:line, all in VisionTransformer.py):
  * ``Attention.forward``  :33-54   query/key/value Dense(512) -> split into 4 heads of 128 (:26-31) -> scores = q k^T
                                    **/ sqrt(num_heads)** (:42: the reference divides by sqrt(4), not sqrt(128)) -> softmax over
                                    keys (:43) -> context = probs v -> merge heads -> out Dense(512); dropout rates are 0.
  * ``Mlp.forward``        :67-73   fc1 Dense(2048) -> exact GELU (tf.keras.activations.gelu, approximate=False) -> fc2 Dense(512)
  * ``Embeddings.forward`` :112-120 ResNest(H, W, 10, radix=3, ksize=3, kpaths=3) (:100) -> 1x1 Conv2D(512) "patch_embeddings"
                                    (:106-107) -> raw reshape to [N, seq_len, 512] (:116) -> + position_embeddings, which is
                                    the CONSTANT tf.zeros (:108), not a variable.
  * ``Block.forward``      :136-147 x + attn(LN(x)); x + ffn(LN(x)); LayerNormalization epsilon 1e-6 (:130-131)
  * ``Encoder.forward``    :163-170 8 blocks, then encoder_norm (LN, eps 1e-6); returns the per-layer attention probabilities
  * ``VisionTransformer.forward`` :220-223  transformer -> DecoderCup(num_classes) on (tokens, features) -> probabilities
  * ``compute_loss``       :225-227 CategoricalCrossentropy(label_smoothing=0.1, reduction=NONE) (:205-206) on the
                                    PROBABILITIES, then tf.nn.compute_average_loss(global_batch_size=batch_size): sum over
                                    every pixel of every image / batch_size.  Keras' CCE on probabilities: y <- y(1-ls) + ls/C;
                                    p <- p / sum(p); p <- clip(p, 1e-7, 1-1e-7); loss = -sum_c y log p.
  * ``train_step``         :235-246 gradients of that scalar -> tf.clip_by_global_norm(., 1.0) (:244: g * 1 / max(||g||, 1))
                                    -> Keras Adam (:204, defaults beta 0.9 / 0.999, eps 1e-7).
Variable names: the attribute paths of the reference objects (``transformer/embeddings/hybrid_model/...`` = the ResNest encoder,
``transformer/embeddings/patch_embeddings``, ``transformer/encoder/layer_{i}/{attention_norm,attn/{query,key,value,out},
ffn_norm,ffn/{fc1,fc2}}``, ``transformer/encoder/encoder_norm``, ``decoder/...`` = DecoderCup); Dense kernels are stored as
1x1 HWIO conv kernels [1,1,in,out] (a Dense over the last axis is exactly that).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import resnest_decoder_oracle as B
from .tbi_resnest_oracle import conv2d_same

VIT_LN_EPS = 1e-6
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-7
ENC = "transformer/embeddings/hybrid_model/"
DEC = "decoder/"
TR = "transformer/encoder/"


def vit_param_shapes(hidden: int = 512, mlp_dim: int = 2048, num_layers: int = 8) -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def dense(name, cin, cout):
        s[name + "/kernel"] = (1, 1, cin, cout); s[name + "/bias"] = (cout,)

    def ln(name, c):
        s[name + "/gamma"] = (c,); s[name + "/beta"] = (c,)

    dense("transformer/embeddings/patch_embeddings", hidden, hidden)
    for i in range(num_layers):
        p = f"{TR}layer_{i}/"
        ln(p + "attention_norm", hidden)
        for nm in ("query", "key", "value", "out"):
            dense(p + "attn/" + nm, hidden, hidden)
        ln(p + "ffn_norm", hidden)
        dense(p + "ffn/fc1", hidden, mlp_dim); dense(p + "ffn/fc2", mlp_dim, hidden)
    ln(TR + "encoder_norm", hidden)
    return s


def model_param_shapes(num_classes: int = 3, channel: int = 10, grid: Tuple[int, int] = (16, 5), hidden: int = 512, mlp_dim: int = 2048,
                       num_layers: int = 8) -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for k, v in B.encoder_param_shapes(channel, 3, 3, 3).items():
        s[ENC + k] = v
    s.update(vit_param_shapes(hidden, mlp_dim, num_layers))
    for k, v in B.decoder_param_shapes(num_classes, hidden=hidden, grid=grid).items():
        s[DEC + k] = v
    return s


def is_trainable(name: str) -> bool:
    return not (name.endswith("/moving_mean") or name.endswith("/moving_variance"))


def init_params(shapes, seed: int = 2240, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """conv kernels HeNormal / perturbed norm parameters as in the Variant B oracle; the Dense and patch-embedding kernels
    Keras' default glorot_uniform scaled so that 8 residual blocks keep O(1) activations"""
    out = B.init_params(shapes, seed=seed, perturb=True, dtype=torch.float64)
    g = torch.Generator().manual_seed(seed + 1)
    for name, shp in shapes.items():
        if name.endswith("/kernel") and ("/attn/" in name or "/ffn/" in name or "patch_embeddings" in name):
            lim = math.sqrt(6.0 / (shp[2] + shp[3]))
            out[name] = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * lim
    return OrderedDict((k, v.to(dtype)) for k, v in out.items())


def layernorm(x, gamma, beta, eps=VIT_LN_EPS):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * gamma + beta


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def cce_label_smoothing(y_true, probs, label_smoothing: float = 0.1):
    """tf.keras.losses.CategoricalCrossentropy(label_smoothing, reduction=NONE) on probabilities -> per-pixel loss"""
    c = y_true.shape[-1]
    y = y_true * (1.0 - label_smoothing) + label_smoothing / c
    p = probs / probs.sum(-1, keepdim=True)
    p = p.clamp(1e-7, 1.0 - 1e-7)
    return -(y * torch.log(p)).sum(-1)


class VisionTransformerOracle:
    def __init__(self, batch_size, img_size=(256, 80), num_classes=3, learning_rate=1e-3, params: Optional[Dict[str, torch.Tensor]] = None,
                 dtype=torch.float64, num_heads=4, hidden=512, mlp_dim=2048, num_layers=8, seed: int = 2240):
        self.batch_size, self.img_size, self.num_classes, self.learning_rate = batch_size, tuple(img_size), num_classes, learning_rate
        self.dtype, self.num_heads, self.hidden, self.num_layers = dtype, num_heads, hidden, num_layers
        self.grid = (img_size[0] // 16, img_size[1] // 16)           # the encoder's x_4 grid (16 x 5 for 256 x 80)
        shapes = model_param_shapes(num_classes, 10, self.grid, hidden, mlp_dim, num_layers)
        src = params if params is not None else init_params(shapes, seed=seed, dtype=dtype)
        self.params: "OrderedDict[str, torch.Tensor]" = OrderedDict(
            (k, src[k].detach().clone().to(dtype).requires_grad_(is_trainable(k))) for k in shapes)
        self.adam_m = {k: torch.zeros_like(v) for k, v in self.params.items() if is_trainable(k)}
        self.adam_v = {k: torch.zeros_like(v) for k, v in self.params.items() if is_trainable(k)}
        self.adam_t = 0

    def _sub(self, prefix):
        return {k[len(prefix):]: v for k, v in self.params.items() if k.startswith(prefix)}

    def _dense(self, x, name):
        p = self.params
        return x @ p[name + "/kernel"][0, 0] + p[name + "/bias"]

    def attention(self, x, p):
        n, t, c = x.shape
        h, d = self.num_heads, c // self.num_heads
        split = lambda z: z.reshape(n, t, h, d).permute(0, 2, 1, 3)
        q, k, v = (split(self._dense(x, p + nm)) for nm in ("query", "key", "value"))
        scores = q @ k.transpose(-1, -2) / math.sqrt(float(h))                   # VisionTransformer.py:42
        probs = torch.softmax(scores, dim=3)
        ctx = (probs @ v).permute(0, 2, 1, 3).reshape(n, t, c)
        return self._dense(ctx, p + "out"), probs

    def transformer(self, x):
        p = self.params
        enc = B.ResNestEncoderOracle(10, 3, 3, 3, {}, dtype=self.dtype)
        enc.p = self._sub(ENC)                                                   # the live (differentiable) tensors, not detached copies
        x4, feats = enc(x.to(self.dtype))
        t = conv2d_same(x4, p["transformer/embeddings/patch_embeddings/kernel"], p["transformer/embeddings/patch_embeddings/bias"])
        h = t.reshape(t.shape[0], -1, self.hidden)                              # + position_embeddings == 0
        weights = []
        for i in range(self.num_layers):
            q = f"{TR}layer_{i}/"
            a, w = self.attention(layernorm(h, p[q + "attention_norm/gamma"], p[q + "attention_norm/beta"]), q + "attn/")
            weights.append(w)
            h = h + a
            m = layernorm(h, p[q + "ffn_norm/gamma"], p[q + "ffn_norm/beta"])
            h = h + self._dense(gelu(self._dense(m, q + "ffn/fc1")), q + "ffn/fc2")
        return layernorm(h, p[TR + "encoder_norm/gamma"], p[TR + "encoder_norm/beta"]), weights, feats

    def forward(self, x, logits: bool = False):
        tokens, weights, feats = self.transformer(x)
        dec = B.DecoderCupOracle(self.num_classes, {}, grid=self.grid, dtype=self.dtype)
        dec.p = self._sub(DEC)
        return dec(tokens, feats, logits=logits), weights

    def compute_loss(self, y_true, y_pred):
        return cce_label_smoothing(y_true.to(self.dtype), y_pred).sum() / float(self.batch_size)

    def gradients(self, x, y):
        probs, _ = self.forward(x)
        loss = self.compute_loss(y, probs)
        names = [k for k in self.params if is_trainable(k)]
        gs = torch.autograd.grad(loss, [self.params[k] for k in names])
        return loss.detach(), probs.detach(), dict(zip(names, gs))

    def step(self, x, y):
        with torch.no_grad():
            probs, _ = self.forward(x)
            return self.compute_loss(y, probs), probs

    def train_step(self, x, y, clip_norm: float = 1.0):
        loss, probs, grads = self.gradients(x, y)
        gnorm = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
        scale = clip_norm / max(gnorm, clip_norm)
        self.adam_t += 1
        t = self.adam_t
        lr_t = self.learning_rate * math.sqrt(1 - ADAM_B2 ** t) / (1 - ADAM_B1 ** t)
        with torch.no_grad():
            for k, g in grads.items():
                g = g * scale
                m, v = self.adam_m[k], self.adam_v[k]
                m.mul_(ADAM_B1).add_(g, alpha=1 - ADAM_B1)
                v.mul_(ADAM_B2).addcmul_(g, g, value=1 - ADAM_B2)
                self.params[k].sub_(lr_t * m / (v.sqrt() + ADAM_EPS))
        self.last_gnorm = gnorm
        return loss, probs

    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self.params.items())


def synthetic_labels(n: int, h: int = 256, w: int = 80, num_class: int = 3, seed: int = 2241, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    planes = torch.randn(n, num_class, max(h // 8, 1), max(w // 8, 1), generator=g, dtype=torch.float64)
    planes = F.interpolate(planes, size=(h, w), mode="bilinear", align_corners=False)
    return F.one_hot(planes.argmax(1), num_class).to(dtype)
