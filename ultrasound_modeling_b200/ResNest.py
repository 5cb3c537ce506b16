"""Variant B encoder: drop-in for the reference's ``ResNest.py`` (classes ResNest, residual_S, cardinal, split_attention)
on the B200 path.  Forward / inference only this round (the training step of Variant B runs through the ViT bridge of
``VisionTransformer.py``, scope row 8(f)-1).

Same constructor signatures and call results as the reference:
    ResNest(height, width, channel, ksize, radix=4, kpaths=4, wDecay=None)(x) -> (x_4, [x_3, x_2, x_1])   ResNest.py:7,38-58
Inputs are NHWC (numpy or torch); outputs are NHWC device tensors in the storage dtype (bf16 default, fp32 for parity runs).
Every arithmetic step is a libtbi_sm100.so entry point (ops.py); torch only allocates, reshapes and owns the parameters.
``wDecay`` (a Keras kernel regulariser) only adds a term to the training loss and is ignored here.

Keras builds layers lazily on the first call (input channel counts are not constructor arguments); so does this file: a
kernel is created (HeNormal, ResNest.py:15) the first time its layer runs unless ``load_variables`` supplied it.  Variable
names: ``<attribute path>/<kernel|bias|gamma|beta|moving_mean|moving_variance>`` with the reference's attribute names
(``initial_conv``, ``convtmp_1``, ``conv_1/cardinal_0/conv1``, ``conv_1/cardinal_0/split/dense1`` ...), Keras layouts (HWIO).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from ._lib import ACT_LRELU, ACT_NONE


class VariableStore:
    """name -> fp32 device tensor, created on first use like a Keras layer's build()."""

    def __init__(self, device, seed: int = 0):
        self.device = torch.device(device)
        self.vars: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.gen = torch.Generator().manual_seed(seed)

    def kernel(self, name: str, shape) -> torch.Tensor:
        t = self.vars.get(name)
        if t is None:
            kh, kw, a, _ = shape
            sigma = math.sqrt(2.0 / (kh * kw * a)) / 0.87962566103423978           # HeNormal (Keras fan_in = rf * shape[-2])
            t = torch.empty(shape, dtype=torch.float32)
            torch.nn.init.trunc_normal_(t, 0.0, sigma, -2 * sigma, 2 * sigma, generator=self.gen)
            t = self.vars[name] = t.to(self.device)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: stored shape {tuple(t.shape)} != layer shape {tuple(shape)}")
        return t

    def vector(self, name: str, c: int, value: float) -> torch.Tensor:
        t = self.vars.get(name)
        if t is None:
            t = self.vars[name] = torch.full((c,), value, dtype=torch.float32, device=self.device)
        if t.numel() != c:
            raise ValueError(f"{name}: stored length {t.numel()} != {c}")
        return t

    def load(self, variables: Dict[str, torch.Tensor]):
        for k, v in variables.items():
            self.vars[k] = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).detach().to(device=self.device, dtype=torch.float32).contiguous()


class _Layer:
    def __init__(self, store: VariableStore, prefix: str):
        self._s, self._p = store, prefix

    # -- Keras layer equivalents on NHWC device tensors --------------------------------------
    def _conv(self, x, name, k, cout, *, dilation=1, bn=None, act=ACT_NONE, residual=None, x2=None, out=None, out_coff=0, out_f32=False):
        cin = x.shape[3] + (x2.shape[3] if x2 is not None else 0)
        w = self._s.kernel(f"{self._p}{name}/kernel", (k, k, cin, cout))
        b = self._s.vector(f"{self._p}{name}/bias", cout, 0.0)
        bnp = None
        if bn is not None:
            q = f"{self._p}{bn}/"
            bnp = (self._s.vector(q + "gamma", cout, 1.0), self._s.vector(q + "beta", cout, 0.0),
                   self._s.vector(q + "moving_mean", cout, 0.0), self._s.vector(q + "moving_variance", cout, 1.0))
        return ops.conv2d(x, w, b, dilation=dilation, bn=bnp, act=act, residual=residual, x2=x2, out=out, out_coff=out_coff, out_f32=out_f32)

    def _ln(self, x, name, act=ACT_LRELU, coff=0, c=None):
        cc = x.shape[3] if c is None else c
        g = self._s.vector(f"{self._p}{name}/gamma", cc, 1.0)
        b = self._s.vector(f"{self._p}{name}/beta", cc, 0.0)
        return ops.layernorm_c(x, g, b, act=act, inplace=True, coff=coff, c=c)


class split_attention(_Layer):
    """ResNest.py:153-202.  The R inputs of the reference are identical tensors (cardinal.forward builds them with the same
    layers) and dense2 is shared, so one tensor U is enough: V = R * U * a (tbi_splitatt_shared_fwd)."""

    def __init__(self, inchannel, radix, atrous=1, wDecay=None, *, _store=None, _prefix=""):
        super().__init__(_store, _prefix)
        self.inchannel, self.radix, self.atrous, self.wDecay = inchannel, radix, atrous, wDecay

    def params(self):
        c, c2, s, p = self.inchannel, self.inchannel // 2, self._s, self._p
        return (s.kernel(p + "dense1/kernel", (1, 1, c, c2)).reshape(c, c2), s.vector(p + "dense1/bias", c2, 0.0),
                s.vector(p + "dense1_bn/gamma", c2, 1.0), s.vector(p + "dense1_bn/beta", c2, 0.0),
                s.kernel(p + "dense2/kernel", (1, 1, c2, c)).reshape(c2, c), s.vector(p + "dense2/bias", c, 0.0))


class cardinal(_Layer):
    """ResNest.py:107-150: conv1x1 -> LN -> LeakyReLU -> conv kxk -> LN -> LeakyReLU (once: the R repetitions are identical)."""

    def __init__(self, ksize, outchannel, radix, kpaths, atrous=1, wDecay=None, *, _store=None, _prefix=""):
        super().__init__(_store, _prefix)
        self.ksize, self.outchannel, self.radix, self.kpaths, self.atrous, self.wDecay = ksize, outchannel, radix, kpaths, atrous, wDecay
        self.cv11 = int(outchannel / radix / kpaths)
        self.cvkk = int(outchannel / kpaths)
        self.split = split_attention(self.cvkk, radix, atrous, wDecay, _store=_store, _prefix=_prefix + "split/")

    def features(self, x, out, out_coff):
        """U of this cardinal written into channels [out_coff, out_coff+cvkk) of ``out``"""
        y = self._conv(x, "conv1", 1, self.cv11, dilation=self.atrous)
        y = self._ln(y, "conv1_bn")
        self._conv(y, "conv2", self.ksize, self.cvkk, dilation=self.atrous, out=out, out_coff=out_coff)
        self._ln(out, "conv2_bn", coff=out_coff, c=self.cvkk)


class residual_S(_Layer):
    """ResNest.py:61-104."""

    def __init__(self, ksize, outchannel, radix, kpaths, atrous=1, wDecay=None, *, _store=None, _prefix=""):
        super().__init__(_store, _prefix)
        self.ksize, self.outchannel, self.radix, self.kpaths, self.atrous, self.wDecay = ksize, outchannel, radix, kpaths, atrous, wDecay
        self.cardinal_blocks = [cardinal(ksize, outchannel // 2, radix, kpaths, atrous, wDecay, _store=_store, _prefix=f"{_prefix}cardinal_{k}/")
                                for k in range(kpaths)]

    def forward(self, x):
        n, h, w, _ = x.shape
        K, c = self.kpaths, self.cardinal_blocks[0].cvkk
        u = torch.empty(n, h, w, K * c, dtype=x.dtype, device=x.device)            # the K cardinals write their slice: no concat pass
        for k, blk in enumerate(self.cardinal_blocks):
            blk.features(x, u, k * c)
        per = [blk.split.params() for blk in self.cardinal_blocks]
        stacked = [torch.stack([p[i] for p in per]).contiguous() for i in range(6)]
        v = ops.splitatt_shared(u, K, self.radix, *stacked, act=ACT_LRELU)
        sc = self._conv(x, "convtmp_sc", 1, self.outchannel, dilation=self.atrous)
        sc = self._ln(sc, "convtmp_scbn")
        return self._conv(v, "concats_2", self.ksize, self.outchannel, dilation=self.atrous, residual=sc)

    __call__ = forward


class ResNest(_Layer):
    """ResNest.py:4-58."""

    def __init__(self, height, width, channel, ksize, radix=4, kpaths=4, wDecay=None, *, dtype="bf16", device="cuda", seed=0):
        store = VariableStore(device, seed)
        super().__init__(store, "")
        self.height, self.width, self.channel, self.ksize, self.radix, self.kpaths, self.wDecay = height, width, channel, ksize, radix, kpaths, wDecay
        self.tdtype = torch.bfloat16 if dtype in ("bf16", torch.bfloat16) else torch.float32
        self.device = torch.device(device)
        mk = lambda name, out: residual_S(ksize=ksize, outchannel=out, radix=radix, kpaths=kpaths, wDecay=wDecay, _store=store, _prefix=name + "/")
        self.conv_1, self.conv_2, self.conv_3, self.conv_4 = mk("conv_1", 64), mk("conv_2", 128), mk("conv_3", 256), mk("conv_4", 512)

    def load_variables(self, variables):
        self._s.load(variables)

    def variables(self):
        return OrderedDict(self._s.vars)

    def forward(self, x):
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(device=self.device, dtype=self.tdtype).contiguous()
        x = self._conv(x, "initial_conv", 3, 16, act=ACT_LRELU)
        x = self._conv(x, "convtmp_1", 3, 32, bn="convtmp_1bn", act=ACT_LRELU)
        x = self._conv(x, "convtmp_2", 3, 32, bn="convtmp_2bn", act=ACT_LRELU)
        x_1 = self.conv_1(ops.avgpool2x2(x))
        x_2 = self.conv_2(ops.avgpool2x2(x_1))
        x_3 = self.conv_3(ops.avgpool2x2(x_2))
        x_4 = self.conv_4(ops.avgpool2x2(x_3))
        return x_4, [x_3, x_2, x_1]

    def __call__(self, x, *args, **kwargs):
        return self.forward(x)
