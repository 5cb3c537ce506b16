"""Variant B decoder: drop-in for the reference's ``Decoder.py`` (DecoderBlock :7-94, DecoderCup :98-146) on the B200 path.
Forward / inference only this round (see ResNest.py in this package for the conventions: lazy variables, Keras names and
layouts, NHWC in and out, every arithmetic step a libtbi_sm100.so entry point).

    DecoderCup(num_classes, wDecay=None)(hidden_states [N,T,hidden], features=[x_3,x_2,x_1]) -> probs [N,16*gh,16*gw,num_classes]

Differences from the reference, both opt-in: ``grid`` (the reference hard-codes the 16 x 5 token grid, Decoder.py:128,140) and
``logits=True`` on forward.  Nothing is concatenated in memory: the skip, and the raw reshape of the tokens the reference
appends after every block, enter the next convolution as a second source (virtual concat); the four dilated branches of a
block write their quarter of the output tensor directly.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import ops
from ._lib import ACT_LRELU, ACT_NONE
from .ResNest import VariableStore, _Layer

_DIL = (1, 2, 4, 8)          # conv*_0: 1x1; conv*_1..3: 3x3 dilated 2, 4, 8 (Decoder.py:11-25,35-49)


class DecoderBlock(_Layer):
    """Decoder.py:7-94: up (Conv2DTranspose k3 s2) -> [concat skip] -> 4 branches + BN -> concat -> LeakyReLU -> again."""

    def __init__(self, out_channels, wDecay=None, *, _store=None, _prefix="", device="cuda", seed=0):
        super().__init__(_store if _store is not None else VariableStore(device, seed), _prefix)
        self.out_channels, self.wDecay = out_channels, wDecay

    def _up(self, x, x2):
        cin = x.shape[3] + (x2.shape[3] if x2 is not None else 0)
        w = self._s.kernel(self._p + "up/kernel", (3, 3, self.out_channels, cin))
        b = self._s.vector(self._p + "up/bias", self.out_channels, 0.0)
        return ops.conv2d_transpose_s2(x, w, b, x2=x2)

    def _branches(self, half, x, x2):
        n, h, w, _ = x.shape
        q = self.out_channels // 4
        y = torch.empty(n, h, w, 4 * q, dtype=x.dtype, device=x.device)
        for j, d in enumerate(_DIL):
            self._conv(x, f"conv{half}_{j}", 1 if j == 0 else 3, q, dilation=1 if j == 0 else d, bn=f"bn{half}_{j}", act=ACT_LRELU,
                       x2=x2, out=y, out_coff=j * q)
        return y

    def forward(self, x, skip=None, x_extra=None):
        """x_extra: channels the caller would have concatenated to x (DecoderCup's token reshape), fed as a second source"""
        x = self._up(x, x_extra)
        x = self._branches(1, x, skip)
        return self._branches(2, x, None)

    def __call__(self, x, skip=None, *args, **kwargs):
        return self.forward(x, skip, *args, **kwargs)


class DecoderCup(_Layer):
    """Decoder.py:98-146."""

    def __init__(self, num_classes, wDecay=None, *, grid=(16, 5), dtype="bf16", device="cuda", seed=0):
        super().__init__(VariableStore(device, seed), "")
        self.num_classes, self.wDecay, self.grid = num_classes, wDecay, tuple(grid)
        self.tdtype = torch.bfloat16 if dtype in ("bf16", torch.bfloat16) else torch.float32
        self.device = torch.device(device)
        self.blocks = [DecoderBlock(c, wDecay, _store=self._s, _prefix=f"block_{i}/") for i, c in enumerate((256, 128, 64))]

    def load_variables(self, variables):
        self._s.load(variables)

    def variables(self):
        return OrderedDict(self._s.vars)

    def _dev(self, t):
        return torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t).to(device=self.device, dtype=self.tdtype).contiguous()

    def forward(self, hidden_states, features=None, logits=False):
        y = self._dev(hidden_states)
        n = y.shape[0]
        gh, gw = self.grid
        x = y.reshape(n, gh, gw, -1)
        x = self._conv(x, "conv_more", 3, 256)
        x = self._ln(x, "bn1")
        extra = None
        for i, blk in enumerate(self.blocks):
            skip = self._dev(features[i]) if (features is not None and i < 3) else None
            x = blk.forward(x, skip, extra)
            extra = y.reshape(n, gh * 2 ** (i + 1), gw * 2 ** (i + 1), -1)      # Decoder.py:140-141, consumed by the next layer
        cin = x.shape[3] + extra.shape[3]
        w = self._s.kernel("head/kernel", (3, 3, self.num_classes, cin))
        b = self._s.vector("head/bias", self.num_classes, 0.0)
        z = ops.conv2d_transpose_s2(x, w, b, x2=extra, out_f32=True)
        if logits:
            return z
        probs, _, _, _ = ops.softmax_loss(z, torch.zeros_like(z))
        return probs

    def __call__(self, hidden_states, features=None, *args, **kwargs):
        return self.forward(hidden_states, features, *args, **kwargs)
