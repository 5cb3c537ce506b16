"""Variant B decoder: drop-in for the reference's ``Decoder.py`` (DecoderBlock :7-94, DecoderCup :98-146) on the B200 path.
Forward, and backward through the tape of ResNest.py in this package (``forward(..., record=True)`` then
``backward(dlogits)`` -> (dL/dhidden_states, [dL/dskips]); parameter gradients in ``gradients()``).  Conventions as in
ResNest.py: lazy variables, Keras names and layouts, NHWC in and out, every arithmetic step a libtbi_sm100.so entry point.

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


def _convt(layer, x, x2, name, cout, out_f32=False):
    """Conv2DTranspose k3 s2 'same' over the virtual concat (x, x2), no norm / activation (Decoder.py:57-59,120); recorded for
    backward when the store is recording.  bf16 storage: a second source whose width is not a multiple of 16 (the 8 token
    channels in front of the head, Decoder.py:140-141) is copied into zero-padded 16-channel records and the matching kernel
    rows are zero, and a gradient narrower than 8 channels (the 3-class head) is widened to 16-channel records, so that
    forward, data gradient and weight gradient all stay on the tensor-core path."""
    s = layer._s
    c0, c2 = x.shape[3], (x2.shape[3] if x2 is not None else 0)
    w = s.kernel(name + "/kernel", (3, 3, cout, c0 + c2))
    b = s.vector(name + "/bias", cout, 0.0)
    c2p = ops.pad_channels(c2, x.dtype) if x2 is not None else 0
    x2p, wk = x2, w
    if c2p != c2:
        x2p = torch.zeros(*x2.shape[:3], c2p, dtype=x2.dtype, device=x2.device)
        x2p[..., :c2] = x2
        wk = torch.zeros(3, 3, cout, c0 + c2p, dtype=torch.float32, device=w.device)
        wk[..., :c0 + c2] = w
    y = ops.conv2d_transpose_s2(x, wk, b, x2=x2p, out_f32=out_f32)
    if s.recording:
        def bwd():
            dz = s.gget(y)
            if dz is None:
                return
            dz = dz.to(x.dtype)
            if x.dtype == torch.bfloat16 and cout % 8 != 0:
                dzp = torch.zeros(*dz.shape[:3], ops.pad_channels(cout, x.dtype), dtype=dz.dtype, device=dz.device)
                dzp[..., :cout] = dz
                dz = dzp
            dxs, dw, db = ops.conv2d_transpose_s2_grads(x, wk, dz, x2=x2p)
            s.pacc(name + "/kernel", dw[..., :c0 + c2].contiguous()); s.pacc(name + "/bias", db)
            if x2 is None:
                s.gacc(x, dxs)
            else:
                s.gacc(x, dxs[0]); s.gacc(x2, dxs[1][..., :c2].contiguous().reshape(x2.shape))
        s.tape.append(bwd)
    return y


class DecoderBlock(_Layer):
    """Decoder.py:7-94: up (Conv2DTranspose k3 s2) -> [concat skip] -> 4 branches + BN -> concat -> LeakyReLU -> again."""

    def __init__(self, out_channels, wDecay=None, *, _store=None, _prefix="", device="cuda", seed=0):
        super().__init__(_store if _store is not None else VariableStore(device, seed), _prefix)
        self.out_channels, self.wDecay = out_channels, wDecay

    def _up(self, x, x2):
        return _convt(self, x, x2, self._p + "up", self.out_channels)

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

    def __init__(self, num_classes, wDecay=None, *, grid=(16, 5), dtype="bf16", device="cuda", seed=0, _store=None, _prefix=""):
        super().__init__(_store if _store is not None else VariableStore(device, seed), _prefix)
        self.num_classes, self.wDecay, self.grid = num_classes, wDecay, tuple(grid)
        self.tdtype = torch.bfloat16 if dtype in ("bf16", torch.bfloat16) else torch.float32
        self.device = torch.device(device)
        self.blocks = [DecoderBlock(c, wDecay, _store=self._s, _prefix=f"{_prefix}block_{i}/") for i, c in enumerate((256, 128, 64))]

    def load_variables(self, variables):
        self._s.load(variables)

    def variables(self):
        return OrderedDict(self._s.vars)

    def _dev(self, t):
        return torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t).to(device=self.device, dtype=self.tdtype).contiguous()

    def forward(self, hidden_states, features=None, logits=False, record=False):
        """record=True keeps what backward() needs (the tape of this call)"""
        y = self._dev(hidden_states)
        n = y.shape[0]
        gh, gw = self.grid
        s = self._s
        if record:
            s.start_recording()
        views = []                                                                  # raw reshapes of the token tensor

        def token_view(*shape):
            v = y.reshape(*shape)
            views.append(v)
            return v

        if s.recording:
            def gather():                                                           # first on the tape = last to run
                g = None
                for v in views:
                    gv = s.gget(v)
                    if gv is not None:
                        g = gv.reshape(y.shape) if g is None else g + gv.reshape(y.shape)
                self._dhidden = g
                if g is not None:
                    s.gacc(y, g)                                                    # for a producer of the tokens on the same tape (the ViT bridge)
            s.tape.append(gather)
        x = token_view(n, gh, gw, -1)
        x = self._conv(x, "conv_more", 3, 256)
        x = self._ln(x, "bn1")
        extra = None
        skips = []
        for i, blk in enumerate(self.blocks):
            skip = self._dev(features[i]) if (features is not None and i < 3) else None
            skips.append(skip)
            x = blk.forward(x, skip, extra)
            extra = token_view(n, gh * 2 ** (i + 1), gw * 2 ** (i + 1), -1)         # Decoder.py:140-141, consumed by the next layer
        z = _convt(self, x, extra, self._p + "head", self.num_classes, out_f32=True)
        self._io = (z, skips)
        if logits:
            return z
        probs, _, _, _ = ops.softmax_loss(z, torch.zeros_like(z))
        return probs

    def backward(self, dlogits):
        """dL/dlogits of the last forward(..., record=True) -> (dL/dhidden_states, [dL/dfeatures]); parameter gradients in
        gradients().  (The reference's loss is applied to the probabilities by the caller, VisionTransformer.py:205-206,225-254.)"""
        if not self._s.recording:
            raise RuntimeError("DecoderCup.backward: call forward(..., record=True) first")
        z, skips = self._io
        self._s.gacc(z, torch.as_tensor(dlogits).to(device=self.device, dtype=z.dtype).contiguous())
        grads_of = [None if t is None else id(t) for t in skips]
        tg = self._s.tgrads
        self._s.run_backward()
        return self._dhidden, [None if k is None else tg.get(k) for k in grads_of]

    def gradients(self):
        """variable name -> fp32 gradient of the last backward() (trainable variables only: no moving statistics)"""
        return OrderedDict(self._s.grads)

    def __call__(self, hidden_states, features=None, *args, **kwargs):
        return self.forward(hidden_states, features, *args, **kwargs)
