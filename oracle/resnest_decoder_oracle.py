"""CPU oracle for Variant B of the hot path: ``ResNest.py`` (encoder) + ``Decoder.py`` (DecoderCup/DecoderBlock).

TEST INFRASTRUCTURE ONLY (same rule as ``tbi_resnest_oracle.py``): only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import it.

PINNED AGAINST THE REFERENCE'S OWN CODE (not TensorFlow's binaries): the unmodified ``/root/reference/ResNest.py`` and
``Decoder.py`` run under ``oracle/tfshim`` as parts of ``VisionTransformer.py`` (``tests/golden/make_golden_ref.py`` ->
``tests/golden/ref_vit_256x80.npz``); ``tests/test_oracle_pinned.py`` holds ``vit_oracle.py`` -- which calls the two classes
below for the encoder and the decoder -- to what that code returned at [1,256,80,10] in float64 (probabilities, loss, every
gradient of two training steps, every variable after them) to 1e-9.  The shim's primitives are restated from TF/Keras'
documented definitions; that layer of arithmetic is the part no TensorFlow binary has confirmed.
This file restates the two Keras modules with plain PyTorch CPU ops (fp32/fp64); autograd supplies the gradients.

What it follows (reference file:line):
  * ``ResNest.forward``      ResNest.py:38-55   stem (LeakyReLU, BN on conv 2 and 3), 4 x residual_S with pools between;
                                               returns (x_4, [x_3, x_2, x_1]) -- the skips are the PRE-pool stage outputs.
  * ``residual_S.forward``   ResNest.py:89-104  K cardinals -> concat -> concats_2 (no norm/act); shortcut conv1x1 -> LN ->
                                               LeakyReLU ALWAYS; add.
  * ``cardinal.forward``     ResNest.py:136-147 the SAME conv1/conv2 applied R times => R identical tensors.
  * ``split_attention``      ResNest.py:171-199 S = sum of the R inputs; GAP; dense1 -> LN -> LeakyReLU; ONE dense2 applied R
                                               times; softmax over channels (sigmoid if R == 1); V = sum_r U_r * a.
  * ``DecoderBlock.forward`` Decoder.py:61-91   up (convT k3 s2) -> concat skip -> {1x1, 3x3 d2, d4, d8} each + BN -> concat ->
                                               LeakyReLU -> the same again.
  * ``DecoderCup.forward``   Decoder.py:124-143 tokens [N,T,hidden] -> [N,gh,gw,-1] -> conv3x3 -> LN -> LeakyReLU -> 3 blocks,
                                               after each a RAW RESHAPE of the token tensor is concatenated -> head convT k3
                                               s2 + softmax.  The reference hard-codes the grid 16 x 5 (:128,140); here it is
                                               a parameter with that default.
Keras semantics (SURVEY 8a tail): LayerNormalization axis -1, eps 1e-3, biased variance; LeakyReLU slope 0.3;
BatchNormalization in inference mode eps 1e-3; HeNormal = truncated normal, sigma = sqrt(2/fan_in)/0.8796.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .tbi_resnest_oracle import avgpool2, batchnorm_inference, conv2d_same, conv2d_transpose_s2_same

LN_EPS = 1e-3
ENC_STAGES = (("conv_1", 64), ("conv_2", 128), ("conv_3", 256), ("conv_4", 512))
SKIP_CHANNELS = (256, 128, 64)
DILATIONS = (1, 2, 4, 8)           # conv*_0 is 1x1, conv*_1..3 are 3x3 with these dilation rates (Decoder.py:11-25)


def leaky(x):
    return F.leaky_relu(x, 0.3)


def layernorm_c(x, gamma, beta, eps: float = LN_EPS):
    mean = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta


# --------------------------------------------------------------------------------------
# parameter inventory
# --------------------------------------------------------------------------------------
def cardinal_channels(stage_out: int, radix: int, kpaths: int) -> Tuple[int, int]:
    oc = stage_out // 2                                  # ResNest.py:72 passes outchannel // 2
    return int(oc / radix / kpaths), int(oc / kpaths)    # ResNest.py:120-121


def encoder_param_shapes(channel: int, ksize: int, radix: int, kpaths: int) -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def conv(name, k, cin, cout):
        s[name + "/kernel"] = (k, k, cin, cout); s[name + "/bias"] = (cout,)

    def bn(name, c):
        for f in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"{name}/{f}"] = (c,)

    def ln(name, c):
        s[name + "/gamma"] = (c,); s[name + "/beta"] = (c,)

    conv("initial_conv", 3, channel, 16)
    conv("convtmp_1", 3, 16, 32); bn("convtmp_1bn", 32)
    conv("convtmp_2", 3, 32, 32); bn("convtmp_2bn", 32)
    cin = 32
    for stage, out in ENC_STAGES:
        cv11, cvkk = cardinal_channels(out, radix, kpaths)
        for k in range(kpaths):
            p = f"{stage}/cardinal_{k}"
            conv(p + "/conv1", 1, cin, cv11); ln(p + "/conv1_bn", cv11)
            conv(p + "/conv2", ksize, cv11, cvkk); ln(p + "/conv2_bn", cvkk)
            conv(p + "/split/dense1", 1, cvkk, cvkk // 2); ln(p + "/split/dense1_bn", cvkk // 2)
            conv(p + "/split/dense2", 1, cvkk // 2, cvkk)
        conv(stage + "/concats_2", ksize, kpaths * cvkk, out)
        conv(stage + "/convtmp_sc", 1, cin, out); ln(stage + "/convtmp_scbn", out)
        cin = out
    return s


def decoder_param_shapes(num_classes: int, hidden: int = 512, grid: Tuple[int, int] = (16, 5), tokens: Optional[int] = None,
                         skip_in: Sequence[int] = (256, 128, 64)) -> "OrderedDict[str, Tuple[int, ...]]":
    """skip_in: channels of features[0..2] (= x_3, x_2, x_1 of the encoder)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    tokens = grid[0] * grid[1] if tokens is None else tokens
    s["conv_more/kernel"] = (3, 3, tokens * hidden // (grid[0] * grid[1]), 256); s["conv_more/bias"] = (256,)
    s["bn1/gamma"] = (256,); s["bn1/beta"] = (256,)
    cin = 256
    for i, out in enumerate(SKIP_CHANNELS):
        p = f"block_{i}"
        s[p + "/up/kernel"] = (3, 3, out, cin); s[p + "/up/bias"] = (out,)
        ccat = out + skip_in[i]
        for half, cc in ((1, ccat), (2, out)):
            for j in range(4):
                k = 1 if j == 0 else 3
                s[f"{p}/conv{half}_{j}/kernel"] = (k, k, cc, out // 4); s[f"{p}/conv{half}_{j}/bias"] = (out // 4,)
                for f in ("gamma", "beta", "moving_mean", "moving_variance"):
                    s[f"{p}/bn{half}_{j}/{f}"] = (out // 4,)
        extra = tokens * hidden // (grid[0] * grid[1] * 4 ** (i + 1))      # raw reshape of the tokens (Decoder.py:140)
        cin = out + extra
    s["head/kernel"] = (3, 3, num_classes, cin); s["head/bias"] = (num_classes,)
    return s


def init_params(shapes: "OrderedDict[str, Tuple[int, ...]]", seed: int = 2236, perturb: bool = True,
                dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """HeNormal kernels (truncated normal, sigma = sqrt(2/fan_in)/0.8796, cut at 2 sigma); with ``perturb`` the biases and
    the norm parameters/statistics are moved off their Keras initial values so every fused term is exercised."""
    g = torch.Generator().manual_seed(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shp in shapes.items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "kernel":
            kh, kw, a, b = shp
            fan_in = kh * kw * a            # Keras computes fans from the kernel shape: receptive field x shape[-2] (also for HWOI)
            sigma = math.sqrt(2.0 / fan_in) / 0.87962566103423978
            t = torch.empty(shp, dtype=torch.float64)
            torch.nn.init.trunc_normal_(t, 0.0, sigma, -2 * sigma, 2 * sigma, generator=g)
        elif leaf == "bias":
            t = torch.randn(shp, generator=g, dtype=torch.float64) * 0.05 if perturb else torch.zeros(shp, dtype=torch.float64)
        elif leaf == "gamma":
            t = 1.0 + (torch.randn(shp, generator=g, dtype=torch.float64) * 0.1 if perturb else 0.0) * torch.ones(shp, dtype=torch.float64)
        elif leaf in ("beta", "moving_mean"):
            t = torch.randn(shp, generator=g, dtype=torch.float64) * 0.1 if perturb else torch.zeros(shp, dtype=torch.float64)
        elif leaf == "moving_variance":
            t = 0.5 + torch.rand(shp, generator=g, dtype=torch.float64) if perturb else torch.ones(shp, dtype=torch.float64)
        else:
            raise KeyError(name)
        out[name] = t.to(dtype)
    return out


# --------------------------------------------------------------------------------------
# forward passes
# --------------------------------------------------------------------------------------
class ResNestEncoderOracle:
    """ResNest.py:4-58 with a name->tensor parameter dict (shapes: encoder_param_shapes)."""

    def __init__(self, channel, ksize, radix, kpaths, params: Dict[str, torch.Tensor], dtype=torch.float64):
        self.ksize, self.radix, self.kpaths = ksize, radix, kpaths
        self.p = {k: v.detach().to(dtype) for k, v in params.items()}
        self.dtype = dtype

    def _conv(self, x, name, dilation=1):
        return conv2d_same(x, self.p[name + "/kernel"], self.p[name + "/bias"], dilation)

    def _bn(self, x, name):
        q = self.p
        return batchnorm_inference(x, q[name + "/gamma"], q[name + "/beta"], q[name + "/moving_mean"], q[name + "/moving_variance"])

    def _ln(self, x, name):
        return layernorm_c(x, self.p[name + "/gamma"], self.p[name + "/beta"])

    def split_attention(self, inputs: Sequence[torch.Tensor], p: str):
        holder = inputs[0]
        for t in inputs[1:]:
            holder = holder + t                                        # :174-178 (tf tensors are immutable: no aliasing)
        y = holder.mean(dim=(1, 2))[:, None, None, :]                  # :180-181
        y = leaky(self._ln(self._conv(y, p + "/dense1"), p + "/dense1_bn"))
        out = None
        for r in range(self.radix):
            z = self._conv(y, p + "/dense2")                           # the one dense2, R times (:188)
            z = torch.sigmoid(z) if self.radix == 1 else torch.softmax(z, dim=-1)
            out = inputs[r] * z if out is None else out + inputs[r] * z
        return out

    def cardinal(self, x, p: str):
        inputs = []
        for _ in range(self.radix):                                    # same layers every time (:138-145)
            y = leaky(self._ln(self._conv(x, p + "/conv1"), p + "/conv1_bn"))
            y = leaky(self._ln(self._conv(y, p + "/conv2"), p + "/conv2_bn"))
            inputs.append(y)
        return self.split_attention(inputs, p + "/split")

    def residual_S(self, x, stage: str):
        cat = torch.cat([self.cardinal(x, f"{stage}/cardinal_{k}") for k in range(self.kpaths)], dim=3)
        c2 = self._conv(cat, stage + "/concats_2")
        sc = leaky(self._ln(self._conv(x, stage + "/convtmp_sc"), stage + "/convtmp_scbn"))
        return sc + c2

    def forward(self, x: torch.Tensor):
        x = x.to(self.dtype)
        x = leaky(self._conv(x, "initial_conv"))
        x = leaky(self._bn(self._conv(x, "convtmp_1"), "convtmp_1bn"))
        x = leaky(self._bn(self._conv(x, "convtmp_2"), "convtmp_2bn"))
        feats = []
        for stage, _ in ENC_STAGES:
            x = avgpool2(x)
            x = self.residual_S(x, stage)
            feats.append(x)
        return feats[3], [feats[2], feats[1], feats[0]]

    __call__ = forward


class DecoderCupOracle:
    """Decoder.py:99-146 (+ DecoderBlock :8-94)."""

    def __init__(self, num_classes, params: Dict[str, torch.Tensor], grid: Tuple[int, int] = (16, 5), dtype=torch.float64):
        self.num_classes, self.grid, self.dtype = num_classes, grid, dtype
        self.p = {k: v.detach().to(dtype) for k, v in params.items()}

    def _bn(self, x, name):
        q = self.p
        return batchnorm_inference(x, q[name + "/gamma"], q[name + "/beta"], q[name + "/moving_mean"], q[name + "/moving_variance"])

    def block(self, x, skip, p: str):
        x = conv2d_transpose_s2_same(x, self.p[p + "/up/kernel"], self.p[p + "/up/bias"])
        if skip is not None:
            x = torch.cat([x, skip], dim=3)
        for half in (1, 2):
            parts = []
            for j, d in enumerate(DILATIONS):
                y = conv2d_same(x, self.p[f"{p}/conv{half}_{j}/kernel"], self.p[f"{p}/conv{half}_{j}/bias"], 1 if j == 0 else d)
                parts.append(self._bn(y, f"{p}/bn{half}_{j}"))
            x = leaky(torch.cat(parts, dim=3))
        return x

    def forward(self, hidden_states: torch.Tensor, features: Optional[Sequence[torch.Tensor]] = None, logits: bool = False):
        y = hidden_states.to(self.dtype)
        n = y.shape[0]
        gh, gw = self.grid
        x = y.reshape(n, gh, gw, -1)
        x = conv2d_same(x, self.p["conv_more/kernel"], self.p["conv_more/bias"])
        x = leaky(layernorm_c(x, self.p["bn1/gamma"], self.p["bn1/beta"]))
        for i in range(3):
            skip = features[i].to(self.dtype) if features is not None else None
            x = self.block(x, skip, f"block_{i}")
            x0 = y.reshape(n, gh * 2 ** (i + 1), gw * 2 ** (i + 1), -1)
            x = torch.cat([x, x0], dim=3)
        z = conv2d_transpose_s2_same(x, self.p["head/kernel"], self.p["head/bias"])
        return z if logits else torch.softmax(z, dim=-1)

    __call__ = forward


def synthetic_input(n: int, h: int = 256, w: int = 80, c: int = 10, seed: int = 2234, dtype=torch.float32):
    """[n,h,w,c] frames: clipped Gaussian noise with a zeroed elliptical exterior (as the Variant A generator)."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, h, w, c, generator=g, dtype=torch.float64) * 0.35).clamp(-1, 1)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, h, dtype=torch.float64), torch.linspace(-1, 1, w, dtype=torch.float64), indexing="ij")
    mask = ((yy / 0.9) ** 2 + (xx / 0.8) ** 2 <= 1.0).to(torch.float64)
    return (x * mask[None, :, :, None]).to(dtype)


def synthetic_tokens(n: int, tokens: int = 80, hidden: int = 512, seed: int = 2235, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, tokens, hidden, generator=g, dtype=torch.float64) * 0.5).to(dtype)
