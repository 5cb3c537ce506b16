"""CPU oracle for Variant A of the hot path: the TBI_ResNest network, loss and training step.

TEST INFRASTRUCTURE ONLY.  Nothing in ``ultrasound_modeling_b200/`` imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may, and there only as the checker / the timed CPU baseline.

PINNED AGAINST THE REFERENCE'S OWN CODE, NOT AGAINST TENSORFLOW'S BINARIES.  The reference
(silverlight6/Ultrasound_Modeling) ships no tests, golden vectors or saved weights and TensorFlow
cannot be installed here, so ``oracle/tfshim`` provides a stand-in ``tensorflow`` package under which
the UNMODIFIED ``/root/reference/TBI_ResNest.py`` imports and runs: its functional-API graph, layer
creation order and Keras auto-names, ``step``, ``my_loss_cat``, GradientTape -> Adam.
``tests/golden/make_golden_ref.py`` recorded what that code returns for radix/kpaths 2/1, 3/4, 4/4 and
1/1 (``tests/golden/ref_tbi_resnest_*.npz``) and ``tests/test_oracle_pinned.py`` holds this file to it
in float64: variable inventory (names, shapes, order), probabilities, loss, accuracy, every gradient
of two consecutive training steps, every variable after them, an evaluation call -- all to 1e-9.
What stays unpinned is the layer arithmetic underneath the reference's code: the shim's primitives are
restated from TF/Keras' documented definitions (explicit tap sums over the SAME padding rule, the
transposed convolution as a scatter -- deliberately not the formulations used below), not TF itself.
This file restates the Keras graph of ``/root/reference/TBI_ResNest.py`` with plain PyTorch CPU ops
(fp32 or fp64); the two non-obvious Keras semantics (TF "SAME" transposed convolution,
inference-mode BatchNorm) are also cross-checked against first-principles numpy definitions in
``tests/test_oracle.py``.

Layouts follow Keras: activations NHWC, Conv2D kernels HWIO ``[kh,kw,Cin,Cout]``,
Conv2DTranspose kernels HWOI ``[kh,kw,Cout,Cin]``.  Parameter names follow the Keras layer names
the reference assigns (``TBI_ResNest.py:83-124,135,143-144,163-169,189-195,210``); layers the
reference leaves unnamed get the Keras auto-names in creation order (``conv2d``, ``conv2d_1``, ...,
``batch_normalization``, ...).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # Keras BatchNormalization default epsilon
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-7   # Keras Adam defaults (TBI_ResNest.py:28)

STAGES = (("conv2_1", 64), ("conv2_2", 128), ("conv3_1", 256), ("conv3_2", 512), ("conv4_1", 512))
UPSAMPLES = ((512, True), (512, True), (512, True), (256, False), (128, False))


# --------------------------------------------------------------------------------------
# Keras layer semantics on NHWC tensors
# --------------------------------------------------------------------------------------
def conv2d_same(x: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor],
                dilation: int = 1) -> torch.Tensor:
    """tf.keras.layers.Conv2D(strides=1, padding='SAME'); x NHWC, kernel HWIO."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    w = kernel.permute(3, 2, 0, 1)
    y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, stride=1,
                 padding=((kh // 2) * dilation, (kw // 2) * dilation), dilation=dilation)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose_s2_same(x: torch.Tensor, kernel: torch.Tensor,
                             bias: Optional[torch.Tensor]) -> torch.Tensor:
    """tf.keras.layers.Conv2DTranspose(k, strides=2, padding='same'); kernel HWOI, k in {3,4}.

    TF defines it as the input-gradient of a SAME strided conv.  k=4: torch padding=1.
    k=3: torch padding=0 then crop to [:2H,:2W] (SURVEY 8c pitfall 2).
    """
    k = kernel.shape[0]
    n, h, w_, _ = x.shape
    w = kernel.permute(3, 2, 0, 1)                      # [Cin, Cout, kh, kw]
    xin = x.permute(0, 3, 1, 2)
    if k == 4:
        y = F.conv_transpose2d(xin, w, bias, stride=2, padding=1)
    elif k == 3:
        y = F.conv_transpose2d(xin, w, bias, stride=2, padding=0)[..., : 2 * h, : 2 * w_]
    else:
        raise ValueError("k must be 3 or 4")
    return y.permute(0, 2, 3, 1)


def batchnorm_inference(x, gamma, beta, mean, var, eps: float = BN_EPS):
    """BatchNormalization called without training=True: affine with the moving statistics
    (TBI_ResNest.py:40 calls the Keras model without ``training``)."""
    return (x - mean) * (gamma / torch.sqrt(var + eps)) + beta


def avgpool2(x):
    """AveragePooling2D(pool_size=2, strides=2), padding valid."""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)


# --------------------------------------------------------------------------------------
# parameter inventory (creation order == Keras trainable order)
# --------------------------------------------------------------------------------------
def cardinal_channels(stage_out: int, radix: int, kpaths: int) -> Tuple[int, int]:
    oc = stage_out // 2                                   # TBI_ResNest.py:134
    return int(oc / radix / kpaths), int(oc / kpaths)      # TBI_ResNest.py:157-158


def param_shapes(channel: int, num_class: int, ksize: int, radix: int, kpaths: int
                 ) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape for every variable of the Keras model, trainable or not."""
    P: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def conv(name, k, cin, cout):
        P[name + "/kernel"] = (k, k, cin, cout)
        P[name + "/bias"] = (cout,)

    def convt(name, k, cin, cout):
        P[name + "/kernel"] = (k, k, cout, cin)
        P[name + "/bias"] = (cout,)

    def bn(name, c):
        for s in ("gamma", "beta", "moving_mean", "moving_variance"):
            P[f"{name}/{s}"] = (c,)

    conv("Conv1", 3, channel, 16)
    conv("conv2_1_1", 3, 16, 32)
    conv("conv2_1_2", 3, 32, 32)
    bn("conv2_1_2bn", 32)
    cin = 32
    n_auto_conv = 0
    for stage, out in STAGES:
        cv11, cvkk = cardinal_channels(out, radix, kpaths)
        for k in range(kpaths):
            nc = f"{stage}_car_k{k}"
            for r in range(radix):
                conv(f"{nc}1_r{r}", 1, cin, cv11)
                bn(f"{nc}1_{r}bn", cv11)                  # "%s1_%rbn" % (name, idx_r)
                conv(f"{nc}2_r{r}", ksize, cv11, cvkk)
                bn(f"{nc}2_{r}bn", cvkk)
            att = f"{nc}_att"
            conv(f"{att}1", 1, cvkk, cvkk // 2)
            bn(f"{att}_bn", cvkk // 2)
            for r in range(radix):
                conv(f"{att}2_r{r}", 1, cvkk // 2, cvkk)
        auto = "conv2d" if n_auto_conv == 0 else f"conv2d_{n_auto_conv}"
        n_auto_conv += 1
        conv(auto, ksize, kpaths * cvkk, out)             # concats_2, TBI_ResNest.py:140
        if cin != out:
            conv(f"{stage}_cc", 1, cin, out)
            bn(f"{stage}_scbn", out)
        cin = out
    skips = [512, 512, 256, 128, 64, 32]                   # conv6..conv1 pooled widths
    cin = skips[0]
    for i, (out, _) in enumerate(UPSAMPLES):
        convt(f"upsample_{i}_t_conv", 4, cin, out)
        bn("batch_normalization" if i == 0 else f"batch_normalization_{i}", out)
        cin = out + skips[i + 1]
    convt("f_tran", 4, cin, num_class)
    return P


def is_trainable(name: str) -> bool:
    return not (name.endswith("/moving_mean") or name.endswith("/moving_variance"))


def auto_conv_name(i: int) -> str:
    return "conv2d" if i == 0 else f"conv2d_{i}"


def auto_bn_name(i: int) -> str:
    return "batch_normalization" if i == 0 else f"batch_normalization_{i}"


def init_params(channel: int, num_class: int, ksize: int, radix: int, kpaths: int, seed: int = 1236,
                perturb: bool = True, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Keras defaults (glorot_uniform kernels, zero bias, BN gamma=1 beta=0 mean=0 var=1); with
    ``perturb`` the biases and BN statistics are randomised so that fused
    epilogues are exercised by parity tests."""
    g = torch.Generator().manual_seed(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shp in param_shapes(channel, num_class, ksize, radix, kpaths).items():
        if name.endswith("/kernel"):
            rf = shp[0] * shp[1]
            limit = math.sqrt(6.0 / (rf * shp[2] + rf * shp[3]))
            t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * limit
        elif name.endswith("/bias"):
            t = torch.randn(shp, generator=g, dtype=torch.float64) * 0.05 if perturb else torch.zeros(shp, dtype=torch.float64)
        elif name.endswith("/gamma"):
            t = 1 + torch.randn(shp, generator=g, dtype=torch.float64) * 0.1 if perturb else torch.ones(shp, dtype=torch.float64)
        elif name.endswith("/beta") or name.endswith("/moving_mean"):
            t = torch.randn(shp, generator=g, dtype=torch.float64) * 0.1 if perturb else torch.zeros(shp, dtype=torch.float64)
        elif name.endswith("/moving_variance"):
            t = 0.5 + torch.rand(shp, generator=g, dtype=torch.float64) if perturb else torch.ones(shp, dtype=torch.float64)
        else:
            raise AssertionError(name)
        out[name] = t.to(dtype)
    return out


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synthetic_batch(n: int, h: int, w: int, c: int = 1, num_class: int = 3, seed: int = 1234,
                    dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, h, w, c, generator=g, dtype=torch.float64) * 0.35).clamp_(-1, 1)
    yy = (torch.arange(h, dtype=torch.float64) - (h - 1) / 2) / (h * 0.46)
    xx = (torch.arange(w, dtype=torch.float64) - (w - 1) / 2) / (w * 0.42)
    inside = (yy[:, None] ** 2 + xx[None, :] ** 2) <= 1.0        # elliptical "brain mask"
    x = x * inside[None, :, :, None]
    g2 = torch.Generator().manual_seed(seed + 1)
    planes = torch.randn(n, num_class, h // 8, w // 8, generator=g2, dtype=torch.float64)
    planes = F.interpolate(planes, size=(h, w), mode="bilinear", align_corners=False)
    y = F.one_hot(planes.argmax(1), num_class).to(dtype)          # [n,h,w,num_class]
    return x.to(dtype), y


def dropout_masks(n: int, h: int, w: int, seed: int = 1237) -> List[torch.Tensor]:
    """Keep-masks (uint8 0/1) for the three always-on dropouts (TBI_ResNest.py:215-216)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(3):
        s = 2 ** (5 - i)                                     # upsample_0 output is H/32
        out.append((torch.rand(n, h // s, w // s, 512, generator=g) < 0.5).to(torch.uint8))
    return out


# --------------------------------------------------------------------------------------
# the graph (TBI_ResNest.py:80-220)
# --------------------------------------------------------------------------------------
class TBIResNestOracle:
    def __init__(self, height, width, channel, num_class, ksize, radix=4, kpaths=4,
                 learning_rate=1e-3, params: Optional[Dict[str, torch.Tensor]] = None,
                 dtype=torch.float32, seed: int = 1236, perturb: bool = True):
        self.height, self.width, self.channel, self.num_class = height, width, channel, num_class
        self.ksize, self.radix, self.kpaths = ksize, radix, kpaths
        self.learning_rate = learning_rate
        self.dtype = dtype
        src = params if params is not None else init_params(channel, num_class, ksize, radix, kpaths,
                                                            seed=seed, perturb=perturb, dtype=dtype)
        self.params: "OrderedDict[str, torch.Tensor]" = OrderedDict(
            (k, v.detach().clone().to(dtype).requires_grad_(is_trainable(k))) for k, v in src.items())
        self.adam_m = {k: torch.zeros_like(v) for k, v in self.params.items() if is_trainable(k)}
        self.adam_v = {k: torch.zeros_like(v) for k, v in self.params.items() if is_trainable(k)}
        self.adam_t = 0
        self._inter = None

    # -- helpers -----------------------------------------------------------------------
    def _conv(self, x, name, dilation=1):
        return conv2d_same(x, self.params[name + "/kernel"], self.params[name + "/bias"], dilation)

    def _bn(self, x, name):
        p = self.params
        return batchnorm_inference(x, p[name + "/gamma"], p[name + "/beta"],
                                   p[name + "/moving_mean"], p[name + "/moving_variance"])

    def split_attention(self, inputs: Sequence[torch.Tensor], name: str) -> torch.Tensor:
        """TBI_ResNest.py:175-207 (softmax over the CHANNEL axis, quirk kept)."""
        holder = inputs[0]
        for u in inputs[1:]:
            holder = holder + u
        ga = holder.mean(dim=(1, 2), keepdim=True)
        d1 = F.elu(self._bn(self._conv(ga, f"{name}1"), f"{name}_bn"))
        out = None
        for r, u in enumerate(inputs):
            d2 = self._conv(d1, f"{name}2_r{r}")
            a = torch.sigmoid(d2) if len(inputs) == 1 else torch.softmax(d2, dim=-1)
            out = u * a if out is None else out + u * a
        return out

    def cardinal(self, x, name):
        """TBI_ResNest.py:153-173."""
        us = []
        for r in range(self.radix):
            t = F.elu(self._bn(self._conv(x, f"{name}1_r{r}"), f"{name}1_{r}bn"))
            if self._inter is not None:
                self._inter[f"{name}/T1_r{r}"] = t
            t = F.elu(self._bn(self._conv(t, f"{name}2_r{r}"), f"{name}2_{r}bn"))
            if self._inter is not None:
                self._inter[f"{name}/U_r{r}"] = t
            us.append(t)
        return self.split_attention(us, f"{name}_att")

    def residual_S(self, x, stage, out, auto_idx):
        """TBI_ResNest.py:130-151."""
        cards = [self.cardinal(x, f"{stage}_car_k{k}") for k in range(self.kpaths)]
        c1 = torch.cat(cards, dim=3)
        if self._inter is not None:
            self._inter[f"{stage}/V"] = c1
        c2 = self._conv(c1, auto_conv_name(auto_idx))
        if x.shape[-1] != out:
            x = F.elu(self._bn(self._conv(x, f"{stage}_cc"), f"{stage}_scbn"))
        return x + c2

    def upsample(self, x, i, mask):
        """TBI_ResNest.py:209-220: convT k4 s2 -> BN -> [dropout 0.5, always on] -> ReLU."""
        p = self.params
        y = conv2d_transpose_s2_same(x, p[f"upsample_{i}_t_conv/kernel"], p[f"upsample_{i}_t_conv/bias"])
        y = self._bn(y, auto_bn_name(i))
        if mask is not None:
            y = y * (mask.to(y.dtype) * 2.0)               # tf.nn.dropout scales kept values by 1/(1-rate)
        return F.relu(y)

    def forward(self, x: torch.Tensor, masks: Optional[Sequence[Optional[torch.Tensor]]] = None,
                return_intermediates: bool = False):
        """x NHWC -> probabilities NHWC.  ``masks``: three keep-masks for upsample_0..2, or None to
        disable dropout (deterministic parity)."""
        x = x.to(self.dtype)
        inter = {}
        self._inter = inter if return_intermediates else None
        t = F.elu(self._conv(x, "Conv1"))
        t = F.elu(self._conv(t, "conv2_1_1"))
        t = F.elu(self._bn(self._conv(t, "conv2_1_2"), "conv2_1_2bn"))
        pools = [avgpool2(t)]                                 # conv1_pool
        for i, (stage, out) in enumerate(STAGES):
            s = self.residual_S(pools[-1], stage, out, i)
            inter[stage] = s
            inter[f"pool_{i + 1}"] = pools[-1]
            pools.append(avgpool2(s))                         # conv2_pool .. conv6_pool
        up = pools[5]
        for i, (_, drop) in enumerate(UPSAMPLES):
            m = masks[i] if (drop and masks is not None) else None
            up = self.upsample(up, i, m)
            inter[f"upsample_{i}"] = up
            up = torch.cat([up, pools[4 - i]], dim=3)
        p = self.params
        logits = conv2d_transpose_s2_same(up, p["f_tran/kernel"], p["f_tran/bias"])
        inter["f_tran"] = logits
        probs = torch.softmax(logits, dim=-1)
        return (probs, inter) if return_intermediates else probs

    # -- loss / step -------------------------------------------------------------------
    def my_loss_cat(self, y_true, y_pred):
        """TBI_ResNest.py:234-248; returns the [H,W] map (3 classes hard-coded like the reference)."""
        ce = 0
        for c in range(3):
            sf = 1.0 / (y_true[..., c].sum(dim=0) + 1.0)
            sf = sf / (self.height * self.width)
            ce = ce + (y_true[..., c] * torch.log(y_pred[..., c] + 1e-7)).sum(dim=0) * sf
        return -ce

    def step(self, x, y, train: bool = False, masks=None):
        """TBI_ResNest.py:35-55 -> (loss[H,W], accuracy, probs); gradient is of sum(loss)."""
        y = y.to(self.dtype)
        probs = self.forward(x, masks)
        loss = self.my_loss_cat(y, probs)
        grads = None
        if train:
            names = [k for k in self.params if is_trainable(k)]
            gs = torch.autograd.grad(loss.sum(), [self.params[k] for k in names])
            grads = dict(zip(names, gs))
            self.apply_adam(grads)
        acc = (probs.argmax(-1) == y.argmax(-1)).to(torch.float32).mean()
        self.last_grads = grads
        return loss.detach(), acc, probs.detach()

    def gradients(self, x, y, masks=None) -> Dict[str, torch.Tensor]:
        probs = self.forward(x, masks)
        loss = self.my_loss_cat(y.to(self.dtype), probs)
        names = [k for k in self.params if is_trainable(k)]
        gs = torch.autograd.grad(loss.sum(), [self.params[k] for k in names])
        return dict(zip(names, gs))

    def apply_adam(self, grads: Dict[str, torch.Tensor]):
        """Keras optimizer_v2 Adam (non-amsgrad) dense update."""
        self.adam_t += 1
        t = self.adam_t
        lr_t = self.learning_rate * math.sqrt(1 - ADAM_B2 ** t) / (1 - ADAM_B1 ** t)
        with torch.no_grad():
            for k, g in grads.items():
                m, v = self.adam_m[k], self.adam_v[k]
                m.mul_(ADAM_B1).add_(g, alpha=1 - ADAM_B1)
                v.mul_(ADAM_B2).addcmul_(g, g, value=1 - ADAM_B2)
                self.params[k].sub_(lr_t * m / (v.sqrt() + ADAM_EPS))

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, v.detach().clone()) for k, v in self.params.items())


def forward_flops_per_image(height, width, channel, num_class, ksize, radix, kpaths) -> float:
    """2*MAC count of the conv layers."""
    fl = 0.0
    h, w = height, width
    fl += 2 * h * w * 9 * (channel * 16 + 16 * 32 + 32 * 32)
    h //= 2; w //= 2
    cin = 32
    for _, out in STAGES:
        cv11, cvkk = cardinal_channels(out, radix, kpaths)
        fl += 2 * h * w * kpaths * radix * (cin * cv11 + ksize * ksize * cv11 * cvkk)
        fl += 2 * h * w * ksize * ksize * kpaths * cvkk * out
        if cin != out:
            fl += 2 * h * w * cin * out
        cin = out
        h //= 2; w //= 2
    skips = [512, 512, 256, 128, 64, 32]
    cin = 512
    for i, (out, _) in enumerate(UPSAMPLES):
        fl += 2 * h * w * 16 * cin * out                      # k4 s2: 16 taps per input pixel
        h *= 2; w *= 2
        cin = out + skips[i + 1]
    fl += 2 * h * w * 16 * cin * num_class
    return fl
