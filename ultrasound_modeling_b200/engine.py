"""Host-side executor of the TBI_ResNest graph (reference: TBI_ResNest.py:80-220, loss :234-248,
step :35-55) on top of the C ABI in include/tbi_sm100.h.

PyTorch is used for device memory, streams, CUDA graphs and torch.distributed only; every
arithmetic op of the path is a kernel of libtbi_sm100.so.  There is no torch/CPU fallback: if the
library or an sm_100 device is missing, construction raises.

Design (DESIGN.md has the long form):
  * all trainable parameters live in ONE flat fp32 buffer (creation order == Keras trainable order),
    gradients and Adam moments in parallel flat buffers -> one Adam launch, contiguous NCCL buckets;
  * the K*R cardinal branches of a stage are executed as ONE 1x1 conv (Cout = K*R*cv11) and ONE
    grouped 3x3 conv (groups = K*R); the master tensors are stored fused and sliced back into the
    Keras-named tensors only in state_dict()/load_state_dict();
  * BatchNorm (inference-affine) is folded into the packed compute weights + a
    per-channel bias each step; gamma/beta gradients are recovered from the raw weight gradient
    (tbi_bn_param_grad), so no pre-activation tensor is ever stored;
  * tf.concat is never materialised: consumers take two source views, dgrads split their output;
  * forward/backward are static "programs" (lists of prebuilt ctypes calls) -> CUDA-graph friendly.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (ACT_ELU, ACT_NONE, ACT_RELU, BF16, F32, IMPL_AUTO, Epilogue, SplitAtt, TapGemm, TapWgrad, View, check)

BN_EPS = 1e-3
STAGES = (("conv2_1", 64), ("conv2_2", 128), ("conv3_1", 256), ("conv3_2", 512), ("conv4_1", 512))
UPSAMPLES = ((512, True), (512, True), (512, True), (256, False), (128, False))
SKIP_C = (32, 64, 128, 256, 512, 512)          # channels of conv1_pool .. conv6_pool


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def view(t: torch.Tensor, c: Optional[int] = None, coff: int = 0) -> View:
    """tbi_view of an NHWC tensor (or a channel slice [coff, coff+c) of it)."""
    assert t.dim() == 4 and t.is_contiguous()
    n, h, w, ct = t.shape
    return View(t.data_ptr(), h, w, ct if c is None else c, ct, coff)


NULL_VIEW = View(None, 0, 0, 0, 0, 0)


def cardinal_channels(stage_out: int, radix: int, kpaths: int) -> Tuple[int, int]:
    oc = stage_out // 2                                    # TBI_ResNest.py:134
    return int(oc / radix / kpaths), int(oc / kpaths)       # TBI_ResNest.py:157-158


class _Spec:
    """one tensor in a flat buffer"""
    __slots__ = ("name", "shape", "offset", "numel")

    def __init__(self, name, shape):
        self.name, self.shape = name, tuple(shape)
        self.numel = int(np.prod(shape))
        self.offset = -1


class FlatStore:
    """A flat fp32 device buffer carved into named tensors (each 16-byte aligned)."""

    def __init__(self):
        self.specs: "OrderedDict[str, _Spec]" = OrderedDict()
        self.total = 0

    def add(self, name, shape) -> str:
        assert name not in self.specs, name
        s = _Spec(name, shape)
        s.offset = self.total
        self.total += (s.numel + 3) // 4 * 4
        self.specs[name] = s
        return name

    def alloc(self, device) -> torch.Tensor:
        return torch.zeros(max(self.total, 4), dtype=torch.float32, device=device)

    def get(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        s = self.specs[name]
        return buf[s.offset:s.offset + s.numel].view(s.shape)


class ConvLayer:
    """A Conv2D / Conv2DTranspose (+ optional inference BN + activation), possibly several Keras
    layers fused along Cout."""

    def __init__(self, name, kind, k, cin, cout, groups=1, bn=False, act=ACT_NONE, keras=None, keras_bn=None):
        self.name, self.kind, self.k = name, kind, k
        self.cin, self.cout, self.groups, self.bn, self.act = cin, cout, groups, bn, act
        self.cin_g = cin // groups
        self.keras = keras or []        # [(keras_layer_name, co_start, co_stop)]
        self.keras_bn = keras_bn or []  # [(keras_bn_name, co_start, co_stop)]
        self.wshape = (k, k, self.cin_g, cout) if kind == "conv" else (k, k, cout, cin)


class Engine:
    def __init__(self, height, width, channel, num_class, ksize, radix, kpaths, dtype="bf16", device="cuda",
                 impl=IMPL_AUTO, seed=0, layout_only=False):
        """layout_only=True builds just the parameter inventory / Keras name map (host logic, no device):
        used by CPU tests; such an engine cannot execute anything."""
        if not layout_only:
            if not torch.cuda.is_available():
                raise _lib.TbiError("ultrasound_modeling_b200 needs a CUDA device (sm_100); there is no CPU path")
            self.L = _lib.lib()
            if not self.L.tbi_device_ok():
                raise _lib.TbiError("libtbi_sm100.so: device is not sm_100 (B200)")
        assert height % 64 == 0 and width % 64 == 0, "TBI_ResNest needs H, W multiples of 64 (6 poolings)"
        assert ksize in (1, 3)
        self.H, self.W, self.Cin, self.num_class = height, width, channel, num_class
        self.ksize, self.radix, self.kpaths = ksize, radix, kpaths
        self.device = torch.device(device)
        self.dt = BF16 if dtype in ("bf16", torch.bfloat16) else F32
        self.tdtype = torch.bfloat16 if self.dt == BF16 else torch.float32
        self.impl = impl
        self.N = 0
        self.overlap_prepare = os.environ.get("TBI_NO_PREP_OVERLAP") is None
        self.overlap_bwd = os.environ.get("TBI_NO_BWD_OVERLAP") is None
        self._side_bwd = None
        self._side = None
        self._prep_pending = False
        self.gen = 0
        self.evicted: List[int] = []       # generations of builds whose buffers were released (graph caches purge these)
        self._define_layers()
        if not layout_only:
            self._alloc_params(seed)

    # ------------------------------------------------------------------ parameters
    def _define_layers(self):
        R, K, ks = self.radix, self.kpaths, self.ksize
        self.P = FlatStore()       # trainable
        self.S = FlatStore()       # moving statistics (not trainable)
        self.convs: "OrderedDict[str, ConvLayer]" = OrderedDict()
        self.atts: List[dict] = []

        def add(layer: ConvLayer):
            self.convs[layer.name] = layer
            self.P.add(layer.name + "/w", layer.wshape)
            self.P.add(layer.name + "/b", (layer.cout,))
            if layer.bn:
                self.P.add(layer.name + "/gamma", (layer.cout,))
                self.P.add(layer.name + "/beta", (layer.cout,))
                self.S.add(layer.name + "/mean", (layer.cout,))
                self.S.add(layer.name + "/var", (layer.cout,))
            return layer

        add(ConvLayer("Conv1", "conv", 3, self.Cin, 16, act=ACT_ELU, keras=[("Conv1", 0, 16)]))
        add(ConvLayer("conv2_1_1", "conv", 3, 16, 32, act=ACT_ELU, keras=[("conv2_1_1", 0, 32)]))
        add(ConvLayer("conv2_1_2", "conv", 3, 32, 32, bn=True, act=ACT_ELU, keras=[("conv2_1_2", 0, 32)],
                      keras_bn=[("conv2_1_2bn", 0, 32)]))
        cin = 32
        self.stage_info = []
        for si, (stage, out) in enumerate(STAGES):
            cv11, cvkk = cardinal_channels(out, R, K)
            assert cv11 >= 1 and cvkk >= 2, "radix*kpaths too large for this stage width"
            G = K * R
            k1, k1bn, k2, k2bn = [], [], [], []
            for k in range(K):
                for r in range(R):
                    g = k * R + r
                    nc = f"{stage}_car_k{k}"
                    k1.append((f"{nc}1_r{r}", g * cv11, (g + 1) * cv11)); k1bn.append((f"{nc}1_{r}bn", g * cv11, (g + 1) * cv11))
                    k2.append((f"{nc}2_r{r}", g * cvkk, (g + 1) * cvkk)); k2bn.append((f"{nc}2_{r}bn", g * cvkk, (g + 1) * cvkk))
            # The fused 1x1 conv produces G*cv11 channels; when that is not a multiple of 16 (radix 3: 24/60/120/252) the layer is
            # declared with its output width padded up to one: pad kernels/bias are zero and no Keras variable maps onto them
            # (their gradients are exactly zero, so Adam never moves them), the stored tensor T1 then has 16-byte aligned,
            # MMA-K-sized pixel records with zero pad lanes, and the grouped 3x3 conv behind it runs as a block-diagonal dense
            # tcgen05 conv over the padded width (tbi_conv_dense_expand) instead of on CUDA cores.
            c1w = (G * cv11 + 15) // 16 * 16
            c1 = add(ConvLayer(f"{stage}/c1", "conv", 1, cin, c1w, bn=True, act=ACT_ELU, keras=k1, keras_bn=k1bn))
            c2 = add(ConvLayer(f"{stage}/c2", "conv", ks, G * cv11, G * cvkk, groups=G, bn=True, act=ACT_ELU, keras=k2, keras_bn=k2bn))
            c = cvkk
            att = dict(stage=stage, c=c, name=f"{stage}/att")
            self.P.add(att["name"] + "/w1", (K, c, c // 2)); self.P.add(att["name"] + "/b1", (K, c // 2))
            self.P.add(att["name"] + "/gamma", (K, c // 2)); self.P.add(att["name"] + "/beta", (K, c // 2))
            self.S.add(att["name"] + "/mean", (K, c // 2)); self.S.add(att["name"] + "/var", (K, c // 2))
            self.P.add(att["name"] + "/w2", (K, R, c // 2, c)); self.P.add(att["name"] + "/b2", (K, R, c))
            self.atts.append(att)
            auto = "conv2d" if si == 0 else f"conv2d_{si}"
            cc2 = add(ConvLayer(f"{stage}/cc2", "conv", ks, K * cvkk, out, keras=[(auto, 0, out)]))
            sc = None
            if cin != out:
                sc = add(ConvLayer(f"{stage}/sc", "conv", 1, cin, out, bn=True, act=ACT_ELU, keras=[(f"{stage}_cc", 0, out)],
                                   keras_bn=[(f"{stage}_scbn", 0, out)]))
            self.stage_info.append(dict(stage=stage, cin=cin, out=out, cv11=cv11, cvkk=cvkk, G=G, c1w=c1w, c1=c1, c2=c2, att=att, cc2=cc2, sc=sc))
            cin = out
        cin = SKIP_C[5]
        self.ups = []
        for i, (out, drop) in enumerate(UPSAMPLES):
            bnname = "batch_normalization" if i == 0 else f"batch_normalization_{i}"
            up = add(ConvLayer(f"upsample_{i}", "convt", 4, cin, out, bn=True, act=ACT_RELU,
                               keras=[(f"upsample_{i}_t_conv", 0, out)], keras_bn=[(bnname, 0, out)]))
            self.ups.append(dict(layer=up, drop=drop, cin=cin, out=out))
            cin = out + SKIP_C[4 - i]
        self.head = add(ConvLayer("f_tran", "convt", 4, cin, self.num_class, keras=[("f_tran", 0, self.num_class)]))
        # all-reduce bucket boundary for data parallelism (parallel.GradSync), as a flat offset: [conv3_2 .. head] (~100 MiB: the
        # decoder and the two deepest encoder stages, final ~70 % of the way through the backward) and the rest (~5 MiB) at the
        # end.  Measured on 2 x B200 (step 6.94 ms on one GPU): this single cut 7.16 ms; cut at conv4_1 7.18; at conv3_1 7.22; no
        # cut (one all-reduce after the backward, fully exposed) 7.21; the earlier three buckets [upsample_0 | conv3_2 | rest] 7.28
        # -- every boundary also joins the weight-gradient side stream into the main stream, and the NCCL kernel costs the
        # persistent compute kernels about as much time as it overlaps, so fewer boundaries win.
        self.bucket_cuts = (self.P.specs["conv3_2/c1/w"].offset,)
        if os.environ.get("TBI_BUCKET_CUTS"):                   # sweep knob: comma-separated variable names (first variable of each later bucket)
            self.bucket_cuts = tuple(self.P.specs[nm].offset for nm in os.environ["TBI_BUCKET_CUTS"].split(","))

    def _alloc_params(self, seed):
        dev = self.device
        self.params = self.P.alloc(dev)
        self.grads = self.P.alloc(dev)
        self.adam_m = self.P.alloc(dev)
        self.adam_v = self.P.alloc(dev)
        self.stats = self.S.alloc(dev)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.draw_count = torch.zeros(1, dtype=torch.int32, device=dev)     # advances on every dropout draw (train or eval)
        self.hyper = torch.tensor([1e-3, 1.0, 0.0], dtype=torch.float32, device=dev)     # lr, grad_scale, clip_norm
        self._hyper_host = (1e-3, 1.0, 0.0)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.dropout_seed = 0x5EED
        # Keras defaults: glorot_uniform kernels, zero bias, BN gamma=1 beta=0 mean=0 var=1
        g = torch.Generator().manual_seed(seed)
        host = torch.zeros(self.params.numel(), dtype=torch.float32)
        hstat = torch.zeros(self.stats.numel(), dtype=torch.float32)

        def hv(store, buf, name):
            s = store.specs[name]
            return buf[s.offset:s.offset + s.numel].view(s.shape)

        for L in self.convs.values():
            w = hv(self.P, host, L.name + "/w")
            rf = L.k * L.k
            # per Keras layer fans (a fused tensor is several Keras layers side by side along Cout)
            for (_, a, b) in L.keras:
                if L.kind == "conv":
                    fan = rf * L.cin_g + rf * (b - a)
                    lim = math.sqrt(6.0 / fan)
                    w[..., a:b] = (torch.rand(w[..., a:b].shape, generator=g) * 2 - 1) * lim
                else:
                    fan = rf * L.cin + rf * L.cout
                    lim = math.sqrt(6.0 / fan)
                    w.copy_((torch.rand(w.shape, generator=g) * 2 - 1) * lim)
            if L.bn:
                hv(self.P, host, L.name + "/gamma").fill_(1.0)
                hv(self.S, hstat, L.name + "/var").fill_(1.0)
        for att in self.atts:
            c = att["c"]
            w1 = hv(self.P, host, att["name"] + "/w1"); w2 = hv(self.P, host, att["name"] + "/w2")
            w1.copy_((torch.rand(w1.shape, generator=g) * 2 - 1) * math.sqrt(6.0 / (c + c // 2)))
            w2.copy_((torch.rand(w2.shape, generator=g) * 2 - 1) * math.sqrt(6.0 / (c + c // 2)))
            hv(self.P, host, att["name"] + "/gamma").fill_(1.0)
            hv(self.S, hstat, att["name"] + "/var").fill_(1.0)
        self.params.copy_(host)
        self.stats.copy_(hstat)

    def p(self, name): return self.P.get(self.params, name)
    def g(self, name): return self.P.get(self.grads, name)
    def s(self, name): return self.S.get(self.stats, name)

    # ------------------------------------------------------------------ Keras-named state dict
    def _keras_items(self):
        """yield (keras_name, store, flat_name, index) where index slices the fused tensor"""
        R, K = self.radix, self.kpaths
        for L in self.convs.values():
            for (kn, a, b) in L.keras:
                if L.kind == "conv":
                    yield kn + "/kernel", "P", L.name + "/w", (Ellipsis, slice(a, b))
                else:
                    yield kn + "/kernel", "P", L.name + "/w", (Ellipsis,)
                yield kn + "/bias", "P", L.name + "/b", (slice(a, b),)
            for (kn, a, b) in L.keras_bn:
                yield kn + "/gamma", "P", L.name + "/gamma", (slice(a, b),)
                yield kn + "/beta", "P", L.name + "/beta", (slice(a, b),)
                yield kn + "/moving_mean", "S", L.name + "/mean", (slice(a, b),)
                yield kn + "/moving_variance", "S", L.name + "/var", (slice(a, b),)
        for att in self.atts:
            st, nm = att["stage"], att["name"]
            for k in range(K):
                an = f"{st}_car_k{k}_att"
                yield f"{an}1/kernel", "P", nm + "/w1", (k, None, None)          # -> [1,1,c,c/2]
                yield f"{an}1/bias", "P", nm + "/b1", (k,)
                yield f"{an}_bn/gamma", "P", nm + "/gamma", (k,)
                yield f"{an}_bn/beta", "P", nm + "/beta", (k,)
                yield f"{an}_bn/moving_mean", "S", nm + "/mean", (k,)
                yield f"{an}_bn/moving_variance", "S", nm + "/var", (k,)
                for r in range(R):
                    yield f"{an}2_r{r}/kernel", "P", nm + "/w2", (k, r, None, None)
                    yield f"{an}2_r{r}/bias", "P", nm + "/b2", (k, r)

    def _named(self, pbuf, sbuf, trainable_only=False):
        out = OrderedDict()
        for kn, store, flat, idx in self._keras_items():
            if store != "P" and (trainable_only or sbuf is None):
                continue
            t = self.P.get(pbuf, flat) if store == "P" else self.S.get(sbuf, flat)
            out[kn] = t[idx]
        return out

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        """Keras-named, Keras-layout (HWIO / HWOI) copies of every variable."""
        return OrderedDict((k, v.detach().clone()) for k, v in self._named(self.params, self.stats).items())

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        named = self._named(self.params, self.stats)
        missing = [k for k in named if k not in sd]
        if missing:
            raise KeyError(f"load_state_dict: missing {missing[:4]}... ({len(missing)})")
        for k, dst in named.items():
            src = torch.as_tensor(sd[k]).to(device=self.device, dtype=torch.float32)
            dst.copy_(src.reshape(dst.shape))

    def grad_dict(self) -> "OrderedDict[str, torch.Tensor]":
        """Keras-named gradients of the last backward (trainable variables only)."""
        named = self._named(self.grads, self.stats)
        return OrderedDict((k, v.detach().clone()) for k, v in named.items()
                           if not (k.endswith("/moving_mean") or k.endswith("/moving_variance")))

    # ------------------------------------------------------------------ buffers + programs
    MAX_CACHED_BUILDS = 3

    def build(self, n: int):
        """allocate activations for batch n and assemble the static call programs.

        A build owns every activation/input buffer and the ctypes programs that point into them, so anything captured
        against it (CUDA graphs) is only valid for THAT build: `self.gen` identifies the live build and callers key their
        graph caches on it.  The reference loop alternates batch sizes (train 64 / test 1, TBI_ResNest.py:395,421), so the
        last few builds are kept (buffers stay allocated, their graphs stay valid) and switching back is free; an evicted
        build's generation is reported through `self.evicted` so graph caches can drop what pointed into it."""
        if n == self.N:
            return
        cache = self.__dict__.setdefault("_builds", OrderedDict())
        if self.N:
            cache[self.N] = {k: self.__dict__[k] for k in self._build_keys}
            cache.move_to_end(self.N)
        if n in cache:
            self.__dict__.update(cache.pop(n))
            return
        while len(cache) >= self.MAX_CACHED_BUILDS:
            _, old = cache.popitem(last=False)
            self.evicted.append(old["gen"])
        first = "_build_keys" not in self.__dict__
        before = set(self.__dict__)
        self._gen_counter = self.__dict__.get("_gen_counter", 0) + 1
        self.gen = self._gen_counter
        self._build_fresh(n)
        if first:                          # the attribute names a build owns (the same for every batch size)
            self._build_keys = (set(self.__dict__) - before - {"_gen_counter", "_build_keys"}) | {"N", "gen"}

    def _build_fresh(self, n: int):
        self.N = n
        dev, td = self.device, self.tdtype
        H, W = self.H, self.W
        E = lambda *shape: torch.empty(shape, dtype=td, device=dev)
        self.x_in = torch.zeros(n, H, W, self.Cin, dtype=torch.float32, device=dev)
        self.y_in = torch.zeros(n, H, W, self.num_class, dtype=torch.float32, device=dev)
        self.x0 = E(n, H, W, self.Cin)
        self.t = [E(n, H, W, 16), E(n, H, W, 32), E(n, H, W, 32)]
        self.dstem = [E(n, H, W, 16), E(n, H, W, 32), E(n, H, W, 32)]
        self.pool = [E(n, H >> (i + 1), W >> (i + 1), SKIP_C[i]) for i in range(6)]
        self.dpool = [E(n, H >> (i + 1), W >> (i + 1), SKIP_C[i]) for i in range(6)]
        self.stage_buf = []
        for si, info in enumerate(self.stage_info):
            h, w = H >> (si + 1), W >> (si + 1)
            G, cv11, cvkk, out, K, R = info["G"], info["cv11"], info["cvkk"], info["out"], self.kpaths, self.radix
            c1w = info["c1w"]
            Z = lambda *shape: torch.zeros(shape, dtype=td, device=dev)      # pad lanes a kernel may not write must read as 0
            b = dict(T1=(Z if c1w != G * cv11 else E)(n, h, w, c1w), U=E(n, h, w, G * cvkk), V=E(n, h, w, K * cvkk), Y=E(n, h, w, out),
                     dZ1=(Z if c1w != G * cv11 else E)(n, h, w, c1w), dZ2=E(n, h, w, G * cvkk), dV=E(n, h, w, K * cvkk), dY=E(n, h, w, out),
                     gap=torch.empty(n, K, cvkk, dtype=torch.float32, device=dev),
                     h1=torch.empty(n, K, cvkk // 2, dtype=torch.float32, device=dev),
                     att=torch.empty(n, K, R, cvkk, dtype=torch.float32, device=dev))
            if info["sc"] is not None:
                b["SC"] = E(n, h, w, out); b["dZsc"] = E(n, h, w, out)
            self.stage_buf.append(b)
        att_scratch = max(n * self.kpaths * (self.radix * i["cvkk"] + 2 * i["cvkk"]) for i in self.stage_info)
        self.att_scratch = torch.empty(att_scratch, dtype=torch.float32, device=dev)
        self.up = []; self.dup = []; self.keep = []
        for i, u in enumerate(self.ups):
            h, w = H >> (5 - i), W >> (5 - i)
            self.up.append(E(n, h, w, u["out"])); self.dup.append(E(n, h, w, u["out"]))
            self.keep.append(torch.ones(n, h, w, u["out"], dtype=torch.uint8, device=dev) if u["drop"] else None)
        self.logits = torch.empty(n, H, W, self.num_class, dtype=torch.float32, device=dev)
        self.probs = torch.empty(n, H, W, self.num_class, dtype=torch.float32, device=dev)
        # bf16: the head's backward runs on the GATHERED gradient G[n,H/2,W/2, 16 taps x num_class (padded to 64 channels)]
        self.head_gather = self.dt == BF16 and 16 * self.num_class <= 64
        # bf16 head gradient: 4 channels per pixel (8-byte records: 3 real + a zero) when only the gather and the bias column sum
        # read it -- 33 MB instead of the 134 MB of 16-channel records at 64 x 256 x 256, which the loss kernel wrote and the
        # gather re-read sector by sector; 16 channels (16-byte aligned pixel records for TMA) when the transposed-conv kernels
        # read it directly.  The loss kernel writes the real channels only, the padding stays zero.
        self.dl_c = (4 if self.head_gather else 16) if self.dt == BF16 else self.num_class
        self.dlogits = torch.zeros(n, H, W, self.dl_c, dtype=td, device=dev)
        if self.head_gather:
            self.head_g = torch.zeros(n, H // 2, W // 2, 64, dtype=td, device=dev)
            self.head_wg = torch.zeros(self.head.cin * 64, dtype=td, device=dev)
        # The head's forward as ONE GEMM per input pixel, Y[n,H/2,W/2, 16 taps x num_class] in fp32, followed by the 4-tap
        # scatter into the logits (tbi_convt_scatter_y).  With 3 output channels the per-phase transposed conv is
        # MMA-dispatch-bound (160 N=16 MMAs per 128 outputs: 263 us against a 60 us HBM floor); the GEMM form issues 16x fewer
        # MMAs and its 48 fp32 columns leave through the wide fp32 epilogue (16-byte stores from the accumulator registers).
        # fp32 Y keeps the logits exact to fp32 rounding (with Y in bf16 every logit is a sum of four bf16-rounded terms and
        # the r4k4 gradient test leaves its 2e-2 bar: TBI_HEAD_FWD_GEMM=bf16, measurement only).  TBI_HEAD_FWD_GEMM=0: off.
        mode = os.environ.get("TBI_HEAD_FWD_GEMM", "1")
        self.head_fwd_gemm = self.head_gather and mode in ("1", "bf16")
        if self.head_fwd_gemm:
            self.head_yc = 64 if mode == "bf16" else 16 * self.num_class          # fp32: exactly the real columns (narrow epilogue)
            self.head_y = torch.empty(n, H // 2, W // 2, self.head_yc, dtype=td if mode == "bf16" else torch.float32, device=dev)
            self.head_wf = torch.zeros(self.head_yc * self.head.cin, dtype=td, device=dev)
        self.loss_map = torch.empty(H, W, dtype=torch.float32, device=dev)
        self.correct = torch.zeros(1, dtype=torch.int32, device=dev)
        # packed compute weights + folded BN
        esz = 2 if self.dt == BF16 else 4
        self.packed = {}
        ws_need = 0
        for L in self.convs.values():
            nel = L.k * L.k * L.cin_g * L.cout
            if L.kind == "conv":          # small-group convs are packed as block-diagonal dense kernels (tbi_conv_dense_expand)
                nel = int(self.L.tbi_conv_packed_elems(self.dt, L.k, L.groups, L.cin_g, L.cout))
                ws_need = max(ws_need, int(self.L.tbi_conv2d_wgrad_workspace(self.dt, L.k, L.groups, L.cin, L.cout)))
            nel_b = L.k * L.k * L.cin * self.dl_c if L is self.head else nel
            self.packed[L.name] = dict(wf=torch.zeros(nel, dtype=td, device=dev), wb=torch.zeros(nel_b, dtype=td, device=dev),
                                       scale=torch.empty(L.cout, dtype=torch.float32, device=dev),
                                       fbias=torch.empty(L.cout, dtype=torch.float32, device=dev))
        self.wgrad_ws = torch.empty(ws_need, dtype=torch.uint8, device=dev) if ws_need else None
        self._assemble()

    # -- program assembly helpers: each appends (fn, args) ; stream is appended at call time
    def _assemble(self):
        L = self.L
        dt, impl, n = self.dt, self.impl, self.N
        self.prog_prepare, self.prog_fwd, self.prog_loss, self.prog_bwd = [], [], [], []
        self._keepalive = []
        ws_bytes = 0
        wgrads = []
        # (number of prog_bwd calls issued, lowest flat offset above which every gradient is final):
        # lets the data-parallel wrapper all-reduce finished buckets while backward continues
        self.bwd_marks: List[Tuple[int, int]] = []
        self.bwd_side = set()          # indices of prog_bwd entries that may run beside the data-gradient chain (see run_bwd)
        done = set()
        order = list(self.P.specs.values())

        def mark(prefix):
            for s_ in order:
                if s_.name.startswith(prefix + "/"):
                    done.add(s_.name)
            wm = self.P.total
            for s_ in reversed(order):
                if s_.name not in done:
                    break
                wm = s_.offset
            self.bwd_marks.append((len(self.prog_bwd), wm))

        def keep(*objs):
            self._keepalive.extend(objs)
            return objs[0] if len(objs) == 1 else objs

        def bref(v):
            return C.byref(keep(v))

        def epi(**kw):
            e = Epilogue()
            for k, v in kw.items():
                setattr(e, k, v)
            if e.out_stride == 0:
                e.out_stride = 1
            return e

        # weight preparation: two item tables (prep.py) per part -- the stem's on the main stream (needed at once), everything
        # else on a side stream under the stem's forward.  Each part is ONE fold launch + ONE pack launch.
        from . import prep as P_
        esz = 2 if dt == BF16 else 4

        def prep_rows(Lr: ConvLayer):
            pk = self.packed[Lr.name]
            bnp = [_ptr(self.p(Lr.name + "/gamma")), _ptr(self.p(Lr.name + "/beta")), _ptr(self.s(Lr.name + "/mean")), _ptr(self.s(Lr.name + "/var"))] if Lr.bn else [None] * 4
            fold = (Lr.cout, *bnp, _ptr(self.p(Lr.name + "/b")), _ptr(pk["scale"]), _ptr(pk["fbias"]))
            g_, v_ = (bnp[0], bnp[3]) if Lr.bn else (None, None)
            wp = _ptr(self.p(Lr.name + "/w"))
            items = []
            if Lr.kind == "conv":
                items += P_.conv_items(L, dt, esz, 0, Lr.k, Lr.groups, Lr.cin_g, Lr.cout, wp, g_, v_, _ptr(pk["wf"]))
                items += P_.conv_items(L, dt, esz, 1, Lr.k, Lr.groups, Lr.cin_g, Lr.cout, wp, g_, v_, _ptr(pk["wb"]))
            else:
                cpad = self.dl_c if Lr is self.head else 0
                items += P_.convt_items(L, esz, 0, Lr.k, Lr.cin, Lr.cout, 0, wp, g_, v_, _ptr(pk["wf"]))
                items += P_.convt_items(L, esz, 1, Lr.k, Lr.cin, Lr.cout, cpad, wp, g_, v_, _ptr(pk["wb"]))
                if Lr is self.head and self.head_gather:
                    items += P_.convt_items(L, esz, 2, Lr.k, Lr.cin, Lr.cout, 64, wp, g_, v_, _ptr(self.head_wg))
                if Lr is self.head and self.head_fwd_gemm:
                    items += P_.convt_items(L, 2 if self.head_wf.dtype == torch.bfloat16 else 4, 3, Lr.k, Lr.cin, Lr.cout, self.head_yc, wp, g_, v_, _ptr(self.head_wf))
            return fold, items

        stem = ("Conv1", "conv2_1_1", "conv2_1_2")
        parts = ([self.convs[nm] for nm in stem], [Lr for Lr in self.convs.values() if Lr.name not in stem])
        self.prep_tables = []
        for layers in parts:
            folds, items = [], []
            for Lr in layers:
                f, it = prep_rows(Lr)
                folds.append(f); items += it
            self.prep_tables.append((P_.FoldTable(folds, self.device), P_.PrepTable(items, self.device)))

        def conv_fwd(Lr, h, w, src0, src1, e):
            pk = self.packed[Lr.name]
            e.bias = _ptr(pk["fbias"]); e.act = Lr.act
            if Lr.kind == "conv":
                self.prog_fwd.append((L.tbi_conv2d_fwd, (dt, impl, n, h, w, Lr.k, 1, Lr.groups, bref(src0), bref(src1) if src1 is not None else None,
                                                         Lr.cout, _ptr(pk["wf"]), bref(e))))
            else:
                self.prog_fwd.append((L.tbi_conv2d_transpose_s2_fwd, (dt, impl, n, h, w, Lr.k, bref(src0), bref(src1) if src1 is not None else None,
                                                                      Lr.cout, _ptr(pk["wf"]), bref(e))))

        def conv_bwd(Lr, h, w, x0, x1, dz, e_dgrad):
            """wgrad (+dbias) [+ BN param grads]; then dgrad through e_dgrad (None: no input gradient)."""
            nonlocal ws_bytes
            pk = self.packed[Lr.name]
            dw, db = _ptr(self.g(Lr.name + "/w")), _ptr(self.g(Lr.name + "/b"))
            if Lr.kind == "conv":
                wargs = [dt, impl, n, h, w, Lr.k, 1, Lr.groups, bref(x0), bref(x1) if x1 is not None else None, bref(dz), dw, db,
                         _ptr(self.wgrad_ws) if self.wgrad_ws is not None else None, self.wgrad_ws.numel() if self.wgrad_ws is not None else 0]
                self.prog_bwd.append((L.tbi_conv2d_wgrad, wargs))
                self.bwd_side.add(len(self.prog_bwd) - 1)
            else:
                wargs = [dt, impl, n, h, w, Lr.k, bref(x0), bref(x1) if x1 is not None else None, bref(dz), Lr.cout, dw, db, None, 0]
                self.prog_bwd.append((L.tbi_conv2d_transpose_s2_wgrad, wargs))
                self.bwd_side.add(len(self.prog_bwd) - 1)
            wgrads.append((Lr, wargs, n, h, w, x0, x1, dz))
            if Lr.bn:
                if Lr.kind == "conv":
                    lay = (Lr.k * Lr.k * Lr.cin_g, 1, Lr.cout, 1)
                else:
                    lay = (Lr.k * Lr.k, Lr.cin, Lr.cout * Lr.cin, Lr.cin)
                self.prog_bwd.append((L.tbi_bn_param_grad, (Lr.cout, *lay, _ptr(self.p(Lr.name + "/w")), dw, _ptr(self.p(Lr.name + "/b")), db,
                                                            _ptr(self.p(Lr.name + "/gamma")), _ptr(self.s(Lr.name + "/mean")),
                                                            _ptr(self.s(Lr.name + "/var")), BN_EPS, _ptr(self.g(Lr.name + "/gamma")),
                                                            _ptr(self.g(Lr.name + "/beta")))))
                self.bwd_side.add(len(self.prog_bwd) - 1)
            mark(Lr.name)
            if e_dgrad is not None:
                if Lr.kind == "conv":
                    self.prog_bwd.append((L.tbi_conv2d_dgrad, (dt, impl, n, h, w, Lr.k, 1, Lr.groups, bref(dz), Lr.cin, _ptr(pk["wb"]), bref(e_dgrad))))
                else:
                    self.prog_bwd.append((L.tbi_conv2d_transpose_s2_dgrad, (dt, impl, n, h, w, Lr.k, bref(dz), Lr.cin, _ptr(pk["wb"]), bref(e_dgrad))))

        H, W = self.H, self.W
        cv = self.convs
        # ---------------- forward ----------------
        self.prog_fwd.append((L.tbi_cast, (1, dt, self.x_in.numel(), _ptr(self.x_in), _ptr(self.x0))))
        conv_fwd(cv["Conv1"], H, W, view(self.x0), None, epi(out=view(self.t[0])))
        conv_fwd(cv["conv2_1_1"], H, W, view(self.t[0]), None, epi(out=view(self.t[1])))
        conv_fwd(cv["conv2_1_2"], H, W, view(self.t[1]), None, epi(out=view(self.t[2])))
        self.prog_fwd.append((L.tbi_avgpool2x2_fwd, (dt, n, H, W, bref(view(self.t[2])), bref(view(self.pool[0])))))
        self.fwd_split = len(self.prog_fwd)                  # everything after this needs the side-stream packs
        self.att_desc = []
        for si, info in enumerate(self.stage_info):
            h, w = H >> (si + 1), W >> (si + 1)
            b = self.stage_buf[si]
            pin = self.pool[si]
            conv_fwd(info["c1"], h, w, view(pin), None, epi(out=view(b["T1"])))
            conv_fwd(info["c2"], h, w, view(b["T1"], c=info["G"] * info["cv11"]), None, epi(out=view(b["U"])))
            nm = info["att"]["name"]
            sa = keep(SplitAtt(dt, n, h, w, self.kpaths, self.radix, info["cvkk"], ACT_ELU, BN_EPS,
                               _ptr(self.p(nm + "/w1")), _ptr(self.p(nm + "/b1")), _ptr(self.p(nm + "/gamma")), _ptr(self.p(nm + "/beta")),
                               _ptr(self.s(nm + "/mean")), _ptr(self.s(nm + "/var")), _ptr(self.p(nm + "/w2")), _ptr(self.p(nm + "/b2")),
                               _ptr(b["gap"]), _ptr(b["h1"]), _ptr(b["att"])))
            self.att_desc.append(sa)
            self.prog_fwd.append((L.tbi_split_attention_fwd, (C.byref(sa), bref(view(b["U"])), bref(view(b["V"])))))
            if info["sc"] is not None:
                conv_fwd(info["sc"], h, w, view(pin), None, epi(out=view(b["SC"])))
                res = view(b["SC"])
            else:
                res = view(pin)
            conv_fwd(info["cc2"], h, w, view(b["V"]), None, epi(out=view(b["Y"]), residual=res))
            self.prog_fwd.append((L.tbi_avgpool2x2_fwd, (dt, n, h, w, bref(view(b["Y"])), bref(view(self.pool[si + 1])))))
        for i, u in enumerate(self.ups):
            h, w = H >> (6 - i), W >> (6 - i)          # input dims
            src0 = view(self.pool[5]) if i == 0 else view(self.up[i - 1])
            src1 = None if i == 0 else view(self.pool[5 - i])
            conv_fwd(u["layer"], h, w, src0, src1, epi(out=view(self.up[i]), drop_keep=_ptr(self.keep[i])))
        if self.head_fwd_gemm:
            gf = keep(TapGemm())
            gf.dtype = dt; gf.impl = impl; gf.n = n; gf.gh = H // 2; gf.gw = W // 2; gf.groups = 1
            gf.cin_g = self.head.cin; gf.cout_g = self.head_yc; gf.src[0] = view(self.up[4]); gf.src[1] = view(self.pool[0]); gf.in_stride = 1; gf.ntaps = 1
            gf.w = _ptr(self.head_wf); gf.epi = epi(out=view(self.head_y), out_f32=int(self.head_y.dtype == torch.float32))
            self.prog_fwd.append((L.tbi_tapgemm_run, (C.byref(gf),)))
            # fp32 tap products: the loss kernel forms the logits itself (tbi_softmax_loss_fwd_bwd_taps) -- the 4-tap scatter launch
            # and the [N,H,W,3] fp32 logits round trip disappear (80 + 68 us -> one kernel; TBI_HEAD_FUSED_LOSS=0 keeps the scatter)
            self.head_fused_loss = self.head_y.dtype == torch.float32 and os.environ.get("TBI_HEAD_FUSED_LOSS", "1") != "0"
            if not self.head_fused_loss:
                self.prog_fwd.append((L.tbi_convt_scatter_y, (F32 if self.head_y.dtype == torch.float32 else BF16, n, H // 2, W // 2, 4, self.num_class, bref(view(self.head_y)),
                                                              _ptr(self.packed[self.head.name]["fbias"]), bref(view(self.logits)))))
        else:
            self.head_fused_loss = False
            conv_fwd(self.head, H // 2, W // 2, view(self.up[4]), view(self.pool[0]), epi(out=view(self.logits), out_f32=1))
        # ---------------- loss ----------------
        if self.head_fused_loss:
            self.prog_loss.append((L.tbi_softmax_loss_fwd_bwd_taps, (dt, n, H, W, self.num_class, bref(view(self.head_y)), _ptr(self.packed[self.head.name]["fbias"]),
                                                                     _ptr(self.y_in), _ptr(self.probs), _ptr(self.loss_map), _ptr(self.correct), _ptr(self.dlogits), self.dl_c)))
        else:
            self.prog_loss.append((L.tbi_softmax_loss_fwd_bwd, (dt, n, H, W, self.num_class, _ptr(self.logits), _ptr(self.y_in), _ptr(self.probs),
                                                                _ptr(self.loss_map), _ptr(self.correct), _ptr(self.dlogits), self.dl_c)))
        # ---------------- backward ----------------
        # head: d(up4) gets ReLU' of up4 (no dropout on upsample_4); d(pool[0]) plain write
        head_epi = epi(out=view(self.dup[4]), dact=ACT_RELU, dact_ref=view(self.up[4]), split_c=self.ups[4]["out"], out2=view(self.dpool[0]))
        if self.head_gather:
            # gather the 16 strided taps of dlogits once, then dW = X^T G (lands in HWOI order) and dX = G W' are 1-tap GEMMs
            hh, hw_ = H // 2, W // 2
            nc = self.num_class
            self.prog_bwd.append((L.tbi_colsum, (dt, n * H * W, bref(view(self.dlogits, c=nc)), _ptr(self.g("f_tran/b")))))
            self.bwd_side.add(len(self.prog_bwd) - 1)
            self.prog_bwd.append((L.tbi_convt_gather_dz, (dt, n, hh, hw_, 4, nc, bref(view(self.dlogits)), bref(view(self.head_g)))))
            wd = keep(TapWgrad())
            wd.dtype = dt; wd.impl = impl; wd.n = n; wd.gh = hh; wd.gw = hw_; wd.groups = 1
            wd.cin_g = self.head.cin; wd.cout_g = 16 * nc
            wd.a_src[0] = view(self.up[4]); wd.a_src[1] = view(self.pool[0]); wd.b_src = view(self.head_g)
            wd.a_stride = 1; wd.b_stride = 1; wd.ntaps = 1
            wd.dw = _ptr(self.g("f_tran/w")); wd.tap_stride = 0; wd.ci_stride = 1; wd.co_stride = self.head.cin
            self.prog_bwd.append((L.tbi_tapwgrad_run, (C.byref(wd),)))
            self.bwd_side.add(len(self.prog_bwd) - 1)
            mark("f_tran")
            gd = keep(TapGemm())
            gd.dtype = dt; gd.impl = impl; gd.n = n; gd.gh = hh; gd.gw = hw_; gd.groups = 1
            gd.cin_g = 64; gd.cout_g = self.head.cin; gd.src[0] = view(self.head_g); gd.in_stride = 1; gd.ntaps = 1
            gd.w = _ptr(self.head_wg); gd.epi = head_epi
            self.prog_bwd.append((L.tbi_tapgemm_run, (C.byref(gd),)))
        else:
            conv_bwd(self.head, H // 2, W // 2, view(self.up[4]), view(self.pool[0]), view(self.dlogits), head_epi)
        for i in range(4, -1, -1):
            u = self.ups[i]
            h, w = H >> (6 - i), W >> (6 - i)
            if i == 0:
                e = epi(out=view(self.dpool[5]))
                conv_bwd(u["layer"], h, w, view(self.pool[5]), None, view(self.dup[0]), e)
            else:
                e = epi(out=view(self.dup[i - 1]), dact=ACT_RELU, dact_ref=view(self.up[i - 1]), dact_keep=_ptr(self.keep[i - 1]),
                        split_c=self.ups[i - 1]["out"], out2=view(self.dpool[5 - i]))
                conv_bwd(u["layer"], h, w, view(self.up[i - 1]), view(self.pool[5 - i]), view(self.dup[i]), e)
        for si in range(4, -1, -1):
            info, b = self.stage_info[si], self.stage_buf[si]
            h, w = H >> (si + 1), W >> (si + 1)
            pin, dpin = self.pool[si], self.dpool[si]
            self.prog_bwd.append((L.tbi_avgpool2x2_bwd, (dt, n, h, w, bref(view(self.dpool[si + 1])), bref(view(b["dY"])), 0, ACT_NONE, None)))
            conv_bwd(info["cc2"], h, w, view(b["V"]), None, view(b["dY"]), epi(out=view(b["dV"])))
            if info["sc"] is not None:
                self.prog_bwd.append((L.tbi_act_bwd, (dt, n * h * w, ACT_ELU, bref(view(b["dY"])), bref(view(b["SC"])), None, bref(view(b["dZsc"])))))
                conv_bwd(info["sc"], h, w, view(pin), None, view(b["dZsc"]), epi(out=view(dpin), residual=view(dpin)))
            else:
                self.prog_bwd.append((L.tbi_accumulate, (dt, n * h * w, bref(view(b["dY"])), bref(view(dpin)))))
            nm = info["att"]["name"]
            self.prog_bwd.append((L.tbi_split_attention_bwd, (C.byref(self.att_desc[si]), bref(view(b["U"])), bref(view(b["dV"])), bref(view(b["dZ2"])),
                                                              _ptr(self.g(nm + "/w1")), _ptr(self.g(nm + "/b1")), _ptr(self.g(nm + "/gamma")),
                                                              _ptr(self.g(nm + "/beta")), _ptr(self.g(nm + "/w2")), _ptr(self.g(nm + "/b2")),
                                                              _ptr(self.att_scratch))))
            mark(nm)
            creal = info["G"] * info["cv11"]
            conv_bwd(info["c2"], h, w, view(b["T1"], c=creal), None, view(b["dZ2"]),
                     epi(out=view(b["dZ1"], c=creal), dact=ACT_ELU, dact_ref=view(b["T1"], c=creal)))
            conv_bwd(info["c1"], h, w, view(pin), None, view(b["dZ1"]), epi(out=view(dpin), residual=view(dpin)))
        self.prog_bwd.append((L.tbi_avgpool2x2_bwd, (dt, n, H, W, bref(view(self.dpool[0])), bref(view(self.dstem[2])), 0, ACT_ELU, bref(view(self.t[2])))))
        conv_bwd(cv["conv2_1_2"], H, W, view(self.t[1]), None, view(self.dstem[2]), epi(out=view(self.dstem[1]), dact=ACT_ELU, dact_ref=view(self.t[1])))
        conv_bwd(cv["conv2_1_1"], H, W, view(self.t[0]), None, view(self.dstem[1]), epi(out=view(self.dstem[0]), dact=ACT_ELU, dact_ref=view(self.t[0])))
        conv_bwd(cv["Conv1"], H, W, view(self.x0), None, view(self.dstem[0]), None)
        self.ws = None
        self._wgrads = wgrads

    # ------------------------------------------------------------------ execution
    def _run(self, prog, stream):
        for fn, args in prog:
            rc = fn(*args, stream)
            if rc != 0:
                check(rc, fn.__name__)

    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def prepare(self):
        """fold BN and refresh the compute copies of the weights: per part one fold launch + one pack launch over a prebuilt item
        table (prep.py).  The stem's part on the current stream; the rest on a side stream that forward() joins after the stem --
        fork and join are event waits, so this also captures into a CUDA graph."""
        (f0, p0), (f1, p1) = self.prep_tables
        st = self.stream()
        f0.run(self.L, BN_EPS, st); p0.run(self.L, self.dt, BN_EPS, st)
        if not self.overlap_prepare:
            f1.run(self.L, BN_EPS, st); p1.run(self.L, self.dt, BN_EPS, st)
            return
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        self._side.wait_stream(torch.cuda.current_stream(self.device))
        f1.run(self.L, BN_EPS, self._side.cuda_stream); p1.run(self.L, self.dt, BN_EPS, self._side.cuda_stream)
        self._prep_pending = True

    def _join_prepare(self):
        if self._prep_pending:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
            self._prep_pending = False

    def draw_dropout(self, seed: Optional[int] = None):
        """fresh keep-masks for upsample_0..2.  tf.nn.dropout draws on EVERY call, eval included (TBI_ResNest.py:215-216), so
        the stream position is a dedicated device counter that advances per draw (not the Adam step, which only moves when
        training); `dropout_seed` is mixed with the data-parallel rank by GradSync so replicas draw independent masks."""
        seed = self.dropout_seed if seed is None else seed
        st = self.stream()
        for i, k in enumerate(self.keep):
            if k is not None:
                check(self.L.tbi_dropout_mask(k.data_ptr(), k.numel(), seed + 7919 * i, self.draw_count.data_ptr(), st), "dropout_mask")
        check(self.L.tbi_adam_advance(self.draw_count.data_ptr(), st), "draw_advance")

    def set_dropout(self, masks: Optional[Sequence[Optional[torch.Tensor]]]):
        """explicit 0/1 keep-masks (parity runs); None -> dropout disabled.  The device buffers hold the
        MULTIPLIER the epilogue applies: 0 dropped, 2 kept, 1 disabled."""
        for i, k in enumerate(self.keep):
            if k is None:
                continue
            if masks is None or masks[i] is None:
                k.fill_(1)
            else:
                k.copy_(torch.as_tensor(masks[i]).to(device=self.device, dtype=torch.uint8) * 2)

    def forward(self):
        self._run(self.prog_fwd[:self.fwd_split], self.stream())
        self._join_prepare()
        self._run(self.prog_fwd[self.fwd_split:], self.stream())

    def loss(self):
        self.correct.zero_()
        self._run(self.prog_loss, self.stream())

    def run_bwd(self, a: int, b: int):
        """prog_bwd[a:b] on the current stream, with the weight-gradient side of every conv layer (wgrad + bias sum + BN
        parameter gradients: needs only the layer's input and its output gradient, writes only parameter gradients) on a
        second stream beside the data-gradient chain.  Every gradient buffer is written once and never reused, so the only
        ordering needed is `side waits for whatever main has issued so far` in front of each side entry; the side stream is
        joined at the end of the slice, i.e. parameter gradients are final at every bucket boundary of parallel.py."""
        if not self.overlap_bwd or not self.bwd_side:
            self._run(self.prog_bwd[a:b], self.stream())
            return
        main = torch.cuda.current_stream(self.device)
        if self._side_bwd is None:
            self._side_bwd = torch.cuda.Stream(self.device)
        side = self._side_bwd
        dirty, used = True, False
        for i in range(a, b):
            fn, args = self.prog_bwd[i]
            if i in self.bwd_side:
                if dirty:
                    side.wait_stream(main)
                    dirty = False
                rc = fn(*args, side.cuda_stream)
                used = True
            else:
                rc = fn(*args, main.cuda_stream)
                dirty = True
            if rc != 0:
                check(rc, fn.__name__)
        if used:
            main.wait_stream(side)

    def backward(self):
        self.grads.zero_()
        self.run_bwd(0, len(self.prog_bwd))

    def set_hyper(self, lr: float, grad_scale: float = 1.0, clip_norm: float = 0.0):
        """learning rate / gradient scale / global-norm clip live in a 3-float device buffer that the (possibly captured) Adam
        launch reads, so changing `.learning_rate` between steps reaches a replayed CUDA graph.  Call OUTSIDE graph capture."""
        want = (float(lr), float(grad_scale), float(clip_norm))
        if want != self._hyper_host:
            self.hyper.copy_(torch.tensor(want, dtype=torch.float32), non_blocking=False)
            self._hyper_host = want

    def adam(self, lr: Optional[float] = None, grad_scale: float = 1.0):
        """one Keras-Adam update of the flat parameter buffer from the flat gradient buffer.  lr given: eager convenience form
        (uploads the hyper-parameters first); lr None: uses whatever set_hyper() last uploaded (the graph-capturable form)."""
        if lr is not None:
            self.set_hyper(lr, grad_scale, self._hyper_host[2])
        st = self.stream()
        gn = None
        if self._hyper_host[2] > 0.0:
            self.gnorm_sq.zero_()
            check(self.L.tbi_sumsq(self.P.total, self.grads.data_ptr(), self.gnorm_sq.data_ptr(), st), "sumsq")
            gn = self.gnorm_sq.data_ptr()
        check(self.L.tbi_adam_multi_dev(self.P.total, self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                                        self.adam_v.data_ptr(), self.step_count.data_ptr(), self.hyper.data_ptr(), gn,
                                        0.9, 0.999, 1e-7, st), "adam")
        check(self.L.tbi_adam_advance(self.step_count.data_ptr(), st), "adam_advance")

    def fallback_report(self, reset: bool = False) -> dict:
        """tap-GEMM launches (bf16, automatic dispatch) that ran on CUDA cores because the tcgen05 path refused their shape,
        cumulative for this process, with the reason of the last one (tbi_fallback_stats / tbi_last_fallback)."""
        a, b = C.c_int64(0), C.c_int64(0)
        check(self.L.tbi_fallback_stats(C.byref(a), C.byref(b), 0), "fallback_stats")
        why = self.L.tbi_last_fallback().decode(errors="replace")
        if reset:
            check(self.L.tbi_fallback_stats(None, None, 1), "fallback_stats")
        return dict(tapgemm_simt=int(a.value), tapwgrad_simt=int(b.value), last_reason=why)

    def launches_per_step(self, train: bool = True) -> int:
        """kernel launches of one step as the tensor-core build issues them (checked against the ncu launch list in
        profiles/): conv / convT wgrad = wgrad kernel + bias column sum (the 1->16 stem wgrad does both in one kernel; a
        block-diagonal-expanded grouped conv adds the gather), split-attention fwd = 1 fused kernel, bwd = fused kernel +
        parameter-gradient kernel; everything else is one kernel per entry point.  Memsets are not counted."""
        def count(prog):
            c = 0
            for fn, args in prog:
                nm = fn.__name__
                k = {"tbi_conv2d_transpose_s2_wgrad": 2, "tbi_conv2d_wgrad": 2, "tbi_split_attention_bwd": 2}.get(nm, 1)
                if nm == "tbi_conv2d_wgrad" and args[-1] and self.L.tbi_conv2d_wgrad_workspace(args[0], args[5], args[7], args[8]._obj.c, args[10]._obj.c):
                    k += 1
                c += k
            return c
        n = 4 + count(self.prog_fwd) + count(self.prog_loss)          # weight preparation: (fold + pack) x (stem, rest)
        n += sum(1 for k in self.keep if k is not None) + 1       # dropout masks + the draw counter
        if train:
            n += count(self.prog_bwd) + 2 - 1                      # Adam + its counter; -1: the stem wgrad needs no separate column sum
        return n
