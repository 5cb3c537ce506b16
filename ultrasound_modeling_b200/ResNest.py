"""Variant B encoder: drop-in for the reference's ``ResNest.py`` (classes ResNest, residual_S, cardinal, split_attention)
on the B200 path: forward, and backward (``forward(x, record=True)`` then ``backward(dx_4, [dx_3, dx_2, dx_1])`` -> dL/dx, with
the parameter gradients in ``gradients()``).  The loss of Variant B's training step sits behind the ViT bridge of
``VisionTransformer.py`` (scope row 8(f)-1, not built): backward therefore takes the gradients of the encoder's four outputs.

Same constructor signatures and call results as the reference:
    ResNest(height, width, channel, ksize, radix=4, kpaths=4, wDecay=None)(x) -> (x_4, [x_3, x_2, x_1])   ResNest.py:7,38-58
Inputs are NHWC (numpy or torch); outputs are NHWC device tensors in the storage dtype (bf16 default, fp32 for parity runs).
Every arithmetic step is a libtbi_sm100.so entry point (ops.py); torch only allocates, reshapes, owns the parameters and, in
the backward pass, slices / accumulates gradient buffers (the tape below).
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
        # backward tape (forward(..., record=True)): closures run in reverse; gradients of activation BUFFERS are kept per buffer
        # (a conv that writes a channel slice of a concat buffer reads the same slice of that buffer's gradient)
        self.recording = False
        self.tape = []
        self.tgrads: Dict[int, torch.Tensor] = {}
        self._alive = []
        self.grads: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.flat = None             # set by freeze(): dict(params, grads, m, v, names, gviews)
        self._rows = {}
        self.generation = 0          # bumped whenever a variable's storage moves (freeze, load): captured CUDA graphs key on it

    def start_recording(self):
        self.recording = True
        self.tape.clear(); self.tgrads.clear(); self._alive.clear(); self.grads.clear()
        if self.flat is not None:
            self.flat["grads"].zero_()

    def freeze(self):
        """Move every TRAINABLE variable into one flat fp32 buffer (the tensors in ``vars`` become views of it) with parallel
        flat buffers for gradients and the two Adam moments: one fused optimizer launch, one contiguous all-reduce, and the
        global gradient norm is one reduction.  Call after the first forward pass (variables are created lazily)."""
        if self.flat is not None:
            return self.flat
        names = [k for k in self.vars if not (k.endswith("/moving_mean") or k.endswith("/moving_variance"))]
        offs, total = {}, 0
        for k in names:
            offs[k] = total
            total += (self.vars[k].numel() + 3) // 4 * 4                 # 16-byte aligned slices
        dev = self.device
        P = torch.zeros(total, dtype=torch.float32, device=dev)
        G = torch.zeros(total, dtype=torch.float32, device=dev)
        gviews = {}
        for k in names:
            t = self.vars[k]
            view = P[offs[k]:offs[k] + t.numel()].view(t.shape)
            view.copy_(t)
            self.vars[k] = view
            gviews[k] = G[offs[k]:offs[k] + t.numel()].view(t.shape)
        self.generation += 1
        self.flat = dict(params=P, grads=G, m=torch.zeros_like(P), v=torch.zeros_like(P), names=names, gviews=gviews, offsets=offs,
                         step=torch.zeros(1, dtype=torch.int32, device=dev))
        return self.flat

    def index_rows(self, key) -> torch.Tensor:
        """device index vector for a channel map (tuple of positions, or an int = arange), built once: a host list turned into
        a device tensor inside a CUDA-graph capture would be a pageable host-to-device copy, which capture refuses"""
        t = self._rows.get(key)
        if t is None:
            t = self._rows[key] = (torch.arange(key, device=self.device) if isinstance(key, int)
                                   else torch.tensor(list(key), dtype=torch.long).to(self.device))
        return t

    def gacc(self, t: torch.Tensor, g: torch.Tensor, coff: int = 0):
        """add g into channels [coff, coff + g.channels) of the gradient buffer of activation buffer t"""
        G = self.tgrads.get(id(t))
        if G is None:
            G = self.tgrads[id(t)] = torch.zeros_like(t)
            self._alive.append(t)                              # ids stay unique while the tape lives
        G[..., coff:coff + g.shape[-1]] += g.to(G.dtype)

    def gset(self, t: torch.Tensor, g: torch.Tensor, coff: int = 0):
        self.tgrads[id(t)][..., coff:coff + g.shape[-1]] = g.to(self.tgrads[id(t)].dtype)

    def gget(self, t: torch.Tensor, coff: int = 0, c: Optional[int] = None) -> Optional[torch.Tensor]:
        G = self.tgrads.get(id(t))
        if G is None:
            return None
        c = G.shape[-1] - coff if c is None else c
        return G if (coff == 0 and c == G.shape[-1]) else G[..., coff:coff + c].contiguous()

    def pacc(self, name: str, g: torch.Tensor):
        if self.flat is not None and name in self.flat["gviews"]:
            view = self.flat["gviews"][name]
            view.add_(g.reshape(view.shape))
            self.grads[name] = view
            return
        self.grads[name] = self.grads[name] + g if name in self.grads else g.clone()

    def run_backward(self):
        for fn in reversed(self.tape):
            fn()
        self.tape.clear(); self._alive.clear()
        self.recording = False

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
            t = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).detach().to(device=self.device, dtype=torch.float32).contiguous()
            if self.flat is not None and k in self.vars and k in self.flat["gviews"]:
                self.vars[k].copy_(t.reshape(self.vars[k].shape))        # frozen: the variable is a view of the flat buffer
            else:
                self.vars[k] = t
                self.generation += 1


class _Layer:
    def __init__(self, store: VariableStore, prefix: str):
        self._s, self._p = store, prefix

    # -- Keras layer equivalents on NHWC device tensors --------------------------------------
    def _conv(self, x, name, k, cout, *, dilation=1, bn=None, act=ACT_NONE, residual=None, x2=None, out=None, out_coff=0, out_f32=False,
              need_dx=True, cin=None, cmap=None, out_slot=None):
        """Conv2D + [BN] + [act] (+ residual).  Stored bf16 activations have their channels zero-padded to a multiple of 16
        (ops.pad_channels), so x may be physically wider than its ``cin`` logical channels and the output is allocated wider
        than ``cout``: the Keras-shaped master kernel is scattered into a zero kernel of the physical shape each call (rows
        ``cmap`` = physical position of every logical input channel, default the first cin), the convolution runs on the
        tensor-core path at the physical widths, pad output lanes come out exactly 0 (zero kernel columns, zero bias, identity
        BN, act(0) = 0), and the gradients are gathered back to the Keras shapes.  ``out``/``out_coff``/``out_slot``: write a
        slot of out_slot >= cout physical channels of a wider buffer."""
        cx = x.shape[3]
        cx_real = cx if cin is None else cin
        c2 = x2.shape[3] if x2 is not None else 0
        w = self._s.kernel(f"{self._p}{name}/kernel", (k, k, cx_real + c2, cout))
        b = self._s.vector(f"{self._p}{name}/bias", cout, 0.0)
        bnp = None
        if bn is not None:
            q = f"{self._p}{bn}/"
            bnp = (self._s.vector(q + "gamma", cout, 1.0), self._s.vector(q + "beta", cout, 0.0),
                   self._s.vector(q + "moving_mean", cout, 0.0), self._s.vector(q + "moving_variance", cout, 1.0))
        cout_p = (out_slot or cout) if out is not None else (cout if out_f32 else ops.pad_channels(cout, x.dtype))
        padded = cx != cx_real or cmap is not None or cout_p != cout
        rows = None
        if padded:
            dev = x.device
            rows = self._s.index_rows(tuple(cmap) if cmap is not None else cx_real)
            wk = torch.zeros(k, k, cx + c2, cout_p, dtype=torch.float32, device=dev)
            wk[:, :, rows, :cout] = w[:, :, :cx_real]
            if c2:
                wk[:, :, cx:, :cout] = w[:, :, cx_real:]
            bk = torch.zeros(cout_p, dtype=torch.float32, device=dev); bk[:cout] = b
            bnk = None
            if bnp is not None:
                fill = (1.0, 0.0, 0.0, 1.0)                      # identity BN on the pad lanes
                bnk = tuple(torch.full((cout_p,), f, dtype=torch.float32, device=dev) for f in fill)
                for dst, src in zip(bnk, bnp):
                    dst[:cout] = src
        else:
            wk, bk, bnk = w, b, bnp
        y = ops.conv2d(x, wk, bk, dilation=dilation, bn=bnk, act=act, residual=residual, x2=x2, out=out, out_coff=out_coff, out_f32=out_f32)
        if self._s.recording:
            s, pre = self._s, f"{self._p}{name}/"

            def bwd():
                buf, off = (out, out_coff) if out is not None else (y, 0)
                dy = s.gget(buf, off, cout_p)
                if dy is None:
                    return
                if residual is not None:
                    s.gacc(residual, dy)
                dz = ops.act_bwd(dy, buf if buf.shape[3] == cout_p else buf[..., off:off + cout_p].contiguous(), act) if act != ACT_NONE else dy
                scale = ops.fold_bn(cout_p, bk, bnk, x.device)[0] if bnk is not None else None
                dxs, dw, db = ops.conv2d_grads(x, wk, dz, dilation=dilation, scale=scale, x2=x2, need_dx=need_dx)
                if bnk is not None:
                    dgam, dbet = ops.bn_param_grad(wk, dw, bk, db, bnk)            # dw, db: raw -> true gradients, in place
                    s.pacc(f"{self._p}{bn}/gamma", dgam[:cout]); s.pacc(f"{self._p}{bn}/beta", dbet[:cout])
                if padded:                                                         # back to the Keras shapes
                    dwr = dw[:, :, rows, :cout]
                    dw = torch.cat([dwr, dw[:, :, cx:, :cout]], 2) if c2 else dwr
                    db = db[:cout]
                s.pacc(pre + "kernel", dw.contiguous()); s.pacc(pre + "bias", db.contiguous())
                if need_dx:
                    if x2 is None:
                        s.gacc(x, dxs)
                    else:
                        s.gacc(x, dxs[0]); s.gacc(x2, dxs[1])
            s.tape.append(bwd)
        return y

    def _ln(self, x, name, act=ACT_LRELU, coff=0, c=None):
        """LayerNormalization + activation IN PLACE on channels [coff, coff+c) of x"""
        cc = x.shape[3] if c is None else c
        g = self._s.vector(f"{self._p}{name}/gamma", cc, 1.0)
        b = self._s.vector(f"{self._p}{name}/beta", cc, 0.0)
        x_in = x[..., coff:coff + cc].clone(memory_format=torch.contiguous_format) if self._s.recording else None    # the LN input
        y = ops.layernorm_c(x, g, b, act=act, inplace=True, coff=coff, c=c)
        if self._s.recording:
            s = self._s

            def bwd():
                dy = s.gget(x, coff, cc)
                if dy is None:
                    return
                dx, dg, db = ops.layernorm_c_bwd(x_in, x[..., coff:coff + cc].contiguous(), dy, g, act=act)
                s.pacc(f"{self._p}{name}/gamma", dg); s.pacc(f"{self._p}{name}/beta", db)
                s.gset(x, dx, coff)        # in-place layer: the slice's gradient becomes the gradient w.r.t. the LN input
            s.tape.append(bwd)
        return y

    def _pool(self, x):
        y = ops.avgpool2x2(x)
        if self._s.recording:
            s = self._s

            def bwd():
                dy = s.gget(y)
                if dy is not None:
                    s.gacc(x, ops.avgpool2x2_bwd(dy))
            s.tape.append(bwd)
        return y


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

    def features(self, x, out, out_coff, out_slot=None):
        """U of this cardinal written into channels [out_coff, out_coff+cvkk) of ``out`` (a slot of out_slot >= cvkk channels)"""
        y = self._conv(x, "conv1", 1, self.cv11, dilation=self.atrous)       # stored with cv11 padded to a multiple of 16 (bf16)
        y = self._ln(y, "conv1_bn", c=self.cv11)
        self._conv(y, "conv2", self.ksize, self.cvkk, dilation=self.atrous, out=out, out_coff=out_coff, out_slot=out_slot, cin=self.cv11)
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
        cp = ops.pad_channels(c, x.dtype)                                          # slot width of one cardinal in u / v (bf16: multiple of 16)
        u = torch.zeros(n, h, w, K * cp, dtype=x.dtype, device=x.device)           # the K cardinals write their slot: no concat pass
        for k, blk in enumerate(self.cardinal_blocks):
            blk.features(x, u, k * cp, cp)
        per = [blk.split.params() for blk in self.cardinal_blocks]
        stacked = [torch.stack([p[i] for p in per]).contiguous() for i in range(6)]
        v, att = ops.splitatt_shared(u, K, self.radix, *stacked, act=ACT_LRELU, return_att=True, c=c)
        if self._s.recording:
            s = self._s

            def bwd():
                dv = s.gget(v)
                if dv is None:
                    return
                du, g = ops.splitatt_shared_bwd(u, dv, att, K, self.radix, *stacked[:5], act=ACT_LRELU, c=c)
                s.gacc(u, du)
                c2 = c // 2
                for k_, blk in enumerate(self.cardinal_blocks):
                    q = blk.split._p
                    s.pacc(q + "dense1/kernel", g["w1"][k_].reshape(1, 1, c, c2)); s.pacc(q + "dense1/bias", g["b1"][k_])
                    s.pacc(q + "dense1_bn/gamma", g["ln_gamma"][k_]); s.pacc(q + "dense1_bn/beta", g["ln_beta"][k_])
                    s.pacc(q + "dense2/kernel", g["w2"][k_].reshape(1, 1, c2, c)); s.pacc(q + "dense2/bias", g["b2"][k_])
            s.tape.append(bwd)
        sc = self._conv(x, "convtmp_sc", 1, self.outchannel, dilation=self.atrous)
        sc = self._ln(sc, "convtmp_scbn")
        cmap = [k * cp + i for k in range(K) for i in range(c)] if cp != c else None      # logical channel -> position in the slotted record
        return self._conv(v, "concats_2", self.ksize, self.outchannel, dilation=self.atrous, residual=sc, cin=K * c, cmap=cmap)

    __call__ = forward


class ResNest(_Layer):
    """ResNest.py:4-58."""

    def __init__(self, height, width, channel, ksize, radix=4, kpaths=4, wDecay=None, *, dtype="bf16", device="cuda", seed=0,
                 _store=None, _prefix=""):
        store = _store if _store is not None else VariableStore(device, seed)
        super().__init__(store, _prefix)
        self.height, self.width, self.channel, self.ksize, self.radix, self.kpaths, self.wDecay = height, width, channel, ksize, radix, kpaths, wDecay
        self.tdtype = torch.bfloat16 if dtype in ("bf16", torch.bfloat16) else torch.float32
        self.device = torch.device(device)
        mk = lambda name, out: residual_S(ksize=ksize, outchannel=out, radix=radix, kpaths=kpaths, wDecay=wDecay, _store=store, _prefix=_prefix + name + "/")
        self.conv_1, self.conv_2, self.conv_3, self.conv_4 = mk("conv_1", 64), mk("conv_2", 128), mk("conv_3", 256), mk("conv_4", 512)

    def load_variables(self, variables):
        self._s.load(variables)

    def variables(self):
        return OrderedDict(self._s.vars)

    def forward(self, x, record=False):
        """record=True keeps what backward() needs (the tape of this call)"""
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(device=self.device, dtype=self.tdtype).contiguous()
        cin = x.shape[3]
        cpad = ops.pad_channels(cin, self.tdtype)
        if cpad != cin:                                    # the 10 input planes are stored as 16-channel records (zero pad lanes)
            xp = torch.zeros(*x.shape[:3], cpad, dtype=x.dtype, device=x.device)
            xp[..., :cin] = x
            x = xp
        if record:
            self._s.start_recording()
        x0 = x
        self._cin = cin
        x = self._conv(x, "initial_conv", 3, 16, act=ACT_LRELU, cin=cin)
        x = self._conv(x, "convtmp_1", 3, 32, bn="convtmp_1bn", act=ACT_LRELU)
        x = self._conv(x, "convtmp_2", 3, 32, bn="convtmp_2bn", act=ACT_LRELU)
        x_1 = self.conv_1(self._pool(x))
        x_2 = self.conv_2(self._pool(x_1))
        x_3 = self.conv_3(self._pool(x_2))
        x_4 = self.conv_4(self._pool(x_3))
        self._io = (x0, x_4, [x_3, x_2, x_1])
        return x_4, [x_3, x_2, x_1]

    def backward(self, dx_4, dfeatures=None):
        """gradients of the four outputs of the last forward(x, record=True) -> dL/dx; parameter gradients in gradients()"""
        if not self._s.recording:
            raise RuntimeError("ResNest.backward: call forward(x, record=True) first")
        x0, x_4, feats = self._io
        dev = lambda t: torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t).to(device=self.device, dtype=self.tdtype).contiguous()
        self._s.gacc(x_4, dev(dx_4))
        for f, d in zip(feats, dfeatures or []):
            if d is not None:
                self._s.gacc(f, dev(d))
        self._s.run_backward()
        g = self._s.gget(x0)
        return g if g is None or g.shape[3] == self._cin else g[..., :self._cin].contiguous()

    def gradients(self):
        """variable name -> fp32 gradient of the last backward() (trainable variables only: no moving statistics)"""
        return OrderedDict(self._s.grads)

    def __call__(self, x, *args, **kwargs):
        return self.forward(x)
