"""Functional wrappers over the C ABI (include/tbi_sm100.h) for torch CUDA tensors.

Each function allocates its outputs with torch and enqueues libtbi_sm100.so kernels on the current
stream.  Layouts are the reference's (Keras): activations NHWC, Conv2D kernels HWIO (grouped:
[k,k,cin/groups,cout]), Conv2DTranspose kernels HWOI.  Used by the parity tests and microbenches;
the model executor (engine.py) calls the same C entry points with prebuilt descriptors.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT_ELU, ACT_NONE, BF16, F32, IMPL_AUTO, Epilogue, SplitAtt, View, check

BN_EPS = 1e-3


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported storage dtype {t.dtype}")


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def view(t: torch.Tensor, c: Optional[int] = None, coff: int = 0) -> View:
    assert t.is_cuda and t.dim() == 4 and t.is_contiguous(), "NHWC contiguous CUDA tensor expected"
    _, h, w, ct = t.shape
    return View(t.data_ptr(), h, w, ct if c is None else c, ct, coff)


def _vp(v: Optional[View]):
    return None if v is None else C.byref(v)


def _p(t):
    return None if t is None else t.data_ptr()


def _f32(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def fold_bn(cout, bias, bn, device):
    """-> (scale or None, folded bias) on device, via tbi_bn_fold"""
    L = _lib.lib()
    scale = torch.empty(cout, dtype=torch.float32, device=device)
    fbias = torch.empty(cout, dtype=torch.float32, device=device)
    b = _f32(bias)
    if bn is not None:
        g, be, m, v = (_f32(t) for t in bn)
        check(L.tbi_bn_fold(cout, _p(g), _p(be), _p(m), _p(v), _p(b), BN_EPS, _p(scale), _p(fbias), _st()), "bn_fold")
        return scale, fbias
    check(L.tbi_bn_fold(cout, None, None, None, None, _p(b), BN_EPS, _p(scale), _p(fbias), _st()), "bn_fold")
    return None, fbias


def pack_conv(w_hwio: torch.Tensor, groups: int, mode: int, dtype: torch.dtype, scale=None) -> torch.Tensor:
    L = _lib.lib()
    k, _, cin_g, cout = w_hwio.shape
    w32 = _f32(w_hwio)
    dtc = F32 if dtype == torch.float32 else BF16
    out = torch.empty(int(L.tbi_conv_packed_elems(dtc, k, groups, cin_g, cout)), dtype=dtype, device=w32.device)
    check(L.tbi_pack_conv_weights(F32 if dtype == torch.float32 else BF16, mode, k, groups, cin_g, cout, _p(w32), _p(scale), _p(out), _st()), "pack_conv")
    return out


def pack_convt(w_hwoi: torch.Tensor, mode: int, dtype: torch.dtype, scale=None, cout_pad: int = 0) -> torch.Tensor:
    L = _lib.lib()
    k, _, cout, cin = w_hwoi.shape
    w32 = _f32(w_hwoi)
    cp = cout_pad if (mode == 1 and cout_pad > cout) else cout
    out = torch.empty(k * k * cin * cp, dtype=dtype, device=w32.device)
    check(L.tbi_pack_convt_weights(F32 if dtype == torch.float32 else BF16, mode, k, cin, cout, cout_pad, _p(w32), _p(scale), _p(out), _st()), "pack_convt")
    return out


def _epi(out, bias=None, act=ACT_NONE, residual=None, keep=None, dact=ACT_NONE, dact_ref=None, dact_keep=None,
         out2=None, split_c=0, residual2=None, out_f32=0) -> Epilogue:
    e = Epilogue()
    e.bias = _p(bias); e.act = act; e.out = view(out); e.out_stride = 1; e.out_f32 = out_f32
    e.drop_keep = _p(keep)
    if residual is not None:
        e.residual = view(residual)
    e.dact = dact
    if dact_ref is not None:
        e.dact_ref = view(dact_ref)
    e.dact_keep = _p(dact_keep)
    if out2 is not None:
        e.out2 = view(out2); e.split_c = split_c
    if residual2 is not None:
        e.residual2 = view(residual2)
    return e


# ------------------------------------------------------------------------------------------ Conv2D
def conv2d(x, w_hwio, bias=None, *, dilation=1, groups=1, bn=None, act=ACT_NONE, residual=None, x2=None,
           impl=IMPL_AUTO, out_f32=False, out=None, out_coff=0):
    """Conv2D(strides=1, padding='SAME') [+ BN-inference] [+ act] [+ residual]; x2 = second source of a
    virtual channel concat.  (TBI_ResNest.py:83-91,140-148,162-170)"""
    L = _lib.lib()
    n, h, w, _ = x.shape
    k, cout = w_hwio.shape[0], w_hwio.shape[3]
    scale, fbias = fold_bn(cout, bias if bias is not None else torch.zeros(cout, device=x.device), bn, x.device)
    wp = pack_conv(w_hwio, groups, 0, x.dtype, scale)
    y = out if out is not None else torch.empty(n, h, w, cout, dtype=torch.float32 if out_f32 else x.dtype, device=x.device)
    e = _epi(y, fbias, act, residual, out_f32=int(out_f32))
    if out is not None:                       # write channels [out_coff, out_coff+cout) of a wider tensor (concat-free)
        e.out = view(out, c=cout, coff=out_coff)
    check(L.tbi_conv2d_fwd(_dt(x), impl, n, h, w, k, dilation, groups, _vp(view(x)), _vp(view(x2)) if x2 is not None else None,
                           cout, _p(wp), C.byref(e), _st()), "conv2d_fwd")
    return y


def conv2d_grads(x, w_hwio, dz, *, dilation=1, groups=1, scale=None, x2=None, impl=IMPL_AUTO, need_dx=True,
                 dact=ACT_NONE, dact_ref=None, wgrad_impl=None):
    """gradients of the conv w.r.t. (input[s], HWIO kernel, bias) given dz = dL/d(conv output, pre-activation).
    ``scale`` = folded BN scale the forward used (dx flows through W*scale; dw/db returned are RAW: A^T dz)."""
    L = _lib.lib()
    n, h, w, c0 = x.shape
    k, _, cin_g, cout = w_hwio.shape
    cin = c0 + (x2.shape[3] if x2 is not None else 0)
    dw = torch.zeros(w_hwio.shape, dtype=torch.float32, device=x.device)
    db = torch.zeros(cout, dtype=torch.float32, device=x.device)
    wsb = int(L.tbi_conv2d_wgrad_workspace(_dt(x), k, groups, cin, cout))
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device) if wsb else None
    check(L.tbi_conv2d_wgrad(_dt(x), impl if wgrad_impl is None else wgrad_impl, n, h, w, k, dilation, groups, _vp(view(x)), _vp(view(x2)) if x2 is not None else None,
                             _vp(view(dz)), _p(dw), _p(db), _p(ws), wsb, _st()), "conv2d_wgrad")
    if not need_dx:
        return None, dw, db
    wb = pack_conv(w_hwio, groups, 1, x.dtype, scale)
    dx = torch.empty_like(x)
    dx2 = torch.empty_like(x2) if x2 is not None else None
    e = _epi(dx, dact=dact, dact_ref=dact_ref, out2=dx2, split_c=c0 if x2 is not None else 0)
    check(L.tbi_conv2d_dgrad(_dt(x), impl, n, h, w, k, dilation, groups, _vp(view(dz)), cin, _p(wb), C.byref(e), _st()), "conv2d_dgrad")
    return (dx if x2 is None else (dx, dx2)), dw, db


# ------------------------------------------------------------------------------------------ Conv2DTranspose s2
def conv2d_transpose_s2(x, w_hwoi, bias=None, *, bn=None, act=ACT_NONE, keep=None, x2=None, impl=IMPL_AUTO, out_f32=False):
    """Conv2DTranspose(k in {3,4}, strides=2, padding='same') [+BN] [+dropout multiplier] [+act]
    (TBI_ResNest.py:124,209-220; Decoder.py:57-59,120)"""
    L = _lib.lib()
    n, h, w, _ = x.shape
    k, cout = w_hwoi.shape[0], w_hwoi.shape[2]
    scale, fbias = fold_bn(cout, bias if bias is not None else torch.zeros(cout, device=x.device), bn, x.device)
    wp = pack_convt(w_hwoi, 0, x.dtype, scale)
    y = torch.empty(n, 2 * h, 2 * w, cout, dtype=torch.float32 if out_f32 else x.dtype, device=x.device)
    e = _epi(y, fbias, act, keep=keep, out_f32=int(out_f32))
    check(L.tbi_conv2d_transpose_s2_fwd(_dt(x), impl, n, h, w, k, _vp(view(x)), _vp(view(x2)) if x2 is not None else None, cout,
                                        _p(wp), C.byref(e), _st()), "convT_fwd")
    return y


def conv2d_transpose_s2_grads(x, w_hwoi, dz, *, scale=None, x2=None, impl=IMPL_AUTO, need_dx=True, wgrad_impl=None):
    """dz may carry more channels than the kernel's cout (zero-padded gradient records, e.g. the head)."""
    L = _lib.lib()
    n, h, w, c0 = x.shape
    k, _, cout, cin = w_hwoi.shape
    cpad = dz.shape[3]
    dw = torch.zeros(w_hwoi.shape, dtype=torch.float32, device=x.device)
    db = torch.zeros(cout, dtype=torch.float32, device=x.device)
    check(L.tbi_conv2d_transpose_s2_wgrad(_dt(x), impl if wgrad_impl is None else wgrad_impl, n, h, w, k, _vp(view(x)), _vp(view(x2)) if x2 is not None else None,
                                          _vp(view(dz)), cout, _p(dw), _p(db), None, 0, _st()), "convT_wgrad")
    if not need_dx:
        return None, dw, db
    wb = pack_convt(w_hwoi, 1, x.dtype, scale, cout_pad=cpad)
    dx = torch.empty_like(x)
    dx2 = torch.empty_like(x2) if x2 is not None else None
    e = _epi(dx, out2=dx2, split_c=c0 if x2 is not None else 0)
    check(L.tbi_conv2d_transpose_s2_dgrad(_dt(x), impl, n, h, w, k, _vp(view(dz)), cin, _p(wb), C.byref(e), _st()), "convT_dgrad")
    return (dx if x2 is None else (dx, dx2)), dw, db


def bn_param_grad(w, dw_raw, bias, dbias_raw, bn, kind="conv"):
    """in place: dw_raw -> dw, dbias_raw -> dbias; returns (dgamma, dbeta).  w: HWIO (conv) or HWOI (convt)."""
    L = _lib.lib()
    g, _, m, v = (_f32(t) for t in bn)
    if kind == "conv":
        k, _, cin_g, cout = w.shape
        lay = (k * k * cin_g, 1, cout, 1)
    else:
        k, _, cout, cin = w.shape
        lay = (k * k, cin, cout * cin, cin)
    dgamma = torch.zeros(cout, dtype=torch.float32, device=w.device)
    dbeta = torch.zeros(cout, dtype=torch.float32, device=w.device)
    w32, b32 = _f32(w), _f32(bias)
    check(L.tbi_bn_param_grad(cout, *lay, _p(w32), _p(dw_raw), _p(b32), _p(dbias_raw), _p(g), _p(m), _p(v), BN_EPS,
                              _p(dgamma), _p(dbeta), _st()), "bn_param_grad")
    return dgamma, dbeta


# ------------------------------------------------------------------------------------------ pooling etc.
def avgpool2x2(x):
    L = _lib.lib()
    n, h, w, c = x.shape
    y = torch.empty(n, h // 2, w // 2, c, dtype=x.dtype, device=x.device)
    check(L.tbi_avgpool2x2_fwd(_dt(x), n, h, w, _vp(view(x)), _vp(view(y)), _st()), "avgpool_fwd")
    return y


def avgpool2x2_bwd(dy, *, dact=ACT_NONE, dact_ref=None, accumulate_into=None):
    L = _lib.lib()
    n, ho, wo, c = dy.shape
    dx = accumulate_into if accumulate_into is not None else torch.empty(n, 2 * ho, 2 * wo, c, dtype=dy.dtype, device=dy.device)
    check(L.tbi_avgpool2x2_bwd(_dt(dy), n, 2 * ho, 2 * wo, _vp(view(dy)), _vp(view(dx)), int(accumulate_into is not None), dact,
                               _vp(view(dact_ref)) if dact_ref is not None else None, _st()), "avgpool_bwd")
    return dx


def act_bwd(dy, y_ref, act, keep=None):
    L = _lib.lib()
    n, h, w, c = dy.shape
    dz = torch.empty_like(dy)
    check(L.tbi_act_bwd(_dt(dy), n * h * w, act, _vp(view(dy)), _vp(view(y_ref)), _p(keep), _vp(view(dz)), _st()), "act_bwd")
    return dz


def colsum(x):
    L = _lib.lib()
    n, h, w, c = x.shape
    out = torch.zeros(c, dtype=torch.float32, device=x.device)
    check(L.tbi_colsum(_dt(x), n * h * w, _vp(view(x)), _p(out), _st()), "colsum")
    return out


class SplitAttention:
    """Radix split-attention tail for K cardinal groups at once (TBI_ResNest.py:175-207).
    u: [n,h,w,K*R*c] (channel order k,r,c) -> v: [n,h,w,K*c].  Parameters are fp32 tensors:
    w1 [K,c,c/2], b1 [K,c/2], gamma/beta/mean/var [K,c/2], w2 [K,R,c/2,c], b2 [K,R,c]."""

    def __init__(self, kpaths, radix, c, w1, b1, gamma, beta, mean, var, w2, b2, act=ACT_ELU):
        self.K, self.R, self.c, self.act = kpaths, radix, c, act
        self.params = [_f32(t) for t in (w1, b1, gamma, beta, mean, var, w2, b2)]

    def _desc(self, u):
        n, h, w, _ = u.shape
        dev = u.device
        self.gap = torch.empty(n, self.K, self.c, dtype=torch.float32, device=dev)
        self.h1 = torch.empty(n, self.K, self.c // 2, dtype=torch.float32, device=dev)
        self.att = torch.empty(n, self.K, self.R, self.c, dtype=torch.float32, device=dev)
        return SplitAtt(_dt(u), n, h, w, self.K, self.R, self.c, self.act, BN_EPS, *[_p(t) for t in self.params],
                        _p(self.gap), _p(self.h1), _p(self.att))

    def forward(self, u, out=None):
        L = _lib.lib()
        n, h, w, _ = u.shape
        self.desc = self._desc(u)
        v = torch.empty(n, h, w, self.K * self.c, dtype=u.dtype, device=u.device) if out is None else out
        check(L.tbi_split_attention_fwd(C.byref(self.desc), _vp(view(u)), _vp(view(v)), _st()), "split_attention_fwd")
        return v

    def new_grads(self, dev):
        K, R, c = self.K, self.R, self.c
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        return dict(w1=z(K, c, c // 2), b1=z(K, c // 2), gamma=z(K, c // 2), beta=z(K, c // 2), w2=z(K, R, c // 2, c), b2=z(K, R, c))

    def backward(self, u, dv, out=None, grads=None, scratch=None):
        """-> (dz_u = dL/du * act'(u), dict of parameter gradients).  out / grads / scratch: caller-kept buffers (the parameter
        gradients are ADDED into grads, as the engine adds them into its flat gradient buffer)."""
        L = _lib.lib()
        n = u.shape[0]
        K, R, c = self.K, self.R, self.c
        dev = u.device
        du = torch.empty_like(u) if out is None else out
        g = self.new_grads(dev) if grads is None else grads
        if scratch is None:
            scratch = torch.empty(n * K * (R * c + 2 * c), dtype=torch.float32, device=dev)
        check(L.tbi_split_attention_bwd(C.byref(self.desc), _vp(view(u)), _vp(view(dv)), _vp(view(du)), _p(g["w1"]), _p(g["b1"]),
                                        _p(g["gamma"]), _p(g["beta"]), _p(g["w2"]), _p(g["b2"]), _p(scratch), _st()), "split_attention_bwd")
        return du, g


def softmax_loss(logits, y, dlogits_dtype=torch.float32):
    """softmax + my_loss_cat + accuracy count + dlogits (TBI_ResNest.py:125,234-248,48-51)"""
    L = _lib.lib()
    n, h, w, nc = logits.shape
    dev = logits.device
    probs = torch.empty_like(logits)
    loss = torch.empty(h, w, dtype=torch.float32, device=dev)
    correct = torch.zeros(1, dtype=torch.int32, device=dev)
    dlog = torch.empty(n, h, w, nc, dtype=dlogits_dtype, device=dev)
    check(L.tbi_softmax_loss_fwd_bwd(F32 if dlogits_dtype == torch.float32 else BF16, n, h, w, nc, _p(logits), _p(y), _p(probs),
                                     _p(loss), _p(correct), _p(dlog), nc, _st()), "softmax_loss")
    return probs, loss, correct, dlog


def softmax_loss_from_taps(ytaps, bias, y, dlogits_dtype=torch.float32):
    """the same with the logits formed in the kernel from the head's tap products ytaps fp32 [n, h/2, w/2, 16*nc]
    (column (ky*4+kx)*nc + c) + bias: tbi_softmax_loss_fwd_bwd_taps; y = labels [n, h, w, nc]"""
    L = _lib.lib()
    n, h, w, nc = y.shape
    dev = y.device
    probs = torch.empty(n, h, w, nc, dtype=torch.float32, device=dev)
    loss = torch.empty(h, w, dtype=torch.float32, device=dev)
    correct = torch.zeros(1, dtype=torch.int32, device=dev)
    dlog = torch.empty(n, h, w, nc, dtype=dlogits_dtype, device=dev)
    check(L.tbi_softmax_loss_fwd_bwd_taps(F32 if dlogits_dtype == torch.float32 else BF16, n, h, w, nc, _vp(view(ytaps)), _p(bias), _p(y), _p(probs),
                                          _p(loss), _p(correct), _p(dlog), nc, _st()), "softmax_loss_taps")
    return probs, loss, correct, dlog


def adam_step(p, g, m, v, step_count, lr, b1=0.9, b2=0.999, eps=1e-7, grad_scale=1.0):
    L = _lib.lib()
    check(L.tbi_adam_multi(p.numel(), _p(p), _p(g), _p(m), _p(v), _p(step_count), lr, b1, b2, eps, grad_scale, _st()), "adam")
    check(L.tbi_adam_advance(_p(step_count), _st()), "adam_advance")


# ------------------------------------------------------------------------------------------ Variant B ops
LN_EPS = 1e-3


def layernorm_c(x, gamma, beta, act=ACT_NONE, eps=LN_EPS, inplace=False, coff=0, c=None):
    """LayerNormalization over channels + activation (ResNest.py:86-87,99-101,126-133; Decoder.py:112-113,130-131).
    coff/c: normalise only the channel slice [coff, coff+c) of x (each slice of a concat buffer is its own layer)."""
    L = _lib.lib()
    n, h, w, ct = x.shape
    c = ct if c is None else c
    y = x if inplace else torch.empty_like(x)
    check(L.tbi_layernorm_c_fwd(_dt(x), n * h * w, c, _vp(view(x, c=c, coff=coff)), _p(_f32(gamma)), _p(_f32(beta)), eps, act,
                                _vp(view(y, c=c, coff=coff)), _st()), "layernorm_c_fwd")
    return y


def layernorm_c_bwd(x, y, dy, gamma, act=ACT_NONE, eps=LN_EPS):
    """-> (dx, dgamma, dbeta); x = LayerNorm input, y = its activated output"""
    L = _lib.lib()
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    dg = torch.zeros(c, dtype=torch.float32, device=x.device)
    db = torch.zeros(c, dtype=torch.float32, device=x.device)
    check(L.tbi_layernorm_c_bwd(_dt(x), n * h * w, c, _vp(view(x)), _vp(view(y)), _vp(view(dy)), _p(_f32(gamma)), eps, act,
                                _vp(view(dx)), _p(dg), _p(db), _st()), "layernorm_c_bwd")
    return dx, dg, db


def splitatt_shared(u, kpaths, radix, w1, b1, ln_gamma, ln_beta, w2, b2, act, eps=LN_EPS, return_att=False, c=None):
    """Variant-B split attention (ResNest.py:171-199): R identical inputs, shared dense2.  u: [n,h,w,K*cg] with the K cardinals'
    c channels at offsets k*cg (cg = c unless the slices are zero-padded to a 16-channel multiple, pass c then);
    w1 [K,c,c/2], b1 [K,c/2], ln_gamma/ln_beta [K,c/2], w2 [K,c/2,c], b2 [K,c] (fp32)."""
    L = _lib.lib()
    n, h, w, C_ = u.shape
    cg = C_ // kpaths
    c = cg if c is None else c
    v = torch.zeros_like(u) if cg != c else torch.empty_like(u)          # pad lanes are never written: they must read 0
    att = torch.empty(n, kpaths * c, dtype=torch.float32, device=u.device)
    args = [_f32(t) for t in (w1, b1, ln_gamma, ln_beta, w2, b2)]
    check(L.tbi_splitatt_shared_fwd(_dt(u), n, h, w, kpaths, radix, c, _vp(view(u)), _vp(view(v)), _p(args[0]), _p(args[1]), _p(args[2]),
                                    _p(args[3]), eps, act, _p(args[4]), _p(args[5]), _p(att), cg, _st()), "splitatt_shared_fwd")
    return (v, att) if return_att else v


def splitatt_shared_bwd(u, dv, att, kpaths, radix, w1, b1, ln_gamma, ln_beta, w2, act, eps=LN_EPS, c=None):
    """backward of splitatt_shared: -> (du, dict of parameter gradients stacked over the K cardinals like the parameters).
    att = the buffer the forward returned with return_att=True."""
    L = _lib.lib()
    n, h, w, C_ = u.shape
    cg = C_ // kpaths
    c = cg if c is None else c
    du = torch.zeros_like(u) if cg != c else torch.empty_like(u)
    args = [_f32(t) for t in (w1, b1, ln_gamma, ln_beta, w2)]
    z = lambda t: torch.zeros(t.shape, dtype=torch.float32, device=u.device)
    g = dict(w1=z(args[0]), b1=z(args[1]), ln_gamma=z(args[2]), ln_beta=z(args[3]), w2=z(args[4]),
             b2=torch.zeros(kpaths, c, dtype=torch.float32, device=u.device))
    scratch = torch.empty(2 * n * kpaths * c, dtype=torch.float32, device=u.device)
    check(L.tbi_splitatt_shared_bwd(_dt(u), n, h, w, kpaths, radix, c, _vp(view(u)), _vp(view(dv)), _vp(view(du)), _p(args[0]), _p(args[1]),
                                    _p(args[2]), _p(args[3]), eps, act, _p(args[4]), _p(att), _p(g["w1"]), _p(g["b1"]), _p(g["ln_gamma"]),
                                    _p(g["ln_beta"]), _p(g["w2"]), _p(g["b2"]), _p(scratch), cg, _st()), "splitatt_shared_bwd")
    return du, g


def pad_channels(c: int, dtype) -> int:
    """physical channel count of a stored activation with c logical channels: bf16 tensors are stored with their channels
    rounded up to a multiple of 16 (zero pad lanes) so that every pixel record is 16-byte aligned and a whole number of MMA
    K steps -- what keeps odd-width layers (Variant B: 3/7/10/21/30/63/85/126/255 channels) on the tcgen05 path"""
    return (c + 15) // 16 * 16 if dtype == torch.bfloat16 else c


# ------------------------------------------------------------------------------------------ ViT bridge (VisionTransformer.py)
def attention(q, k, v, heads, scale):
    """softmax(q k^T * scale) v per head; q, k, v: [n, tokens, heads*d] -> (ctx [n, tokens, heads*d], probs fp32 [n, heads, tokens, tokens])"""
    L = _lib.lib()
    n, t, c = q.shape
    ctx = torch.empty_like(q)
    probs = torch.empty(n, heads, t, t, dtype=torch.float32, device=q.device)
    check(L.tbi_attention_fwd(_dt(q), n, t, heads, c // heads, float(scale), _p(q), _p(k), _p(v), _p(ctx), _p(probs), _st()), "attention_fwd")
    return ctx, probs


def attention_bwd(q, k, v, probs, dctx, heads, scale):
    L = _lib.lib()
    n, t, c = q.shape
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    check(L.tbi_attention_bwd(_dt(q), n, t, heads, c // heads, float(scale), _p(q), _p(k), _p(v), _p(probs), _p(dctx.contiguous()), _p(dq), _p(dk), _p(dv), _st()),
          "attention_bwd")
    return dq, dk, dv


def gelu(x):
    L = _lib.lib()
    y = torch.empty_like(x)
    check(L.tbi_gelu_fwd(_dt(x), x.numel(), _p(x), _p(y), _st()), "gelu_fwd")
    return y


def gelu_bwd(x, dy):
    L = _lib.lib()
    dx = torch.empty_like(x)
    check(L.tbi_gelu_bwd(_dt(x), x.numel(), _p(x), _p(dy.contiguous()), _p(dx), _st()), "gelu_bwd")
    return dx


def softmax_cce(logits, y, label_smoothing, global_batch, need_grad=True):
    """softmax + label-smoothed CategoricalCrossentropy / global_batch (VisionTransformer.py:205-206,225-227)
    -> (probs, scalar loss tensor, dlogits or None)"""
    L = _lib.lib()
    nc = logits.shape[-1]
    npix = logits.numel() // nc
    probs = torch.empty_like(logits)
    loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
    dlog = torch.empty_like(logits) if need_grad else None
    check(L.tbi_softmax_cce_fwd_bwd(npix, nc, float(label_smoothing), float(global_batch), _p(logits), _p(y), _p(probs), _p(loss), _p(dlog), _st()), "softmax_cce")
    return probs, loss, dlog
