"""Per-kernel parity: every C-ABI op against the CPU oracle ops (torch fp64 on the host) on seeded inputs.
Bars (north_star): fp32 storage 1e-4, bf16 storage 2e-2, relative to the reference tensor's max-abs."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import tbi_resnest_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
DTYPES = [torch.float32, torch.bfloat16]


@pytest.fixture(scope="module")
def ops(cuda_device):
    from ultrasound_modeling_b200 import ops as _ops
    return _ops


def rel(got, want):
    want = want.detach().double().cpu()
    got = got.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


def q(t, dtype):
    """quantise a host fp64 tensor to the storage dtype and back (so both sides see identical inputs)"""
    return t.to(dtype).double()


def dev(t, dtype, device="cuda"):
    return t.to(dtype).to(device).contiguous()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", [
    dict(n=2, h=9, w=7, cin=5, cout=11, k=3, d=1, g=1),
    dict(n=1, h=12, w=10, cin=8, cout=12, k=3, d=1, g=2),
    dict(n=2, h=17, w=16, cin=6, cout=70, k=3, d=2, g=1),
    dict(n=1, h=20, w=9, cin=3, cout=4, k=3, d=4, g=1),
    dict(n=2, h=8, w=8, cin=32, cout=48, k=1, d=1, g=1),
    dict(n=3, h=6, w=5, cin=1, cout=16, k=3, d=1, g=1),
    dict(n=1, h=24, w=24, cin=4, cout=8, k=3, d=8, g=1),
])
def test_conv2d_fwd_and_grads(ops, dtype, case):
    torch.manual_seed(1)
    n, h, w, cin, cout, k, d, g = (case[x] for x in "n h w cin cout k d g".split())
    x = q(torch.randn(n, h, w, cin, dtype=torch.float64), dtype)
    wt = torch.randn(k, k, cin // g, cout, dtype=torch.float64) * (1 / math.sqrt(k * k * cin / g))
    b = torch.randn(cout, dtype=torch.float64) * 0.1
    bn = [1 + 0.1 * torch.randn(cout, dtype=torch.float64), 0.1 * torch.randn(cout, dtype=torch.float64),
          0.1 * torch.randn(cout, dtype=torch.float64), 0.5 + torch.rand(cout, dtype=torch.float64)]
    res = q(torch.randn(n, h, w, cout, dtype=torch.float64), dtype)
    xr = x.clone().requires_grad_(True); wr = wt.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    conv = F.conv2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), br, padding=(k // 2) * d, dilation=d, groups=g).permute(0, 2, 3, 1)
    z = O.batchnorm_inference(conv, *bn)
    want = F.elu(z) + res
    got = ops.conv2d(dev(x, dtype), dev(wt, torch.float32), dev(b, torch.float32), dilation=d, groups=g,
                     bn=[dev(t, torch.float32) for t in bn], act=ops.ACT_ELU, residual=dev(res, dtype))
    assert rel(got, want) < TOL[dtype]
    # gradients w.r.t. the pre-activation z -> conv input / raw weight grads / BN params
    dz = q(torch.randn(n, h, w, cout, dtype=torch.float64), dtype)
    gx, gw, gb = torch.autograd.grad((z * dz).sum(), [xr, wr, br])
    scale = (bn[0] / torch.sqrt(bn[3] + O.BN_EPS))
    dx, dw_raw, db_raw = ops.conv2d_grads(dev(x, dtype), dev(wt, torch.float32), dev(dz, dtype), dilation=d, groups=g,
                                          scale=dev(scale, torch.float32))
    assert rel(dx, gx) < TOL[dtype] * 2
    dgamma, dbeta = ops.bn_param_grad(dev(wt, torch.float32), dw_raw, dev(b, torch.float32), db_raw,
                                      [dev(t, torch.float32) for t in bn], "conv")
    gam = bn[0].clone().requires_grad_(True); bet = bn[1].clone().requires_grad_(True)
    z2 = O.batchnorm_inference(conv.detach(), gam, bet, bn[2], bn[3])
    ggam, gbet = torch.autograd.grad((z2 * dz).sum(), [gam, bet])
    assert rel(dw_raw, gw) < TOL[dtype] * 2 and rel(db_raw, gb) < TOL[dtype] * 2
    assert rel(dgamma, ggam) < TOL[dtype] * 4 and rel(dbeta, gbet) < TOL[dtype] * 2


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv2d_two_sources_is_concat(ops, dtype):
    torch.manual_seed(2)
    x1 = q(torch.randn(2, 8, 6, 12, dtype=torch.float64), dtype); x2 = q(torch.randn(2, 8, 6, 5, dtype=torch.float64), dtype)
    wt = torch.randn(3, 3, 17, 9, dtype=torch.float64) * 0.1
    xc = torch.cat([x1, x2], 3).requires_grad_(True)
    want = O.conv2d_same(xc, wt, None)
    got = ops.conv2d(dev(x1, dtype), dev(wt, torch.float32), None, x2=dev(x2, dtype))
    assert rel(got, want) < TOL[dtype]
    dz = q(torch.randn_like(want), dtype)
    gx, = torch.autograd.grad((want * dz).sum(), [xc])
    (dx1, dx2), dw, _ = ops.conv2d_grads(dev(x1, dtype), dev(wt, torch.float32), dev(dz, dtype), x2=dev(x2, dtype))
    assert rel(dx1, gx[..., :12]) < TOL[dtype] * 2 and rel(dx2, gx[..., 12:]) < TOL[dtype] * 2


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k", [4, 3])
@pytest.mark.parametrize("two", [False, True])
def test_conv2d_transpose_s2(ops, dtype, k, two):
    torch.manual_seed(3 + k)
    n, h, w, c1, c2, cout = 2, 5, 7, 10, 6, 9
    x1 = q(torch.randn(n, h, w, c1, dtype=torch.float64), dtype)
    x2 = q(torch.randn(n, h, w, c2, dtype=torch.float64), dtype) if two else None
    cin = c1 + (c2 if two else 0)
    wt = torch.randn(k, k, cout, cin, dtype=torch.float64) * 0.1
    b = torch.randn(cout, dtype=torch.float64) * 0.1
    bn = [1 + 0.1 * torch.randn(cout, dtype=torch.float64), 0.1 * torch.randn(cout, dtype=torch.float64),
          0.1 * torch.randn(cout, dtype=torch.float64), 0.5 + torch.rand(cout, dtype=torch.float64)]
    keep01 = (torch.rand(n, 2 * h, 2 * w, cout) < 0.5).to(torch.uint8)
    xc = (torch.cat([x1, x2], 3) if two else x1).clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    z = O.batchnorm_inference(O.conv2d_transpose_s2_same(xc, wr, br), *bn)
    want = F.relu(z * keep01.double() * 2)
    got = ops.conv2d_transpose_s2(dev(x1, dtype), dev(wt, torch.float32), dev(b, torch.float32),
                                  bn=[dev(t, torch.float32) for t in bn], act=ops._lib.ACT_RELU, keep=(keep01 * 2).cuda(),
                                  x2=dev(x2, dtype) if two else None)
    assert rel(got, want) < TOL[dtype]
    dz = q(torch.randn_like(z), dtype)
    gx, gw, gb = torch.autograd.grad((z * dz).sum(), [xc, wr, br])
    scale = bn[0] / torch.sqrt(bn[3] + O.BN_EPS)
    dx, dw_raw, db_raw = ops.conv2d_transpose_s2_grads(dev(x1, dtype), dev(wt, torch.float32), dev(dz, dtype),
                                                       scale=dev(scale, torch.float32), x2=dev(x2, dtype) if two else None)
    if two:
        assert rel(dx[0], gx[..., :c1]) < TOL[dtype] * 2 and rel(dx[1], gx[..., c1:]) < TOL[dtype] * 2
    else:
        assert rel(dx, gx) < TOL[dtype] * 2
    ops.bn_param_grad(dev(wt, torch.float32), dw_raw, dev(b, torch.float32), db_raw, [dev(t, torch.float32) for t in bn], "convt")
    assert rel(dw_raw, gw) < TOL[dtype] * 2 and rel(db_raw, gb) < TOL[dtype] * 2


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c", [3, 32])
def test_avgpool_and_act_bwd(ops, dtype, c):
    torch.manual_seed(4)
    x = q(torch.randn(2, 8, 6, c, dtype=torch.float64), dtype).requires_grad_(True)
    want = O.avgpool2(F.elu(x))
    y = F.elu(x).detach()
    got = ops.avgpool2x2(dev(y, dtype))
    assert rel(got, want) < TOL[dtype]
    dy = q(torch.randn_like(want), dtype)
    gx, = torch.autograd.grad((want * dy).sum(), [x])
    yq = q(y, dtype)
    dx = ops.avgpool2x2_bwd(dev(dy, dtype), dact=ops.ACT_ELU, dact_ref=dev(y, dtype))
    # derivative through the QUANTISED activation output, as the kernel sees it
    dref = dy.repeat_interleave(2, 1).repeat_interleave(2, 2) / 4 * torch.where(yq > 0, torch.ones_like(yq), yq + 1)
    assert rel(dx, dref) < TOL[dtype]
    assert rel(dx, gx) < 3e-2
    acc = dev(torch.ones(2, 8, 6, c, dtype=torch.float64), dtype)
    ops.avgpool2x2_bwd(dev(dy, dtype), accumulate_into=acc)
    assert rel(acc, 1 + dy.repeat_interleave(2, 1).repeat_interleave(2, 2) / 4) < TOL[dtype]
    dz = ops.act_bwd(dev(dy.repeat_interleave(2, 1).repeat_interleave(2, 2), dtype), dev(y, dtype), ops.ACT_ELU)
    assert rel(dz, dref * 4) < TOL[dtype]
    assert rel(ops.colsum(dev(y, dtype)), yq.sum((0, 1, 2))) < 1e-4


def _splitatt_ref(us, P, K, R, c, act=F.elu):
    """oracle split_attention (TBI_ResNest.py:175-207) for K cardinals with stacked params"""
    outs = []
    for k in range(K):
        ins = us[k]
        g = sum(ins).mean(dim=(1, 2))
        qv = g @ P["w1"][k] + P["b1"][k]
        bn = (qv - P["mean"][k]) * P["gamma"][k] / torch.sqrt(P["var"][k] + O.BN_EPS) + P["beta"][k]
        h1 = act(bn)
        o = 0
        for r in range(R):
            z = h1 @ P["w2"][k, r] + P["b2"][k, r]
            a = torch.sigmoid(z) if R == 1 else torch.softmax(z, -1)
            o = o + ins[r] * a[:, None, None, :]
        outs.append(o)
    return torch.cat(outs, 3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("K,R,c", [(1, 2, 32), (4, 4, 8), (2, 1, 16), (3, 3, 10), (1, 2, 256)])
def test_split_attention_fwd_bwd(ops, dtype, K, R, c):
    _check_split_attention(ops, dtype, K, R, c, 3, 6, 5)


# the cluster-per-image kernels (csrc/splitatt_fused.cu): cluster sizes 1..16, both thread counts, chunk kept in shared memory
# or re-read from L2, images whose pixel count does not divide by the cluster size, FC layers redundant per CTA (slice 1) or
# sliced over the cluster (slice 2) incl. slices that are empty or larger than the thread count, weights staged in shared
# memory or read from global (stage 0, or too large to stage)
@pytest.mark.parametrize("K,R,c,h,w,cs,nt,cache_kb,slice_mode,stage", [
    (1, 2, 32, 16, 16, 8, 256, 100, 0, 1), (1, 2, 32, 40, 33, 8, 512, 0, 2, 1), (4, 4, 8, 9, 7, 2, 256, 100, 2, 0),
    (1, 2, 64, 24, 20, 16, 256, 100, 0, 1), (2, 1, 16, 12, 12, 4, 256, 0, 1, 0), (1, 3, 16, 20, 20, 8, 512, 100, 2, 1),
    (1, 2, 256, 8, 8, 4, 256, 100, 0, 1), (1, 2, 256, 5, 3, 1, 256, 0, 0, 1), (1, 2, 128, 32, 32, 0, 0, 100, 0, 1),
    (4, 3, 16, 16, 16, 0, 0, 0, 0, 1), (1, 1, 8, 16, 16, 8, 256, 100, 2, 1), (1, 2, 256, 16, 16, 8, 256, 100, 0, 1),
    (4, 4, 64, 16, 16, 8, 256, 0, 0, 1), (1, 2, 64, 64, 64, 0, 0, 100, 0, 1)])
def test_split_attention_cluster_kernels(ops, monkeypatch, K, R, c, h, w, cs, nt, cache_kb, slice_mode, stage):
    monkeypatch.setenv("TBI_SA_CS", str(cs))
    monkeypatch.setenv("TBI_SA_NT", str(nt))
    monkeypatch.setenv("TBI_SA_CACHE_KB", str(cache_kb))
    monkeypatch.setenv("TBI_SA_SLICE", str(slice_mode))
    monkeypatch.setenv("TBI_SA_STAGE", str(stage))
    _check_split_attention(ops, torch.bfloat16, K, R, c, 5, h, w)


@pytest.mark.parametrize("h,c", [(128, 32), (32, 128)])
def test_split_attention_full_size_properties(ops, h, c):
    """BASELINE config 2 sizes (batch 32, radix 2, bf16), where the oracle is too slow: size-independent properties.
    Images are independent and the cluster kernels use no atomics, so (i) running the batch in two halves gives BIT-identical
    V and dU, (ii) parameter gradients of the halves add up, (iii) V is a convex combination per (pixel, channel) when R > 1
    only through the per-image attention: sum_r a_r over radix is NOT 1 (softmax is over channels, the reference's quirk) but
    sum_c a_r[c] is 1 for every r -- checked through V of an all-ones U."""
    torch.manual_seed(2000 + c)
    K, R, N = 1, 2, 32
    D = lambda *s_: torch.randn(*s_, device="cuda")
    mk = lambda: ops.SplitAttention(K, R, c, D(K, c, c // 2) * 0.2, D(K, c // 2) * 0.1, 1 + 0.1 * D(K, c // 2), 0.1 * D(K, c // 2), 0.1 * D(K, c // 2),
                                    0.5 + torch.rand(K, c // 2, device="cuda"), D(K, R, c // 2, c) * 0.2, D(K, R, c) * 0.1)
    sa = mk()
    u = torch.randn(N, h, h, K * R * c, device="cuda").to(torch.bfloat16)
    dv = torch.randn(N, h, h, K * c, device="cuda").to(torch.bfloat16)
    v = sa.forward(u)
    du, pg = sa.backward(u, dv)
    halves, grads = [], []
    for sl in (slice(0, N // 2), slice(N // 2, N)):
        vh = sa.forward(u[sl].contiguous())
        duh, pgh = sa.backward(u[sl].contiguous(), dv[sl].contiguous())
        halves.append((vh, duh)); grads.append(pgh)
    assert torch.equal(torch.cat([a for a, _ in halves]), v) and torch.equal(torch.cat([b for _, b in halves]), du)
    for k_ in pg:
        assert rel(grads[0][k_] + grads[1][k_], pg[k_]) < 1e-5, k_
    ones = torch.ones(2, 8, 8, K * R * c, device="cuda", dtype=torch.bfloat16)
    vo = sa.forward(ones).float()                       # V[c] = sum_r a_r[c]  ->  sum_c V[c] = R
    assert float((vo.sum(-1) - R).abs().max()) < 2e-2 * R


def _check_split_attention(ops, dtype, K, R, c, n, h, w):
    torch.manual_seed(5)
    u_pre = torch.randn(n, h, w, K * R * c, dtype=torch.float64, requires_grad=True)
    u = F.elu(u_pre)                                          # u is an ELU output in the network
    uq = q(u.detach(), dtype)
    D = lambda *s: torch.randn(*s, dtype=torch.float64)
    P = dict(w1=D(K, c, c // 2) * 0.3, b1=D(K, c // 2) * 0.1, gamma=1 + 0.1 * D(K, c // 2), beta=0.1 * D(K, c // 2),
             mean=0.1 * D(K, c // 2), var=0.5 + torch.rand(K, c // 2, dtype=torch.float64), w2=D(K, R, c // 2, c) * 0.3, b2=D(K, R, c) * 0.1)
    Pr = {k_: v.clone().requires_grad_(k_ not in ("mean", "var")) for k_, v in P.items()}
    ur = uq.clone().requires_grad_(True)
    us = [[ur[..., (k * R + r) * c:(k * R + r + 1) * c] for r in range(R)] for k in range(K)]
    want = _splitatt_ref(us, Pr, K, R, c)
    f32 = lambda t: t.to(torch.float32).cuda()
    sa = ops.SplitAttention(K, R, c, f32(P["w1"]), f32(P["b1"]), f32(P["gamma"]), f32(P["beta"]), f32(P["mean"]), f32(P["var"]),
                            f32(P["w2"]), f32(P["b2"]))
    got = sa.forward(dev(uq, dtype))
    assert rel(got, want) < TOL[dtype]
    dv = q(torch.randn_like(want), dtype)
    names = ["w1", "b1", "gamma", "beta", "w2", "b2"]
    grads = torch.autograd.grad((want * dv).sum(), [ur] + [Pr[k_] for k_ in names])
    du, pg = sa.backward(dev(uq, dtype), dev(dv, dtype))
    want_du = grads[0] * torch.where(uq > 0, torch.ones_like(uq), uq + 1)     # kernel returns dL/du * ELU'(u)
    assert rel(du, want_du) < TOL[dtype] * 2
    for k_, gref in zip(names, grads[1:]):
        assert rel(pg[k_], gref) < 5e-4, k_                   # fp32 FC math either way (inputs already quantised)


@pytest.mark.parametrize("nc", [3, 4])
def test_softmax_loss(ops, nc):
    torch.manual_seed(6)
    n, h, w = 5, 8, 8
    logits = torch.randn(n, h, w, nc, dtype=torch.float64, requires_grad=True) * 2
    y = F.one_hot(torch.randint(0, nc, (n, h, w)), nc).double()
    o = O.TBIResNestOracle.__new__(O.TBIResNestOracle); o.height, o.width = h, w
    probs = torch.softmax(logits, -1)
    loss = o.my_loss_cat(y, probs)
    g, = torch.autograd.grad(loss.sum(), [logits])
    p, l, correct, dl = ops.softmax_loss(logits.detach().float().cuda(), y.float().cuda())
    assert rel(p, probs) < 1e-5 and rel(l, loss) < 1e-5 and rel(dl, g) < 1e-4
    assert int(correct.item()) == int((probs.argmax(-1) == y.argmax(-1)).sum())


@pytest.mark.parametrize("shape", [(5, 8, 8), (3, 20, 36), (2, 64, 96)])
def test_softmax_loss_from_head_taps(ops, shape):
    """tbi_softmax_loss_fwd_bwd_taps: the logits formed inside the loss kernel from the head's per-input-pixel tap products
    (f_tran as a GEMM, TBI_ResNest.py:124) must be what the 4-tap scatter writes -- so probabilities, loss map, accuracy count and
    dlogits equal scatter + tbi_softmax_loss_fwd_bwd BIT FOR BIT -- and the logits must equal the transposed conv of the oracle."""
    torch.manual_seed(16)
    n, h, w = shape                                            # OUTPUT grid; the head's input grid is h/2 x w/2
    nc, cin = 3, 32
    x = torch.randn(n, h // 2, w // 2, cin, dtype=torch.float64)
    wt = torch.randn(4, 4, nc, cin, dtype=torch.float64) * 0.2          # HWOI
    b = torch.randn(nc, dtype=torch.float64) * 0.1
    ytaps = torch.einsum("nhwi,abci->nhwabc", x, wt).reshape(n, h // 2, w // 2, 16 * nc).float().cuda().contiguous()
    y = F.one_hot(torch.randint(0, nc, (n, h, w)), nc).float().cuda()
    L = ops._lib.lib()
    bd = b.float().cuda()
    logits = torch.empty(n, h, w, nc, dtype=torch.float32, device="cuda")
    ops.check(L.tbi_convt_scatter_y(ops.F32, n, h // 2, w // 2, 4, nc, ops._vp(ops.view(ytaps)), bd.data_ptr(),
                                    ops._vp(ops.view(logits)), ops._st()), "scatter")
    want = O.conv2d_transpose_s2_same(x, wt, b)
    assert rel(logits, want) < 1e-5
    ref = ops.softmax_loss(logits, y)
    got = ops.softmax_loss_from_taps(ytaps, bd, y)
    torch.cuda.synchronize()
    for a_, b_ in zip(got, ref):
        assert torch.equal(a_, b_)


def test_adam_matches_oracle(ops):
    torch.manual_seed(7)
    n = 1003
    p = torch.randn(n, dtype=torch.float64); g1 = torch.randn(n, dtype=torch.float64); g2 = torch.randn(n, dtype=torch.float64)
    pd = torch.zeros(1004, dtype=torch.float32, device="cuda"); pd[:n] = p.float()
    gd = torch.zeros_like(pd); m = torch.zeros_like(pd); v = torch.zeros_like(pd)
    sc = torch.zeros(1, dtype=torch.int32, device="cuda")
    pr, mr, vr = p.clone(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    for t, g in enumerate([g1, g2], 1):
        gd[:n] = g.float()
        ops.adam_step(pd, gd, m, v, sc, 5e-3)
        lr_t = 5e-3 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        mr = 0.9 * mr + 0.1 * g; vr = 0.999 * vr + 0.001 * g * g
        pr = pr - lr_t * mr / (vr.sqrt() + 1e-7)
    assert int(sc.item()) == 2
    assert rel(pd[:n], pr) < 1e-5


def test_dropout_multiplier_stream(ops):
    L = ops._lib.lib()
    keep = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    sc = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ops.check(L.tbi_dropout_mask(keep.data_ptr(), keep.numel(), 123, sc.data_ptr(), st), "mask")
    a = keep.clone()
    assert set(a.unique().tolist()) == {0, 2}
    assert abs(float((a == 2).float().mean()) - 0.5) < 5e-3
    sc.fill_(1)
    ops.check(L.tbi_dropout_mask(keep.data_ptr(), keep.numel(), 123, sc.data_ptr(), st), "mask")
    assert 0.45 < float((a != keep).float().mean()) < 0.55       # a fresh draw per step


def test_unsupported_shape_is_an_error_not_a_fallback(ops):
    x = torch.randn(1, 4, 4, 5, device="cuda", dtype=torch.bfloat16)
    wt = torch.randn(3, 3, 5, 7, device="cuda")
    with pytest.raises(ops._lib.TbiError):
        ops.conv2d(x, wt, None, impl=ops._lib.IMPL_TCGEN05)       # 5 input channels: tcgen05 path refuses
    with pytest.raises(ops._lib.TbiError):
        ops.conv2d(x, wt, None, dilation=3)
