"""tcgen05 tap-GEMM path vs the CUDA-core path of the same library on identical bf16 inputs (they must
agree to fp32 summation order), at the layer shapes of the r2k1 network, plus direct checks vs the oracle."""
import pytest
import torch
import torch.nn.functional as F

from oracle import tbi_resnest_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def ops(cuda_device):
    from ultrasound_modeling_b200 import ops as _ops
    return _ops


def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(BF)


@pytest.mark.parametrize("case", [
    # n, h, w, cin, cout, k, groups   (stage / layer it mirrors)
    (2, 16, 16, 64, 128, 3, 1),      # concats_2 stage 2 (KC=64)
    (2, 32, 32, 32, 64, 3, 1),       # concats_2 stage 1 (KC=32, 64B swizzle)
    (2, 32, 32, 16, 32, 3, 1),       # stem-like 16->32 (KC=16, 32B swizzle)
    (3, 8, 8, 128, 256, 1, 1),       # shortcut 1x1
    (2, 16, 16, 64, 128, 3, 2),      # grouped cardinal conv2 (G=2, 32->64 per group)
    (2, 8, 8, 256, 512, 3, 2),       # grouped, 128->256 per group, two N tiles per group
    (5, 4, 4, 512, 256, 3, 1),       # tiny spatial: tile spans images
    (1, 64, 64, 32, 16, 1, 1),       # narrow N (BN=16)
    (2, 20, 12, 64, 72, 3, 1),       # non power-of-two spatial + partial N tile
])
def test_conv_fwd_tc_vs_simt(ops, case):
    torch.manual_seed(0)
    n, h, w, cin, cout, k, g = case
    x = rnd(n, h, w, cin)
    wt = torch.randn(k, k, cin // g, cout, device="cuda") * (1.0 / (k * k * cin / g) ** 0.5)
    b = torch.randn(cout, device="cuda") * 0.1
    res = rnd(n, h, w, cout)
    kw = dict(groups=g, act=ops.ACT_ELU, residual=res)
    y_tc = ops.conv2d(x, wt, b, impl=ops._lib.IMPL_TCGEN05, **kw)
    y_si = ops.conv2d(x, wt, b, impl=ops._lib.IMPL_SIMT, **kw)
    torch.cuda.synchronize()
    assert rel(y_tc, y_si) < 1e-2                       # both round to bf16; fp32 sums differ only in order
    conv = F.conv2d(x.double().cpu().permute(0, 3, 1, 2), wt.to(BF).double().cpu().permute(3, 2, 0, 1), b.double().cpu(),
                    padding=k // 2, groups=g).permute(0, 2, 3, 1)
    want = F.elu(conv) + res.double().cpu()
    assert rel(y_tc, want) < 1e-2


@pytest.mark.parametrize("case", [(2, 16, 16, 64, 128, 3, 1), (2, 8, 8, 256, 512, 3, 2), (2, 32, 32, 32, 64, 1, 1)])
def test_conv_dgrad_tc_vs_simt(ops, case):
    torch.manual_seed(1)
    n, h, w, cin, cout, k, g = case
    x = rnd(n, h, w, cin); dz = rnd(n, h, w, cout)
    wt = torch.randn(k, k, cin // g, cout, device="cuda") * 0.05
    yref = rnd(n, h, w, cin)
    outs = []
    for impl in (ops._lib.IMPL_TCGEN05, ops._lib.IMPL_SIMT):
        dx, _, _ = ops.conv2d_grads(x, wt, dz, groups=g, impl=impl, dact=ops.ACT_ELU, dact_ref=yref)
        outs.append(dx)
    torch.cuda.synchronize()
    assert rel(outs[0], outs[1]) < 1e-2


@pytest.mark.parametrize("case", [(2, 8, 8, 64, 64, 128), (4, 4, 4, 512, 0, 512), (2, 16, 16, 128, 32, 64), (2, 8, 8, 512, 256, 256)])
def test_convt_fwd_dgrad_tc_vs_simt(ops, case):
    torch.manual_seed(2)
    n, h, w, c1, c2, cout = case
    x1 = rnd(n, h, w, c1); x2 = rnd(n, h, w, c2) if c2 else None
    cin = c1 + c2
    wt = torch.randn(4, 4, cout, cin, device="cuda") * (1.0 / (4 * cin) ** 0.5)
    b = torch.randn(cout, device="cuda") * 0.1
    keep = (torch.rand(n, 2 * h, 2 * w, cout, device="cuda") < 0.5).to(torch.uint8) * 2
    ys, dxs = [], []
    dz = rnd(n, 2 * h, 2 * w, cout)
    for impl in (ops._lib.IMPL_TCGEN05, ops._lib.IMPL_SIMT):
        ys.append(ops.conv2d_transpose_s2(x1, wt, b, act=ops._lib.ACT_RELU, keep=keep, x2=x2, impl=impl))
        dx, _, _ = ops.conv2d_transpose_s2_grads(x1, wt, dz, x2=x2, impl=impl)
        dxs.append(dx)
    torch.cuda.synchronize()
    assert rel(ys[0], ys[1]) < 1e-2
    if c2:
        assert rel(dxs[0][0], dxs[1][0]) < 1e-2 and rel(dxs[0][1], dxs[1][1]) < 1e-2
    else:
        assert rel(dxs[0], dxs[1]) < 1e-2
    xc = torch.cat([x1, x2], 3) if c2 else x1
    want = F.relu((O.conv2d_transpose_s2_same(xc.double().cpu(), wt.to(BF).double().cpu(), b.double().cpu())) * keep.double().cpu())
    assert rel(ys[0], want) < 1e-2


@pytest.mark.parametrize("case", [
    # n, h, w, cin(c1,c2), cout, k, groups
    (2, 16, 16, (64, 0), 128, 3, 1),
    (2, 16, 16, (128, 0), 256, 1, 1),
    (3, 8, 8, (256, 0), 512, 3, 2),       # grouped 128->256
    (2, 32, 32, (32, 0), 64, 3, 2),       # grouped 16->32: a 64-channel box spans several groups (masked rows)
    (2, 16, 16, (128, 64), 96, 3, 1),     # two sources, partial N tile
    (5, 4, 4, (512, 0), 256, 3, 1),       # pixel tile spans images
    (1, 40, 24, (64, 0), 64, 3, 1),       # non power-of-two spatial
    (2, 32, 32, (32, 0), 32, 3, 1),       # small-channel kernel: stem 32->32 (SW64 x, 3 dx taps packed in M)
    (2, 32, 24, (16, 0), 32, 3, 1),       # small-channel kernel: stem 16->32 (SW32 x, 8 M blocks)
    (2, 32, 32, (32, 0), 64, 3, 1),       # small-channel kernel: 32->64 (N = 64)
    (3, 16, 16, (32, 0), 64, 1, 1),       # small-channel kernel: 1x1 shortcut 32->64
    (2, 32, 32, (32, 0), 64, 3, 2),       # small-channel kernel, grouped 16->32 (stage-1 cardinal conv2)
    (1, 20, 12, (24, 0), 40, 3, 1),       # small-channel kernel: odd channel counts + partial tiles
])
def test_conv_wgrad_tc_vs_simt(ops, case):
    torch.manual_seed(3)
    n, h, w, (c1, c2), cout, k, g = case
    x1 = rnd(n, h, w, c1); x2 = rnd(n, h, w, c2) if c2 else None
    dz = rnd(n, h, w, cout)
    wt = torch.zeros(k, k, (c1 + c2) // g, cout, device="cuda")
    res = []
    for impl in (ops._lib.IMPL_TCGEN05, ops._lib.IMPL_SIMT):
        _, dw, db = ops.conv2d_grads(x1, wt, dz, groups=g, x2=x2, impl=impl, need_dx=False)
        res.append((dw, db))
    torch.cuda.synchronize()
    assert rel(res[0][0], res[1][0]) < 1e-4 and rel(res[0][1], res[1][1]) < 1e-4      # both accumulate bf16 products in fp32
    xc = (torch.cat([x1, x2], 3) if c2 else x1).double().cpu()
    wr = wt.double().cpu().requires_grad_(True)
    y = F.conv2d(xc.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), None, padding=k // 2, groups=g).permute(0, 2, 3, 1)
    gw, = torch.autograd.grad((y * dz.double().cpu()).sum(), [wr])
    assert rel(res[0][0], gw) < 1e-4


@pytest.mark.parametrize("case", [(2, 8, 8, 64, 64, 128), (4, 4, 4, 512, 0, 512), (2, 16, 16, 128, 32, 64)])
def test_convt_wgrad_tc_vs_simt(ops, case):
    torch.manual_seed(4)
    n, h, w, c1, c2, cout = case
    x1 = rnd(n, h, w, c1); x2 = rnd(n, h, w, c2) if c2 else None
    dz = rnd(n, 2 * h, 2 * w, cout)
    wt = torch.zeros(4, 4, cout, c1 + c2, device="cuda")
    res = []
    for impl in (ops._lib.IMPL_TCGEN05, ops._lib.IMPL_SIMT):
        _, dw, db = ops.conv2d_transpose_s2_grads(x1, wt, dz, x2=x2, impl=impl, need_dx=False)
        res.append((dw, db))
    torch.cuda.synchronize()
    assert rel(res[0][0], res[1][0]) < 1e-4 and rel(res[0][1], res[1][1]) < 1e-4
    xc = (torch.cat([x1, x2], 3) if c2 else x1).double().cpu()
    wr = wt.double().cpu().requires_grad_(True)
    y = O.conv2d_transpose_s2_same(xc, wr, None)
    gw, = torch.autograd.grad((y * dz.double().cpu()).sum(), [wr])
    assert rel(res[0][0], gw) < 1e-4


def test_head_shapes_on_tensor_cores(ops):
    """f_tran: Conv2DTranspose 160 -> 3 (TBI_ResNest.py:124): narrow fp32-out epilogue, 16-channel padded gradient."""
    torch.manual_seed(5)
    n, h, w, c1, c2, nc = 2, 16, 16, 128, 32, 3
    x1 = rnd(n, h, w, c1); x2 = rnd(n, h, w, c2)
    wt = torch.randn(4, 4, nc, c1 + c2, device="cuda") * 0.05
    b = torch.randn(nc, device="cuda") * 0.1
    ys = [ops.conv2d_transpose_s2(x1, wt, b, x2=x2, impl=impl, out_f32=True) for impl in (ops._lib.IMPL_TCGEN05, ops._lib.IMPL_SIMT)]
    assert ys[0].dtype == torch.float32 and rel(ys[0], ys[1]) < 1e-5
    xc = torch.cat([x1, x2], 3).double().cpu()
    want = O.conv2d_transpose_s2_same(xc, wt.to(BF).double().cpu(), b.double().cpu())
    assert rel(ys[0], want) < 1e-5
    dz = torch.zeros(n, 2 * h, 2 * w, 16, device="cuda", dtype=BF)
    dz[..., :nc] = rnd(n, 2 * h, 2 * w, nc)
    res = []
    for impl in (ops._lib.IMPL_TCGEN05, ops._lib.IMPL_SIMT):
        (dx1, dx2), dw, db = ops.conv2d_transpose_s2_grads(x1, wt, dz, x2=x2, impl=impl)
        res.append((dx1, dx2, dw, db))
    torch.cuda.synchronize()
    assert rel(res[0][0], res[1][0]) < 1e-2 and rel(res[0][1], res[1][1]) < 1e-2
    assert rel(res[0][2], res[1][2]) < 1e-4 and rel(res[0][3], res[1][3]) < 1e-4
    wr = wt.double().cpu().requires_grad_(True)
    xr = xc.clone().requires_grad_(True)
    y = O.conv2d_transpose_s2_same(xr, wr, None)
    gx, gw = torch.autograd.grad((y * dz[..., :nc].double().cpu()).sum(), [xr, wr])
    assert rel(res[0][2], gw) < 1e-4 and res[0][2].shape == (4, 4, nc, c1 + c2)


@pytest.mark.parametrize("dtype,hw", [(torch.float32, (40, 36)), (BF, (40, 36)), (BF, (24, 64))])
def test_stem_conv1_direct_kernels(ops, dtype, hw):
    """Conv1: 1 -> 16 (TBI_ResNest.py:83) runs on the direct few-channel kernels under IMPL_AUTO
    (width a multiple of 32 in bf16: the row-walking weight-gradient kernel)"""
    torch.manual_seed(6)
    n, (h, w) = 2, hw
    x = torch.randn(n, h, w, 1, device="cuda").to(dtype)
    wt = torch.randn(3, 3, 1, 16, device="cuda") * 0.3
    b = torch.randn(16, device="cuda") * 0.1
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    y = ops.conv2d(x, wt, b, act=ops.ACT_ELU)
    y_s = ops.conv2d(x, wt, b, act=ops.ACT_ELU, impl=ops._lib.IMPL_SIMT)
    assert rel(y, y_s) < tol
    dz = torch.randn(n, h, w, 16, device="cuda").to(dtype)
    _, dw, db = ops.conv2d_grads(x, wt, dz, need_dx=False)
    _, dw_s, db_s = ops.conv2d_grads(x, wt, dz, need_dx=False, impl=ops._lib.IMPL_SIMT)
    assert rel(dw, dw_s) < 1e-4 and rel(db, db_s) < 1e-4
