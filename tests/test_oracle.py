"""CPU tests of the oracle itself (nothing external pins it -- SURVEY 8c -- so it is pinned by
first-principles definitions, fp64/fp32 agreement, gradcheck and a committed golden fixture)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import tbi_resnest_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tbi_resnest_r2k1_64.npz")


def tf_same_strided_conv_np(x, w, stride):
    """tf.nn.conv2d(x NHWC, w HWIO, strides=stride, padding='SAME') from the TF padding rule, plain loops."""
    n, h, wd, cin = x.shape
    kh, kw, _, cout = w.shape
    oh, ow = math.ceil(h / stride), math.ceil(wd / stride)
    pt = max((oh - 1) * stride + kh - h, 0); pl = max((ow - 1) * stride + kw - wd, 0)
    top, left = pt // 2, pl // 2
    y = np.zeros((n, oh, ow, cout))
    for oy in range(oh):
        for ox in range(ow):
            for ky in range(kh):
                for kx in range(kw):
                    iy, ix = oy * stride + ky - top, ox * stride + kx - left
                    if 0 <= iy < h and 0 <= ix < wd:
                        y[:, oy, ox, :] += x[:, iy, ix, :] @ w[ky, kx]
    return y


def tf_conv2d_transpose_np(x, w_hwoi, stride):
    """Conv2DTranspose(padding='same') as TF defines it: the input-gradient of the SAME strided conv whose
    HWIO kernel is w_hwoi read as [kh,kw,out_of_transpose(=in of conv),in_of_transpose(=out of conv)]."""
    n, h, wd, cin = x.shape
    kh, kw, cout, _ = w_hwoi.shape
    H, W = h * stride, wd * stride
    pt = max((h - 1) * stride + kh - H, 0); pl = max((wd - 1) * stride + kw - W, 0)
    top, left = pt // 2, pl // 2
    y = np.zeros((n, H, W, cout))
    for oy in range(h):              # positions of the strided conv's output == transpose's input
        for ox in range(wd):
            for ky in range(kh):
                for kx in range(kw):
                    iy, ix = oy * stride + ky - top, ox * stride + kx - left
                    if 0 <= iy < H and 0 <= ix < W:
                        y[:, iy, ix, :] += x[:, oy, ox, :] @ w_hwoi[ky, kx].T
    return y


@pytest.mark.parametrize("k", [3, 4])
def test_tf_same_transposed_conv_identity(k):
    rng = np.random.default_rng(k)
    x = rng.standard_normal((2, 5, 3, 4)); w = rng.standard_normal((k, k, 6, 4))
    ref = tf_conv2d_transpose_np(x, w, 2)
    got = O.conv2d_transpose_s2_same(torch.from_numpy(x), torch.from_numpy(w), None).numpy()
    assert np.abs(ref - got).max() < 1e-12
    # and it really is the VJP of the SAME strided conv: <conv(u), x> == <u, convT(x)>
    u = rng.standard_normal((2, 10, 6, 6))
    lhs = (tf_same_strided_conv_np(u, w, 2) * x).sum()
    assert abs(lhs - (u * ref).sum()) < 1e-9 * max(1.0, abs(lhs))


def test_conv_same_matches_tf_rule():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 6, 5, 3)); w = rng.standard_normal((3, 3, 3, 4))
    ref = tf_same_strided_conv_np(x, w, 1)
    got = O.conv2d_same(torch.from_numpy(x), torch.from_numpy(w), None).numpy()
    assert np.abs(ref - got).max() < 1e-12


def test_bn_inference_init_stats():
    x = torch.randn(2, 3, 3, 4, dtype=torch.float64)
    one, zero = torch.ones(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64)
    assert torch.allclose(O.batchnorm_inference(x, one, zero, zero, one), x / math.sqrt(1.001))


def test_shape_walk_and_tables():
    # decoder input channels 512/1024/768/640/320/160, FLOP and parameter tables
    shp = O.param_shapes(1, 3, 3, 4, 4)
    assert [shp[f"upsample_{i}_t_conv/kernel"][3] for i in range(5)] == [512, 1024, 768, 640, 320]
    assert shp["f_tran/kernel"] == (4, 4, 3, 160)
    for (r, k), (gf, mp) in {(2, 1): (21.519, 26.93), (4, 4): (20.556, 25.78), (3, 4): (20.513, 25.74)}.items():
        assert abs(O.forward_flops_per_image(256, 256, 1, 3, 3, r, k) / 1e9 - gf) < 2e-3
        n = sum(int(np.prod(s)) for nme, s in O.param_shapes(1, 3, 3, r, k).items() if O.is_trainable(nme))
        assert abs(n / 1e6 - mp) < 6e-3
    assert O.cardinal_channels(128, 3, 4) == (5, 16) and O.cardinal_channels(512, 3, 4) == (21, 64)
    assert "conv2_1_car_k01_0bn/gamma" in shp           # "%s1_%rbn" % (name, idx_r), TBI_ResNest.py:164


def test_fp64_fp32_agree_and_probs_normalised():
    x, y = O.synthetic_batch(2, 64, 64)
    m = O.dropout_masks(2, 64, 64)
    p32 = O.TBIResNestOracle(64, 64, 1, 3, 3, 2, 1, dtype=torch.float32).forward(x, m)
    p64 = O.TBIResNestOracle(64, 64, 1, 3, 3, 2, 1, dtype=torch.float64).forward(x.double(), m)
    assert (p32.double() - p64).abs().max() < 2e-5
    assert (p64.sum(-1) - 1).abs().max() < 1e-12


def test_split_attention_analytic_backward():
    """SURVEY 8a backward formulas (what the CUDA kernel implements) == autograd of the restatement."""
    torch.manual_seed(0)
    o = O.TBIResNestOracle(64, 64, 1, 3, 3, 2, 1, dtype=torch.float64)
    name, R, c = "conv2_1_car_k0_att", 2, 32
    us = [torch.randn(2, 5, 4, c, dtype=torch.float64, requires_grad=True) for _ in range(R)]
    v = o.split_attention(us, name)
    dv = torch.randn_like(v)
    gu = torch.autograd.grad((v * dv).sum(), us)
    p = o.params
    with torch.no_grad():
        g = sum(us).mean(dim=(1, 2))
        w1 = p[f"{name}1/kernel"][0, 0]; q = g @ w1 + p[f"{name}1/bias"]
        istd = 1 / torch.sqrt(p[f"{name}_bn/moving_variance"] + O.BN_EPS)
        bn = (q - p[f"{name}_bn/moving_mean"]) * p[f"{name}_bn/gamma"] * istd + p[f"{name}_bn/beta"]
        h1 = torch.nn.functional.elu(bn)
        a = [torch.softmax(h1 @ p[f"{name}2_r{r}/kernel"][0, 0] + p[f"{name}2_r{r}/bias"], -1) for r in range(R)]
        da = [(dv * us[r]).sum(dim=(1, 2)) for r in range(R)]
        dz = [a[r] * (da[r] - (a[r] * da[r]).sum(-1, keepdim=True)) for r in range(R)]
        dh1 = sum(dz[r] @ p[f"{name}2_r{r}/kernel"][0, 0].T for r in range(R))
        dbn = dh1 * torch.where(h1 > 0, torch.ones_like(h1), h1 + 1)
        dg = (dbn * p[f"{name}_bn/gamma"] * istd) @ w1.T
        for r in range(R):
            du = dv * a[r][:, None, None, :] + dg[:, None, None, :] / 20
            assert (du - gu[r]).abs().max() < 1e-10


def test_gradcheck_tiny_model():
    o = O.TBIResNestOracle(64, 64, 1, 3, 3, 2, 1, dtype=torch.float64)
    x, y = O.synthetic_batch(1, 64, 64, dtype=torch.float64)
    g = o.gradients(x, y, None)
    # finite-difference a handful of scalar parameters
    rng = np.random.default_rng(0)
    for name in ["Conv1/kernel", "conv3_1_car_k0_att2_r1/kernel", "upsample_3_t_conv/kernel", "conv2d_2/bias",
                 "batch_normalization_1/gamma", "f_tran/kernel"]:
        t = o.params[name]
        idx = tuple(int(rng.integers(0, s)) for s in t.shape)
        eps = 1e-5
        with torch.no_grad():
            old = t[idx].item()
            t[idx] = old + eps; lp = o.my_loss_cat(y, o.forward(x)).sum().item()
            t[idx] = old - eps; lm = o.my_loss_cat(y, o.forward(x)).sum().item()
            t[idx] = old
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - g[name][idx].item()) <= 1e-5 * max(1.0, abs(fd)) + 1e-8, name


def test_adam_matches_keras_formula():
    o = O.TBIResNestOracle(64, 64, 1, 3, 3, 2, 1, learning_rate=1e-2, dtype=torch.float64)
    k = "Conv1/bias"
    p0 = o.params[k].detach().clone()
    g = {n: torch.zeros_like(v) for n, v in o.params.items() if O.is_trainable(n)}
    g[k] = torch.full_like(p0, 0.5)
    o.apply_adam(g)
    lr_t = 1e-2 * math.sqrt(1 - 0.999) / (1 - 0.9)
    want = p0 - lr_t * (0.1 * 0.5) / (math.sqrt(0.001 * 0.25) + 1e-7)
    assert torch.allclose(o.params[k].detach(), want, atol=1e-15)


def test_golden_fixture():
    """the committed vectors (tests/golden/make_golden.py) pin the oracle against silent edits"""
    gz = np.load(GOLDEN)
    o = O.TBIResNestOracle(64, 64, 1, 3, 3, 2, 1, dtype=torch.float64)
    x, y = O.synthetic_batch(2, 64, 64, dtype=torch.float64)
    m = O.dropout_masks(2, 64, 64)
    probs = o.forward(x, m)
    loss = o.my_loss_cat(y, probs)
    assert np.abs(probs.detach().numpy() - gz["probs"]).max() < 1e-9
    assert np.abs(loss.detach().numpy() - gz["loss"]).max() < 1e-12
    g = o.gradients(x, y, m)
    for name in gz["grad_names"]:
        name = str(name)
        assert abs(float(g[name].norm()) - float(gz["gradnorm__" + name.replace("/", "__")])) < 1e-9 * max(1.0, float(g[name].norm()))
