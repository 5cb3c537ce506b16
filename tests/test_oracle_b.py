"""CPU self-checks of the Variant B oracle (oracle/resnest_decoder_oracle.py): nothing external pins it (PARITY UNPINNED),
so it is checked against first principles, the identities SURVEY 8c lists, and the committed golden vectors."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import resnest_decoder_oracle as B

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resnest_decoder_r3k3_64x32.npz")


def _small(dtype=torch.float64, H=64, W=32):
    grid = (H // 16, W // 16)
    pe = B.init_params(B.encoder_param_shapes(10, 3, 3, 3), seed=2236, dtype=dtype)
    pd = B.init_params(B.decoder_param_shapes(3, grid=grid), seed=2237, dtype=dtype)
    enc = B.ResNestEncoderOracle(10, 3, 3, 3, pe, dtype=dtype)
    dec = B.DecoderCupOracle(3, pd, grid=grid, dtype=dtype)
    x = B.synthetic_input(2, H, W, 10, dtype=dtype)
    tok = B.synthetic_tokens(2, grid[0] * grid[1], 512, dtype=dtype)
    return enc, dec, x, tok


def test_shape_walk_matches_survey_tables():
    """skip shapes at [N,256,80,10], cardinal widths, decoder input channels 384/160/72"""
    assert [B.cardinal_channels(o, 3, 3) for _, o in B.ENC_STAGES] == [(3, 10), (7, 21), (14, 42), (28, 85)]
    ds = B.decoder_param_shapes(3)
    assert ds["block_1/up/kernel"][3] == 384 and ds["block_2/up/kernel"][3] == 160 and ds["head/kernel"][3] == 72
    es = B.encoder_param_shapes(10, 3, 3, 3)
    assert es["conv_1/concats_2/kernel"][2] == 30 and es["conv_4/concats_2/kernel"][2] == 255
    enc = B.ResNestEncoderOracle(10, 3, 3, 3, B.init_params(es), dtype=torch.float32)
    x4, feats = enc(B.synthetic_input(1))
    assert [tuple(t.shape) for t in [x4] + feats] == [(1, 16, 5, 512), (1, 32, 10, 256), (1, 64, 20, 128), (1, 128, 40, 64)]


def test_layernorm_matches_definition():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 5, 7, 21, generator=g, dtype=torch.float64)
    ga, be = torch.randn(21, generator=g, dtype=torch.float64), torch.randn(21, generator=g, dtype=torch.float64)
    want = F.layer_norm(x, (21,), ga, be, eps=1e-3)
    assert (B.layernorm_c(x, ga, be) - want).abs().max() < 1e-12


def test_split_attention_identity():
    """R identical inputs + one dense2  =>  V = R * U * softmax_c(z)"""
    enc, _, x, _ = _small()
    p = "conv_1/cardinal_0"
    u = torch.randn(2, 8, 4, 10, dtype=torch.float64)
    v = enc.split_attention([u, u, u], p + "/split")
    g = (3 * u).mean(dim=(1, 2))[:, None, None, :]
    h = B.leaky(B.layernorm_c(enc._conv(g, p + "/split/dense1"), enc.p[p + "/split/dense1_bn/gamma"], enc.p[p + "/split/dense1_bn/beta"]))
    a = torch.softmax(enc._conv(h, p + "/split/dense2"), dim=-1)
    assert (v - 3 * u * a).abs().max() < 1e-12


def test_decoder_token_reshape_is_raw():
    """the tensor concatenated after block i is a plain reshape of the tokens (Decoder.py:140)"""
    _, dec, _, tok = _small()
    n = tok.shape[0]
    x0 = tok.reshape(n, dec.grid[0] * 2, dec.grid[1] * 2, -1)
    assert x0.shape[-1] == 128 and torch.equal(x0.reshape(n, -1), tok.reshape(n, -1))


def test_fp32_fp64_agree_and_probs_normalised():
    enc64, dec64, x, tok = _small(torch.float64)
    enc32, dec32, _, _ = _small(torch.float32)
    p64 = dec64(tok, enc64(x)[1])
    p32 = dec32(tok.float(), enc32(x.float())[1])
    assert (p64.sum(-1) - 1).abs().max() < 1e-12
    assert (p64 - p32.double()).abs().max() < 5e-5


def test_golden_fixture():
    gz = np.load(GOLDEN)
    enc, dec, x, tok = _small()
    x4, feats = enc(x)
    probs = dec(tok, feats)
    assert np.abs(probs.numpy() - gz["probs"]).max() < 1e-9
    assert abs(float(x4.norm()) - float(gz["x4_norm"])) < 1e-9 * float(gz["x4_norm"])
    for f, w in zip(feats, gz["feat_norms"]):
        assert abs(float(f.norm()) - float(w)) < 1e-9 * float(w)
