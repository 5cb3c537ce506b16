"""Generates tests/golden/resnest_decoder_r3k3_64x32.npz from the Variant B CPU oracle (fp64).
The reference itself (TensorFlow) cannot run here, so these vectors pin the ORACLE, not TensorFlow:
run once, commit the .npz; tests/test_oracle_b.py::test_golden_fixture re-derives and compares."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import resnest_decoder_oracle as B  # noqa: E402

H, W, C, R, K = 64, 32, 10, 3, 3
grid = (H // 16, W // 16)
pe = B.init_params(B.encoder_param_shapes(C, 3, R, K), seed=2236, dtype=torch.float64)
pd = B.init_params(B.decoder_param_shapes(3, grid=grid), seed=2237, dtype=torch.float64)
enc = B.ResNestEncoderOracle(C, 3, R, K, pe)
dec = B.DecoderCupOracle(3, pd, grid=grid)
x = B.synthetic_input(2, H, W, C, dtype=torch.float64)
tok = B.synthetic_tokens(2, grid[0] * grid[1], 512, dtype=torch.float64)
x4, feats = enc(x)
probs = dec(tok, feats)
out = {"probs": probs.numpy(), "x4_norm": np.float64(x4.norm().item()), "feat_norms": np.array([f.norm().item() for f in feats])}
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resnest_decoder_r3k3_64x32.npz"), **out)
print("wrote probs", probs.shape, "x4", x4.shape)
