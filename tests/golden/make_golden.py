"""Generates tests/golden/tbi_resnest_r2k1_64.npz from the CPU oracle (fp64), seeds as in SURVEY 8d.
The reference itself (TensorFlow) cannot run here, so these vectors pin the ORACLE, not TensorFlow:
run once, commit the .npz; tests/test_oracle.py::test_golden_fixture re-derives and compares."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tbi_resnest_oracle as O  # noqa: E402

o = O.TBIResNestOracle(64, 64, 1, 3, 3, 2, 1, dtype=torch.float64)
x, y = O.synthetic_batch(2, 64, 64, dtype=torch.float64)
m = O.dropout_masks(2, 64, 64)
probs = o.forward(x, m)
loss = o.my_loss_cat(y, probs)
g = o.gradients(x, y, m)
names = sorted(g)
out = {"probs": probs.detach().numpy(), "loss": loss.detach().numpy(), "grad_names": np.array(names)}
for n in names:
    out["gradnorm__" + n.replace("/", "__")] = np.float64(g[n].norm().item())
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tbi_resnest_r2k1_64.npz"), **out)
print("wrote", len(names), "gradient norms; probs", probs.shape)
