"""Writes tests/golden/vit_64x32.npz from the ViT-bridge oracle (fp64, 2 layers, 64x32x10 input, batch 2): the probabilities,
the loss and the norm of every 7th parameter gradient.  Run from the repo root: python tests/golden/make_golden_vit.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import vit_oracle as V  # noqa: E402

o = V.VisionTransformerOracle(2, img_size=(64, 32), num_classes=3, dtype=torch.float64, num_layers=2)
x = V.B.synthetic_input(2, 64, 32, 10).double(); y = V.synthetic_labels(2, 64, 32).double()
loss, probs, grads = o.gradients(x, y)
names = sorted(grads)[::7]
out = {"probs": probs.numpy(), "loss": np.asarray(float(loss)), "grad_names": np.asarray(names)}
for n in names:
    out["gradnorm__" + n.replace("/", "__")] = np.asarray(float(grads[n].norm()))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vit_64x32.npz"), **out)
print("wrote", len(names), "gradient norms; loss", float(loss))
