"""Golden vectors for the data path FROM THE REFERENCE'S OWN CODE: imports /root/reference/DataAugs.py and Dataset_2.py (they
import tensorflow at module level but use only numpy / random on this path, so an empty module stands in for it), runs
dataAug / label2vec on seeded inputs and records inputs, the decisions drawn (random.randint is wrapped to log them) and the
outputs.  Writes tests/golden/data_aug.npz.  Run from the repo root in the container that has /root/reference."""
import os
import random
import sys
import types

import numpy as np

sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
sys.path.insert(0, "/root/reference")
import DataAugs                       # noqa: E402  (the reference, unmodified)
import Dataset_2                      # noqa: E402

H, W, C, NS = 40, 36, 2, 16
rs = np.random.RandomState(7)
out = {}
images, labels, params, aug_images, aug_labels, noises = [], [], [], [], [], []
real_normal = np.random.normal
for s in range(NS):
    img = rs.uniform(-1, 1, (H, W, C))
    lab = np.round(rs.uniform(0, 2.2, (H, W)) * 2) / 2 * (rs.uniform(0, 1, (H, W)) > 0.3)        # values 0, .5, 1, 1.5, 2 with holes of 0
    random.seed(1000 + s)
    drawn = []
    real_randint = random.randint
    def logged(a, b, _d=drawn, _r=real_randint):
        v = _r(a, b); _d.append(v); return v
    DataAugs.random.randint = logged
    noise_box = []
    def fake_normal(mean, sigma, shape, _n=noise_box):
        g = rs.normal(mean, sigma, shape); _n.append(g.copy()); return g
    DataAugs.np.random.normal = fake_normal
    try:
        ai, al = DataAugs.dataAug(img.copy(), lab.copy())
    finally:
        DataAugs.random.randint = real_randint
        DataAugs.np.random.normal = real_normal
    # decisions in the layout of ultrasound_modeling_b200/data.py draw_params
    r, t = drawn[0], drawn[1]
    p = np.zeros(16, dtype=np.int32); k = 2
    p[0] = int(r % 3 != 0); p[1] = r % 3
    for c in range(r % 3):
        p[2 + 4 * c:6 + 4 * c] = drawn[k:k + 4]; k += 4
    if t % 2:
        p[10] = 1; p[11:14] = drawn[k:k + 3]; k += 3
    p[14] = int(t % 3 != 0)
    assert k == len(drawn)
    images.append(img); labels.append(lab); params.append(p); aug_images.append(np.asarray(ai, dtype=np.float64)); aug_labels.append(np.asarray(al, dtype=np.float64))
    noises.append(noise_box[0] if noise_box else np.zeros((H, W, C)))
out.update(image=np.stack(images), label=np.stack(labels), params=np.stack(params), aug_image=np.stack(aug_images), aug_label=np.stack(aug_labels),
           noise=np.stack(noises))
lab4 = np.stack(labels)
out["label2vec_3"] = Dataset_2.label2vec(lab4, 3).astype(np.float32)
lab2 = (lab4 > 0.9).astype(np.float64)
out["label2vec_2_in"] = lab2
out["label2vec_2"] = Dataset_2.label2vec(lab2, 2).astype(np.float32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_aug.npz"), **out)
print("wrote", NS, "samples; decisions:", np.stack(params)[:, [0, 1, 10, 14]].sum(0))
