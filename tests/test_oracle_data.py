"""The data-path oracle (oracle/data_oracle.py) against vectors produced by the REFERENCE'S OWN CODE
(tests/golden/data_aug.npz, written by tests/golden/make_golden_data.py from /root/reference/DataAugs.py and Dataset_2.py),
and the host-side decision drawing against the reference's draw order.  No GPU."""
import os
import random

import numpy as np

from oracle import data_oracle as D

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_aug.npz")


def test_data_aug_restatement_equals_the_reference_outputs():
    g = np.load(GOLDEN)
    n = g["image"].shape[0]
    used = np.zeros(4, dtype=int)
    for s in range(n):
        p = g["params"][s]
        img, lab = D.data_aug(g["image"][s], g["label"][s], p, noise=g["noise"][s])
        assert np.array_equal(lab, g["aug_label"][s]), s
        assert np.allclose(img, g["aug_image"][s], rtol=0, atol=1e-15), s
        used += [p[0], p[1] > 0, p[10], p[14]]
    assert (used > 0).all(), used                       # every stage occurs in the fixture


def test_label2vec_equals_the_reference_outputs():
    g = np.load(GOLDEN)
    assert np.array_equal(D.label2vec(g["label"], 3), g["label2vec_3"])
    assert np.array_equal(D.label2vec(g["label2vec_2_in"], 2), g["label2vec_2"])
    y = D.label2vec(g["label"], 3)
    assert set(np.unique(g["label"])) >= {0.0, 1.0, 1.5, 2.0} and np.allclose(y.sum(-1)[g["label"] > 0.95], 1.0)


def test_decisions_are_drawn_in_the_reference_order():
    from ultrasound_modeling_b200.data import draw_params
    g = np.load(GOLDEN)
    for s in range(g["params"].shape[0]):
        rng = random.Random(1000 + s)                   # the seed make_golden_data.py gave the reference for sample s
        assert np.array_equal(draw_params(1, rng)[0], g["params"][s]), s


def test_evaluator_maps():
    p = np.random.RandomState(0).dirichlet((1, 1, 1), (2, 5, 4))
    po, pO = D.prob_maps(p)
    assert np.allclose(po, p[..., 2]) and np.allclose(pO, 1 - p[..., 0] - 0.5 * p[..., 1] + p[..., 2])
    x = np.ones((1, 5, 4, 3)); m = np.zeros((1, 5, 4, 2)); m[0, 1, 2, 0] = 0.7; m[0, 3, 3, 0] = 0.4
    out = D.apply_brain_mask(x, m)
    assert out[0, 1, 2].sum() == 0 and out.sum() == x.sum() - 3
