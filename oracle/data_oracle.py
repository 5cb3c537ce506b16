"""CPU oracle (numpy) for the device-side data path and the evaluator epilogue: ``DataAugs.py``, ``Dataset.py:41-52`` /
``Dataset_2.py:6-20`` (label2vec) and ``TBIEvaluator.py:225-252`` of the reference.

TEST INFRASTRUCTURE ONLY.  PINNED: unlike the network oracles this one is checked against the REFERENCE'S OWN CODE -- these
functions need only numpy and ``random``, so ``tests/golden/make_golden_data.py`` imports ``/root/reference/DataAugs.py`` and
``Dataset_2.py`` (with an empty stand-in for the ``tensorflow`` module they import but do not use on this path), runs them on
seeded inputs and commits inputs, drawn decisions and outputs as ``tests/golden/data_aug.npz``; ``tests/test_oracle_data.py``
holds this restatement to those vectors exactly.

The restatement is vectorised but keeps the reference's loop bounds and its one surprising property:
  * ``imageReduc`` (DataAugs.py:52-79): the erosion loop tests ``mask[i, j] > 1`` on a 0/1 mask, so it never fires; after the
    first pass the mask is all zeros and the function reduces to "zero every input plane where the LABEL plane is 0".
  * ``clip`` (:26-37) and ``shift`` (:6-23) iterate ``range(0, H-1)`` x ``range(0, W-1)``: the last row and column are never
    clipped, and never receive a shifted value.
"""
from __future__ import annotations

import numpy as np

AUG_WORDS = 16


def label2vec(label: np.ndarray, num_classes: int = 3) -> np.ndarray:
    """label [N,H,W] -> [N,H,W,num_classes] (Dataset_2.py:6-20; Dataset.py:41-52 is the 3-class branch)"""
    if num_classes == 3:
        c2 = np.where(label >= 1.05, label - 1, 0)
        c2 = np.where(c2 > 1, 1, c2)
        c1 = np.where(label > 0.95, 1 - c2, 0)
        c0 = np.where(label <= 0.95, 1, 0)
        return np.stack([c0, c1, c2], axis=-1).astype(np.float32)
    return np.stack([1 - label, label], axis=-1).astype(np.float32)


def image_reduc(image: np.ndarray, label: np.ndarray):
    """[H,W,C], [H,W] -> (label, image): input planes zeroed where the label is 0"""
    out = np.where((label == 0)[:, :, None], 0, image)
    return label, out


def clip(image: np.ndarray, label: np.ndarray, r: int, c: int, ra: int, ca: int):
    h, w = label.shape
    i = np.arange(h)[:, None]; j = np.arange(w)[None, :]
    m = (i < h - 1) & (j < w - 1) & (r + ra > i) & (i > r - ra) & (c + ca > j) & (j > c - ca)
    return np.where(m, 0, label), np.where(m[:, :, None], 0, image)


def shift(image: np.ndarray, label: np.ndarray, r: int, c: int, direction: int):
    h, w = label.shape
    sg = 1 if direction else -1
    i = np.arange(h)[:, None]; j = np.arange(w)[None, :]
    si, sj = i + sg * r, j + sg * c
    ok = (i < h - 1) & (j < w - 1) & (si >= 0) & (si < h) & (sj >= 0) & (sj < w)
    sic, sjc = np.clip(si, 0, h - 1), np.clip(sj, 0, w - 1)
    lab = np.where(ok, label[sic, sjc], 0)
    img = np.where(ok[:, :, None], image[sic, sjc, :], 0)
    return lab, img


def data_aug(image: np.ndarray, label: np.ndarray, p: np.ndarray, noise: np.ndarray = None):
    """one sample, decisions p (see ultrasound_modeling_b200/data.py draw_params), DataAugs.dataAug :82-102 -> (image, label);
    ``noise``: the N(0,1) field the reference would draw ([H,W,C]) or None for no noise even if p[14] is set"""
    image = image.astype(np.float64); label = label.astype(np.float64)
    if p[0]:
        label, image = image_reduc(image, label)
    for k in range(int(p[1])):
        label, image = clip(image, label, *[int(v) for v in p[2 + 4 * k:6 + 4 * k]])
    if p[10]:
        label, image = shift(image, label, int(p[11]), int(p[12]), int(p[13]))
    if p[14] and noise is not None:
        image = image + noise / 5000
    return image, label


def prob_maps(prob: np.ndarray):
    """TBIEvaluator.py:239-252: prob [N,H,W,3] -> (probOut = last class, probO = 1 - p0 - 0.5 p1 + p2)"""
    return prob[..., -1], 1.0 - prob[..., 0] - 0.5 * prob[..., 1] + prob[..., 2]


def apply_brain_mask(x: np.ndarray, mask_prob: np.ndarray):
    """TBIEvaluator.py:225-231: zero every input plane where round(mask[..., 0]) == 1"""
    return np.where((np.round(mask_prob[..., 0]) == 1)[..., None], 0.0, x)
