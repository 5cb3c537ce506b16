"""Device-side data path: the reference's per-sample Python loops -- ``label2vec`` (Dataset.py:41-52,
Dataset_2.py:6-20) and the ``DataAugs.py`` augmentations -- as kernels over device-resident batches.

    from ultrasound_modeling_b200.data import label2vec, DataAug
    aug = DataAug(seed=0)                               # draws every decision with Python's `random`, in the reference's order
    x_aug, label_aug = aug(x, label)                    # x [N,256,80,10], label [N,256,80] (numpy or torch) -> device fp32
    y = label2vec(label_aug, 3)                         # [N,256,80,3] soft one-hot

With the same ``random.seed`` the decisions (which samples are reduced / clipped / shifted, by how much) are the ones
``DataAugs.dataAug`` would draw sample by sample; only the Gaussian noise comes from a different generator (a counter-based
hash on the device instead of ``np.random.normal``).
"""
from __future__ import annotations

import random
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

AUG_WORDS = 16


def _dev(t, device, dtype=torch.float32):
    return torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t).to(device=device, dtype=dtype).contiguous()


def label2vec(label, num_classes: int = 3, device="cuda"):
    """label [N,H,W] (or [N,H,W,1]) -> fp32 [N,H,W,num_classes] on the device"""
    L = _lib.lib()
    lab = _dev(label, device)
    if lab.dim() == 4:
        lab = lab[..., 0].contiguous()
    y = torch.empty(*lab.shape, num_classes, dtype=torch.float32, device=lab.device)
    _lib.check(L.tbi_label2vec(lab.numel(), num_classes, lab.data_ptr(), y.data_ptr(), torch.cuda.current_stream(lab.device).cuda_stream), "label2vec")
    return y


def draw_params(n: int, rng=random) -> np.ndarray:
    """one row per sample, drawn exactly as DataAugs.dataAug draws (r, t, then clip / shift parameters in call order)"""
    out = np.zeros((n, AUG_WORDS), dtype=np.int32)
    for i in range(n):
        r = rng.randint(0, 100000); t = rng.randint(0, 100000)
        out[i, 0] = int(r % 3 != 0)
        out[i, 1] = r % 3
        for k in range(r % 3):                                   # clip(): r, c, ra, ca
            out[i, 2 + 4 * k:6 + 4 * k] = (rng.randint(0, 256), rng.randint(0, 80), rng.randint(20, 40), rng.randint(10, 20))
        if t % 2:                                                # shift(): r, c, direction
            out[i, 10] = 1
            out[i, 11], out[i, 12], out[i, 13] = rng.randint(0, 30), rng.randint(0, 12), rng.randint(0, 1)
        out[i, 14] = int(t % 3 != 0)
    return out


class DataAug:
    def __init__(self, seed: Optional[int] = None, device="cuda", noise: bool = True):
        self.rng = random.Random(seed) if seed is not None else random
        self.device = torch.device(device)
        self.noise = noise
        self._calls = 0
        self._seed = 0x0DA7A if seed is None else int(seed)

    def __call__(self, image, label, params: Optional[Sequence] = None):
        """image [N,H,W,C], label [N,H,W] -> (augmented image, augmented label), fp32 device tensors.  params: explicit decisions
        ([N,16] int32, see draw_params) instead of fresh draws."""
        L = _lib.lib()
        x = _dev(image, self.device); lab = _dev(label, self.device)
        n, h, w, c = x.shape
        p = np.asarray(params if params is not None else draw_params(n, self.rng), dtype=np.int32).reshape(n, AUG_WORDS).copy()
        if not self.noise:
            p[:, 14] = 0
        pd = torch.from_numpy(p).to(self.device)
        xo, lo = torch.empty_like(x), torch.empty_like(lab)
        self._calls += 1
        _lib.check(L.tbi_data_aug(n, h, w, c, x.data_ptr(), lab.data_ptr(), pd.data_ptr(), self._seed * 1000003 + self._calls, xo.data_ptr(), lo.data_ptr(),
                                  torch.cuda.current_stream(self.device).cuda_stream), "data_aug")
        return xo, lo
