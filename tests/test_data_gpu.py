"""Device-side data path and evaluator epilogue against (a) vectors produced by the reference's own DataAugs.py / Dataset_2.py
(tests/golden/data_aug.npz) and (b) the numpy oracle at the reference's real batch shape."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as D

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_aug.npz")


def test_data_aug_kernel_equals_the_reference_outputs(cuda_device):
    from ultrasound_modeling_b200.data import DataAug, label2vec
    g = np.load(GOLDEN)
    aug = DataAug(seed=3, noise=False)
    xo, lo = aug(g["image"], g["label"], params=g["params"])
    # noise off: the kernel is a pure gather, so it must reproduce the reference bit for bit (in fp32)
    want = np.stack([D.data_aug(g["image"][s], g["label"][s], g["params"][s])[0] for s in range(g["image"].shape[0])])
    assert np.array_equal(xo.cpu().numpy(), want.astype(np.float32))
    assert np.array_equal(lo.cpu().numpy(), g["aug_label"].astype(np.float32))
    # ... and with noise the difference to the noise-free result is N(0, 1/5000^2) on the samples that asked for it
    xn, _ = DataAug(seed=3, noise=True)(g["image"], g["label"], params=g["params"])
    d = (xn - xo).cpu().numpy() * 5000
    on = g["params"][:, 14] == 1
    assert np.all(d[~on] == 0) and abs(d[on].mean()) < 0.02 and abs(d[on].std() - 1) < 0.02
    assert np.array_equal(label2vec(g["label"], 3).cpu().numpy(), g["label2vec_3"])
    assert np.array_equal(label2vec(g["label2vec_2_in"], 2).cpu().numpy(), g["label2vec_2"])


def test_data_aug_full_batch_vs_oracle(cuda_device):
    """the reference's batch shape [64,256,80,10] with freshly drawn decisions (every sample differs)"""
    from ultrasound_modeling_b200.data import DataAug, draw_params
    import random
    rs = np.random.RandomState(11)
    x = rs.uniform(-1, 1, (64, 256, 80, 10)).astype(np.float32)
    lab = (np.round(rs.uniform(0, 2.2, (64, 256, 80)) * 2) / 2 * (rs.uniform(0, 1, (64, 256, 80)) > 0.3)).astype(np.float32)
    p = draw_params(64, random.Random(5))
    xo, lo = DataAug(seed=1, noise=False)(x, lab, params=p)
    for s in range(0, 64, 7):
        wi, wl = D.data_aug(x[s], lab[s], p[s])
        assert np.array_equal(lo[s].cpu().numpy(), wl.astype(np.float32)) and np.array_equal(xo[s].cpu().numpy(), wi.astype(np.float32)), s
    assert p[:, 0].any() and p[:, 1].any() and p[:, 10].any()


def test_evaluator_service(cuda_device):
    """one resident model serves a batch of requests; prob / probOut / probO equal the oracle's maps of the model's own
    probabilities; the brain-mask pre-pass zeroes exactly the pixels the mask model's rounded first class marks"""
    from ultrasound_modeling_b200.VisionTransformer import VisionTransformer
    from ultrasound_modeling_b200.evaluator import Evaluator
    from oracle import vit_oracle as V
    net = VisionTransformer(4, img_size=(64, 32), dtype="fp32", num_layers=1, seed=1)
    x = V.B.synthetic_input(5, 64, 32, 10)
    ev = Evaluator(net, max_batch=2)                           # 5 requests in batches of 2, 2, 1
    out = ev(x)
    probs, _ = net.forward(x)
    po, pO = D.prob_maps(probs.double().cpu().numpy())
    assert tuple(out["prob"].shape) == (5, 64, 32, 3)
    assert np.allclose(out["prob"].cpu().numpy(), probs.cpu().numpy(), atol=2e-6)
    assert np.allclose(out["probOut"].cpu().numpy(), po, atol=2e-6) and np.allclose(out["probO"].cpu().numpy(), pO, atol=5e-6)
    mask_net = VisionTransformer(4, img_size=(64, 32), num_classes=2, dtype="fp32", num_layers=1, seed=2)
    mp, _ = mask_net.forward(x)
    xm = D.apply_brain_mask(x.numpy(), mp.cpu().numpy())
    want, _ = net.forward(torch.from_numpy(xm).float())
    got = Evaluator(net, brain_mask_model=mask_net, max_batch=8)(x)
    assert 0 < float((torch.from_numpy(xm) == 0).float().mean()) < 1          # the mask removes something, not everything
    assert np.allclose(got["prob"].cpu().numpy(), want.cpu().numpy(), atol=1e-5)
