"""The oracles held to the REFERENCE'S OWN CODE (CPU).

``tests/golden/ref_*.npz`` were produced by ``tests/golden/make_golden_ref.py``: the unmodified ``/root/reference/TBI_ResNest.py``
and ``VisionTransformer.py`` (+ ``ResNest.py``, ``Decoder.py``) executed under ``oracle/tfshim`` (a stand-in ``tensorflow``
package whose primitives are restated from TF/Keras' documented definitions, independently of the oracles' formulations).
Here every oracle re-derives the same quantities from the same seeded parameters and inputs:

  * the variable inventory -- every Keras name (explicit and auto-generated), shape, and the reference's trainable ORDER;
  * probabilities, loss, accuracy of ``ResNest.step`` / ``VisionTransformer.train_step`` / ``step``;
  * every gradient (L2 norm, sum and 4 probe entries per variable) of two consecutive training steps;
  * every variable after those two optimizer steps (Adam; global-norm clip for the ViT model);
  * an evaluation call afterwards (Variant A: dropout still on, TBI_ResNest.py:215-216).

Everything is float64 on both sides, so the bar is 1e-9 of each tensor's scale.  When ``/root/reference`` is present (the build
container) the reference itself is re-run and must reproduce the committed fixture; on the GPU box that part is skipped.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import tbi_resnest_oracle as O
from oracle import vit_oracle as V

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
import make_golden_ref as G  # noqa: E402

TOL = 1e-9


def close(a, b, scale=None, tol=TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = max(float(np.abs(b).max()), 1e-300) if scale is None else scale
    return float(np.abs(a - b).max()) <= tol * s


def check_stats(got_tensors, want_stats, names, tol=TOL):
    """per-variable [norm, sum, 4 probes]; the scale of a variable is the largest norm of the set (a mathematically zero gradient,
    e.g. the key bias under softmax, is rounding noise on both sides)"""
    scale = float(want_stats[:, 0].max())
    bad = []
    for t, w, n in zip(got_tensors, want_stats, names):
        g = G.tensor_stats(t)
        ref = max(float(w[0]), 1e-6 * scale)
        if not (abs(g[0] - w[0]) <= tol * ref and np.abs(g[2:] - w[2:]).max() <= tol * ref and abs(g[1] - w[1]) <= 1e-7 * ref * max(1.0, np.sqrt(t.numel()))):
            bad.append((n, g.tolist(), w.tolist()))
    assert not bad, bad[:3]


@pytest.mark.parametrize("fname", sorted(G.CASES_A))
def test_variant_a_oracle_matches_reference_code(fname):
    r, k = G.CASES_A[fname]
    z = np.load(os.path.join(GOLD, fname))
    shapes = O.param_shapes(1, 3, 3, r, k)
    # inventory: names, shapes and the trainable order are the reference's
    assert list(z["variable_names"]) == list(shapes)
    assert [tuple(int(d) for d in s.split(",")) for s in z["variable_shapes"]] == [tuple(v) for v in shapes.values()]
    trainable = [n for n in shapes if O.is_trainable(n)]
    assert list(z["trainable_names"]) == trainable
    o = O.TBIResNestOracle(64, 64, 1, 3, 3, r, k, dtype=torch.float64)
    x, y = O.synthetic_batch(2, 64, 64, dtype=torch.float64)
    for s in range(2):
        loss, acc, probs = o.step(x, y, train=True, masks=O.dropout_masks(2, 64, 64, seed=1237 + s))
        want = z[f"probs_{s}"]
        assert close(probs.numpy() if s == 0 else probs.numpy()[:, ::2, ::2, :], want)
        assert close(loss.numpy(), z[f"loss_{s}"])
        assert abs(float(acc) - float(z[f"acc_{s}"])) < 1e-12
        check_stats([o.last_grads[n] for n in trainable], z[f"grad_stats_{s}"], trainable)
    check_stats([o.params[n] for n in shapes], z["final_stats"], list(shapes))
    loss, acc, probs = o.step(x, y, train=False, masks=O.dropout_masks(2, 64, 64, seed=1299))
    assert close(probs.numpy()[:, ::2, ::2, :], z["eval_probs_sub"]) and close(loss.numpy(), z["eval_loss"])
    assert abs(float(acc) - float(z["eval_acc"])) < 1e-12


def test_variant_b_vit_oracle_matches_reference_code():
    z = np.load(os.path.join(GOLD, "ref_vit_256x80.npz"))
    shapes = V.model_param_shapes()
    assert set(z["variable_names"]) == set(shapes)
    for n, s in zip(z["variable_names"], z["variable_shapes"]):
        shp = tuple(int(d) for d in s.split(","))
        want = tuple(shapes[n])
        assert shp == want or (len(shp) == 2 and want == (1, 1) + shp), (n, shp, want)     # Keras Dense kernels are [in,out]
    trainable = list(z["trainable_names"])
    assert sorted(trainable) == sorted(n for n in shapes if V.is_trainable(n))
    o = V.VisionTransformerOracle(1, img_size=(256, 80), dtype=torch.float64)
    x = V.B.synthetic_input(1, 256, 80, 10).double(); y = V.synthetic_labels(1, 256, 80).double()
    for s in range(2):
        ref = V.VisionTransformerOracle(1, img_size=(256, 80), dtype=torch.float64, params=o.state_dict())
        _, _, grads = ref.gradients(x, y)
        loss, probs = o.train_step(x, y)
        assert abs(float(loss) - float(z[f"loss_{s}"])) <= TOL * abs(float(z[f"loss_{s}"]))
        if s == 0:
            assert close(probs.numpy(), z["probs_0"])
        check_stats([grads[n] for n in trainable], z[f"grad_stats_{s}"], trainable)
    check_stats([o.params[n] for n in z["variable_names"]], z["final_stats"], list(z["variable_names"]))
    loss, probs = o.step(x, y)
    assert abs(float(loss) - float(z["eval_loss"])) <= TOL * abs(float(z["eval_loss"]))
    assert close(probs.numpy()[:, ::4, ::4, :], z["eval_probs_sub"])
    _, weights = o.forward(x)
    check_stats(weights, z["attn_weights_stats"], [f"attn_{i}" for i in range(len(weights))])


# ---------------------------------------------------------------------------------------------------------------
# the shim's primitives against loop definitions of TensorFlow's padding rule (they are what the reference code ran on)
# ---------------------------------------------------------------------------------------------------------------
def conv_loops(x, w, stride, dil, same=True):
    n, h, wd, _ = x.shape
    kh, kw, _, cout = w.shape
    if same:
        oh, ow = -(-h // stride), -(-wd // stride)
        pt = max((oh - 1) * stride + (kh - 1) * dil + 1 - h, 0) // 2
        pl = max((ow - 1) * stride + (kw - 1) * dil + 1 - wd, 0) // 2
    else:
        oh, ow, pt, pl = (h - (kh - 1) * dil - 1) // stride + 1, (wd - (kw - 1) * dil - 1) // stride + 1, 0, 0
    y = np.zeros((n, oh, ow, cout))
    for oy in range(oh):
        for ox in range(ow):
            for ky in range(kh):
                for kx in range(kw):
                    iy, ix = oy * stride + ky * dil - pt, ox * stride + kx * dil - pl
                    if 0 <= iy < h and 0 <= ix < wd:
                        y[:, oy, ox, :] += x[:, iy, ix, :] @ w[ky, kx]
    return y


def shim():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "tfshim")
    if p not in sys.path:
        sys.path.insert(0, p)
    import tensorflow as tf
    assert tf.__version__ == "2.shim"
    return tf


@pytest.mark.parametrize("k,stride,dil,same", [(3, 1, 1, True), (1, 1, 1, True), (3, 1, 2, True), (3, 1, 8, True), (4, 2, 1, True),
                                               (3, 2, 1, True), (1, 1, 1, False)])
def test_shim_conv_is_tf_same_conv(k, stride, dil, same):
    tf = shim()
    rng = np.random.default_rng(k * 10 + dil)
    x = rng.standard_normal((2, 9, 6, 3)); w = rng.standard_normal((k, k, 3, 5))
    got = tf._conv2d(torch.from_numpy(x), torch.from_numpy(w), (stride, stride), "SAME" if same else "VALID", (dil, dil)).numpy()
    assert np.abs(got - conv_loops(x, w, stride, dil, same)).max() < 1e-12


@pytest.mark.parametrize("k", [3, 4])
def test_shim_transposed_conv_is_the_gradient_of_the_same_conv(k):
    """tf.nn.conv2d_transpose is DEFINED as the input-gradient of conv2d: <conv(u, W), x> == <u, convT(x, W)> for all u"""
    tf = shim()
    rng = np.random.default_rng(k)
    x = torch.from_numpy(rng.standard_normal((2, 5, 3, 4)))          # transpose input  == conv output
    w = torch.from_numpy(rng.standard_normal((k, k, 6, 4)))          # HWOI of the transpose == HWIO of the conv (6 -> 4)
    u = torch.from_numpy(rng.standard_normal((2, 10, 6, 6))).requires_grad_(True)
    conv = torch.from_numpy(conv_loops(u.detach().numpy(), w.numpy(), 2, 1))
    y = tf._conv2d_transpose(x, w, (2, 2), "SAME")
    assert tuple(y.shape) == (2, 10, 6, 6)
    assert abs(float((conv * x).sum()) - float((u.detach() * y).sum())) < 1e-9
    (g,) = torch.autograd.grad((tf._conv2d(u, w, (2, 2), "SAME", (1, 1)) * x).sum(), u)
    assert float((g - y).abs().max()) < 1e-12


def test_shim_keras_auto_names_and_immutability():
    tf = shim()
    tf.shim_reset_names()
    L = tf.keras.layers
    names = [L.Conv2D(4, 1).name, L.Conv2D(4, 1).name, L.Conv2D(4, 1, name="x").name, L.Conv2D(4, 1).name,
             L.BatchNormalization().name, L.BatchNormalization().name, L.Conv2DTranspose(4, 4).name, L.LeakyReLU().name,
             L.AveragePooling2D().name, L.LayerNormalization().name]
    assert names == ["conv2d", "conv2d_1", "x", "conv2d_2", "batch_normalization", "batch_normalization_1", "conv2d_transpose",
                     "leaky_re_lu", "average_pooling2d", "layer_normalization"]
    a = tf.convert_to_tensor(np.ones((2, 2))); b = a
    b += a                                                             # rebinding, not an in-place write (ResNest.py:177)
    assert float(a.numpy().sum()) == 4.0 and float(b.numpy().sum()) == 8.0


@pytest.mark.skipif(not os.path.isdir(G.REF), reason="the reference tree exists only in the build container")
def test_reference_rerun_reproduces_the_committed_fixtures():
    got = G.run_reference_a(2, 1)
    z = np.load(os.path.join(GOLD, "ref_tbi_resnest_r2k1_64.npz"))
    for key in z.files:
        if z[key].dtype.kind == "f":
            assert close(got[key], z[key], tol=1e-12), key
        else:
            assert list(got[key]) == list(z[key]), key
    got = G.run_reference_b()
    z = np.load(os.path.join(GOLD, "ref_vit_256x80.npz"))
    for key in ("loss_0", "loss_1", "probs_0", "eval_loss", "final_stats"):
        assert close(got[key], z[key], tol=1e-11), key
    assert list(got["keras_names"]) == list(z["keras_names"])
