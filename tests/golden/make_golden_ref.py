"""Golden vectors produced by the REFERENCE'S OWN CODE (run in the build container only: needs /root/reference).

``oracle/tfshim`` supplies a stand-in ``tensorflow`` package, so ``/root/reference/TBI_ResNest.py`` and
``/root/reference/VisionTransformer.py`` (+ ``ResNest.py``, ``Decoder.py``) import unmodified and their own ``ResNest.step`` /
``VisionTransformer.train_step`` / ``step`` run: functional-API graph, layer creation order and Keras auto-names, ``my_loss_cat``,
``compute_loss``, GradientTape -> (clip_by_global_norm) -> Adam.  This script feeds them the oracles' seeded parameters and inputs
and records what THEY return; ``tests/test_oracle_pinned.py`` then holds the oracles (and through them the CUDA path) to it.

    python tests/golden/make_golden_ref.py            # writes tests/golden/ref_*.npz

Recorded per case: the reference's variable inventory (names in ITS trainable order, shapes), probabilities, loss, accuracy,
for every trainable variable the gradient's L2 norm / sum / 4 probe entries, and the same three statistics of every variable
after the optimizer steps the case runs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ULTRASOUND_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def import_reference():
    """the shim first, then the reference directory, on sys.path; returns the shim module"""
    for p in (REF, os.path.join(ROOT, "oracle", "tfshim")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import tensorflow as tf
    assert tf.__version__ == "2.shim", "a real tensorflow shadows oracle/tfshim"
    return tf


def probes(numel: int):
    return [0, numel // 3, (2 * numel) // 3, numel - 1]


def tensor_stats(t: torch.Tensor):
    f = t.detach().reshape(-1).double()
    return np.array([float(f.norm()), float(f.sum())] + [float(f[i]) for i in probes(f.numel())])


# ------------------------------------------------------------------------------------------------------------------
# Variant A: TBI_ResNest.py
# ------------------------------------------------------------------------------------------------------------------
def run_reference_a(radix, kpaths, size=64, n=2, steps=2):
    tf = import_reference()
    import TBI_ResNest as ref
    from oracle import tbi_resnest_oracle as O
    tf.shim_reset_names()
    net = ref.ResNest(size, size, 1, 3, 3, radix=radix, kpaths=kpaths)
    params = O.init_params(1, 3, 3, radix, kpaths, dtype=torch.float64)
    for v in net.resModel.variables:
        v.assign(params[v.name[:-2]])
    x, y = O.synthetic_batch(n, size, size, dtype=torch.float64)
    out = {"trainable_names": np.array([v.name[:-2] for v in net.resModel.trainable_variables]),
           "variable_names": np.array([v.name[:-2] for v in net.resModel.variables]),
           "variable_shapes": np.array([",".join(map(str, v.shape)) for v in net.resModel.variables])}
    for s in range(steps):
        masks = O.dropout_masks(n, size, size, seed=1237 + s)
        tf.shim_queue_dropout_masks([m.bool() for m in masks])
        loss, acc, probs = net.step(x.numpy(), tf.convert_to_tensor(y.numpy(), dtype=tf.float32), train=True)
        out[f"loss_{s}"] = loss.numpy(); out[f"acc_{s}"] = np.float64(float(acc)); out[f"probs_{s}"] = probs.numpy() if s == 0 else probs.numpy()[:, ::2, ::2, :]
        out[f"grad_stats_{s}"] = np.stack([tensor_stats(tf.shim_last_gradients[v.name]) for v in net.resModel.trainable_variables])
    out["final_stats"] = np.stack([tensor_stats(v._t) for v in net.resModel.variables])
    # evaluation call of the same object (train=False): no update, dropout still on (TBI_ResNest.py:215-216)
    tf.shim_queue_dropout_masks([m.bool() for m in O.dropout_masks(n, size, size, seed=1299)])
    loss, acc, probs = net.step(x.numpy(), tf.convert_to_tensor(y.numpy(), dtype=tf.float32), train=False)
    out["eval_loss"] = loss.numpy(); out["eval_acc"] = np.float64(float(acc)); out["eval_probs_sub"] = probs.numpy()[:, ::2, ::2, :]
    return out


# ------------------------------------------------------------------------------------------------------------------
# Variant B + ViT bridge: VisionTransformer.py (ResNest.py, Decoder.py)
# ------------------------------------------------------------------------------------------------------------------
LIST_NAMES = {"cardinal_blocks": "cardinal", "Transformer_layers": "layer", "blocks": "block"}


def walk_reference_b(tf, obj, prefix, out):
    """attribute paths of the reference object tree -> its Keras variables (the naming of oracle/vit_oracle.py)"""
    for attr, val in vars(obj).items():
        if isinstance(val, tf.keras.layers.Layer):
            name = "initial_conv" if (attr == "conv1" and prefix.endswith("hybrid_model/")) else attr   # ResNest.py:14 names it
            for v in val._weights:
                out[prefix + name + "/" + v.name[:-2].rsplit("/", 1)[1]] = v
        elif isinstance(val, tf.Module):
            walk_reference_b(tf, val, prefix + attr + "/", out)
        elif isinstance(val, list) and val and isinstance(val[0], tf.Module):
            for i, e in enumerate(val):
                walk_reference_b(tf, e, prefix + f"{LIST_NAMES[attr]}_{i}/", out)
    return out


def run_reference_b(n=1, steps=2):
    tf = import_reference()
    import VisionTransformer as refv
    from oracle import vit_oracle as V
    tf.shim_reset_names()
    net = refv.VisionTransformer(batch_size=n, img_size=(256, 80))
    byname = walk_reference_b(tf, net, "", {})
    params = V.init_params(V.model_param_shapes(), dtype=torch.float64)
    assert set(byname) == set(params), (set(byname) ^ set(params))
    for k, v in byname.items():
        p = params[k]
        v.assign(p[0, 0] if (p.dim() == 4 and len(v.shape) == 2) else p)           # Dense kernels are [in,out] in Keras
    path_of = {id(v): k for k, v in byname.items()}
    tv = net.visionModel.trainable_variables
    x = V.B.synthetic_input(n, 256, 80, 10).double(); y = V.synthetic_labels(n, 256, 80).double()
    out = {"trainable_names": np.array([path_of[id(v)] for v in tv]),
           "keras_names": np.array([v.name[:-2] for v in tv]),
           "variable_names": np.array(list(byname)),
           "variable_shapes": np.array([",".join(map(str, v.shape)) for v in byname.values()])}
    for s in range(steps):
        loss, probs = net.train_step(x.numpy(), tf.convert_to_tensor(y.numpy(), dtype=tf.float32))
        out[f"loss_{s}"] = np.float64(float(loss))
        if s == 0:
            out["probs_0"] = probs.numpy()
        out[f"grad_stats_{s}"] = np.stack([tensor_stats(tf.shim_last_gradients[v.name]) for v in tv])
    out["final_stats"] = np.stack([tensor_stats(v._t) for v in byname.values()])
    loss, probs = net.step(x.numpy(), tf.convert_to_tensor(y.numpy(), dtype=tf.float32))
    out["eval_loss"] = np.float64(float(loss)); out["eval_probs_sub"] = probs.numpy()[:, ::4, ::4, :]
    logits, weights = net(x.numpy())
    out["attn_weights_stats"] = np.stack([tensor_stats(w._v) for w in weights])
    return out


CASES_A = {"ref_tbi_resnest_r2k1_64.npz": (2, 1), "ref_tbi_resnest_r3k4_64.npz": (3, 4), "ref_tbi_resnest_r4k4_64.npz": (4, 4),
           "ref_tbi_resnest_r1k1_64.npz": (1, 1)}

if __name__ == "__main__":
    for fname, (r, k) in CASES_A.items():
        o = run_reference_a(r, k)
        np.savez_compressed(os.path.join(HERE, fname), **o)
        print(fname, len(o["trainable_names"]), "trainable variables, loss sum", float(o["loss_0"].sum()))
    o = run_reference_b()
    np.savez_compressed(os.path.join(HERE, "ref_vit_256x80.npz"), **o)
    print("ref_vit_256x80.npz", len(o["trainable_names"]), "trainable variables, loss", float(o["loss_0"]))
