"""Pins of the ViT-bridge oracle (oracle/vit_oracle.py) that need no GPU: every building block against an independent
definition, the variable inventory, the loss against torch's label-smoothed cross entropy, clip + Adam against the formulas,
and the committed golden fixture (tests/golden/make_golden_vit.py)."""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import vit_oracle as V

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vit_64x32.npz")


def small(dtype=torch.float64, lr=1e-3):
    return V.VisionTransformerOracle(2, img_size=(64, 32), num_classes=3, learning_rate=lr, dtype=dtype, num_layers=2)


def test_blocks_against_independent_definitions():
    torch.manual_seed(0)
    x = torch.randn(3, 8, 512, dtype=torch.float64)
    g, b = torch.randn(512, dtype=torch.float64), torch.randn(512, dtype=torch.float64)
    assert torch.allclose(V.layernorm(x, g, b), F.layer_norm(x, (512,), g, b, eps=1e-6), atol=1e-12)
    assert torch.allclose(V.gelu(x), F.gelu(x), atol=1e-12)
    o = small()
    out, probs = o.attention(x, V.TR + "layer_0/attn/")
    p = o.params
    lin = lambda z, nm: z @ p[V.TR + "layer_0/attn/" + nm + "/kernel"][0, 0] + p[V.TR + "layer_0/attn/" + nm + "/bias"]
    heads = lambda z: z.reshape(3, 8, 4, 128).transpose(1, 2)
    ref = F.scaled_dot_product_attention(heads(lin(x, "query")), heads(lin(x, "key")), heads(lin(x, "value")), scale=1.0 / math.sqrt(4.0))
    assert torch.allclose(out, lin(ref.transpose(1, 2).reshape(3, 8, 512), "out"), atol=1e-10)
    assert torch.allclose(probs.sum(-1), torch.ones(3, 4, 8, dtype=torch.float64), atol=1e-12)      # softmax over the keys


def test_loss_is_label_smoothed_cross_entropy_over_global_batch():
    torch.manual_seed(1)
    z = torch.randn(2, 6, 5, 3, dtype=torch.float64)
    y = F.one_hot(torch.randint(0, 3, (2, 6, 5)), 3).double()
    o = small()
    got = o.compute_loss(y, torch.softmax(z, -1))
    want = F.cross_entropy(z.reshape(-1, 3), y.reshape(-1, 3).argmax(-1), label_smoothing=0.1, reduction="sum") / 2.0
    assert abs(float(got) - float(want)) < 1e-10
    # the clip only matters for saturated probabilities
    p = torch.tensor([[[[1.0, 0.0, 0.0]]]], dtype=torch.float64)
    l = V.cce_label_smoothing(torch.tensor([[[[0.0, 1.0, 0.0]]]], dtype=torch.float64), p)
    assert abs(float(l) - (-(0.1 / 3) * math.log(1 - 1e-7) - (0.9 + 0.1 / 3) * math.log(1e-7) - (0.1 / 3) * math.log(1e-7))) < 1e-9


def test_inventory_and_shapes():
    s = V.model_param_shapes()
    assert s["transformer/embeddings/patch_embeddings/kernel"] == (1, 1, 512, 512)
    assert s[V.TR + "layer_7/ffn/fc1/kernel"] == (1, 1, 512, 2048) and s[V.TR + "encoder_norm/gamma"] == (512,)
    assert s[V.DEC + "conv_more/kernel"] == (3, 3, 512, 256) and s[V.DEC + "head/kernel"] == (3, 3, 3, 72)
    vit = sum(int(np.prod(v)) for k, v in s.items() if k.startswith(V.TR))
    assert vit == 8 * (4 * (512 * 512 + 512) + 2 * 512 * 2048 + 2048 + 512 + 4 * 512) + 2 * 512      # 8 blocks + encoder_norm
    o = small()
    probs, weights = o.forward(V.B.synthetic_input(2, 64, 32, 10).double())
    assert tuple(probs.shape) == (2, 64, 32, 3) and len(weights) == 2 and tuple(weights[0].shape) == (2, 4, 8, 8)


def test_train_step_is_clip_then_keras_adam():
    o = small(lr=1e-2)
    x = V.B.synthetic_input(2, 64, 32, 10).double(); y = V.synthetic_labels(2, 64, 32).double()
    before = o.state_dict()
    loss, probs, grads = o.gradients(x, y)
    gnorm = math.sqrt(sum(float((g ** 2).sum()) for g in grads.values()))
    o.train_step(x, y)
    after = o.state_dict()
    assert abs(o.last_gnorm - gnorm) < 1e-9 * gnorm and gnorm > 1.0                  # the clip is active in this configuration
    k = V.TR + "layer_0/ffn/fc1/kernel"
    g = grads[k] / gnorm                                                              # clip_by_global_norm(., 1.0)
    m, v = 0.1 * g, 0.001 * g * g
    lr_t = 1e-2 * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert torch.allclose(after[k], before[k] - lr_t * m / (v.sqrt() + 1e-7), atol=1e-12)
    # finite-difference check of one gradient entry
    eps = 1e-6
    with torch.no_grad():
        o2 = small(); o2.params[k][0, 0, 3, 5] += eps
        o3 = small(); o3.params[k][0, 0, 3, 5] -= eps
        lp = o2.compute_loss(y, o2.forward(x)[0]); lm = o3.compute_loss(y, o3.forward(x)[0])
    assert abs(float((lp - lm) / (2 * eps)) - float(grads[k][0, 0, 3, 5])) < 1e-6 * max(1.0, abs(float(grads[k][0, 0, 3, 5])))


def test_golden_fixture():
    gz = np.load(GOLDEN)
    o = small()
    x = V.B.synthetic_input(2, 64, 32, 10).double(); y = V.synthetic_labels(2, 64, 32).double()
    loss, probs, grads = o.gradients(x, y)
    assert np.allclose(probs.numpy(), gz["probs"], atol=1e-10) and abs(float(loss) - float(gz["loss"])) < 1e-10
    for name in gz["grad_names"]:
        name = str(name)
        assert abs(float(grads[name].norm()) - float(gz["gradnorm__" + name.replace("/", "__")])) < 1e-9 * max(1.0, float(grads[name].norm())), name
