"""The ViT bridge and the Variant B training step (VisionTransformer.py of the reference) on the GPU against the CPU oracle:
the three new kernels alone, then forward / every gradient / two optimizer steps of the whole model.
Bars: fp32 storage 1e-4 (kernels), 2e-4 on whole-model gradients (LeakyReLU kinks, as in test_variant_b_gpu.py); bf16 2e-2."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vit_oracle as V

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def rel(got, want):
    want = want.detach().double().cpu(); got = got.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,t,heads,d", [(3, 80, 4, 128), (2, 8, 4, 128), (2, 33, 2, 64)])
def test_attention_fwd_bwd(cuda_device, dtype, n, t, heads, d):
    from ultrasound_modeling_b200 import ops
    g = torch.Generator().manual_seed(5 + t)
    q, k, v, do = (torch.randn(n, t, heads * d, generator=g, dtype=torch.float64).to(dtype).double() * 0.5 for _ in range(4))
    scale = 1.0 / math.sqrt(heads)
    qr, kr, vr = (z.clone().requires_grad_(True) for z in (q, k, v))
    sp = lambda z: z.reshape(n, t, heads, d).transpose(1, 2)
    pw = torch.softmax(sp(qr) @ sp(kr).transpose(-1, -2) * scale, -1)
    want = (pw @ sp(vr)).transpose(1, 2).reshape(n, t, heads * d)
    want.backward(do)
    dev = lambda z: z.to(cuda_device, dtype)
    ctx, probs = ops.attention(dev(q), dev(k), dev(v), heads, scale)
    assert rel(ctx, want) < TOL[dtype] and rel(probs, pw) < 1e-4
    dq, dk, dv = ops.attention_bwd(dev(q), dev(k), dev(v), probs, dev(do), heads, scale)
    assert max(rel(dq, qr.grad), rel(dk, kr.grad), rel(dv, vr.grad)) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gelu_and_loss(cuda_device, dtype):
    from ultrasound_modeling_b200 import ops
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(4, 7, 5, 64, generator=g, dtype=torch.float64) * 2).to(dtype).double()
    dy = torch.randn(4, 7, 5, 64, generator=g, dtype=torch.float64).to(dtype).double()
    xr = x.clone().requires_grad_(True)
    want = V.gelu(xr); want.backward(dy)
    y = ops.gelu(x.to(cuda_device, dtype))
    dx = ops.gelu_bwd(x.to(cuda_device, dtype), dy.to(cuda_device, dtype))
    assert rel(y, want) < TOL[dtype] and rel(dx, xr.grad) < TOL[dtype]
    if dtype == torch.float32:
        z = torch.randn(3, 9, 6, 3, generator=g, dtype=torch.float64) * 3
        z[0, 0, 0] = torch.tensor([40.0, -40.0, 0.0])                      # saturated: exercises the probability clip
        yl = F.one_hot(torch.randint(0, 3, (3, 9, 6), generator=g), 3).double()
        zr = z.clone().requires_grad_(True)
        lw = V.cce_label_smoothing(yl, torch.softmax(zr, -1)).sum() / 5.0
        lw.backward()
        probs, loss, dz = ops.softmax_cce(z.float().to(cuda_device), yl.float().to(cuda_device), 0.1, 5.0)
        assert rel(probs, torch.softmax(z, -1)) < 1e-5 and abs(float(loss) - float(lw)) < 1e-5 * float(lw)
        assert rel(dz, zr.grad) < 1e-4


def build(cuda_device, dtype, img=(64, 32), n=2, layers=2, lr=1e-3):
    from ultrasound_modeling_b200.VisionTransformer import VisionTransformer
    o = V.VisionTransformerOracle(n, img_size=img, num_classes=3, learning_rate=lr, dtype=torch.float64, num_layers=layers)
    net = VisionTransformer(n, img_size=img, num_classes=3, learning_rate=lr, dtype="fp32" if dtype == torch.float32 else "bf16",
                            device=str(cuda_device), num_layers=layers)
    net.load_variables(o.state_dict())
    x = V.B.synthetic_input(n, img[0], img[1], 10); y = V.synthetic_labels(n, img[0], img[1])
    return o, net, x, y


@pytest.mark.parametrize("dtype,img,n,layers", [(torch.float32, (64, 32), 2, 2), (torch.float32, (256, 80), 2, 8), (torch.bfloat16, (64, 32), 2, 2),
                                                (torch.bfloat16, (256, 80), 2, 8)])
def test_forward_parity(cuda_device, dtype, img, n, layers):
    o, net, x, y = build(cuda_device, dtype, img, n, layers)
    probs, weights = net.forward(x)
    with torch.no_grad():
        zw, ww = o.forward(x.double(), logits=True)
    pw = torch.softmax(zw, -1)
    tol = TOL[dtype]
    ptol = tol * max(1.0, float(zw.abs().max()) / 2)                       # a softmax moves a probability by at most |dz| / 2
    print("probs rel", rel(probs, pw), "| attention weights rel", [round(rel(a, b), 6) for a, b in zip(weights, ww)], "| max|z|", float(zw.abs().max()))
    assert rel(probs, pw) < 3 * ptol                                       # three chained drop-in units (encoder, bridge, decoder)
    assert len(weights) == layers and tuple(weights[0].shape) == tuple(ww[0].shape)
    # the attention probabilities are a side output (~1/T each); in bf16 the token errors of 8 chained blocks show in them
    assert max(rel(a, b) for a, b in zip(weights, ww)) < (1e-3 if dtype == torch.float32 else 0.15)
    loss, p2 = net.step(x, y)
    lw, _ = o.step(x.double(), y.double())
    assert abs(float(loss) - float(lw)) < (1e-4 if dtype == torch.float32 else 2e-2) * float(lw)
    # the public loss entry point on probabilities
    assert abs(float(net.compute_loss(y, pw.float())) - float(lw)) < 1e-4 * float(lw)
    # variables created by the product cover exactly the oracle's inventory (names and shapes)
    assert {k: tuple(v.shape) for k, v in net.variables().items()} == {k: tuple(v.shape) for k, v in o.params.items()}


def test_gradients_and_train_steps_fp32(cuda_device):
    o, net, x, y = build(cuda_device, torch.float32, lr=1e-3)
    lw, pw, gw = o.gradients(x.double(), y.double())
    loss, probs = net.backward(x, y)
    got = net.gradients()
    assert set(got) == set(gw)
    # the key bias has NO influence on the output (q.b_k is the same for every key, and softmax is shift invariant): its true
    # gradient is 0 and both sides hold rounding noise; errors are therefore measured against max(|tensor|, 1e-6 of the largest
    # gradient entry of the model)
    floor = 1e-6 * max(float(g.abs().max()) for g in gw.values())
    relf = lambda a, b: float((a.detach().double().cpu() - b).abs().max() / max(float(b.abs().max()), floor))
    keyb = [k for k in gw if k.endswith("attn/key/bias")]
    assert len(keyb) == 2 and all(float(got[k].abs().max()) < 1e-5 * floor / 1e-6 and float(gw[k].abs().max()) < 1e-9 * floor / 1e-6 for k in keyb)
    errs = sorted(((relf(got[k], gw[k]), k) for k in gw if k not in keyb), reverse=True)
    print("largest relative gradient errors:", errs[:5], "| loss", float(loss), float(lw))
    assert errs[0][0] < 2e-4, errs[:5]
    assert abs(float(loss) - float(lw)) < 1e-5 * float(lw)
    # two optimizer steps: clip_by_global_norm(1.0) + Keras Adam on the flat buffer
    o2, net2, _, _ = build(cuda_device, torch.float32, lr=1e-3)
    for _ in range(2):
        l1, _ = net2.train_step(x, y)
        l2, _ = o2.train_step(x.double(), y.double())
        assert abs(float(l1) - float(l2)) < 1e-4 * float(l2)
        assert abs(net2.global_grad_norm() - o2.last_gnorm) < 1e-4 * o2.last_gnorm and o2.last_gnorm > 1.0      # the clip is active
    want = o2.state_dict(); gotv = net2.variables()
    diffs = torch.cat([(gotv[k].double().cpu() - want[k]).abs().reshape(-1) for k in want])
    # Adam turns rounding-level gradients into +-lr steps: all but a vanishing fraction within lr*2e-2, none beyond 2 steps
    assert float((diffs > 1e-3 * 2e-2).double().mean()) < 1e-4 and float(diffs.max()) < 4e-3


def test_gradients_bf16_track_the_oracle(cuda_device):
    """bf16 storage: the direction of the whole gradient and its norm (what the clip sees) follow the fp64 oracle; per-tensor
    errors are dominated by LayerNorm over 3-10 channels in the encoder and are reported, not bounded."""
    o, net, x, y = build(cuda_device, torch.bfloat16)
    lw, pw, gw = o.gradients(x.double(), y.double())
    loss, probs = net.backward(x, y)
    got = net.gradients()
    a = torch.cat([got[k].double().cpu().reshape(-1) for k in gw]); b = torch.cat([gw[k].reshape(-1) for k in gw])
    cos = float((a @ b) / (a.norm() * b.norm()))
    errs = sorted(rel(got[k], gw[k]) for k in gw)
    print("bf16 whole-model gradient: cosine", cos, "| norm ratio", float(a.norm() / b.norm()), "| median tensor error", errs[len(errs) // 2])
    assert cos > 0.97 and abs(float(a.norm() / b.norm()) - 1) < 0.1
    assert abs(float(loss) - float(lw)) < 2e-2 * float(lw)
    import ctypes
    from ultrasound_modeling_b200 import _lib
    L = _lib.lib()
    L.tbi_fallback_stats(None, None, 1)
    net.train_step(x, y)
    a_, b_ = ctypes.c_int64(0), ctypes.c_int64(0)
    L.tbi_fallback_stats(ctypes.byref(a_), ctypes.byref(b_), 0)
    assert (a_.value, b_.value) == (0, 0), (a_.value, b_.value, L.tbi_last_fallback())      # the whole step stays on tcgen05


def test_cuda_graph_replay_equals_eager(cuda_device):
    """train_step / step / forward through CUDA-graph replay (two eager calls, capture, replays) track the eager path over six
    optimizer steps, across a batch-size change and back (graphs are keyed by input shape and variable-storage generation)"""
    from ultrasound_modeling_b200.VisionTransformer import VisionTransformer
    o = V.VisionTransformerOracle(2, img_size=(64, 32), num_classes=3, learning_rate=1e-3, dtype=torch.float64, num_layers=2)
    nets = [VisionTransformer(2, img_size=(64, 32), num_classes=3, learning_rate=1e-3, dtype="fp32", device=str(cuda_device), num_layers=2,
                              use_cuda_graph=g) for g in (False, True)]
    for net in nets:
        net.load_variables(o.state_dict())
    x2 = V.B.synthetic_input(2, 64, 32, 10); y2 = V.synthetic_labels(2, 64, 32)
    x4 = V.B.synthetic_input(4, 64, 32, 10, seed=77); y4 = V.synthetic_labels(4, 64, 32, seed=78)
    # (1) at FIXED variables: backward() through replay == eager, tensor by tensor.  Not bit-equal: fp32 sums (split-K partials,
    # atomics) are taken in a different order from one net / call to the next; the key bias gradient is exactly 0 in exact
    # arithmetic (softmax shift invariance), i.e. pure rounding noise, and is skipped.
    for (x, y) in [(x2, y2), (x4, y4)]:
        for _ in range(4):                                   # two eager warm-ups, the capture, one replay
            (la, _), (lb, _) = nets[0].backward(x, y), nets[1].backward(x, y)
        ga, gb = nets[0].gradients(), nets[1].gradients()
        assert abs(float(la) - float(lb)) < 1e-5 * abs(float(la))
        worst = max((rel(gb[k], ga[k]), k) for k in ga if not k.endswith("attn/key/bias"))
        assert worst[0] < 2e-2, worst                           # a stale replay is off by O(1); summation-order noise measured <= 1e-3
    # step() / forward() through replay at the same (identical) variables
    for _ in range(4):
        (la, pa), (lb, pb) = nets[0].step(x2, y2), nets[1].step(x2, y2)
        fa, fb = nets[0].forward(x4), nets[1].forward(x4)
    assert abs(float(la) - float(lb)) < 1e-5 * abs(float(la)) and rel(pb, pa) < 1e-4 and rel(fb[0], fa[0]) < 1e-4
    assert len(fb[1]) == 2 and rel(fb[1][0], fa[1][0]) < 1e-4
    # (2) six optimizer steps across a batch-size change and back.  Two EAGER nets already part ways here (measured on B200,
    # scratch/vit_div.py: loss 1e-7 apart for the first steps, then up to 7e-4 by step 10; ~20 % of the variables more than
    # 2e-5 apart, 1.3e-3 at most): summation-order differences of ~1e-4 in a few cancellation-heavy gradients are turned into
    # full-size steps by global-norm clipping + Adam.  The bounds below are that chaos with margin; a replay that used stale
    # inputs or stale variable storage is off by O(1) (the x2 and x4 losses differ by 2x).
    losses = [[], []]
    for (x, y) in [(x2, y2)] * 4 + [(x4, y4)] * 4 + [(x2, y2)] * 2:
        for i, net in enumerate(nets):
            loss, probs = net.train_step(x, y)
            losses[i].append(float(loss))
    assert len(nets[1]._graphs) >= 2 and all(e["graph"] is not None for e in nets[1]._graphs.values())
    assert max(abs(a - b) / abs(a) for a, b in zip(*losses[:2])) < 2e-2, losses
    assert max(abs(a - b) / abs(a) for a, b in zip(losses[0][:3], losses[1][:3])) < 1e-5, losses
    va, vb = nets[0].variables(), nets[1].variables()
    diffs = torch.cat([(va[k] - vb[k]).abs().reshape(-1) for k in va])
    assert float(diffs.mean()) < 2e-3 and float(diffs.max()) < 3e-2      # 10 Adam steps of at most ~lr each: 2e-2 is the hard bound on a difference
    # the replayed step / forward graphs read the UPDATED variables (same storage): still the eager net's answer up to the divergence above
    for _ in range(2):
        (la, pa), (lb, pb) = nets[0].step(x2, y2), nets[1].step(x2, y2)
        fa, fb = nets[0].forward(x4), nets[1].forward(x4)
    assert abs(float(la) - float(lb)) < 2e-2 * abs(float(la)) and rel(pb, pa) < 0.25 and rel(fb[0], fa[0]) < 0.25     # O(1) if a replay read stale storage
    # a learning-rate change reaches the replayed step (the Adam tail reads a device buffer)
    before = nets[1].variables()["decoder/head/bias"].clone()
    nets[1].optimizer.learning_rate = 0.0
    nets[1].train_step(x2, y2)
    assert torch.equal(before, nets[1].variables()["decoder/head/bias"])
