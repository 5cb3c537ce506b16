"""Parity of the BENCHMARKED arithmetic at the benchmarked shapes (VERDICT r1 "next round" item 1).

(a) whole graph, bf16 storage, 256x256x1, batch 2 -- BASELINE.json configs[0] in the arithmetic configs[2] times -- for
    radix 2 / kpaths 1 (the headline), the reference defaults radix 4 / kpaths 4, and the reference driver's own
    radix 3 / kpaths 4 (TBI_ResNest.py:461): probabilities, loss, accuracy and EVERY parameter gradient against the fp64
    oracle.  At 256x256 every persistent tcgen05 CTA of the full-resolution layers runs several tiles (stem: 512 M tiles over
    <= 296 CTAs), so TMEM buffer reuse, ring parity wrap and resident-slab striding are on the compared path.
(b) single layers on the tcgen05 path at the headline layer shapes with batch >= 8 (each persistent CTA >= 4 tiles, split-K
    weight gradients with many partials) against the oracle's conv definitions evaluated in fp64 on the same bf16 inputs.
(c) the data-parallel definition of SURVEY 8(e) is in tests/test_dp_gpu.py.

Bars (north_star): bf16 probabilities within 2e-2 relative, argmax agreement >= 99.9 % (see check_argmax), every gradient
tensor within 2e-2 of its largest entry.  What is asserted beyond that is printed: the ReLU tie-flip count (bounded), the
worst tensors, and which tensors (if any) sit between 2e-2 and the hard bound.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import tbi_resnest_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_fullres.jsonl")


def report(**kw):
    print("PARITY", json.dumps(kw))
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


@pytest.fixture(scope="module")
def ResNest(cuda_device):
    from ultrasound_modeling_b200.TBI_ResNest import ResNest as R
    return R


@pytest.fixture(scope="module")
def ops(cuda_device):
    from ultrasound_modeling_b200 import ops as _ops
    return _ops


def rel(got, want):
    want = want.detach().double().cpu(); got = got.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


def rel2(got, want):
    want = want.detach().double().cpu(); got = got.detach().double().cpu()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


def count_and_sync_relu_ties(net, inter, tie_tol):
    """ReLU' is discontinuous at 0: a decoder unit whose pre-activation the bf16 path and the fp64 oracle round to opposite
    sides of zero flips its derivative although both values are ~0.  Count such units, assert each IS a tie (|value| below
    bf16 rounding of the tensor's scale), adopt the oracle's side for them, and return (flips, units)."""
    e = net.engine
    flips_total, units = 0, 0
    for i in range(5):
        ref = inter[f"upsample_{i}"].detach()
        got = e.up[i].double().cpu()
        flips = (got > 0) != (ref > 0)
        nf = int(flips.sum())
        units += ref.numel()
        if nf:
            worst = float(torch.maximum(ref[flips].abs(), got[flips].abs()).max())
            assert worst < tie_tol * float(ref.abs().max()), ("a flipped ReLU unit is not a tie", i, worst)
            e.up[i].copy_(torch.where(flips, ref, got).to(e.up[i].dtype))
        flips_total += nf
    return flips_total, units


# (radix, kpaths, max fraction of decoder ReLU units that may be ties).  Measured on B200 (profiles/r2_parity.md): 3575 /
# 3965 / 3652 of 7 667 712 units = 4.7e-4 / 5.2e-4 / 4.8e-4 -- a property of bf16 rounding of pre-activations that sit at ~0;
# the bound is 2x the measurement.
CASES = [(2, 1, 1e-3), (4, 4, 1e-3), (3, 4, 1e-3)]


@pytest.mark.parametrize("radix,kpaths,max_flip_frac", CASES)
def test_whole_graph_bf16_256(ResNest, radix, kpaths, max_flip_frac):
    hw, n = 256, 2
    o = O.TBIResNestOracle(hw, hw, 1, 3, 3, radix, kpaths, dtype=torch.float64)
    net = ResNest(hw, hw, 1, 3, 3, radix=radix, kpaths=kpaths, dtype="bf16", use_cuda_graph=False)
    net.load_state_dict(o.state_dict())
    net.engine.fallback_report(reset=True)
    x, y = O.synthetic_batch(n, hw, hw)
    masks = O.dropout_masks(n, hw, hw)
    loss, acc, probs = net.step(x, y, train=False, dropout_masks=masks)
    probs = probs.clone(); loss = loss.clone()
    want_probs, inter = o.forward(x.double(), masks, return_intermediates=True)
    want_loss = o.my_loss_cat(y.double(), want_probs)
    e_probs = rel(probs, want_probs)
    same = probs.argmax(-1).cpu() == want_probs.argmax(-1)
    top2 = want_probs.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    worst_margin = float(margin[~same].max()) if (~same).any() else 0.0
    agree = float(same.float().mean())
    e_loss = rel(loss, want_loss)
    want_acc = float((want_probs.argmax(-1) == y.argmax(-1)).float().mean())
    # gradients
    flips, units = count_and_sync_relu_ties(net, inter, tie_tol=1e-2)
    net.engine.backward()
    torch.cuda.synchronize()
    got = net.engine.grad_dict()
    want = o.gradients(x.double(), y.double(), masks)
    assert set(got) == set(want)
    errs = sorted((rel(got[k], want[k]), k) for k in want)
    errs2 = sorted((rel2(got[k], want[k]), k) for k in want)
    over = [(round(v, 4), k) for v, k in errs if v >= 2e-2]
    # the same comparison at the granularity the layers are EXECUTED at: the K*R cardinal branches of a stage are one fused
    # 1x1 conv / one grouped 3x3 conv / one split-attention launch, i.e. one weight tensor each in the engine's flat buffer.
    # A branch whose gradient is 20-40x smaller than its siblings' (a nearly dead pair of ELU channels in this random init;
    # profiles/r2_parity.md) carries the same ABSOLUTE bf16 noise as they do, which is a large fraction of its own tiny scale.
    e = net.engine
    want_flat = torch.zeros_like(e.grads)
    for k, dst in e._named(want_flat, None, trainable_only=True).items():
        dst.copy_(want[k].to(dst.dtype).reshape(dst.shape))
    fused = sorted((rel(e.P.get(e.grads, nm), e.P.get(want_flat, nm)), nm) for nm in e.P.specs)
    fb = net.engine.fallback_report()
    report(test="whole_graph_bf16_256", radix=radix, kpaths=kpaths, probs_rel=e_probs, argmax_agree=agree, worst_disagreeing_margin=worst_margin,
           loss_rel=e_loss, relu_ties=flips, relu_units=units, relu_tie_frac=flips / units, grad_tensors=len(errs),
           grad_maxabs_median=errs[len(errs) // 2][0], grad_maxabs_worst=errs[-1], grad_2norm_worst=errs2[-1],
           grad_tensors_over_2e2=over, fused_tensors=len(fused), fused_maxabs_worst=fused[-3:], simt_fallbacks=fb)
    assert e_probs < 2e-2, e_probs
    # a random-init net answers ~(1/3,1/3,1/3): an argmax disagreement must be a top-2 tie of the ORACLE within the tolerance
    assert worst_margin < 4e-2 and agree >= 0.99, (worst_margin, agree)
    assert e_loss < 1e-1 * 1.0 and e_loss < 5 * 2e-2, e_loss
    assert abs(float(acc) - want_acc) < 2e-3
    assert flips <= max_flip_frac * units, (flips, units)
    # every executed (fused) gradient tensor within 2e-2 of its largest entry ...
    assert fused[-1][0] < 2e-2, fused[-5:]
    # ... and per Keras variable: all but a few percent within 2e-2; the exceptions are listed in the report line above
    assert len(over) <= max(3, (3 * len(errs)) // 100), over
    assert fb["tapgemm_simt"] == 0 and fb["tapwgrad_simt"] == 0, fb      # nothing left the tensor cores


def test_argmax_999_on_a_trained_like_head_256(ResNest):
    """the 99.9 % argmax bar on margins like a trained network's (head scaled x30), at full resolution"""
    hw, n = 256, 2
    o = O.TBIResNestOracle(hw, hw, 1, 3, 3, 2, 1, dtype=torch.float64)
    sd = o.state_dict()
    sd["f_tran/kernel"] = sd["f_tran/kernel"] * 30
    o2 = O.TBIResNestOracle(hw, hw, 1, 3, 3, 2, 1, params=sd, dtype=torch.float64)
    net = ResNest(hw, hw, 1, 3, 3, radix=2, kpaths=1, dtype="bf16", use_cuda_graph=False)
    net.load_state_dict(sd)
    x, y = O.synthetic_batch(n, hw, hw)
    masks = O.dropout_masks(n, hw, hw)
    _, _, p = net.step(x, y, train=False, dropout_masks=masks)
    w = o2.forward(x.double(), masks)
    agree = float((p.argmax(-1).cpu() == w.argmax(-1)).float().mean())
    err = float((p.double().cpu() - w).abs().max())
    report(test="argmax_trained_like_head_256", argmax_agree=agree, probs_abs=err)
    assert agree >= 0.999, agree
    assert err < 2e-2, err


# ------------------------------------------------------------------------------------------------------------------
# (b) single layers at headline shapes, batch >= 8
# ------------------------------------------------------------------------------------------------------------------
def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(BF)


def conv_oracle(x, wt, b, k):
    return O.conv2d_same(x.double().cpu(), wt.to(BF).double().cpu(), b.double().cpu() if b is not None else None)


@pytest.mark.parametrize("name,n,h,cin,cout", [
    ("conv2_1_2 (stem 32->32 @256^2)", 8, 256, 32, 32),
    ("conv2_1_1 (stem 16->32 @256^2)", 8, 256, 16, 32),
    ("conv2_2/cc2 (concats_2 64->128 @64^2)", 40, 64, 64, 128),
    ("conv2_1/cc2 (concats_2 32->64 @128^2)", 10, 128, 32, 64),
    ("conv3_1/cc2 (concats_2 128->256 @32^2; CTA-pair weight gradient, M = dz)", 48, 32, 128, 256),
    ("conv3_2/cc2 (concats_2 256->512 @16^2; CTA-pair weight gradient)", 64, 16, 256, 512),
])
def test_conv3x3_layer_tcgen05_vs_oracle(ops, name, n, h, cin, cout):
    """forward (bias + ELU + residual epilogue), data gradient (fused ELU') and weight/bias gradient of a 3x3 conv on the
    tcgen05 path vs fp64 torch on the SAME bf16 inputs.  Tiles per persistent CTA: n*(h/16)*(h/8) M tiles over <= 296 CTAs."""
    x = rnd(n, h, h, cin, seed=1)
    wt = (torch.randn(3, 3, cin, cout, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)) / (9 * cin) ** 0.5)
    b = torch.randn(cout, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 0.1
    res = rnd(n, h, h, cout, seed=4)
    y = ops.conv2d(x, wt, b, act=ops.ACT_ELU, residual=res, impl=ops._lib.IMPL_TCGEN05)
    want = F.elu(conv_oracle(x, wt, b, 3)) + res.double().cpu()
    e_fwd = rel(y, want)
    dz = rnd(n, h, h, cout, seed=5)
    yref = rnd(n, h, h, cin, seed=6)
    dx, dw, db = ops.conv2d_grads(x, wt, dz, impl=ops._lib.IMPL_TCGEN05, dact=ops.ACT_ELU, dact_ref=yref)
    torch.cuda.synchronize()
    xd = x.double().cpu().requires_grad_(True)
    wd = wt.to(BF).double().cpu().requires_grad_(True)        # dgrad uses the bf16-packed weights
    z = O.conv2d_same(xd, wd, None)
    gx, = torch.autograd.grad(z, xd, dz.double().cpu(), retain_graph=True)
    yr = yref.double().cpu()
    want_dx = gx * torch.where(yr > 0, torch.ones_like(yr), yr + 1)
    # the weight gradient does not involve the weights: A^T dz on the bf16 activations
    gw, = torch.autograd.grad(z, wd, dz.double().cpu())
    want_db = dz.double().cpu().sum((0, 1, 2))
    e_dx, e_dw, e_db = rel(dx, want_dx), rel(dw, gw), rel(db, want_db)
    tiles = n * ((h + 15) // 16) * ((h + 7) // 8)
    report(test="conv3x3_layer", layer=name, m_tiles=tiles, fwd=e_fwd, dgrad=e_dx, wgrad=e_dw, dbias=e_db)
    assert tiles >= 4 * 296 or cout >= 256
    assert e_fwd < 1e-2 and e_dx < 1e-2, (e_fwd, e_dx)       # bf16 output rounding
    assert e_dw < 2e-3 and e_db < 2e-3, (e_dw, e_db)         # fp32 outputs: only the summation order differs


@pytest.mark.parametrize("name,n,h,c1,c2,cout,drop", [
    ("upsample_4 [n,64,64,256+64]->128", 8, 64, 256, 64, 128, False),
    ("upsample_3 [n,32,32,512+128]->256", 16, 32, 512, 128, 256, False),
    ("upsample_2 [n,16,16,512+256]->512", 32, 16, 512, 256, 512, True),
])
def test_convt_layer_tcgen05_vs_oracle(ops, name, n, h, c1, c2, cout, drop):
    """Conv2DTranspose k4 s2 over a virtual concat of two sources: forward (bias + dropout multiplier + ReLU), data gradient
    split over both sources, weight + bias gradient; tcgen05 path vs fp64 torch on the same bf16 inputs."""
    x1, x2 = rnd(n, h, h, c1, seed=11), rnd(n, h, h, c2, seed=12)
    cin = c1 + c2
    wt = torch.randn(4, 4, cout, cin, device="cuda", generator=torch.Generator(device="cuda").manual_seed(13)) / (4 * cin) ** 0.5
    b = torch.randn(cout, device="cuda", generator=torch.Generator(device="cuda").manual_seed(14)) * 0.1
    keep = None
    if drop:
        keep = (torch.rand(n, 2 * h, 2 * h, cout, device="cuda", generator=torch.Generator(device="cuda").manual_seed(15)) < 0.5).to(torch.uint8) * 2
    y = ops.conv2d_transpose_s2(x1, wt, b, act=ops._lib.ACT_RELU, keep=keep, x2=x2, impl=ops._lib.IMPL_TCGEN05)
    dz = rnd(n, 2 * h, 2 * h, cout, seed=16)
    (dx1, dx2), dw, db = ops.conv2d_transpose_s2_grads(x1, wt, dz, x2=x2, impl=ops._lib.IMPL_TCGEN05)
    torch.cuda.synchronize()
    xc = torch.cat([x1, x2], 3).double().cpu().requires_grad_(True)
    wd = wt.to(BF).double().cpu().requires_grad_(True)
    z = O.conv2d_transpose_s2_same(xc, wd, None)
    pre = z + b.double().cpu()
    if keep is not None:
        pre = pre * keep.double().cpu()
    e_fwd = rel(y, F.relu(pre))
    gx, gw = torch.autograd.grad(z, (xc, wd), dz.double().cpu())
    e_dx1, e_dx2 = rel(dx1, gx[..., :c1]), rel(dx2, gx[..., c1:])
    e_dw, e_db = rel(dw, gw), rel(db, dz.double().cpu().sum((0, 1, 2)))
    tiles = 4 * n * ((h + 15) // 16) * ((h + 7) // 8) * ((cout + 127) // 128)
    report(test="convt_layer", layer=name, tiles=tiles, fwd=e_fwd, dgrad_src0=e_dx1, dgrad_src1=e_dx2, wgrad=e_dw, dbias=e_db)
    assert e_fwd < 1e-2 and e_dx1 < 1e-2 and e_dx2 < 1e-2, (e_fwd, e_dx1, e_dx2)
    assert e_dw < 2e-3 and e_db < 2e-3, (e_dw, e_db)


def test_graph_survives_batch_size_round_trip(ResNest):
    """train at batch a, evaluate at batch b, train at a again -- with CUDA graphs on.  The engine keeps the
    buffers of the last few batch sizes and graphs are keyed on the build generation, so the replay after the round trip
    must equal an eager engine fed the same sequence."""
    x, y = O.synthetic_batch(4, 64, 64)
    masks = O.dropout_masks(4, 64, 64)
    o = O.TBIResNestOracle(64, 64, 1, 3, 3, 2, 1, dtype=torch.float64)
    nets = []
    for graph in (False, True):
        net = ResNest(64, 64, 1, 3, 3, radix=2, kpaths=1, dtype="fp32", use_cuda_graph=graph, learning_rate=1e-3)
        net.load_state_dict(o.state_dict())
        nets.append(net)
    seq = [4, 4, 4, 1, 1, 4, 4, 2, 3, 5, 4, 4]            # 5 distinct sizes: more than the engine caches -> one eviction
    eager, graph = nets
    for step, n in enumerate(seq):
        outs = []
        for net in nets:
            net.optimizer.learning_rate = 1e-3 if step < 6 else 2e-4         # ADVICE (medium): lr change must reach the graph
            l, a, p = net.step(x[:n], y[:n], train=(n != 1), dropout_masks=[m[:n] for m in masks])
            outs.append((l.clone(), p.clone()))
        assert float((outs[0][1] - outs[1][1]).abs().max()) < 1e-4, (step, n)
        assert float((outs[0][0] - outs[1][0]).abs().max()) < 1e-5, (step, n)
        # the two engines sum their fp32 weight-gradient atomics in different orders and Adam turns a rounding-level gradient
        # into a +-lr step: compare the update of THIS step, then restart both from the same state so that noise does not
        # compound over the 10 training steps
        d = (eager.engine.params - graph.engine.params).abs()
        assert float((d > 1e-4).float().mean()) < 1e-3, (step, n)
        for name in ("params", "adam_m", "adam_v"):
            getattr(graph.engine, name).copy_(getattr(eager.engine, name))
    assert int(eager.engine.step_count.item()) == int(graph.engine.step_count.item()) == sum(1 for n in seq if n != 1)


def test_learning_rate_reaches_a_captured_graph(ResNest):
    x, y = O.synthetic_batch(2, 64, 64)
    masks = O.dropout_masks(2, 64, 64)
    net = ResNest(64, 64, 1, 3, 3, radix=2, kpaths=1, dtype="fp32", use_cuda_graph=True, learning_rate=1e-3)
    for _ in range(3):
        net.step(x, y, train=True, dropout_masks=masks)                      # eager warm-up, capture, replay
    before = net.engine.params.clone()
    net.optimizer.learning_rate = 0.0
    net.step(x, y, train=True, dropout_masks=masks)
    assert float((net.engine.params - before).abs().max()) == 0.0            # a replay with lr = 0 moves nothing
    net.optimizer.learning_rate = 1e-3
    net.step(x, y, train=True, dropout_masks=masks)
    assert float((net.engine.params - before).abs().max()) > 0.0


def test_pageable_host_inputs_are_not_overwritten_in_flight(ResNest):
    """numpy / pageable inputs go through a ring of pinned staging buffers guarded by events; feeding a
    different batch every step without synchronising must give the same result as feeding device tensors."""
    net = ResNest(64, 64, 1, 3, 3, radix=2, kpaths=1, dtype="fp32", use_cuda_graph=True)
    ref = ResNest(64, 64, 1, 3, 3, radix=2, kpaths=1, dtype="fp32", use_cuda_graph=False)
    ref.load_state_dict(net.state_dict())
    masks = O.dropout_masks(2, 64, 64)
    batches = [O.synthetic_batch(2, 64, 64, seed=100 + i) for i in range(6)]
    got = []
    for xb, yb in batches:                                  # no synchronisation between steps
        l, a, p = net.step(xb.numpy().astype("float64"), yb.numpy(), train=False, dropout_masks=masks)
        got.append(p.clone())
    for (xb, yb), p in zip(batches, got):
        _, _, pw = ref.step(xb.cuda(), yb.cuda(), train=False, dropout_masks=masks)
        assert float((p - pw).abs().max()) < 1e-5


def test_checkpoint_round_trip_keras_layout(ResNest, tmp_path):
    """save_params / load_params / resModel.save write .npz archives keyed by the reference's Keras variable names in Keras
    layouts (HWIO / HWOI): names and shapes equal the oracle's variable table, values and Adam state survive the trip."""
    import numpy as np
    a = ResNest(64, 64, 1, 3, 3, radix=3, kpaths=4, dtype="fp32", use_cuda_graph=False, ckpt_dir=str(tmp_path / "ck"), seed=7)
    x, y = O.synthetic_batch(2, 64, 64)
    masks = O.dropout_masks(2, 64, 64)
    a.step(x, y, train=True, dropout_masks=masks)
    a.save_params()
    z = np.load(tmp_path / "ck" / "variables.npz")
    want = O.param_shapes(1, 3, 3, 3, 4)
    assert set(z.files) == set(want)
    for k, shp in want.items():
        assert tuple(z[k].shape) == tuple(shp), k
    b = ResNest(64, 64, 1, 3, 3, radix=3, kpaths=4, dtype="fp32", use_cuda_graph=False, ckpt_dir=str(tmp_path / "ck"), seed=8)
    b.load_params()
    assert float((a.engine.params - b.engine.params).abs().max()) == 0.0
    assert float((a.engine.adam_m - b.engine.adam_m).abs().max()) == 0.0 and float((a.engine.adam_v - b.engine.adam_v).abs().max()) == 0.0
    assert int(b.engine.step_count.item()) == 1
    la, _, pa = a.step(x, y, train=True, dropout_masks=masks)
    lb, _, pb = b.step(x, y, train=True, dropout_masks=masks)
    assert float((pa - pb).abs().max()) < 1e-6
    # the oracle (the reference's graph restated) loads the archive by name and reproduces the device forward
    a.resModel.save(str(tmp_path / "model"))
    zz = np.load(tmp_path / "model" / "variables.npz")
    o = O.TBIResNestOracle(64, 64, 1, 3, 3, 3, 4, params={k: torch.from_numpy(zz[k]).double() for k in zz.files}, dtype=torch.float64)
    _, _, pd = a.step(x, y, train=False, dropout_masks=masks)
    assert rel(pd, o.forward(x.double(), masks)) < 1e-4
    c = ResNest(64, 64, 1, 3, 3, radix=3, kpaths=4, dtype="fp32", use_cuda_graph=False, seed=9)
    c.resModel.load(str(tmp_path / "model"))
    assert float((c.engine.params - a.engine.params).abs().max()) == 0.0


@pytest.mark.parametrize("n,h,w", [(5, 48, 24), (3, 20, 12), (1, 16, 8)])
def test_cta_pair_mode_ragged_shapes(ops, n, h, w):
    """The CTA-pair schedule of the persistent halo kernel (cta_group::2: one MMA spans two pixel tiles on two SMs, each CTA
    streams half of every weight stage; taken for streamed 128-column tiles once a layer has a tile pair per SM) on shapes
    the model never produces: an ODD number of pixel tiles (the last pair's second CTA repeats a tile and its epilogue must
    drop it), partial tiles in x and y, two sources, and a single tile.  The threshold is lowered through the library's test
    knob; results must equal fp64 on the same bf16 inputs and the single-CTA schedule bit for bit (the same products are
    accumulated in the same order per output element)."""
    L = ops._lib.lib()
    import ctypes
    L.tbi_debug_set_pair_min.argtypes = [ctypes.c_int]
    x1, x2 = rnd(n, h, w, 64, seed=21), rnd(n, h, w, 64, seed=22)
    w3 = torch.randn(3, 3, 128, 256, device="cuda", generator=torch.Generator(device="cuda").manual_seed(23)) / (9 * 128) ** 0.5
    wt = torch.randn(4, 4, 128, 128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(24)) / (4 * 128) ** 0.5
    b3 = torch.randn(256, device="cuda", generator=torch.Generator(device="cuda").manual_seed(25)) * 0.1
    bt = torch.randn(128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(26)) * 0.1
    dz = rnd(n, 2 * h, 2 * w, 128, seed=27)
    outs = {}
    try:
        for mode, pair_min in (("pair", 1), ("single", 1 << 30)):
            L.tbi_debug_set_pair_min(pair_min)
            y3 = ops.conv2d(x1, w3, b3, act=ops.ACT_ELU, x2=x2, impl=ops._lib.IMPL_TCGEN05)
            yt = ops.conv2d_transpose_s2(x1, wt, bt, act=ops._lib.ACT_RELU, x2=x2, impl=ops._lib.IMPL_TCGEN05)
            (dx1, dx2), _, _ = ops.conv2d_transpose_s2_grads(x1, wt, dz, x2=x2, impl=ops._lib.IMPL_TCGEN05)
            torch.cuda.synchronize()
            outs[mode] = (y3, yt, dx1, dx2)
    finally:
        L.tbi_debug_set_pair_min(0)
    for a, b in zip(outs["pair"], outs["single"]):
        assert torch.equal(a, b)
    xc = torch.cat([x1, x2], 3).double().cpu().requires_grad_(True)
    want3 = F.elu(O.conv2d_same(xc, w3.to(BF).double().cpu(), b3.double().cpu()))
    wd = wt.to(BF).double().cpu()
    z = O.conv2d_transpose_s2_same(xc, wd, None)
    gx, = torch.autograd.grad(z, xc, dz.double().cpu())
    y3, yt, dx1, dx2 = outs["pair"]
    errs = (rel(y3, want3), rel(yt, F.relu(z + bt.double().cpu())), rel(dx1, gx[..., :64]), rel(dx2, gx[..., 64:]))
    report(test="cta_pair_ragged", n=n, h=h, w=w, conv3x3=errs[0], convt_fwd=errs[1], convt_dgrad=errs[2:])
    assert max(errs) < 1e-2, errs


@pytest.mark.parametrize("case", [
    # n, h, w, (c1, c2), cout, k     weight gradient on CTA pairs (tapwgrad_tc2.cu): even / odd numbers of 128-channel tiles of the M
    (3, 16, 16, (256, 0), 128, 4),       # convT, M = x (2 tiles = 1 pair), N = dz 128
    (2, 16, 24, (256, 64), 128, 4),      # convT, two sources, 320 channels: one pair + the 64-channel rest on the single-CTA kernel
    (2, 8, 8, (512, 128), 256, 4),       # convT, 640 channels: two pairs + rest, two N tiles
    (3, 16, 16, (128, 0), 256, 3),       # 3x3 conv: taps share dz -> M = dz (256 = 1 pair), N = x (128)
    (2, 12, 20, (128, 0), 384, 3),       # 3x3 conv: 3 M tiles -> 1 pair + rest (co tiles), partial pixel tiles
    (5, 8, 8, (128, 0), 256, 1),         # 1x1 conv (one tap): M = dz
    (5, 8, 8, (256, 0), 128, 1),         # 1x1 conv (one tap): M = x
    (2, 16, 16, (64, 64), 256, 3),       # 3x3 conv, two-source x on the N side (64 + 64)
])
def test_wgrad_cta_pairs_vs_fp64(ops, case):
    n, h, w, (c1, c2), cout, k = case
    cin = c1 + c2
    x1 = rnd(n, h, w, c1, seed=31); x2 = rnd(n, h, w, c2, seed=32) if c2 else None
    xc = (torch.cat([x1, x2], 3) if c2 else x1).double().cpu()
    if k == 4:
        dz = rnd(n, 2 * h, 2 * w, cout, seed=33)
        wt = torch.zeros(4, 4, cout, cin, device="cuda")
        _, dw, db = ops.conv2d_transpose_s2_grads(x1, wt, dz, x2=x2, need_dx=False, impl=ops._lib.IMPL_TCGEN05)
        wd = wt.double().cpu().requires_grad_(True)
        z = O.conv2d_transpose_s2_same(xc, wd, None)
    else:
        dz = rnd(n, h, w, cout, seed=33)
        wt = torch.zeros(k, k, cin, cout, device="cuda")
        _, dw, db = ops.conv2d_grads(x1, wt, dz, x2=x2, need_dx=False, impl=ops._lib.IMPL_TCGEN05)
        wd = wt.double().cpu().requires_grad_(True)
        z = O.conv2d_same(xc, wd, None)
    torch.cuda.synchronize()
    gw, = torch.autograd.grad(z, wd, dz.double().cpu())
    e_dw, e_db = rel(dw, gw), rel(db, dz.double().cpu().sum((0, 1, 2)))
    report(test="wgrad_cta_pairs", case=str(case), wgrad=e_dw, dbias=e_db)
    assert e_dw < 2e-3 and e_db < 2e-3, (e_dw, e_db)


@pytest.mark.parametrize("n,h,w,c0,c1", [(16, 128, 64, 128, 32),      # the head's data gradient, 1 024 pixel tiles: 3-4 per persistent CTA
                                         (3, 48, 40, 128, 32),       # partial tiles in x and y
                                         (2, 32, 32, 64, 32)])       # 96 = 3 x 32 columns, split inside the slab list
def test_split_output_all_slabs_vs_fp64(ops, n, h, w, c0, c1):
    """A one-chunk (K = 64) 1x1 data gradient whose output is the virtual concat of two tensors with a ragged column count (the
    head: 160 = 128 + 32, TBI_ResNest.py:124 backward): the halo kernel's ALL-SLABS resident mode (every 32-column weight slab in
    one CTA, one halo per pixel tile, one accumulator per slab), with ReLU' from a stored forward output on the first tensor and
    a plain write on the second, and the side-input boxes prefetched one tile ahead.  Against fp64 on the same bf16 inputs, at
    a size where every persistent CTA walks several pixel tiles (ring wrap, TMEM buffer reuse, prefetch flush at the end)."""
    x, x2 = rnd(n, h, w, c0, seed=31), rnd(n, h, w, c1, seed=32)
    wt = torch.randn(1, 1, c0 + c1, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(33)) / 8.0
    dz = rnd(n, h, w, 64, seed=34)
    (dx, dx2), dw, db = ops.conv2d_grads(x, wt, dz, x2=x2, impl=ops._lib.IMPL_TCGEN05, dact=ops._lib.ACT_RELU, dact_ref=x,
                                         wgrad_impl=ops._lib.IMPL_AUTO)
    torch.cuda.synchronize()
    g = dz.double().cpu() @ wt.to(BF).double().cpu()[0, 0].T                     # [n,h,w,c0+c1]
    want1 = g[..., :c0] * (x.double().cpu() > 0)
    want2 = g[..., c0:]
    e1, e2 = rel(dx, want1), rel(dx2, want2)
    xc = torch.cat([x, x2], 3).double().cpu()
    gw = torch.einsum("nhwi,nhwo->io", xc, dz.double().cpu())
    e3 = rel(dw[0, 0], gw)
    report(test="all_slabs_split_dgrad", n=n, h=h, w=w, c=(c0, c1), dx=e1, dx2=e2, dw=e3)
    assert e1 < 1e-2 and e2 < 1e-2 and e3 < 1e-3, (e1, e2, e3)
