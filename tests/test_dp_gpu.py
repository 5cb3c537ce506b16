"""Data-parallel parity (SURVEY 8e "DP parity definition"): an n-replica step on global batch B must equal the oracle that
evaluates the loss PER SHARD (my_loss_cat normalises by per-pixel class counts over the batch it sees, TBI_ResNest.py:240)
and averages the shard gradients -- all-reduce(SUM) then x 1/world, what compute_average_loss + MirroredStrategy's SUM do
(VisionTransformer.py:225-227, MainParallel.py:130).

  * test_dp_definition_one_gpu: the definition itself on ONE device (runs on the driver's single-GPU box): two shards
    through the same engine, gradients summed in the flat buffer, Adam with grad_scale 1/2, three steps, vs the oracle.
  * test_dp_two_gpus_nccl: the real thing when >= 2 GPUs are visible (gpurun --gpus 2): two processes, NCCL, GradSync with
    small buckets, eager AND graph-segment replay; both must track the per-shard-average oracle and each other.
"""
import os
import socket

import pytest
import torch

from oracle import tbi_resnest_oracle as O

pytestmark = pytest.mark.gpu
HW, R, K, LR = 64, 2, 1, 5e-3


def oracle_dp_steps(shards, masks, steps, world):
    """reference semantics: per-shard loss, gradients averaged over replicas, one Adam update per step"""
    o = O.TBIResNestOracle(HW, HW, 1, 3, 3, R, K, learning_rate=LR, dtype=torch.float64)
    sd0 = o.state_dict()
    grads_first = None
    for _ in range(steps):
        gs = [o.gradients(x.double(), y.double(), m) for (x, y), m in zip(shards, masks)]
        avg = {k: sum(g[k] for g in gs) / world for k in gs[0]}
        if grads_first is None:
            grads_first = avg
        o.apply_adam(avg)
    return sd0, grads_first, o.state_dict()


def make_shards(world, per):
    x, y = O.synthetic_batch(world * per, HW, HW)
    m = O.dropout_masks(world * per, HW, HW)
    shards = [(x[r * per:(r + 1) * per], y[r * per:(r + 1) * per]) for r in range(world)]
    masks = [[t[r * per:(r + 1) * per] for t in m] for r in range(world)]
    return shards, masks


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def sync_relu_ties(e, inter, tie_tol=1e-5, max_flips=8):
    """see tests/test_model_gpu.py: a decoder unit the fp32 path and the fp64 oracle round to opposite sides of 0 is a tie"""
    total = 0
    for i in range(5):
        ref = inter[f"upsample_{i}"].detach()
        got = e.up[i].double().cpu()
        flips = (got > 0) != (ref > 0)
        if int(flips.sum()):
            assert float(torch.maximum(ref[flips].abs(), got[flips].abs()).max()) < tie_tol * float(ref.abs().max())
            e.up[i].copy_(torch.where(flips, ref, got).to(e.up[i].dtype))
        total += int(flips.sum())
    assert total <= max_flips, total


def test_dp_definition_one_gpu(cuda_device):
    from ultrasound_modeling_b200.TBI_ResNest import ResNest
    world, per, steps = 2, 2, 3
    shards, masks = make_shards(world, per)
    o = O.TBIResNestOracle(HW, HW, 1, 3, 3, R, K, learning_rate=LR, dtype=torch.float64)
    net = ResNest(HW, HW, 1, 3, 3, radix=R, kpaths=K, learning_rate=LR, dtype="fp32", use_cuda_graph=False)
    net.load_state_dict(o.state_dict())
    e = net.engine
    e.build(per)
    for s in range(steps):
        total = torch.zeros_like(e.grads)
        want = []
        for (x, y), m in zip(shards, masks):
            net.step(x, y, train=False, dropout_masks=m)         # forward + loss of THIS shard (per-shard class counts)
            _, inter = o.forward(x.double(), m, return_intermediates=True)
            sync_relu_ties(e, inter)
            e.backward()                                         # zeroes, then fills the flat gradient buffer
            total += e.grads
            want.append(o.gradients(x.double(), y.double(), m))
        avg = {k: sum(g[k] for g in want) / world for k in want[0]}
        e.grads.copy_(total)                                     # == all-reduce(SUM)
        got = e.grad_dict()
        worst = max((rel(got[k] / world, avg[k]), k) for k in avg)
        assert worst[0] < 1e-4, (s, worst)
        e.adam(LR, 1.0 / world)
        o.apply_adam(avg)
    got, sd_want = net.state_dict(), o.state_dict()
    worst = max((float((got[k].double().cpu() - sd_want[k]).abs().max()), k) for k in sd_want)
    assert worst[0] < LR * 2e-2, worst


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from ultrasound_modeling_b200.TBI_ResNest import ResNest
        from ultrasound_modeling_b200.parallel import GradSync
        per, steps = 2, 4
        shards, masks = make_shards(world, per)
        sd0, g_first, sd_want = oracle_dp_steps(shards, masks, steps, world)
        out = {}
        for mode, graph in (("eager", False), ("graph", True)):
            gs = GradSync(bucket_bytes=8 << 20)                  # ~13 buckets over the 108 MB of gradients
            # different seeds: attach() must broadcast rank 0's variables
            net = ResNest(HW, HW, 1, 3, 3, radix=R, kpaths=K, learning_rate=LR, dtype="fp32", device=f"cuda:{rank}",
                          use_cuda_graph=graph, grad_sync=gs, seed=100 + rank)
            ref = [torch.zeros_like(net.engine.params) for _ in range(world)]
            dist.all_gather(ref, net.engine.params)
            assert all(float((r - ref[0]).abs().max()) == 0.0 for r in ref), "variables differ across replicas after attach()"
            net.load_state_dict(sd0)
            (x, y), m = shards[rank], masks[rank]
            for s in range(steps):
                net.step(x, y, train=True, dropout_masks=m)
                if s == 0 and mode == "eager":
                    got = net.engine.grad_dict()                 # all-reduced SUM
                    out["grad_err"] = max(rel(got[k] / world, g_first[k]) for k in g_first)
            torch.cuda.synchronize()
            got = net.state_dict()
            diffs = torch.cat([(got[k].double().cpu() - sd_want[k]).abs().reshape(-1) for k in sd_want])
            out[mode + "_param_err"] = float(diffs.max())
            out[mode + "_param_bad_frac"] = float((diffs > LR * 2e-2).double().mean())
            allp = [torch.zeros_like(net.engine.params) for _ in range(world)]
            dist.all_gather(allp, net.engine.params)
            out[mode + "_replica_diff"] = max(float((p - allp[0]).abs().max()) for p in allp)
            out[mode + "_params"] = net.engine.params.clone()
        d = (out.pop("eager_params") - out.pop("graph_params")).abs()
        out["graph_vs_eager_frac_moved"] = float((d > 1e-4).float().mean())
        q.put((rank, out))
    except Exception as exc:                                     # noqa: BLE001
        import traceback
        q.put((rank, {"error": f"{exc!r}\n{traceback.format_exc()}"}))
    finally:
        dist.destroy_process_group()


def test_dp_two_gpus_nccl(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=900) for _ in ps)
    for p in ps:
        p.join(120)
    print("DP2", res)
    for r, out in res.items():
        assert "error" not in out, out["error"]
        # the 1e-4 gradient bar is held (with ReLU ties synced) by test_dp_definition_one_gpu; inside a full step no tie can be
        # synced, and ONE tie flip moves the deep, tiny gradients by ~1e-3
        assert out["grad_err"] < 5e-3, out
        # Adam turns a gradient at rounding-noise level into a full +-lr step, and inside a real step no ReLU tie can be synced
        # with the oracle (the one-GPU test does that and holds lr*2e-2 on EVERY weight): here all but a vanishing fraction of
        # the 26.9 M weights must be within lr*2e-2 after 4 steps, and none further than the 4 steps could carry it
        assert out["eager_param_bad_frac"] < 1e-3 and out["graph_param_bad_frac"] < 1e-3, out      # measured 1.2e-4
        assert out["eager_param_err"] < 4 * 2 * LR and out["graph_param_err"] < 4 * 2 * LR, out
        assert out["eager_replica_diff"] == 0.0 and out["graph_replica_diff"] == 0.0, out      # replicas stay bit-identical
        assert out["graph_vs_eager_frac_moved"] < 1e-3, out
