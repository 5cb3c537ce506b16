"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/tbi_sm100.h
declares, the parameter inventory maps 1:1 onto the reference's Keras variables, bucket planning, and
the world-size-2 gradient exchange over gloo.  No kernel is launched here."""
import os
import re
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tbi_resnest_oracle as O
from ultrasound_modeling_b200 import _lib
from ultrasound_modeling_b200.engine import Engine
from ultrasound_modeling_b200.parallel import plan_buckets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tbi_sm100.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tbi_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = _lib.lib()                      # builds with nvcc if the .so is absent
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in tbi_sm100.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.tbi_version() == 100


def test_no_device_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(_lib.TbiError):
        Engine(64, 64, 1, 3, 3, 2, 1)


@pytest.mark.parametrize("radix,kpaths", [(2, 1), (4, 4), (3, 4), (1, 2)])
def test_parameter_inventory_matches_reference_variables(radix, kpaths):
    e = Engine(256, 256, 1, 3, 3, radix, kpaths, layout_only=True)
    want = O.param_shapes(1, 3, 3, radix, kpaths)
    got = {}
    for kn, store, flat, idx in e._keras_items():
        spec = (e.P if store == "P" else e.S).specs[flat]
        got[kn] = tuple(torch.empty(spec.shape)[idx].shape)
    assert set(got) == set(want)
    for k, shp in want.items():
        assert got[k] == tuple(shp), (k, got[k], shp)
    n_train = sum(torch.Size(s).numel() for n, s in want.items() if O.is_trainable(n))
    mapped = sum(torch.Size(got[kn]).numel() for kn, store, _, _ in e._keras_items() if store == "P")
    assert mapped == n_train
    # what the flat buffer holds beyond the Keras variables: the zero pad channels of the fused 1x1 cardinal convs whose
    # width G*cv11 is not a multiple of 16 (radix 3), each with a kernel column, a bias, gamma and beta
    pad = sum((i["c1w"] - i["G"] * i["cv11"]) * (i["cin"] + 3) for i in e.stage_info)
    assert sum(s.numel for s in e.P.specs.values()) == n_train + pad
    assert (pad > 0) == (radix == 3)


def test_plan_buckets_tiles_the_buffer():
    marks = [(3, 900), (5, 700), (9, 650), (12, 100), (20, 0)]
    plan = plan_buckets(marks, 1000, 200)
    assert plan == [(5, 700, 1000), (12, 100, 700), (20, 0, 100)]
    cover = sorted((lo, hi) for _, lo, hi in plan)
    assert cover[0][0] == 0 and cover[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert [c for c, _, _ in plan] == sorted(c for c, _, _ in plan)
    assert plan_buckets(marks, 1000, 10 ** 9) == [(20, 0, 1000)]


def test_plan_buckets_at_named_boundaries():
    from ultrasound_modeling_b200.parallel import plan_buckets_at
    marks = [(3, 900), (5, 700), (9, 650), (12, 100), (20, 0)]
    assert plan_buckets_at(marks, 1000, (700, 100)) == [(5, 700, 1000), (12, 100, 700), (20, 0, 100)]
    assert plan_buckets_at(marks, 1000, (800,)) == [(5, 800, 1000), (20, 0, 800)]           # a cut between marks waits for the next mark
    assert plan_buckets_at(marks, 1000, ()) == [(20, 0, 1000)]
    # the engine's cut: [decoder + the two deepest encoder stages] | rest, tiling the flat buffer (swept on 2 GPUs, engine.py)
    e = Engine(256, 256, 1, 3, 3, 2, 1, layout_only=True)
    cut, = e.bucket_cuts
    assert 0 < cut < e.P.total and (e.P.total - cut) * 4 > 95 << 20 and cut * 4 < 8 << 20


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ultrasound_modeling_b200.parallel import GradSync
    gs = GradSync(bucket_bytes=4 * 300)
    total = 1000
    torch.manual_seed(rank)
    flat = torch.randn(total)
    mine = flat.clone()
    marks = [(1, 800), (2, 640), (3, 300), (4, 0)]
    for _, lo, hi in plan_buckets(marks, total, gs.bucket_elems):
        gs.allreduce(flat[lo:hi])
    gathered = [torch.zeros(total) for _ in range(world)]
    dist.all_gather(gathered, mine)
    # the evaluation-side collectives of MainParallel.py:159-163: gather along the batch axis, SUM of scalars
    shard = torch.full((2, 3, 1), float(rank))
    g = gs.gather(shard)
    ok_gather = tuple(g.shape) == (2 * world, 3, 1) and all(float(g[2 * r].mean()) == r for r in range(world))
    ok_sum = float(gs.reduce_sum(torch.tensor(float(rank + 1)))) == world * (world + 1) / 2
    q.put((rank, bool(torch.allclose(flat, sum(gathered), atol=1e-6)) and ok_gather and ok_sum, gs.world_size))
    dist.destroy_process_group()


def test_gradient_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert all(ok and ws == 2 for _, ok, ws in res), res


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores: the TensorFlow graph cannot run here) prints
    one JSON line with the driver's keys; under torchrun only rank 0 works."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--size", "64"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env={**os.environ, "RANK": "0"})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env={**os.environ, "RANK": "1"})
    assert other.returncode == 0 and other.stdout.strip() == ""
