#!/usr/bin/env python
"""Headline benchmark: TBI_ResNest bf16 training step (fwd + my_loss_cat + bwd + [all-reduce] + Adam),
batch 64 per GPU, synthetic 1x256x256 frames (BASELINE.json configs[2]); weak scaling over N GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--radix 2 --kpaths 1]

Prints ONE JSON line (rank 0).  `value` = whole-job img/s with inputs resident in HBM; `e2e` = the
same through ResNest.step() with pinned HOST inputs and a device->host read of loss+accuracy inside
the timed region; `roofline` = the dominant kernel timed live with CUDA events; `cpu_baseline` =
the CPU oracle (PyTorch restatement of the reference; TensorFlow itself is not installable here)
timed on the host cores on a bounded sample.  --impl reference times that CPU oracle alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "TBI_ResNest train img/s", "img/s"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def cpu_oracle_rate(radix, kpaths, size, seconds=15.0, batch=2):
    """img/s of the CPU oracle's train step on a bounded sample of the workload"""
    import torch
    from oracle import tbi_resnest_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = O.TBIResNestOracle(size, size, 1, 3, 3, radix, kpaths, learning_rate=5e-3)
    x, y = O.synthetic_batch(batch, size, size, seed=3000)
    m = O.dropout_masks(batch, size, size)
    o.step(x, y, True, m)                                        # warm-up
    t0 = time.perf_counter(); steps = 0
    while steps < 2 or (time.perf_counter() - t0 < seconds and steps < 64):
        o.step(x, y, True, m); steps += 1
    dt = time.perf_counter() - t0
    return batch * steps / dt, cores, f"{steps} train steps of batch {batch} at {size}x{size}x1 (fp32, torch CPU, {cores} threads)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    per_step = []
    import torch
    from oracle import tbi_resnest_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = O.TBIResNestOracle(args.size, args.size, 1, 3, 3, args.radix, args.kpaths, learning_rate=5e-3)
    B = 2                                                        # bounded sample: batch 2 per step (config[0] shape)
    x, y = O.synthetic_batch(B, args.size, args.size, seed=3000)
    m = O.dropout_masks(B, args.size, args.size)
    for _ in range(args.warmup):
        o.step(x, y, True, m)
    for _ in range(args.steps):
        t = time.perf_counter(); o.step(x, y, True, m); per_step.append(time.perf_counter() - t)
    dt = sum(per_step)
    val = B * args.steps / dt
    sample = f"{args.steps} train steps of batch {B} (of the {args.batch}-image batch) at {args.size}x{args.size}x1, fp32 torch CPU oracle, {cores} threads"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": workload_name(args), "note": "reference TF graph not installable; CPU restatement (oracle) timed"},
                      "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0, "wall_s": time.perf_counter() - t0}))


def other_configs(args, dev):
    """The BASELINE.json configurations other than the headline one, measured in the same driver-run process so that the
    record carries them (rank 0, one GPU): configs[1] split-attention microbench (three shapes, fwd / bwd HBM GB/s),
    configs[2] with the reference-default radix 4 x kpaths 4 and the reference driver's own radix 3 x kpaths 4
    (TBI_ResNest.py:461), configs[3] 512x512 r4k4 batch 16, configs[4] the inference sweep (Variant A forward, and Variant B
    encoder+decoder forward at the [N,256,80,10] shape TBIEvaluator.py:186,238 feeds).  Each is a short run: CUDA events,
    3 warm-up calls, device-resident synthetic inputs."""
    import torch
    from oracle import tbi_resnest_oracle as O
    from oracle import resnest_decoder_oracle as B
    from ultrasound_modeling_b200.TBI_ResNest import ResNest
    import bench_splitatt

    def timed(f, reps, warm=3):
        for _ in range(warm):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {}
    sa = bench_splitatt.main(emit=False, reps=5)
    out["splitatt_config1"] = [{"shape_U_r": r["shape_U_r"], "fwd_us": round(r["fwd"]["us"], 2), "fwd_GBps": round(r["fwd"]["GBps"], 1),
                                "fwd_frac_of_hbm_peak": round(r["fwd"]["frac_of_hbm_peak"], 3), "bwd_us": round(r["bwd"]["us"], 2),
                                "bwd_GBps": round(r["bwd"]["GBps"], 1), "bwd_frac_of_hbm_peak": round(r["bwd"]["frac_of_hbm_peak"], 3),
                                "peak_gbs": r["peak_gbs"], "method": r["method"]} for r in sa]
    torch.cuda.empty_cache()
    train = []
    for (hw, n, r, k, tag) in ((256, 64, 4, 4, "configs[2] with reference defaults r4k4"), (256, 64, 3, 4, "configs[2] with the reference driver's r3k4"),
                               (512, 16, 4, 4, "configs[3]: 512x512 r4k4 batch 16")):
        net = ResNest(hw, hw, 1, 3, 3, radix=r, kpaths=k, dtype="bf16", use_cuda_graph=True, device=str(dev))
        x, y = O.synthetic_batch(2, hw, hw, seed=3000)
        x = x.repeat(n // 2, 1, 1, 1).to(dev); y = y.repeat(n // 2, 1, 1, 1).to(dev)
        net.engine.fallback_report(reset=True)
        ms = timed(lambda: net.step(x, y, train=True), 10, warm=4)
        fb = net.engine.fallback_report()
        train.append({"config": tag, "size": hw, "batch": n, "radix": r, "kpaths": k, "ms_per_step": round(ms, 3), "img_per_s": round(n / ms * 1e3, 1),
                      "cuda_core_fallback_launches": fb["tapgemm_simt"] + fb["tapwgrad_simt"]})
        del net, x, y
        torch.cuda.empty_cache()
    out["train_other"] = train
    infer = []
    net = ResNest(256, 256, 1, 3, 3, radix=2, kpaths=1, dtype="bf16", use_cuda_graph=False, device=str(dev))
    for n in (1, 8, 64, 256):
        x, _ = O.synthetic_batch(min(n, 4), 256, 256)
        x = x.repeat((n + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:n].to(dev)
        ms = timed(lambda: net.predict(x, dropout_masks=False), 10)
        infer.append({"model": "Variant A forward 256x256x1 r2k1 bf16", "batch": n, "ms": round(ms, 3), "img_per_s": round(n / ms * 1e3, 1)})
    del net
    torch.cuda.empty_cache()
    try:
        from ultrasound_modeling_b200.ResNest import ResNest as EncB
        from ultrasound_modeling_b200.Decoder import DecoderCup
        for n in (1, 32):
            enc = EncB(256, 80, 10, 3, radix=3, kpaths=3, dtype="bf16", device=str(dev)); dec = DecoderCup(3, dtype="bf16", device=str(dev))
            x = B.synthetic_input(n).to(dev); tok = B.synthetic_tokens(n).to(dev)
            ms = timed(lambda: dec(tok, enc(x)[1]), 5)
            infer.append({"model": "Variant B encoder+decoder forward [N,256,80,10] r3k3 bf16", "batch": n, "ms": round(ms, 3), "img_per_s": round(n / ms * 1e3, 1)})
            del enc, dec
    except Exception as exc:                                     # noqa: BLE001 -- a side measurement must not take the headline down
        infer.append({"model": "Variant B", "error": repr(exc)[:200]})
    out["inference_sweep_config4"] = infer
    # the model MainParallel.py trains (VisionTransformer.py: Variant B encoder + 8-block ViT bridge + DecoderCup), per-replica
    # batches of the reference driver (MainParallel.py:205 uses a global batch of 64), CUDA-graph replay of loss + backward
    vb = []
    try:
        from oracle import vit_oracle as V
        from ultrasound_modeling_b200.VisionTransformer import VisionTransformer
        for n, graph in ((16, True), (64, True), (16, False)):
            net = VisionTransformer(n, dtype="bf16", device=str(dev), use_cuda_graph=graph)
            x = V.B.synthetic_input(2, 256, 80, 10).repeat(n // 2, 1, 1, 1).to(dev); y = V.synthetic_labels(2, 256, 80).repeat(n // 2, 1, 1, 1).to(dev)
            ms = timed(lambda: net.train_step(x, y), 5, warm=4)
            msf = timed(lambda: net.forward(x), 5, warm=4)
            vb.append({"model": "VisionTransformer train_step / forward [N,256,80,10] bf16", "batch": n, "cuda_graph": graph, "train_ms": round(ms, 3),
                       "train_img_per_s": round(n / ms * 1e3, 1), "forward_ms": round(msf, 3), "forward_img_per_s": round(n / msf * 1e3, 1)})
            del net, x, y
            torch.cuda.empty_cache()
    except Exception as exc:                                     # noqa: BLE001
        vb.append({"model": "VisionTransformer", "error": repr(exc)[:300]})
    out["variant_b_vit"] = vb
    torch.cuda.empty_cache()
    return out


def workload_name(args):
    return (f"TBI_ResNest full training step, batch {args.batch}/GPU, {args.size}x{args.size}x1, radix {args.radix} kpaths {args.kpaths} "
            f"(BASELINE.json configs[2])")


def run_ours(args):
    import torch
    import torch.distributed as dist
    from oracle import tbi_resnest_oracle as O
    from ultrasound_modeling_b200.TBI_ResNest import ResNest
    from ultrasound_modeling_b200.parallel import GradSync
    from ultrasound_modeling_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sync = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        sync = GradSync()
    pk, pk_kind = peaks()

    net = ResNest(args.size, args.size, 1, 3, 3, radix=args.radix, kpaths=args.kpaths, learning_rate=5e-3, dtype=args.dtype,
                  device=f"cuda:{local}", use_cuda_graph=not args.no_graph, grad_sync=sync, seed=1236)
    B = args.batch
    x, y = O.synthetic_batch(B, args.size, args.size, seed=3000 + rank)          # host, fp32 NHWC
    x_pin, y_pin = x.pin_memory(), y.pin_memory()
    e = net.engine

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ---------------------------------------------------
    net.step(x_pin, y_pin, train=True)                      # builds buffers, first (eager) call
    for _ in range(max(args.warmup, 3)):
        net._run(True, True)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        net._run(True, True)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if sampler:
        sampler.stop_flag.set(); sampler.join(2)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through ResNest.step with host inputs (`e2e`) -------------------------------
    host_out = torch.empty(args.size * args.size + 1, dtype=torch.float32).pin_memory()
    for _ in range(2):
        net.step(x_pin, y_pin, train=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, acc, _ = net.step(x_pin, y_pin, train=True)
        host_out[:-1].copy_(loss.reshape(-1), non_blocking=True)
        host_out[-1:].copy_(acc.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()           # the caller reads loss/accuracy every step
    e1.record()
    barrier()
    ms_e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e = world * B * args.steps / (float(ms_e.item()) / 1e3)
    h2d = x_pin.numel() * 4 + y_pin.numel() * 4
    d2h = host_out.numel() * 4

    out = None
    if rank == 0:
        # ---- dominant kernel, timed live: the persistent halo tap-GEMM (tapgemm_halo_kernel<128>, ~19 % of the step) on its
        # longest launch of the step, upsample_4's transposed conv forward (one 4-phase launch).  Algorithmic flops per launch
        # = 2 * N*h*w * 16 taps * Cin * Cout; `traffic` = dram read+write bytes of that launch from the ncu
        # --set full capture summarised in profiles/r2_ncu_upsample4_fwd.md (same shape, N = 64 only).
        st = torch.cuda.current_stream().cuda_stream
        name = "upsample_4"
        calls = [(fn, a) for fn, a in e.prog_fwd if fn.__name__ == "tbi_conv2d_transpose_s2_fwd"]
        fn, a = calls[4]
        h, w, cin, cout = args.size >> 2, args.size >> 2, 320, 128
        flops_call = 2.0 * B * h * w * 16 * cin * cout
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        flush_rd = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
        for _ in range(3):
            fn(*a, st)
        reps, kms = 10, 0.0
        for _ in range(reps):                                   # L2 flushed between launches: each launch is timed cold, as in the step
            flush.zero_(); flush_rd.max()                       # write, then read 256 MB: the L2 is left with CLEAN foreign lines
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(); fn(*a, st); k1.record(); torch.cuda.synchronize()
            kms += k0.elapsed_time(k1) / reps
        del flush, flush_rd
        ach = flops_call / (kms / 1e3) / 1e12
        peak = pk["bf16_tflops"]                                 # burst figure: this kernel is timed alone (B200_PROFILING.md)
        peak_sus = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])   # sustained figure: for the whole step
        step_flops = 3.0 * O.forward_flops_per_image(args.size, args.size, 1, 3, 3, args.radix, args.kpaths) * B
        # dram__bytes_read.sum + dram__bytes_write.sum of exactly this launch (N=64, 256x256, bf16) in the committed capture
        # profiles/r2_ncu_upsample4_fwd.md, end-of-round re-capture (169.2 MB read + 223.2 MB written; algorithmic 436 MB: the tail
        # of the output is still in L2 when the kernel ends); null for any other shape
        traffic = 392.37e6 if (args.dtype == "bf16" and B == 64 and args.size == 256) else None
        roof = {"bound": "tensor", "kernel": f"tapgemm_halo_kernel<128, pair>: {name} Conv2DTranspose k4 s2 fwd [{B},{h},{w},{cin}]->[{B},{2*h},{2*w},{cout}] (one 4-phase launch, cta_group::2)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "traffic_source": "profiles/r2_ncu_upsample4_fwd.md (ncu --set full capture of this launch)",
                "peak_source": pk_kind + " (burst bf16 = what cuBLAS reached on this pool, kernel timed alone with L2 flushed between launches; a frac near or above 1 means this kernel runs at the library GEMM's measured rate, not at the hardware limit)",
                "frac_of_nominal_dense_bf16": ach / 2250.0,
                "ms_per_call": kms, "flops_per_call": flops_call, "step_tflops": step_flops * world * args.steps / (ms / 1e3) / 1e12 / world,
                "step_frac_of_peak_sustained": step_flops * args.steps / (ms / 1e3) / 1e12 / peak_sus}
        cpu_val, cores, sample = cpu_oracle_rate(args.radix, args.kpaths, args.size, seconds=args.cpu_seconds)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": args.dtype, "data": "synthetic",
               "config": {"workload": workload_name(args), "global_batch": world * B, "parallelism": f"dp{world}",
                          "l2": "working set (activations+gradients, ~GBs) exceeds the 126 MB L2; no explicit flush",
                          "cuda_graph": not args.no_graph},
               "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": e.launches_per_step(True) * args.steps,
               "roofline": roof,
               "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
               "clocks": sampler.summary() if sampler else None}
        if world == 1 and not args.no_extras and B == 64 and args.size == 256:
            del net, e
            torch.cuda.empty_cache()
            try:
                out["other_configs"] = other_configs(args, dev)
            except Exception as exc:                             # noqa: BLE001 -- side measurements must not take the headline down
                out["other_configs"] = {"error": repr(exc)[:300]}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    if os.environ.get("TBI_BENCH_WATCHDOG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["TBI_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--radix", type=int, default=2)
    ap.add_argument("--kpaths", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (split-attention microbench, r4k4/r3k4/512, inference sweep)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
