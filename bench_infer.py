"""SURVEY 8d configs 4 and 5 (not the driver's bench.py): forward-only sweep of Variant A over batch sizes, the 512x512
r4k4 training configuration, the r4k4 training step at 256x256, and Variant B encoder+decoder forward at [N,256,80,10].
Prints one JSON line per measurement; CUDA events, 3 warm-up calls, inputs resident on the device."""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from oracle import tbi_resnest_oracle as O          # synthetic inputs only
from oracle import resnest_decoder_oracle as B      # synthetic inputs only


def timed(f, reps):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--what", default="all")
    a = ap.parse_args()
    from ultrasound_modeling_b200.TBI_ResNest import ResNest
    if a.what in ("all", "infer_a"):
        for n in (1, 2, 4, 8, 16, 32, 64, 128, 256):
            net = ResNest(256, 256, 1, 3, 3, radix=2, kpaths=1, dtype="bf16", use_cuda_graph=False)
            x, _ = O.synthetic_batch(min(n, 4), 256, 256)
            x = x.repeat((n + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:n].cuda()
            ms = timed(lambda: net.predict(x, dropout_masks=False), a.reps)
            print(json.dumps({"config": "5: Variant A forward, 256x256x1, r2k1, bf16", "batch": n, "ms": round(ms, 3), "img_per_s": round(n / ms * 1e3, 1)}), flush=True)
            del net
    if a.what in ("all", "train_r4k4"):
        for (hw, n, tag) in ((256, 64, "3b: 256x256 r4k4 N=64"), (512, 16, "4: 512x512 r4k4 N=16")):
            net = ResNest(hw, hw, 1, 3, 3, radix=4, kpaths=4, dtype="bf16", use_cuda_graph=True)
            x, y = O.synthetic_batch(2, hw, hw)
            x = x.repeat(n // 2, 1, 1, 1).cuda(); y = y.repeat(n // 2, 1, 1, 1).cuda()
            ms = timed(lambda: net.step(x, y, train=True), a.reps)
            print(json.dumps({"config": tag + " training step, bf16, CUDA graph", "batch": n, "ms": round(ms, 3), "img_per_s": round(n / ms * 1e3, 1),
                              "launches": net.engine.launches_per_step(True)}), flush=True)
            del net
    if a.what in ("all", "infer_b"):
        from ultrasound_modeling_b200.ResNest import ResNest as EncB
        from ultrasound_modeling_b200.Decoder import DecoderCup
        for dt in ("bf16", "fp32"):
            for n in (1, 8, 32):
                enc = EncB(256, 80, 10, 3, radix=3, kpaths=3, dtype=dt); dec = DecoderCup(3, dtype=dt)
                x = B.synthetic_input(n).cuda(); tok = B.synthetic_tokens(n).cuda()
                ms = timed(lambda: dec(tok, enc(x)[1]), a.reps)
                print(json.dumps({"config": "5: Variant B encoder+decoder forward, [N,256,80,10], r3k3, " + dt, "batch": n, "ms": round(ms, 3),
                                  "img_per_s": round(n / ms * 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
