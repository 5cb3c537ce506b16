#!/usr/bin/env python
"""BASELINE.json configs[1]: ResNeSt split-attention tail microbench (radix 2, cardinality 1, batch 32, bf16)
at the three in-network shapes.  Reports achieved HBM GB/s on the ALGORITHMIC bytes of SURVEY 8d:
fwd (2R+1)*NHWC*b, bwd (2R+2)*NHWC*b, against MEASURED_PEAKS.json hbm_gbs.  One JSON line per shape.
L2 note: the largest shape (U = 67 MB) is below the 126 MB L2, so an explicit L2 flush (a 512 MB memset)
runs between timed iterations."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from ultrasound_modeling_b200 import ops  # noqa: E402


def main():
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; src = "measured"
    except Exception:
        peak, src = 6650.0, "fallback"
    R, K, N = 2, 1, 32
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for (h, c) in ((128, 32), (64, 64), (32, 128)):
        torch.manual_seed(2000 + c)
        u = torch.randn(N, h, h, K * R * c, device="cuda").to(torch.bfloat16)
        dv = torch.randn(N, h, h, K * c, device="cuda").to(torch.bfloat16)
        D = lambda *s: torch.randn(*s, device="cuda")
        sa = ops.SplitAttention(K, R, c, D(K, c, c // 2) * 0.2, D(K, c // 2) * 0.1, 1 + 0.1 * D(K, c // 2), 0.1 * D(K, c // 2),
                                0.1 * D(K, c // 2), 0.5 + torch.rand(K, c // 2, device="cuda"), D(K, R, c // 2, c) * 0.2, D(K, R, c) * 0.1)
        nhwc_b = N * h * h * c * 2
        res = {}
        for name, fn, passes in (("fwd", lambda: sa.forward(u), 2 * R + 1), ("bwd", lambda: sa.backward(u, dv), 2 * R + 2)):
            sa.forward(u)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()                    # the engine replays these kernels from a CUDA graph: time them the same way
            with torch.cuda.graph(g):
                fn()
            times = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = sorted(times)[len(times) // 2]
            res[name] = {"us": ms * 1e3, "algorithmic_MB": passes * nhwc_b / 1e6, "GBps": passes * nhwc_b / ms / 1e6,
                         "frac_of_hbm_peak": passes * nhwc_b / ms / 1e6 / peak}
        print(json.dumps({"metric": "split-attention HBM GB/s", "shape_U_r": [N, h, h, c], "radix": R, "kpaths": K, "dtype": "bf16",
                          "peak_gbs": peak, "peak_source": src, "l2": "flushed between iterations (512 MB memset)",
                          "note": "CUDA-graph replay of the whole op (memset + reduce + FC + recombine | memset + reduce + FC-bwd + param-grads + dU)", **res}))


if __name__ == "__main__":
    main()
