"""Phase timeline of the cluster split-attention kernels: thread 0 of every CTA stamps clock64 at the phase boundaries
(csrc/splitatt_fused.cu: stamp()).  Prints median / max over CTAs of every phase in SM cycles and the kernel span in ns."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from ultrasound_modeling_b200 import ops, _lib
L = _lib.lib()
fn = L.tbi_debug_set_splitatt_trace; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
names = ["start", "pass1(t0)", "reduce", "barrierA", "pull", "fc1(+B)", "fc2(+C)", "softmax", "pass2"]
R, K, N = 2, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 32
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda"); flush_rd = torch.zeros(128 << 20, dtype=torch.int32, device="cuda")
for (h, c) in ((128, 32), (64, 64), (32, 128)):
    u = torch.randn(N, h, h, K * R * c, device="cuda").to(torch.bfloat16); dv = torch.randn(N, h, h, K * c, device="cuda").to(torch.bfloat16)
    D = lambda *s: torch.randn(*s, device="cuda")
    sa = ops.SplitAttention(K, R, c, D(K, c, c // 2) * 0.2, D(K, c // 2) * 0.1, 1 + 0.1 * D(K, c // 2), 0.1 * D(K, c // 2), 0.1 * D(K, c // 2),
                            0.5 + torch.rand(K, c // 2, device="cuda"), D(K, R, c // 2, c) * 0.2, D(K, R, c) * 0.1)
    for name, f in (("fwd", lambda: sa.forward(u)), ("bwd", lambda: sa.backward(u, dv))):
        sa.forward(u); f(); f()
        tr = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
        flush.zero_(); flush_rd.max(); torch.cuda.synchronize()
        fn(tr.data_ptr()); f(); torch.cuda.synchronize(); fn(None)
        t = tr.cpu().view(4096, 16); t = t[t[:, 0] > 0]
        span = int(t[:, 15].max() - t[:, 0].min())
        d = (t[:, 2:10] - t[:, 1:9]).float()
        print(f"[{N},{h},{h},{c}] {name}: {t.shape[0]} CTAs, span {span} ns, start spread {int(t[:,0].max()-t[:,0].min())} ns, "
              f"per-CTA total median {float((t[:,15]-t[:,0]).float().median()):.0f} ns")
        print("   " + "  ".join(f"{names[i+1]} {float(d[:, i].median()):.0f}/{float(d[:, i].max()):.0f}" for i in range(8)) + "   (median/max cycles)")
