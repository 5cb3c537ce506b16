"""timeline of block 0 of the streamed halo kernel: python scratch/trace_stream.py up4|up3|cc2"""
import sys, ctypes, torch
sys.path.insert(0, '.')
from ultrasound_modeling_b200 import ops, _lib
case = sys.argv[1] if len(sys.argv) > 1 else "up4"
L = _lib.lib()
BF = torch.bfloat16
if case == "up4":
    x1 = torch.randn(64, 64, 64, 256, device="cuda").to(BF); x2 = torch.randn(64, 64, 64, 64, device="cuda").to(BF)
    w = torch.randn(4, 4, 128, 320, device="cuda") * 0.02; b = torch.zeros(128, device="cuda")
    f = lambda: ops.conv2d_transpose_s2(x1, w, b, act=ops._lib.ACT_RELU, x2=x2)
elif case == "up3":
    x1 = torch.randn(64, 32, 32, 512, device="cuda").to(BF); x2 = torch.randn(64, 32, 32, 128, device="cuda").to(BF)
    w = torch.randn(4, 4, 256, 640, device="cuda") * 0.02; b = torch.zeros(256, device="cuda")
    f = lambda: ops.conv2d_transpose_s2(x1, w, b, act=ops._lib.ACT_RELU, x2=x2)
else:
    x = torch.randn(64, 64, 64, 64, device="cuda").to(BF); w = torch.randn(3, 3, 64, 128, device="cuda") * 0.05; b = torch.zeros(128, device="cuda")
    r = torch.randn(64, 64, 64, 128, device="cuda").to(BF)
    f = lambda: ops.conv2d(x, w, b, residual=r)
for _ in range(2): f()
tr = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
fn = L.tbi_debug_set_halo_trace; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
assert fn(tr.data_ptr()) == 0
f(); torch.cuda.synchronize()
fn(None)
t = tr.cpu().view(3, 64, 8)
ev = t[:, :, :4]
t0 = int(ev[ev > 0].min())
rel = lambda v: (int(v) - t0) if int(v) > 0 else -1
print("MMA per tile(-pair): [loop_start, got_t_empty, last_a_full, committed] | cycles waiting for A, for B | tile cycles")
for i in range(4, 20):
    print("M %2d" % i, [rel(v) for v in t[1, i, :4]], "| wait_a %6d wait_b %6d | %6d" % (int(t[1, i, 4]), int(t[1, i, 5]), int(t[1, i, 3]) - int(t[1, i - 1, 3])))
print("EPI warp0: [loop_start, got_t_full, ld_done, arrived, stores_issued]")
for i in range(4, 12): print("E %2d" % i, [rel(v) for v in t[2, i, :5]])
d = [int(t[1, i + 1, 3]) - int(t[1, i, 3]) for i in range(4, 24)]
wa = [int(t[1, i, 4]) for i in range(4, 24)]; wb = [int(t[1, i, 5]) for i in range(4, 24)]; we = [int(t[1, i, 1]) - int(t[1, i, 0]) for i in range(4, 24)]
print("steady state per tile(-pair): cycles %.0f, waiting t_empty %.0f, A %.0f, B %.0f" % (sum(d) / len(d), sum(we) / len(we), sum(wa) / len(wa), sum(wb) / len(wb)))
