import sys, ctypes, torch
sys.path.insert(0, '.')
from ultrasound_modeling_b200 import ops, _lib
case = sys.argv[1] if len(sys.argv) > 1 else "s32"
L = _lib.lib()
BF = torch.bfloat16
if case == "up4":
    x1 = torch.randn(64, 64, 64, 256, device="cuda").to(BF); x2 = torch.randn(64, 64, 64, 64, device="cuda").to(BF)
    w = torch.randn(4, 4, 128, 320, device="cuda") * 0.02; b = torch.zeros(128, device="cuda")
    f = lambda: ops.conv2d_transpose_s2(x1, w, b, act=ops._lib.ACT_RELU, x2=x2)
elif case == "wide":          # 1x1 conv 64 -> 128 with a residual at 128x128: K tiny, N wide -> epilogue-paced
    x = torch.randn(64, 128, 128, 64, device="cuda").to(BF); w = torch.randn(1, 1, 64, 128, device="cuda") * 0.05
    b = torch.zeros(128, device="cuda"); r = torch.randn(64, 128, 128, 128, device="cuda").to(BF)
    f = lambda: ops.conv2d(x, w, b, residual=r)
elif case == "headd":
    x1 = torch.randn(64, 128, 128, 128, device="cuda").to(BF); x2 = torch.randn(64, 128, 128, 32, device="cuda").to(BF)
    w = torch.randn(4, 4, 3, 160, device="cuda") * 0.02
    dz = torch.zeros(64, 256, 256, 16, device="cuda", dtype=BF); dz[..., :3] = torch.randn(64, 256, 256, 3, device="cuda").to(BF)
    f = lambda: ops.conv2d_transpose_s2_grads(x1, w, dz, x2=x2, wgrad_impl=ops._lib.IMPL_TCGEN05)
else:
    C = int(case[1:]); k = 3 if case[0] == 's' else 1
    x = torch.randn(64, 256, 256, C, device="cuda").to(BF); w = torch.randn(k, k, C, 32, device="cuda") * 0.05; b = torch.zeros(32, device="cuda")
    f = lambda: ops.conv2d(x, w, b, act=ops.ACT_ELU)
for _ in range(2): f()
tr = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
fn = L.tbi_debug_set_halo_trace; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
assert fn(tr.data_ptr()) == 0
f(); torch.cuda.synchronize()
fn(None)
t = tr.cpu().view(3, 64, 8)
t0 = int(t[t > 0].min())
rel = lambda v: (int(v) - t0) if int(v) > 0 else -1
print("cycles relative to first event; producer: [wait_start, got_empty] per halo load")
for i in range(20, 32): print("P %2d" % i, [rel(v) for v in t[0, i, :2]])
print("MMA: [loop_start, got_t_empty, got_a_full, committed]")
for i in range(20, 32): print("M %2d" % i, [rel(v) for v in t[1, i, :4]])
print("EPI warp2: [loop_start, got_t_full, ld_done, arrived, stores_issued]")
for i in range(20, 32): print("E %2d" % i, [rel(v) for v in t[2, i, :5]])
d = [int(t[1, i + 1, 3]) - int(t[1, i, 3]) for i in range(20, 40)]
print("steady-state cycles per tile (MMA commit to commit):", sum(d) / len(d))
