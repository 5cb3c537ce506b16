"""single-layer microbench: python scratch/mb_conv.py <case> [reps]"""
import sys, torch
sys.path.insert(0, '.')
from ultrasound_modeling_b200 import ops
case = sys.argv[1] if len(sys.argv) > 1 else "stem"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(0)
BF = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(BF)
if case == "stem":          # conv2_1_2: 3x3 32->32 +BN+ELU at 256x256, N=64
    x = rnd(64, 256, 256, 32); w = torch.randn(3, 3, 32, 32, device="cuda") * 0.05; b = torch.zeros(32, device="cuda")
    f = lambda: ops.conv2d(x, w, b, act=ops.ACT_ELU)
    flops = 2 * 64 * 256 * 256 * 9 * 32 * 32; bytes_ = x.numel() * 2 * 2
elif case in ("s16", "s32", "s64", "s128"):   # 3x3 C->32 at 256x256 with C input channels (row bytes = 2C)
    C = int(case[1:])
    x = rnd(64, 256, 256, C); w = torch.randn(3, 3, C, 32, device="cuda") * 0.05; b = torch.zeros(32, device="cuda")
    f = lambda: ops.conv2d(x, w, b, act=ops.ACT_ELU)
    flops = 2 * 64 * 256 * 256 * 9 * C * 32; bytes_ = (x.numel() + 64 * 256 * 256 * 32) * 2
elif case in ("p32", "p64"):                  # 1x1 C->32 at 256x256
    C = int(case[1:])
    x = rnd(64, 256, 256, C); w = torch.randn(1, 1, C, 32, device="cuda") * 0.05; b = torch.zeros(32, device="cuda")
    f = lambda: ops.conv2d(x, w, b, act=ops.ACT_ELU)
    flops = 2 * 64 * 256 * 256 * C * 32; bytes_ = (x.numel() + 64 * 256 * 256 * 32) * 2
elif case == "copy":                          # plain device copy of the stem's bytes, for scale
    x = rnd(64, 256, 256, 32); y = torch.empty_like(x)
    f = lambda: y.copy_(x)
    flops = 0; bytes_ = x.numel() * 4
elif case == "cc2":         # concats_2 stage 2: 3x3 64->128 + residual at 64x64
    x = rnd(64, 64, 64, 64); w = torch.randn(3, 3, 64, 128, device="cuda") * 0.05; b = torch.zeros(128, device="cuda"); r = rnd(64, 64, 64, 128)
    f = lambda: ops.conv2d(x, w, b, residual=r)
    flops = 2 * 64 * 64 * 64 * 9 * 64 * 128; bytes_ = (x.numel() + 2 * r.numel()) * 2
elif case == "up3":         # upsample_3: convT 640->256 at 32x32
    x1 = rnd(64, 32, 32, 512); x2 = rnd(64, 32, 32, 128); w = torch.randn(4, 4, 256, 640, device="cuda") * 0.02; b = torch.zeros(256, device="cuda")
    f = lambda: ops.conv2d_transpose_s2(x1, w, b, act=ops._lib.ACT_RELU, x2=x2)
    flops = 2 * 64 * 32 * 32 * 16 * 640 * 256; bytes_ = (x1.numel() + x2.numel() + 64 * 64 * 64 * 256) * 2
elif case == "up4":         # upsample_4: convT 320->128 at 64x64 (the slowest forward launch of the step)
    x1 = rnd(64, 64, 64, 256); x2 = rnd(64, 64, 64, 64); w = torch.randn(4, 4, 128, 320, device="cuda") * 0.02; b = torch.zeros(128, device="cuda")
    f = lambda: ops.conv2d_transpose_s2(x1, w, b, act=ops._lib.ACT_RELU, x2=x2)
    flops = 2 * 64 * 64 * 64 * 16 * 320 * 128; bytes_ = (x1.numel() + x2.numel() + 64 * 128 * 128 * 128) * 2
elif case == "up2":         # upsample_2: convT 768->512 at 16x16
    x1 = rnd(64, 16, 16, 512); x2 = rnd(64, 16, 16, 256); w = torch.randn(4, 4, 512, 768, device="cuda") * 0.02; b = torch.zeros(512, device="cuda")
    f = lambda: ops.conv2d_transpose_s2(x1, w, b, act=ops._lib.ACT_RELU, x2=x2)
    flops = 2 * 64 * 16 * 16 * 16 * 768 * 512; bytes_ = (x1.numel() + x2.numel() + 64 * 32 * 32 * 512) * 2
elif case == "wstem":       # wgrad of conv2_1_2
    x = rnd(64, 256, 256, 32); dz = rnd(64, 256, 256, 32); w = torch.zeros(3, 3, 32, 32, device="cuda")
    f = lambda: ops.conv2d_grads(x, w, dz, need_dx=False)
    flops = 2 * 64 * 256 * 256 * 9 * 32 * 32; bytes_ = x.numel() * 2 * 2
elif case == "wup4":        # wgrad of upsample_4
    x1 = rnd(64, 64, 64, 256); x2 = rnd(64, 64, 64, 64); dz = rnd(64, 128, 128, 128); w = torch.zeros(4, 4, 128, 320, device="cuda")
    f = lambda: ops.conv2d_transpose_s2_grads(x1, w, dz, x2=x2, need_dx=False)
    flops = 2 * 64 * 64 * 64 * 16 * 320 * 128; bytes_ = (x1.numel() + x2.numel() + dz.numel()) * 2
elif case == "wup2":        # wgrad of upsample_2
    x1 = rnd(64, 16, 16, 512); x2 = rnd(64, 16, 16, 256); dz = rnd(64, 32, 32, 512); w = torch.zeros(4, 4, 512, 768, device="cuda")
    f = lambda: ops.conv2d_transpose_s2_grads(x1, w, dz, x2=x2, need_dx=False)
    flops = 2 * 64 * 16 * 16 * 16 * 768 * 512; bytes_ = (x1.numel() + x2.numel() + dz.numel()) * 2
elif case == "wcc3":        # wgrad of concats_2 stage 3: 3x3 128->256 at 32x32
    x = rnd(64, 32, 32, 128); dz = rnd(64, 32, 32, 256); w = torch.zeros(3, 3, 128, 256, device="cuda")
    f = lambda: ops.conv2d_grads(x, w, dz, need_dx=False)
    flops = 2 * 64 * 32 * 32 * 9 * 128 * 256; bytes_ = (x.numel() + dz.numel()) * 2
elif case == "wup3":        # wgrad of upsample_3
    x1 = rnd(64, 32, 32, 512); x2 = rnd(64, 32, 32, 128); dz = rnd(64, 64, 64, 256); w = torch.zeros(4, 4, 256, 640, device="cuda")
    f = lambda: ops.conv2d_transpose_s2_grads(x1, w, dz, x2=x2, need_dx=False)
    flops = 2 * 64 * 32 * 32 * 16 * 640 * 256; bytes_ = (x1.numel() + x2.numel() + dz.numel()) * 2
elif case in ("dstem", "dc2"):   # dgrad alone: 3x3 C->C with act' fused (stem: 32 ch at 256^2, ELU'; dc2: 128 ch at 64^2 + accumulate)
    import ctypes as C_
    from ultrasound_modeling_b200 import _lib
    L = _lib.lib()
    ch, hw = (32, 256) if case == "dstem" else (128, 64)
    dz = rnd(64, hw, hw, ch); ref = rnd(64, hw, hw, ch); w = torch.randn(3, 3, ch, ch, device="cuda") * 0.05
    wb = ops.pack_conv(w, 1, 1, BF, None); dx = torch.zeros_like(dz)
    e = ops._epi(dx, dact=ops.ACT_ELU, dact_ref=ref, residual=dx if case == "dc2" else None)
    f = lambda: _lib.check(L.tbi_conv2d_dgrad(ops._dt(dz), 0, 64, hw, hw, 3, 1, 1, ops._vp(ops.view(dz)), ch, ops._p(wb), C_.byref(e), ops._st()), "dgrad")
    flops = 2 * 64 * hw * hw * 9 * ch * ch; bytes_ = dz.numel() * 2 * (3 if case == "dstem" else 4)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{case}: {ms*1e3:.1f} us/call (incl. pack/fold helpers)  {flops/ms/1e9:.1f} TFLOP/s  {bytes_/ms/1e6:.1f} GB/s algorithmic")
