import sys, json, torch
sys.path.insert(0, '.')
from oracle import tbi_resnest_oracle as O
from ultrasound_modeling_b200.TBI_ResNest import ResNest
def timed(f, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# the reference driver's own configuration (TBI_ResNest.py:461): 256x64x6, radix 3, kpaths 4
net = ResNest(256, 64, 6, 3, 3, radix=3, kpaths=4, learning_rate=5e-3, dtype="bf16")
n = 64
x = (torch.randn(n, 256, 64, 6) * 0.3).cuda(); y = torch.nn.functional.one_hot(torch.randint(0, 3, (n, 256, 64)), 3).float().cuda()
ms = timed(lambda: net.step(x, y, train=True))
print(json.dumps({"config": "reference main(): 256x64x6 r3k4 N=64 training step, bf16", "ms": round(ms, 3), "img_per_s": round(n / ms * 1e3, 1)}))
