import sys, torch
sys.path.insert(0, '.')
from ultrasound_modeling_b200 import ops
def t(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (hw, c) in [(256, 32), (128, 128), (128, 64), (64, 256), (32, 512)]:
    x = torch.randn(64, hw, hw, c, device="cuda").to(torch.bfloat16)
    y = torch.empty_like(x)
    a = t(lambda: (flush.zero_(), ops.colsum(x))) - t(lambda: flush.zero_())
    b = t(lambda: (flush.zero_(), x.view(-1, c).sum(0, dtype=torch.float32))) - t(lambda: flush.zero_())
    cp = t(lambda: (flush.zero_(), y.copy_(x))) - t(lambda: flush.zero_())
    mb = x.numel() * 2 / 1e6
    print(f"[{64*hw*hw} x {c}] {mb:.0f} MB: tbi_colsum {a:.1f} us ({mb/a*1e3/1e3:.2f} TB/s) | torch.sum {b:.1f} us | copy {cp:.1f} us ({2*mb/cp*1e3/1e3:.2f} TB/s r+w)")
