"""How far do the gradients of Variant B move when the activations are merely STORED in bf16 (fp64 math, straight-through
rounding at the points where the product stores a tensor)?  CPU only, oracle only: a property of the network, not of the kernels."""
import sys, torch
sys.path.insert(0, '.')
from oracle import resnest_decoder_oracle as B
H, W, n = 64, 32, 2
grid = (H // 16, W // 16)
pe = B.init_params(B.encoder_param_shapes(10, 3, 3, 3), seed=2236, dtype=torch.float64)
pd = B.init_params(B.decoder_param_shapes(3, grid=grid), seed=2237, dtype=torch.float64)
x = B.synthetic_input(n, H, W, 10, dtype=torch.float64); tok = B.synthetic_tokens(n, grid[0] * grid[1], 512, dtype=torch.float64)
trainable = lambda k: not k.endswith(("moving_mean", "moving_variance"))
def run():
    oe = B.ResNestEncoderOracle(10, 3, 3, 3, pe); od = B.DecoderCupOracle(3, pd, grid=grid)
    oe.p = {k: v.clone().requires_grad_(trainable(k)) for k, v in oe.p.items()}
    od.p = {k: v.clone().requires_grad_(trainable(k)) for k, v in od.p.items()}
    xr = x.clone().requires_grad_(True); tokr = tok.clone().requires_grad_(True)
    x4w, fw = oe(xr); zw = od(tokr, fw, logits=True)
    gen = torch.Generator().manual_seed(97)
    gz = torch.randn(zw.shape, generator=gen, dtype=torch.float64) / zw.numel() ** 0.5
    g4 = torch.randn(x4w.shape, generator=gen, dtype=torch.float64) / x4w.numel() ** 0.5
    ((zw * gz).sum() + (x4w * g4).sum()).backward()
    g = {"dx": xr.grad, "dhidden": tokr.grad}
    g.update({"enc/" + k: v.grad for k, v in oe.p.items() if trainable(k)}); g.update({"dec/" + k: v.grad for k, v in od.p.items() if trainable(k)})
    return g, x4w.detach(), zw.detach()
ref, x4r, zr = run()
class _Q(torch.autograd.Function):            # a tensor stored in bf16 and its gradient stored in bf16
    @staticmethod
    def forward(ctx, t): return t.to(torch.bfloat16).to(t.dtype)
    @staticmethod
    def backward(ctx, g): return g.to(torch.bfloat16).to(g.dtype)
qz = _Q.apply
orig = {k: getattr(B, k) for k in ("conv2d_same", "leaky", "avgpool2", "conv2d_transpose_s2_same")}
for k, f in orig.items():
    setattr(B, k, (lambda f: lambda *a, **kw: qz(f(*a, **kw)))(f))
got, x4q, zq = run()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
print("forward: x_4 rel err %.2e, logits rel err %.2e" % (rel(x4q, x4r), rel(zq, zr)))
errs = sorted(((rel(got[k], ref[k]), k) for k in ref), reverse=True)
for e, k in errs[:25]: print("%.3f %s" % (e, k))
import statistics
print("median %.4f  n>2e-2: %d of %d" % (statistics.median(e for e, _ in errs), sum(e > 2e-2 for e, _ in errs), len(errs)))
keys = [k for k in ref if k not in ("dx", "dhidden")]
a = torch.cat([got[k].flatten() / ref[k].abs().max() for k in keys]); b = torch.cat([ref[k].flatten() / ref[k].abs().max() for k in keys])
print("global cosine (per-tensor normalised) %.4f" % float((a * b).sum() / a.norm() / b.norm()))
