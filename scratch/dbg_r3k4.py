"""r3k4 bf16: tcgen05 engine vs CUDA-core engine (same bf16 storage), buffer by buffer"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import tbi_resnest_oracle as O
from ultrasound_modeling_b200.TBI_ResNest import ResNest
from ultrasound_modeling_b200 import _lib

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
R, K = 3, 4
o = O.TBIResNestOracle(hw, hw, 1, 3, 3, R, K, dtype=torch.float64)
x, y = O.synthetic_batch(2, hw, hw)
masks = O.dropout_masks(2, hw, hw)
nets = {}
for name, impl in (("tc", _lib.IMPL_AUTO), ("simt", _lib.IMPL_SIMT)):
    net = ResNest(hw, hw, 1, 3, 3, radix=R, kpaths=K, dtype="bf16", use_cuda_graph=False, impl=impl)
    net.load_state_dict(o.state_dict())
    net.step(x, y, train=False, dropout_masks=masks)
    net.engine.backward()
    torch.cuda.synchronize()
    nets[name] = net
a, b = nets["tc"].engine, nets["simt"].engine
def rel(p, q):
    p = p.double(); q = q.double()
    return float((p - q).abs().max() / q.abs().max().clamp_min(1e-30))
for si in range(5):
    for k in ("T1", "U", "V", "Y", "dY", "dV", "dZ2", "dZ1"):
        ta, tb = a.stage_buf[si][k], b.stage_buf[si][k]
        line = f"stage {si} {k:4s} {tuple(ta.shape)} rel {rel(ta, tb):.3e}"
        if k in ("T1", "dZ1"):
            creal = a.stage_info[si]["G"] * a.stage_info[si]["cv11"]
            per = [(rel(ta[..., c], tb[..., c])) for c in range(min(6, ta.shape[-1]))]
            line += f" | first channels {['%.1e' % v for v in per]} | pad max tc {float(ta[..., creal:].abs().max()) if creal < ta.shape[-1] else 0:.2e} simt {float(tb[..., creal:].abs().max()) if creal < tb.shape[-1] else 0:.2e}"
        print(line)
ga, gb = a.grad_dict(), b.grad_dict()
errs = sorted(((rel(ga[k], gb[k]), k) for k in ga), reverse=True)[:12]
for e_, k in errs:
    print(f"grad {k}: {e_:.3e}")
want = o.gradients(x.double(), y.double(), masks)
for nm, g in (("tc", ga), ("simt", gb)):
    errs = sorted(((rel(g[k].cpu(), want[k]), k) for k in want), reverse=True)[:6]
    print(nm, "vs oracle:", [(round(v, 4), k) for v, k in errs])
