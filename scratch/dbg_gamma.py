import sys, torch
sys.path.insert(0, '.')
from oracle import tbi_resnest_oracle as O
from ultrasound_modeling_b200.TBI_ResNest import ResNest
def rel(a,b):
    a=a.double().cpu(); b=b.double().cpu(); return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
o = O.TBIResNestOracle(256,256,1,3,3,4,4,dtype=torch.float64)
net = ResNest(256,256,1,3,3,radix=4,kpaths=4,dtype="fp32",use_cuda_graph=False); net.load_state_dict(o.state_dict())
x,y = O.synthetic_batch(2,256,256); m = O.dropout_masks(2,256,256)
net.step(x,y,train=False,dropout_masks=m); net.engine.backward()
got = net.engine.grad_dict(); want = o.gradients(x.double(), y.double(), m)
errs = sorted(((rel(got[k],want[k]),k, float(want[k].abs().max())) for k in want), reverse=True)
for e in errs[:25]: print("%.3e %-40s max|g|=%.3e" % e)
k = errs[0][1]; print(k, got[k].cpu()[:8], want[k][:8])
