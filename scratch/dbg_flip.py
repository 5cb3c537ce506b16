import sys, torch
sys.path.insert(0, '.')
from oracle import tbi_resnest_oracle as O
from ultrasound_modeling_b200.TBI_ResNest import ResNest
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
R,K = 4,4
o = O.TBIResNestOracle(256,256,1,3,3,R,K,dtype=torch.float64)
net = ResNest(256,256,1,3,3,radix=R,kpaths=K,dtype="fp32",use_cuda_graph=False); net.load_state_dict(o.state_dict())
x,y = O.synthetic_batch(2,256,256); m = O.dropout_masks(2,256,256)
net.step(x,y,train=False,dropout_masks=m)
probs, inter = o.forward(x.double(), m, return_intermediates=True)
e = net.engine
for i in range(5):
    a = (e.up[i].cpu() > 0); b = (inter[f'upsample_{i}'] > 0)
    print(f"up{i}: sign flips {(a != b).sum().item()} of {a.numel()}")
want = o.gradients(x.double(), y.double(), m)
e.backward(); got = e.grad_dict()
errs = sorted(((rel(got[k],want[k]),k) for k in want), reverse=True)
print("before:", ["%.2e %s" % t for t in errs[:3]], "n>1e-4:", sum(t[0] > 1e-4 for t in errs))
for i in range(5):
    e.up[i].copy_(inter[f'upsample_{i}'].detach().float())
e.backward(); got = e.grad_dict()
errs = sorted(((rel(got[k],want[k]),k) for k in want), reverse=True)
print("after :", ["%.2e %s" % t for t in errs[:3]], "n>1e-4:", sum(t[0] > 1e-4 for t in errs))
