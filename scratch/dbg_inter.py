import sys, torch
sys.path.insert(0, '.')
import torch.nn.functional as F
from oracle import tbi_resnest_oracle as O
from ultrasound_modeling_b200.TBI_ResNest import ResNest
def rel(a,b):
    a=a.double().cpu(); b=b.double().cpu(); return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
R,K = 4,4
o = O.TBIResNestOracle(256,256,1,3,3,R,K,dtype=torch.float64)
net = ResNest(256,256,1,3,3,radix=R,kpaths=K,dtype="fp32",use_cuda_graph=False); net.load_state_dict(o.state_dict())
x,y = O.synthetic_batch(2,256,256); m = O.dropout_masks(2,256,256)
net.step(x,y,train=False,dropout_masks=m); net.engine.backward()
probs, inter = o.forward(x.double(), m, return_intermediates=True)
loss = o.my_loss_cat(y.double(), probs).sum()
names = [k for k in inter if k.endswith('/V') or '/U_r' in k or '/T1_r' in k or k.startswith('conv') and '/' not in k or k.startswith('pool_')]
gs = dict(zip(names, torch.autograd.grad(loss, [inter[k] for k in names], allow_unused=True)))
e = net.engine
for si,(stage,out) in enumerate(O.STAGES):
    b = e.stage_buf[si]; info = e.stage_info[si]; cv11, cvkk = info['cv11'], info['cvkk']
    print(stage, "fwd Y %.2e V %.2e" % (rel(b['Y'], inter[stage]), rel(b['V'], inter[f'{stage}/V'])),
          "| dY %.2e dV %.2e dPin %.2e" % (rel(b['dY'], gs[stage]), rel(b['dV'], gs[f'{stage}/V']), rel(e.dpool[si], gs[f'pool_{si+1}'])))
    # dZ2 = dU * ELU'(U) ; dZ1 = dT1*ELU'(T1), per (k,r)
    w2=w1=0
    for k in range(K):
        for r in range(R):
            g_ = k*R+r
            U = inter[f'{stage}_car_k{k}/U_r{r}']; dU = gs[f'{stage}_car_k{k}/U_r{r}']
            want = dU*torch.where(U>0, torch.ones_like(U), U+1)
            w2 = max(w2, rel(b['dZ2'][..., g_*cvkk:(g_+1)*cvkk], want))
            T = inter[f'{stage}_car_k{k}/T1_r{r}']; dT = gs[f'{stage}_car_k{k}/T1_r{r}']
            want = dT*torch.where(T>0, torch.ones_like(T), T+1)
            w1 = max(w1, rel(b['dZ1'][..., g_*cv11:(g_+1)*cv11], want))
    print("      worst dZ2 %.2e  dZ1 %.2e   |dV|max %.2e |dZ2|max %.2e" % (w2, w1, float(gs[f'{stage}/V'].abs().max()), float(b['dZ2'].abs().max())))
