import sys, torch
sys.path.insert(0, '.')
from oracle import tbi_resnest_oracle as O
from ultrasound_modeling_b200.TBI_ResNest import ResNest
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
R,K = 2,1
o = O.TBIResNestOracle(256,256,1,3,3,R,K,dtype=torch.float64)
net = ResNest(256,256,1,3,3,radix=R,kpaths=K,dtype="fp32",use_cuda_graph=False); net.load_state_dict(o.state_dict())
x,y = O.synthetic_batch(2,256,256); m = O.dropout_masks(2,256,256)
net.step(x,y,train=False,dropout_masks=m); net.engine.backward()
probs, inter = o.forward(x.double(), m, return_intermediates=True)
loss = o.my_loss_cat(y.double(), probs).sum()
names = ['f_tran'] + [f'upsample_{i}' for i in range(5)] + [f'pool_{i}' for i in range(1,6)]
gs = dict(zip(names, torch.autograd.grad(loss, [inter[k] for k in names])))
e = net.engine
print("dlogits %.2e  (max %.2e)" % (rel(e.dlogits, gs['f_tran']), float(gs['f_tran'].abs().max())))
for i in range(4,-1,-1):
    up = inter[f'upsample_{i}']
    want = gs[f'upsample_{i}'] * (up > 0)
    if i < 3: want = want * (m[i].double()*2)
    print(f"up{i}: fwd %.2e  dz %.2e (max %.2e)" % (rel(e.up[i], up), rel(e.dup[i], want), float(want.abs().max())))
for i in range(1,6):
    print(f"dpool[{i-1}] %.2e (max %.2e)" % (rel(e.dpool[i-1], gs[f'pool_{i}']), float(gs[f'pool_{i}'].abs().max())))
# isolate: recompute head dgrad alone with ops from the ORACLE's exact dlogits
from ultrasound_modeling_b200 import ops
dl = gs['f_tran'].float().cuda().contiguous()
(dx1, dx2), dw, db = ops.conv2d_transpose_s2_grads(e.up[4], o.params['f_tran/kernel'].detach().float().cuda(), dl, x2=e.pool[0])
xc = torch.cat([inter['upsample_4'], inter['pool_1']], 3).detach().requires_grad_(True)
yy = O.conv2d_transpose_s2_same(xc, o.params['f_tran/kernel'].detach(), None)
gx, = torch.autograd.grad((yy*gs['f_tran']).sum(), [xc])
print("head dgrad alone: dx1 %.2e dx2 %.2e  max %.2e" % (rel(dx1, gx[...,:128]), rel(dx2, gx[...,128:]), float(gx.abs().max())))
err = (dx1.double().cpu()-gx[...,:128]).abs()
idx = err.argmax(); print("worst idx", idx.item(), "got", dx1.flatten()[idx].item(), "want", gx[...,:128].flatten()[idx].item())
