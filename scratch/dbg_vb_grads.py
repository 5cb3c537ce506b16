import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from ultrasound_modeling_b200 import ops
def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
torch.manual_seed(0)
for dtype in (torch.float32, torch.bfloat16):
    for (cin, cin2, cout, k, d, h, w) in ((7, 0, 21, 3, 1, 16, 8), (63, 0, 128, 3, 1, 16, 8), (64, 0, 7, 1, 1, 16, 8), (126, 0, 256, 3, 1, 8, 4), (14, 0, 42, 3, 1, 8, 4),
                                          (256, 256, 64, 3, 2, 8, 4), (256, 256, 64, 1, 1, 8, 4), (30, 0, 64, 3, 1, 32, 16), (10, 0, 16, 3, 1, 64, 32), (255, 0, 512, 3, 1, 4, 2)):
        n = 2
        x = torch.randn(n, h, w, cin + cin2, dtype=torch.float64).to(dtype).double()
        wt = torch.randn(k, k, cin + cin2, cout, dtype=torch.float64) * 0.1
        dz = torch.randn(n, h, w, cout, dtype=torch.float64).to(dtype).double()
        xr = x.clone().requires_grad_(True); wr = wt.clone().requires_grad_(True)
        y = F.conv2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), padding=d * (k // 2), dilation=d).permute(0, 2, 3, 1)
        (y * dz).sum().backward()
        xd = x.to(dtype).cuda()
        x1 = xd[..., :cin].contiguous(); x2 = xd[..., cin:].contiguous() if cin2 else None
        dxs, dw, db = ops.conv2d_grads(x1, wt.float().cuda(), dz.to(dtype).cuda(), dilation=d, x2=x2)
        dx = torch.cat(dxs, 3) if cin2 else dxs
        print(str(dtype)[6:], (cin, cin2, cout, k, d, h, w), "dx %.2e dw %.2e db %.2e" % (rel(dx, xr.grad), rel(dw, wr.grad), rel(db, dz.sum((0, 1, 2)))))
    for (cin, cin2, cout, h, w) in ((512, 0, 256, 4, 2), (256, 128, 128, 8, 4), (64, 8, 3, 32, 16)):
        n = 2
        x = torch.randn(n, h, w, cin + cin2, dtype=torch.float64).to(dtype).double()
        wt = torch.randn(3, 3, cout, cin + cin2, dtype=torch.float64) * 0.1
        dz = torch.randn(n, 2 * h, 2 * w, cout, dtype=torch.float64).to(dtype).double()
        xr = x.clone().requires_grad_(True); wr = wt.clone().requires_grad_(True)
        y = F.conv_transpose2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), stride=2, padding=0)[..., :2 * h, :2 * w].permute(0, 2, 3, 1)
        (y * dz).sum().backward()
        xd = x.to(dtype).cuda()
        x1 = xd[..., :cin].contiguous(); x2 = xd[..., cin:].contiguous() if cin2 else None
        dxs, dw, db = ops.conv2d_transpose_s2_grads(x1, wt.float().cuda(), dz.to(dtype).cuda(), x2=x2)
        dx = torch.cat(dxs, 3) if cin2 else dxs
        print(str(dtype)[6:], "convT", (cin, cin2, cout, h, w), "dx %.2e dw %.2e db %.2e" % (rel(dx, xr.grad), rel(dw, wr.grad), rel(db, dz.sum((0, 1, 2)))))
