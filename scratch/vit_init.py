import sys, torch
sys.path.insert(0, '.')
from oracle import vit_oracle as V
from ultrasound_modeling_b200.VisionTransformer import VisionTransformer
o = V.VisionTransformerOracle(2, img_size=(64, 32), num_classes=3, learning_rate=1e-3, dtype=torch.float64, num_layers=2)
net = VisionTransformer(2, img_size=(64, 32), num_classes=3, learning_rate=1e-3, dtype="fp32", device="cuda:0", num_layers=2, use_cuda_graph=False)
net.load_variables(o.state_dict())
x2 = V.B.synthetic_input(2, 64, 32, 10); y2 = V.synthetic_labels(2, 64, 32)
x4 = V.B.synthetic_input(4, 64, 32, 10, seed=77); y4 = V.synthetic_labels(4, 64, 32, seed=78)
for (x, y, tag) in [(x2, y2, "x2"), (x4, y4, "x4")]:
    ref = None
    for it in range(12):
        junk = torch.randn(1 << 20, device="cuda") * (it + 1)    # perturb allocator contents
        loss, probs = net.backward(x, y)
        g = {k: v.clone() for k, v in net.gradients().items()}
        if ref is None: ref = g; continue
        worst = sorted(((float((g[k] - ref[k]).norm() / (ref[k].norm() + 1e-30)), k) for k in g), reverse=True)[:4]
        print(tag, it, ["%s %.1e" % (k, v) for v, k in worst])
        del junk
