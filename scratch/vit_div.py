import sys, torch
sys.path.insert(0, '.')
from oracle import vit_oracle as V
from ultrasound_modeling_b200.VisionTransformer import VisionTransformer
o = V.VisionTransformerOracle(2, img_size=(64, 32), num_classes=3, learning_rate=1e-3, dtype=torch.float64, num_layers=2)
for trial in range(5):
    nets = [VisionTransformer(2, img_size=(64, 32), num_classes=3, learning_rate=1e-3, dtype="fp32", device="cuda:0", num_layers=2, use_cuda_graph=False) for g in range(2)]
    for net in nets: net.load_variables(o.state_dict())
    x2 = V.B.synthetic_input(2, 64, 32, 10); y2 = V.synthetic_labels(2, 64, 32)
    x4 = V.B.synthetic_input(4, 64, 32, 10, seed=77); y4 = V.synthetic_labels(4, 64, 32, seed=78)
    for step, (x, y) in enumerate([(x2, y2)] * 4 + [(x4, y4)] * 2):
        for i, net in enumerate(nets):
            loss, probs = net.train_step(x, y)
        ga, gb = nets[0].gradients(), nets[1].gradients()
        va, vb = nets[0].variables(), nets[1].variables()
        worst = sorted(((float((ga[k] - gb[k]).norm() / (ga[k].norm() + 1e-30)), k) for k in ga if 'key/bias' not in k), reverse=True)[:3]
        wv = sorted(((float((va[k] - vb[k]).abs().max()), k) for k in va), reverse=True)[:3]
        if step >= 3: print(trial, step, "G", ["%s %.1e" % (k[-40:], v) for v, k in worst], "V", ["%s %.1e" % (k[-40:], v) for v, k in wv])
