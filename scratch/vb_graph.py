import sys, time, torch
sys.path.insert(0, '.')
from oracle import vit_oracle as V
from ultrasound_modeling_b200.VisionTransformer import VisionTransformer
dev = torch.device('cuda:0')
def timed(f, reps, warm=4):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n, graph in ((16, True), (64, True), (16, False)):
    net = VisionTransformer(n, dtype="bf16", device=str(dev), use_cuda_graph=graph)
    x = V.B.synthetic_input(2, 256, 80, 10).repeat(n // 2, 1, 1, 1).to(dev); y = V.synthetic_labels(2, 256, 80).repeat(n // 2, 1, 1, 1).to(dev)
    ms = timed(lambda: net.train_step(x, y), 5)
    msf = timed(lambda: net.forward(x), 5)
    print(n, graph, 'train ms', round(ms, 3), 'fwd ms', round(msf, 3), flush=True)
    del net; torch.cuda.empty_cache()
