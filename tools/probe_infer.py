"""GPU probe: single-clip serving latency (deployment.py path), eager module vs CUDA-graph replay."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for (T, S) in ((16, 112), (16, 224)):
    m = vc.LRCN(4, T, 32, 8, cnn_backbone="resnet50", rnn_layers=3).to(dev).eval()
    x = torch.rand(1, T, 3, S, S, device=dev)
    infer = vc.GraphedInference(m, x)
    def timeit(fn, reps=30):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3
    with torch.no_grad():
        eager = timeit(lambda: m(x))
    graphed = timeit(lambda: infer(x))
    print(f"B=1 T={T} {S}x{S} resnet50 LRCN eval: eager {eager:.2f} ms/clip, graph replay {graphed:.2f} ms/clip ({infer.n_launch} kernels per forward)")
