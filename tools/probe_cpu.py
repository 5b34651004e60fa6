"""Host-side cost of one train step: wall time of enqueueing N steps (no sync) vs the GPU time of the same steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc
from video_classif_b200.ingest import ingest_batch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = vc.LRCN(4, 16, 32, 8, cnn_backbone="resnet50", rnn_layers=3, dropout=0.25, precision="bf16").to(dev).train()
opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, fused=True)
x = torch.rand(64, 16, 3, 112, 112, device=dev)
y = torch.randint(0, 4, (64,), device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(model(x), y)
    loss.backward()
    opt.step()
for _ in range(5):
    step()
torch.cuda.synchronize()
for n in (1, 20):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        step()
    t1 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"n={n}: host enqueue {1e3*(t1-t0)/n:.2f} ms/step, gpu {e0.elapsed_time(e1)/n:.2f} ms/step, wall {1e3*(t2-t0)/n:.2f} ms/step")
# split: backbone only (no grad) vs whole
with torch.no_grad():
    t0 = time.perf_counter()
    for _ in range(20):
        f = model._features(x)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"backbone only: host enqueue {1e3*(t1-t0)/20:.2f} ms, wall {1e3*(t2-t0)/20:.2f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
