"""GPU probe: the bench's pipelined train step with the trainable tail (forward, CE, backward, Adam) launched kernel by kernel
vs replayed from ONE CUDA graph (the frozen encoder pass of the next batch runs on the side stream in both)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc
from video_classif_b200.models import _FeatureHandle

dev = torch.device("cuda", 0)
B, T, S = 64, 16, 112
torch.manual_seed(0)
model = vc.LRCN(4, T, 32, 8, cnn_backbone="resnet50", rnn_layers=3, dropout=0.25, precision="bf16").to(dev).train()
model.enable_encoder_graph()
params = [p for p in model.parameters() if p.requires_grad]
xs = [torch.rand(B, T, 3, S, S, device=dev) for _ in range(4)]
ys = [torch.randint(0, 4, (B,), device=dev) for _ in range(4)]


def timed(run, n=40, warm=6):
    run(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); run(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


opt = torch.optim.Adam(params, lr=1e-4, fused=True)


def eager(n):
    h = model.encode_async(xs[0])
    for i in range(n):
        hn = model.encode_async(xs[(i + 1) % 4]) if i + 1 < n else None
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(xs[i % 4], features=h), ys[i % 4])
        loss.backward(); opt.step()
        h = hn


print(f"eager tail, pipelined:   {timed(eager):.3f} ms/step")

opt2 = torch.optim.Adam(params, lr=1e-4, fused=True, capturable=True)
feat0 = model.encode_async(xs[0])
torch.cuda.current_stream().wait_event(feat0.event)
static_feat = feat0.tensor.clone()
static_y = ys[0].clone()
handle = _FeatureHandle(static_feat, None, tuple(xs[0].shape))


def tail():
    loss = torch.nn.functional.cross_entropy(model(xs[0], features=handle), static_y)
    loss.backward(); opt2.step()
    return loss


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        opt2.zero_grad(set_to_none=True); tail()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
opt2.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    static_loss = tail()


def graphed(n):
    h = model.encode_async(xs[0])
    cur = torch.cuda.current_stream()
    for i in range(n):
        hn = model.encode_async(xs[(i + 1) % 4]) if i + 1 < n else None
        cur.wait_event(h.event)
        static_feat.copy_(h.tensor); static_y.copy_(ys[i % 4])
        g.replay()
        h = hn


print(f"graphed tail, pipelined: {timed(graphed):.3f} ms/step   (loss {static_loss.item():.4f})")


def tail_only(n):
    for i in range(n):
        g.replay()


print(f"graphed tail alone:      {timed(tail_only):.3f} ms")


def eager_tail_only(n):
    for i in range(n):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(xs[0], features=handle), static_y)
        loss.backward(); opt.step()


print(f"eager tail alone:        {timed(eager_tail_only):.3f} ms")
