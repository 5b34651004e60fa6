"""GPU probe: the trainable tail of the bench model (adapts -> LSTM stack -> head), forward + backward + Adam,
fed fixed features: CUDA-event time of the whole tail and of its LSTM stack alone (warm caches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc
from video_classif_b200 import ops

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, T = 64, 16
model = vc.LRCN(4, T, 32, 8, cnn_backbone="resnet50", rnn_layers=3, dropout=0.25, precision="bf16").to(dev).train()
opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, fused=True)
feat = torch.randn(B, T, 2048, device=dev)
object.__setattr__(model, "_features", lambda _x: feat)
x = torch.empty(B, T, 3, 8, 8, device=dev)
y = torch.randint(0, 4, (B,), device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def tail_step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(model(x), y)
    loss.backward()
    opt.step()


def tail_fwd():
    with torch.no_grad():
        model(x)


print(torch.cuda.get_device_name(0))
n0 = vc._lib.launch_count()
tail_step()
print("launches per tail step (b2_* only):", vc._lib.launch_count() - n0)
print(f"tail fwd+bwd+Adam: {timeit(tail_step):8.1f} us")
print(f"tail fwd (no grad): {timeit(tail_fwd):8.1f} us")
xin = torch.randn(B, T, 8, device=dev, requires_grad=True)
wgt = torch.randn(B, T, 32, device=dev)


def lstm_fb():
    xin.grad = None
    out = ops.lstm_forward(xin, model.rnn)
    (out * wgt).sum().backward()


def lstm_f():
    with torch.no_grad():
        ops.lstm_forward(xin, model.rnn)


print(f"lstm stack fwd+bwd (+2 torch ops): {timeit(lstm_fb):8.1f} us")
print(f"lstm stack fwd: {timeit(lstm_f):8.1f} us")
ops.LSTM_STACK = False
print(f"per-layer lstm fwd+bwd: {timeit(lstm_fb):8.1f} us")
print(f"per-layer lstm fwd: {timeit(lstm_f):8.1f} us")
