"""Weight-gradient kernel (conv_wgrad_kernel) at the ResNet-50 @112 layer shapes of a 1024-frame batch: CUDA-event time, TFLOP/s,
operand bytes.  `python tools/probe_wgrad.py ncu` runs each shape once (for an `ncu --set full -k regex:conv_wgrad` capture)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_classif_b200 import backbone_train as BT

dev = "cuda"
ncu = len(sys.argv) > 1 and sys.argv[1] == "ncu"
N = 1024
shapes = [("l1.conv2 3x3 64->64", 28, 64, 64, 3, 1), ("l1.conv3 1x1 64->256", 28, 64, 256, 1, 1), ("l2.conv2 3x3 128->128", 14, 128, 128, 3, 1),
          ("l2.conv3 1x1 128->512", 14, 128, 512, 1, 1), ("l3.conv1 1x1 1024->256", 7, 1024, 256, 1, 1), ("l3.conv2 3x3 256->256", 7, 256, 256, 3, 1),
          ("l3.0.conv2 3x3/2 256->256", 14, 256, 256, 3, 2), ("l4.conv2 3x3 512->512", 4, 512, 512, 3, 1), ("l4.conv3 1x1 512->2048", 4, 512, 2048, 1, 1)]
torch.manual_seed(0)
for name, hw, C, Cout, R, s in shapes:
    p = (R - 1) // 2
    x = torch.randn(N, hw, hw, C, device=dev).to(torch.bfloat16)
    P = (hw + 2 * p - R) // s + 1
    dy = torch.randn(N, P, P, Cout, device=dev).to(torch.bfloat16)
    if ncu:
        BT.conv_wgrad(x, dy, R, R, s, p)
        continue
    for _ in range(3):
        BT.conv_wgrad(x, dy, R, R, s, p)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        BT.conv_wgrad(x, dy, R, R, s, p)          # includes the zero-fill of dW
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    flop = 2.0 * N * P * P * Cout * C * R * R
    mb = (x.numel() + dy.numel()) * 2 / 1e6
    print(f"{name:28s} M={N * P * P:7d}  {us:7.1f} us  {flop / us / 1e6:7.1f} TFLOP/s  operands {mb:6.1f} MB -> {mb / us * 1e3:6.0f} GB/s if read once")
torch.cuda.synchronize()
