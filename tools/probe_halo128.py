"""GPU probe: the layer2 3x3 conv (128 -> 128 channels, 1024 x 14 x 14) on the streamed-weight halo kernel vs the
im2col-TMA kernel + its stand-alone BN1 pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_classif_b200 import ops

dev = "cuda"
torch.manual_seed(0)
for (N, H, W) in [(1024, 14, 14), (32, 28, 28), (1024, 28, 28)]:
    C = 128
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = (torch.randn(C, 3, 3, C, device=dev) / (9 * C) ** 0.5).bfloat16()
    a = (torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.3)
    s = torch.zeros(2, C, device=dev)

    def t(fn, reps=30):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    fl = 2.0 * N * H * W * C * C * 9
    if ops.conv3x3_halo_supported(x, w, 1, 1):
        us = t(lambda: ops.conv3x3_halo_bn(x, w, a=a, stats=(s[0], s[1])))
        print(f"[{N}x{H}x{W}x{C}] halo stream (BN1+ReLU folded): {us:7.1f} us  {fl / us / 1e6:7.1f} TF/s")
        us = t(lambda: ops.conv3x3_halo_bn(x, w, stats=(s[0], s[1])))
        print(f"[{N}x{H}x{W}x{C}] halo stream (no transform):     {us:7.1f} us  {fl / us / 1e6:7.1f} TF/s")
    xc = x.clone()
    us1 = t(lambda: ops.scale_shift_apply(xc, a[0], a[1], relu=True))
    us2 = t(lambda: ops.conv2d_bn_nhwc(x, w, 1, 1, stats=(s[0], s[1])))
    print(f"[{N}x{H}x{W}x{C}] im2col-TMA conv {us2:7.1f} us ({fl / us2 / 1e6:7.1f} TF/s) + BN1 pass {us1:6.1f} us = {us1 + us2:7.1f} us")
