"""GPU probe: the layer1 3x3 conv (64 -> 64 channels, 1024 x 28 x 28) on the resident-weight halo kernel, with and without
the BN1 + ReLU fold (ncu captures: `ncu --set full --import-source on -k regex:conv3x3_halo_kernel ...`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_classif_b200 import ops

dev = "cuda"
torch.manual_seed(0)
N, H, W, C = 1024, 28, 28, 64
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
for name, fn in (("BN1+ReLU folded, stats", lambda: ops.conv3x3_halo_bn(x, w, a=a, stats=(s[0], s[1]))),
                 ("no transform, stats", lambda: ops.conv3x3_halo_bn(x, w, stats=(s[0], s[1]))),
                 ("no transform, no stats", lambda: ops.conv3x3_halo_bn(x, w))):
    us = t(fn)
    print(f"[{N}x{H}x{W}x{C}] halo ({name}): {us:7.1f} us  {fl / us / 1e6:7.1f} TF/s  {(2 * N * H * W * C * 2) / us / 1e3:7.1f} GB/s")
