"""Localises fp32 gradient error of the small-CNN path at the cfg-1 shapes: every ConvBnReluPoolFn node (forward,
dx, dw, dgamma, dbeta) and the raw conv kernels against torch fp64 on the same GPU.
    python tools/probe_cfg1_parity.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_classif_b200 as vc  # noqa: E402
from video_classif_b200 import ops  # noqa: E402
from video_classif_b200._lib import call, stream_ptr  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def node(N, Cin, Cout, S, pool, positive):
    conv = torch.nn.Conv2d(Cin, Cout, 3, padding=1).to(dev)
    bn = torch.nn.BatchNorm2d(Cout).to(dev)
    x = torch.randn(N, Cin, S, S, device=dev)
    if positive:
        x = torch.relu(x + 0.5)
    x.requires_grad_(True)
    y = ops.conv_bn_relu_pool(x, conv, bn, pool, True)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    got = dict(y=y.detach(), dx=x.grad.clone(), dw=conv.weight.grad.clone(), dg=bn.weight.grad.clone(), db=bn.bias.grad.clone())
    c64, b64 = torch.nn.Conv2d(Cin, Cout, 3, padding=1).to(dev).double(), torch.nn.BatchNorm2d(Cout).to(dev).double()
    c64.load_state_dict({k: v.double() for k, v in conv.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    z = torch.relu(b64(c64(x64)))
    if pool:
        z = torch.nn.functional.max_pool2d(z, 2, 2)
    (z * w.double()).sum().backward()
    ref = dict(y=z.detach(), dx=x64.grad, dw=c64.weight.grad, dg=b64.weight.grad, db=b64.bias.grad)
    # raw weight-gradient kernel on the fp64 dz (isolates the kernel from the BN backward)
    zc = c64(x64.detach())
    zc.retain_grad()
    zz = torch.relu(b64(zc))
    if pool:
        zz = torch.nn.functional.max_pool2d(zz, 2, 2)
    c64.weight.grad = None
    (zz * w.double()).sum().backward()
    dz = zc.grad.float().contiguous()
    dw = torch.zeros_like(conv.weight)
    xf = x.detach().contiguous()
    call("b2_conv3x3_wgrad_f32", xf.data_ptr(), dz.data_ptr(), dw.data_ptr(), N, Cin, Cout, S, S, stream_ptr())
    torch.cuda.synchronize()
    print(f"N={N} Cin={Cin} Cout={Cout} S={S} pool={pool}: " + "  ".join(f"{k} {rel(got[k], ref[k]):.2e}" for k in got)
          + f"  | wgrad kernel alone {rel(dw, c64.weight.grad):.2e}")


for N in (24, 160):
    node(N, 3, 16, 64, False, True)
    node(N, 16, 32, 64, True, True)
    node(N, 32, 64, 32, True, True)
