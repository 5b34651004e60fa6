"""GPU probe: timings of the BatchNorm-folded conv kernel variants at ResNet-50 @112 layer shapes
(run under gpurun).  `python tools/probe_fused.py [time|ncu]`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc
from video_classif_b200 import ops

dev = "cuda"
torch.manual_seed(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
NF = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
print(torch.cuda.get_device_name(0), "frames", NF, flush=True)


def timeit(fn, reps=5):
    if mode in ("ncu", "ncu2", "ncu3", "ncu4", "gram"):
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def bnbuf(C):
    b = torch.zeros(4 * C + 4, device=dev)
    return b, (b[:C], b[C:2 * C]), (b[2 * C:3 * C], b[3 * C:4 * C]), b[4 * C:]


def case(name, H, C, Cout, R, stride, pad, variants):
    x = torch.randn(NF, H, H, C, device=dev).bfloat16()
    w = (torch.randn(Cout, R, R, C, device=dev) / (C * R * R) ** 0.5).bfloat16()
    a = (torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1)
    o = (torch.rand(Cout, device=dev) + 0.5, torch.randn(Cout, device=dev) * 0.1)
    P = (H + 2 * pad - R) // stride + 1
    res = torch.randn(NF, P, P, Cout, device=dev).bfloat16()
    gamma, beta = torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)
    buf, st, ss, cnt = bnbuf(Cout)
    fin = (gamma, beta, None, None, ss[0], ss[1], cnt, 1e-5, 0.1)
    flops = 2.0 * NF * P * P * Cout * C * R * R

    def run(v):
        buf.zero_()
        if v == "plain":
            return ops.conv2d_bn_nhwc(x, w, stride, pad)
        if v == "stats":
            return ops.conv2d_bn_nhwc(x, w, stride, pad, stats=st, fin=fin)
        if v == "a":
            return ops.conv2d_bn_nhwc(x, w, stride, pad, a=a)
        if v == "a+stats":
            return ops.conv2d_bn_nhwc(x, w, stride, pad, a=a, stats=st, fin=fin)
        if v == "a+statsonly":
            return ops.conv2d_bn_nhwc(x, w, stride, pad, a=a, stats=st, fin=fin, store=False)
        if v == "statsonly":
            return ops.conv2d_bn_nhwc(x, w, stride, pad, stats=st, fin=fin, store=False)
        if v == "halo":
            return ops.conv3x3_halo_bn(x, w, stats=st, fin=fin)
        if v == "halo+a":
            return ops.conv3x3_halo_bn(x, w, a=a, stats=st, fin=fin)
        if v == "gram":
            return ops.conv1x1_gram_bnstats(x, w, a, fin)
        if v == "a+o":
            return ops.conv2d_bn_nhwc(x, w, stride, pad, a=a, o=o, relu=True)
        if v == "a+o+res":
            return ops.conv2d_bn_nhwc(x, w, stride, pad, a=a, o=o, res=res, relu=True)
        if v == "o+res":
            return ops.conv2d_bn_nhwc(x, w, stride, pad, o=o, res=res, relu=True)
        if v == "a+o+res+r":
            return ops.conv2d_bn_nhwc(x, w, stride, pad, a=a, o=o, res=res, r=o, relu=True)
        raise ValueError(v)

    if "apply" in variants:
        variants = [v for v in variants if v != "apply"]
        us = timeit(lambda: ops.scale_shift_apply(x, a[0], a[1], relu=True))
        print(f"{name:10s} {'apply(in)':12s} {us:8.1f} us  {x.numel() * 4 / us / 1e6:7.2f} TB/s", flush=True)
    for v in variants:
        us = timeit(lambda: run(v))
        print(f"{name:10s} {v:12s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)


if mode == "quick":
    case("l1.conv1", 28, 256, 64, 1, 1, 0, ["stats"])
    case("l1.conv2", 28, 64, 64, 3, 1, 1, ["stats"])
    case("l1.conv3", 28, 64, 256, 1, 1, 0, ["plain", "stats", "a+statsonly", "a+o+res"])
    case("l3.conv2", 7, 256, 256, 3, 1, 1, ["plain", "stats"])
    case("l4.conv2", 4, 512, 512, 3, 1, 1, ["plain", "stats"])
elif mode == "gram":
    case("l1.conv3", 28, 64, 256, 1, 1, 0, ["gram", "a+o+res"])
    case("l2.conv3", 14, 128, 512, 1, 1, 0, ["gram", "a+o+res"])
elif mode == "ncu4":
    case("l1.conv2", 28, 64, 64, 3, 1, 1, ["halo+a"])
elif mode == "ncu3":
    case("l3.conv1", 7, 1024, 256, 1, 1, 0, ["stats"])
    case("l2.conv1", 14, 512, 128, 1, 1, 0, ["stats"])
elif mode == "deep":     # long-K tensor-bound layers: 3 vs 4 operand stages (B2_NO_DEEP=1)
    case("l3.conv1", 7, 1024, 256, 1, 1, 0, ["stats"])
    case("l3.conv2", 7, 256, 256, 3, 1, 1, ["stats"])
    case("l4.conv1", 4, 2048, 512, 1, 1, 0, ["stats"])
    case("l4.conv2", 4, 512, 512, 3, 1, 1, ["stats"])
    case("l4.conv3", 4, 512, 2048, 1, 1, 0, ["statsonly"])
elif mode == "tf":       # is the A transform worth it at the 7x7 / 4x4 stages?
    case("l3.conv3", 7, 256, 1024, 1, 1, 0, ["apply", "a+o+res", "o+res", "a+statsonly", "statsonly", "gram"])
    case("l4.conv3", 4, 512, 2048, 1, 1, 0, ["apply", "a+o+res", "o+res", "a+statsonly", "statsonly"])
    case("l2.conv3", 14, 128, 512, 1, 1, 0, ["apply", "a+o+res", "o+res"])
elif mode == "ncu2":     # the 7x7 / 4x4 stages: conv3 + BN3 + shortcut, statistics-only pass, 1x1 reduce
    case("l3.conv3", 7, 256, 1024, 1, 1, 0, ["a+o+res", "plain"])
    case("l4.conv3", 4, 512, 2048, 1, 1, 0, ["a+o+res", "a+statsonly"])
    case("l4.conv2", 4, 512, 512, 3, 1, 1, ["stats"])
elif mode == "ncu":      # one launch each of the representative kernels, for `ncu --set full`
    case("l1.conv3", 28, 64, 256, 1, 1, 0, ["a+o+res", "gram"])      # EPI_POST (HBM bound) + Gram statistics
    case("l3.conv2", 7, 256, 256, 3, 1, 1, ["stats"])                 # EPI_BF16 3x3 implicit GEMM (tensor bound)
    case("l1.conv2", 28, 64, 64, 3, 1, 1, ["stats"])                  # BN = 64 3x3 (L2 -> SM bound)
    case("l3.conv1", 7, 1024, 256, 1, 1, 0, ["stats"])                # 1x1 reduce, K = 1024
else:
    xs = torch.rand(NF, 3, 112, 112, device=dev)
    wk = ops.pack_stem_weight(torch.randn(64, 3, 7, 7, device=dev) / 12)
    b, st, ss, cnt = bnbuf(64)
    print(f"stem       stats        {timeit(lambda: ops.stem_conv(xs, wk, stats=st)):8.1f} us (pack + conv)")
    print(f"stem       plain        {timeit(lambda: ops.stem_conv(xs, wk)):8.1f} us (pack + conv)")
    case("l1.conv1", 28, 256, 64, 1, 1, 0, ["plain", "stats"])
    case("l1.conv2", 28, 64, 64, 3, 1, 1, ["apply", "plain", "stats", "a", "a+stats", "halo", "halo+a"])
    case("l1.conv3", 28, 64, 256, 1, 1, 0, ["plain", "stats", "statsonly", "a", "a+statsonly", "gram", "a+o", "o+res", "a+o+res", "a+o+res+r"])
    case("l2.conv2", 28, 128, 128, 3, 2, 1, ["apply", "plain", "stats"])
    case("l2.conv2", 14, 128, 128, 3, 1, 1, ["apply", "plain", "stats", "a", "a+stats"])
    case("l2.conv3", 14, 128, 512, 1, 1, 0, ["plain", "stats", "a+statsonly", "gram", "a+o+res"])
    case("l3.conv2", 7, 256, 256, 3, 1, 1, ["plain", "a+stats"])
    case("l3.conv3", 7, 256, 1024, 1, 1, 0, ["plain", "stats", "a+statsonly", "a+o+res"])
print("PROBE DONE")
