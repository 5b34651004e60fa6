"""Achieved HBM GB/s of the ingest and element kernels: CUDA events, warm, inputs larger than L2.
Algorithmic bytes = every operand read once + every result written once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc
from video_classif_b200 import ops
from video_classif_b200._lib import call, stream_ptr

dev = "cuda"
BF16 = torch.bfloat16
PEAK = 6552.0      # MEASURED_PEAKS.json copy bandwidth, GB/s


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3        # us


def report(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print(f"{name:74s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of copy peak")


torch.manual_seed(0)
# ingest: 1024 frames, 360x640 BGR uint8 -> 112x112 RGB (cv2-exact bilinear), and the identity-size case of cfg 1
for (F, H0, W0, H, W, dt) in [(1024, 360, 640, 112, 112, BF16), (1024, 360, 640, 112, 112, torch.float32), (2560, 64, 64, 64, 64, torch.float32)]:
    src = torch.randint(0, 256, (F, H0, W0, 3), device=dev, dtype=torch.uint8)
    us = timed(lambda: ops.ingest_u8(src, H, W, out_dtype=dt))
    touched = min(H0 * W0, 4 * H * W) * 3            # only the 4 taps per output pixel are read
    report(f"ingest_u8 {F} x {H0}x{W0} -> {H}x{W} {str(dt)[6:]} (taps actually read)", us, F * (touched + H * W * 3 * (2 if dt == BF16 else 4)))
    del src
M, C = 802816, 256
x = torch.randn(M, C, device=dev).to(BF16)
r = torch.randn(M, C, device=dev).to(BF16)
sc, sh = torch.rand(C, device=dev), torch.rand(C, device=dev)
out = torch.empty_like(x)
report("scale_shift_apply (BN + ReLU)  [802816, 256] bf16", timed(lambda: ops.scale_shift_apply(x, sc, sh, relu=True, out=out)), 2 * M * C * 2)
report("scale_shift_apply (+ shortcut) [802816, 256] bf16", timed(lambda: ops.scale_shift_apply(x, sc, sh, res=r, relu=True, out=out)), 3 * M * C * 2)
# stem tail: BN + ReLU + maxpool 3x3/2 on [1024, 56, 56, 64]
raw = torch.randn(1024, 56, 56, 64, device=dev).to(BF16)
y = torch.empty(1024, 28, 28, 64, device=dev, dtype=BF16)
s = torch.zeros(2, 64, device=dev)
s[0] = raw.float().sum((0, 1, 2)); s[1] = (raw.float() ** 2).sum((0, 1, 2))
g, b, rm, rv = torch.ones(64, device=dev), torch.zeros(64, device=dev), torch.zeros(64, device=dev), torch.ones(64, device=dev)
report("bn_relu_maxpool [1024, 56, 56, 64] -> [1024, 28, 28, 64]",
       timed(lambda: call("b2_bn_relu_maxpool_nhwc", raw.data_ptr(), y.data_ptr(), 1024, 56, 56, 64, s[0].data_ptr(), s[1].data_ptr(), g.data_ptr(),
                          b.data_ptr(), rm.data_ptr(), rv.data_ptr(), 1e-5, 0.0, 1, stream_ptr())), raw.numel() * 2 + y.numel() * 2)
# BatchNorm backward (reduce + apply) with the ReLU mask
dz = torch.randn(M, C, device=dev).to(BF16)
dy = torch.empty_like(x)
st = torch.stack([x.float().sum(0), (x.float() ** 2).sum(0)]).contiguous()
s12 = torch.zeros(2, C, device=dev)
report("bn_bwd (mask + reduce, then apply) [802816, 256] bf16: 6 reads + 1 write",
       timed(lambda: call("b2_bn_bwd_nhwc_bf16", dz.data_ptr(), 0, out.data_ptr(), x.data_ptr(), dy.data_ptr(), sc.data_ptr(), st[0].data_ptr(),
                          st[1].data_ptr(), 0, 0, s12[0].data_ptr(), s12[1].data_ptr(), M, C, M, 1e-5, 1, stream_ptr())), 7 * M * C * 2)
# DenseNet strided apply / depthwise conv
X = torch.randn(802816, 256, device=dev).to(BF16)
a1 = torch.empty(802816, 160, device=dev, dtype=BF16)
ss = torch.rand(2, 192, device=dev)
report("scale_shift_apply_ld X[:, :160] of a 256-wide block buffer -> contiguous",
       timed(lambda: call("b2_scale_shift_apply_ld_bf16", X.data_ptr(), 256, a1.data_ptr(), 160, 802816, 160, ss[0].data_ptr(), ss[1].data_ptr(), 1,
                          stream_ptr())), 2 * 802816 * 160 * 2)
xd = torch.randn(1024, 28, 28, 144, device=dev).to(BF16)
wd = torch.randn(144, 9, device=dev)
yd = torch.empty(1024, 28, 28, 144, device=dev, dtype=BF16)
sd = torch.zeros(2, 144, device=dev)
sc2, sh2 = torch.rand(144, device=dev), torch.rand(144, device=dev)
report("dwconv3x3 (BN + ReLU6 on load, stats) [1024, 28, 28, 144] stride 1",
       timed(lambda: call("b2_dwconv3x3_bn_nhwc_bf16", xd.data_ptr(), sc2.data_ptr(), sh2.data_ptr(), 2, wd.data_ptr(), yd.data_ptr(), sd[0].data_ptr(),
                          sd[1].data_ptr(), 1024, 28, 28, 144, 1, stream_ptr())), 2 * xd.numel() * 2)
report("dwconv3x3 (input already activated, stats) [1024, 28, 28, 144] stride 1",
       timed(lambda: call("b2_dwconv3x3_bn_nhwc_bf16", xd.data_ptr(), 0, 0, 0, wd.data_ptr(), yd.data_ptr(), sd[0].data_ptr(),
                          sd[1].data_ptr(), 1024, 28, 28, 144, 1, stream_ptr())), 2 * xd.numel() * 2)
xd2 = torch.randn(1024, 56, 56, 96, device=dev).to(BF16)
yd2 = torch.empty(1024, 28, 28, 96, device=dev, dtype=BF16)
wd2 = torch.randn(96, 9, device=dev)
report("dwconv3x3 (BN + ReLU6 on load, stats) [1024, 56, 56, 96] stride 2",
       timed(lambda: call("b2_dwconv3x3_bn_nhwc_bf16", xd2.data_ptr(), sc2.data_ptr(), sh2.data_ptr(), 2, wd2.data_ptr(), yd2.data_ptr(), sd[0].data_ptr(),
                          sd[1].data_ptr(), 1024, 56, 56, 96, 2, stream_ptr())), (xd2.numel() + yd2.numel()) * 2)
