"""The hoisted LSTM gate GEMM of cfg 1 (north_star kernel 3): G[B*T, 4H] = X[B*T, 16384] W_ih^T, bf16 operands, fp32 out.
N = 4H = 128 makes it HBM bound (arithmetic intensity ~ 125 FLOP/B): reported as a fraction of the measured copy bandwidth,
with and without split-K, plus its two backward products.
    python tools/probe_gate_gemm.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_classif_b200 import ops  # noqa: E402

dev = "cuda"
peaks = {"hbm_gbs": 6552.3}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                                   # operands out of L2
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for BT in (160, 1280):
    X = torch.randn(BT, 16384, device=dev).to(torch.bfloat16)
    W = torch.randn(128, 16384, device=dev).to(torch.bfloat16)
    b = torch.randn(128, device=dev)
    ref = X.float() @ W.float().t() + b
    for tag, env in (("split-K", None), ("one CTA per tile", "1")):
        if env:
            os.environ["B2_NO_SPLITK"] = env
        else:
            os.environ.pop("B2_NO_SPLITK", None)
        out = ops.gemm_tn(X, W, bias=b)
        err = ((out - ref).abs().max() / ref.abs().max()).item()
        us = timed(lambda: ops.gemm_tn(X, W, bias=b))
        byts = (X.numel() + W.numel()) * 2 + out.numel() * 4
        print(f"gate GEMM [{BT},16384]x[16384,128] {tag:18s}: {us:7.1f} us  {byts / 1e6:6.1f} MB  {byts / us / 1e3:7.0f} GB/s "
              f"= {byts / us / 1e3 / peaks['hbm_gbs']:.2f} of copy peak, {2 * BT * 128 * 16384 / us / 1e6:6.1f} TFLOP/s, rel err {err:.1e}")
    os.environ.pop("B2_NO_SPLITK", None)
    dG = torch.randn(BT, 128, device=dev)
    us = timed(lambda: ops.gemm_tn(ops.cast_bf16(dG), ops.transpose_cast_bf16(W.float()), out_dtype=torch.bfloat16))
    print(f"   dX = dG W_ih   [{BT},128]x[128,16384] (bf16 out, incl. operand casts): {us:7.1f} us")
    us = timed(lambda: ops.gemm_tn(ops.transpose_cast_bf16(dG), ops.transpose_bf16(X), out_dtype=torch.float32))
    print(f"   dW = dG^T X    [128,{BT}]x[{BT},16384] (fp32 out, incl. transposes):   {us:7.1f} us")
    ref_dw = dG.t() @ X.float()
    got_dw = ops.linear_wgrad_tc(X, dG)
    us = timed(lambda: ops.linear_wgrad_tc(X, dG))
    print(f"   dW on the MN-major weight-gradient kernel (no transposes; incl. cast of dG + zero fill): {us:7.1f} us, "
          f"rel err {((got_dw - ref_dw).abs().max() / ref_dw.abs().max()).item():.1e}")
