"""GPU probe: tcgen05 GEMM + im2col-TMA conv against torch fp32 (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import video_classif_b200 as vc
from video_classif_b200 import _lib

torch.manual_seed(0)
dev = "cuda"
print(torch.cuda.get_device_name(0), flush=True)
_lib.call("b2_device_check")
st = lambda: torch.cuda.current_stream().cuda_stream


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def gemm_case(M, N, K, bias=True, out_bf16=True, relu=False, stats=False, reps=0):
    A = (torch.randn(M, K, device=dev)).bfloat16()
    B = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    D = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    s1 = torch.zeros(N, device=dev) if stats else None
    s2 = torch.zeros(N, device=dev) if stats else None
    args = (A.data_ptr(), K, B.data_ptr(), K, D.data_ptr(), N, M, N, K, _lib.ptr(b), 0, int(out_bf16), int(relu),
            _lib.ptr(s1), _lib.ptr(s2), st())
    _lib.call("b2_gemm_bf16_tn", *args)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    if bias: ref = ref + b
    if relu: ref = ref.relu()
    e = relerr(D.float(), ref)
    msg = f"gemm M={M} N={N} K={K} bias={bias} bf16out={out_bf16} relu={relu}: relerr={e:.3e}"
    if stats:
        r = ref.bfloat16().float() if out_bf16 else ref
        msg += f" sum={relerr(s1, r.sum(0)):.2e} sumsq={relerr(s2, (r*r).sum(0)):.2e}"
    if reps:
        for _ in range(3): _lib.call("b2_gemm_bf16_tn", *args)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps): _lib.call("b2_gemm_bf16_tn", *args)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        msg += f"  {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s  {(M*K+M*N)*2/ms/1e6:.0f} GB/s"
    print(msg, flush=True)
    return e


def conv_case(N, H, W, C, Cout, R, stride, pad, stats=True, reps=0):
    x = torch.randn(N, C, H, W, device=dev).bfloat16()
    w = (torch.randn(Cout, C, R, R, device=dev) / (C * R * R) ** 0.5).bfloat16()
    xn = x.permute(0, 2, 3, 1).contiguous()
    wn = w.permute(0, 2, 3, 1).contiguous()
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    y = torch.full((N, P, Q, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    s1 = torch.zeros(Cout, device=dev) if stats else None
    s2 = torch.zeros(Cout, device=dev) if stats else None
    args = (xn.data_ptr(), N, H, W, C, wn.data_ptr(), Cout, R, R, stride, pad, y.data_ptr(), 0, 1, 0,
            _lib.ptr(s1), _lib.ptr(s2), st())
    _lib.call("b2_conv2d_nhwc_bf16", *args)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    e = relerr(y.float(), ref)
    msg = f"conv N={N} {H}x{W}x{C}->{Cout} k{R} s{stride} p{pad}: relerr={e:.3e}"
    if stats:
        r = ref.bfloat16().float().reshape(-1, Cout)
        msg += f" sum={relerr(s1, r.sum(0)):.2e} sumsq={relerr(s2, (r*r).sum(0)):.2e}"
    if reps:
        for _ in range(3): _lib.call("b2_conv2d_nhwc_bf16", *args)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps): _lib.call("b2_conv2d_nhwc_bf16", *args)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        msg += f"  {ms*1e3:.1f} us  {2.0*N*P*Q*Cout*C*R*R/ms/1e9:.1f} TFLOP/s"
    print(msg, flush=True)
    return e

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm"):
    gemm_case(128, 32, 64, bias=False)
    gemm_case(128, 64, 64)
    gemm_case(128, 128, 128)
    gemm_case(256, 256, 256, stats=True)
    gemm_case(300, 200, 136, stats=True)
    gemm_case(1000, 8, 512, out_bf16=False)
    gemm_case(160, 128, 16384, out_bf16=False)
    gemm_case(1920, 1024, 2048, relu=True, stats=True)
    gemm_case(1920, 224, 112, out_bf16=False)
    gemm_case(4096, 50, 640, out_bf16=False)
    gemm_case(8192, 8192, 8192, bias=False, reps=5)
    gemm_case(1505280, 256, 64, bias=False, stats=True, reps=5)
    gemm_case(1505280, 64, 256, bias=False, stats=True, reps=5)
    gemm_case(94080, 2048, 512, bias=False, stats=True, reps=5)
if which in ("all", "layers"):
    for st_ in (False, True):
        gemm_case(802816, 256, 64, bias=False, stats=st_, reps=5)
        gemm_case(802816, 64, 256, bias=False, stats=st_, reps=5)
        gemm_case(200704, 512, 128, bias=False, stats=st_, reps=5)
        gemm_case(200704, 128, 512, bias=False, stats=st_, reps=5)
        gemm_case(50176, 1024, 256, bias=False, stats=st_, reps=5)
        gemm_case(16384, 2048, 512, bias=False, stats=st_, reps=5)
        gemm_case(3211264, 64, 160, bias=False, stats=st_, reps=5)
if which == "one":
    gemm_case(802816, 256, 64, bias=False, stats=False, reps=0)
    gemm_case(802816, 256, 64, bias=False, stats=True, reps=0)
    gemm_case(802816, 64, 256, bias=False, stats=True, reps=0)
if which in ("all", "conv"):
    conv_case(2, 8, 8, 64, 64, 1, 1, 0)
    conv_case(2, 8, 8, 64, 64, 3, 1, 1)
    conv_case(3, 14, 14, 128, 128, 3, 1, 1)
    conv_case(3, 28, 28, 128, 128, 3, 2, 1)
    conv_case(3, 7, 7, 256, 512, 1, 2, 0)
    conv_case(5, 7, 7, 512, 512, 3, 2, 1)
    conv_case(64, 28, 28, 64, 64, 3, 1, 1, reps=5)
    conv_case(1920, 28, 28, 64, 64, 3, 1, 1, reps=5)
    conv_case(1920, 14, 14, 128, 128, 3, 1, 1, reps=5)
    conv_case(1920, 7, 7, 256, 256, 3, 1, 1, reps=5)
    conv_case(1920, 4, 4, 512, 512, 3, 1, 1, reps=5)
print("PROBE DONE", flush=True)
