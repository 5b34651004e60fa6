"""GPU probe: per-tile cost model of the persistent tcgen05 conv kernel.  1x1 conv over [1024 x 7 x 7, C] (M = 50176 rows =
392 row blocks) for C (= K) in 64..2048 and Cout 256 / 1024, statistics-only pass vs bf16 store + statistics:
time per tile = a + b K separates the fixed per-tile cost (epilogue) from the per-k-block cost (operand delivery / MMA).
Run with B2_PAIR=0 for the one-CTA-per-tile form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_classif_b200 import _lib

dev = "cuda"
_lib.call("b2_device_check")
st = lambda: torch.cuda.current_stream().cuda_stream
NI, H, W = 1024, 7, 7
M = NI * H * W
SMS = torch.cuda.get_device_properties(0).multi_processor_count
print(f"pair={os.environ.get('B2_PAIR', '1')}  M={M}  SMs={SMS}")
for Cout in (256, 1024):
    for C in (64, 128, 256, 512, 1024, 2048):
        for store in (0, 1):
            x = torch.randn(NI, H, W, C, device=dev).bfloat16()
            w = (torch.randn(Cout, 1, 1, C, device=dev) / C ** 0.5).bfloat16()
            y = torch.empty(NI, H, W, Cout, device=dev, dtype=torch.bfloat16)
            s1 = torch.zeros(Cout, device=dev); s2 = torch.zeros(Cout, device=dev)
            args = (x.data_ptr(), NI, H, W, C, w.data_ptr(), Cout, 1, 1, 1, 0, y.data_ptr() if store else 0,
                    0, 0, 0, 0, 0, 0, 0, 0, 0, s1.data_ptr(), s2.data_ptr(), 0, 0, 0, 0, 0, 0, 0, 1e-5, 0.1, st())
            for _ in range(3): _lib.call("b2_conv2d_bn_nhwc_bf16", *args)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            reps = 20
            e0.record()
            for _ in range(reps): _lib.call("b2_conv2d_bn_nhwc_bf16", *args)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            tiles = ((M + 127) // 128) * (Cout // 256)
            per_cta = -(-tiles // SMS)
            print(f"Cout={Cout:5d} K={C:5d} store={store}: {us:7.1f} us  tiles/CTA={per_cta:3d}  us/tile={us / per_cta:5.2f}  "
                  f"us/kblock={us / per_cta / (C // 64):5.3f}  {2.0 * M * Cout * C / us / 1e6:7.1f} TF/s")
