"""ingest kernel timing (CUDA events, inputs > L2) for the shapes of SURVEY section 8(d): python tools/probe_ingest.py [once]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_classif_b200 import ops
dev = "cuda"
once = len(sys.argv) > 1
for (F, H0, W0, H, W, dt) in [(1024, 360, 640, 112, 112, torch.bfloat16), (1024, 360, 640, 112, 112, torch.float32),
                              (2560, 64, 64, 64, 64, torch.float32), (1024, 112, 112, 112, 112, torch.float32),
                              (1024, 240, 320, 112, 112, torch.float32)]:
    src = torch.randint(0, 256, (F, H0, W0, 3), device=dev, dtype=torch.uint8)
    ops.ingest_u8(src, H, W, out_dtype=dt)
    torch.cuda.synchronize()
    if once:
        continue
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        ops.ingest_u8(src, H, W, out_dtype=dt)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    rows_touched = min(H0, 2 * H) * W0 * 3 if (H0, W0) != (H, W) else H0 * W0 * 3
    out_b = H * W * 3 * (2 if dt == torch.bfloat16 else 4)
    print(f"ingest {F} x {H0}x{W0} -> {H}x{W} {str(dt)[6:]:8s}: {us:7.1f} us; source rows touched + output = "
          f"{F * (rows_touched + out_b) / 1e6:6.1f} MB -> {F * (rows_touched + out_b) / us / 1e3:6.0f} GB/s")
    del src
