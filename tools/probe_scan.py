"""GPU probe: fused selective scan (BASELINE config 5 shapes) -- achieved HBM GB/s against the algorithmic
(3 D + 2 N) * 4 bytes per token, next to the oracle port (the reference's pure-PyTorch scan) on the host cores."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import lrcn_oracle as O
from video_classif_b200 import ops

dev = "cuda"
print(torch.cuda.get_device_name(0), flush=True)
SHAPES = [(8, 16, 2048, 16, None), (8, 3136, 2048, 16, 256), (64, 16, 2048, 16, None)]
if len(sys.argv) > 1 and sys.argv[1] == "big":          # ncu captures: the config-5 shape only
    SHAPES = SHAPES[1:2]
for (B, L, D, N, chunk) in SHAPES:
    g = torch.Generator().manual_seed(L)
    u = torch.randn(B, L, D, generator=g)
    delta = F.softplus(torch.randn(B, L, D, generator=g))
    A = -torch.exp(torch.randn(D, N, generator=g))
    Bm, Cm = torch.randn(B, L, N, generator=g), torch.randn(B, L, N, generator=g)
    args = [t.to(dev) for t in (u, delta, A, Bm, Cm)]
    for _ in range(3):
        y = ops.selective_scan(*args, chunk_reset=chunk)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = 20
    e0.record()
    for _ in range(reps):
        y = ops.selective_scan(*args, chunk_reset=chunk)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    nbytes = (3 * D + 2 * N) * 4 * B * L
    line = f"scan B={B} L={L} D={D} N={N} chunk={chunk}: {us:9.1f} us  {nbytes / us / 1e3:8.1f} GB/s algorithmic"
    if B * L <= 8 * 3136:
        Ls = min(L, 256)
        t0 = time.perf_counter()
        ref = O.selective_scan(u[:, :Ls], delta[:, :Ls], A, Bm[:, :Ls], Cm[:, :Ls], chunk_reset=chunk)
        cpu_s = time.perf_counter() - t0
        err = (y[:, :Ls].cpu() - ref).abs().max().item() / ref.abs().max().item()
        line += f" | oracle port on {os.cpu_count()} host cores: {cpu_s * 1e6 / (B * Ls):8.1f} us/token vs GPU {us / (B * L):.4f} us/token, rel err {err:.1e}"
    print(line, flush=True)
    # forward + backward (BPTT over recomputed states; chunks in parallel)
    gargs = [t.clone().requires_grad_(True) for t in args]
    w = torch.randn_like(args[0])
    for _ in range(2):
        for t in gargs:
            t.grad = None
        (ops.selective_scan(*gargs, chunk_reset=chunk) * w).sum().backward()
    e0.record()
    for _ in range(5):
        for t in gargs:
            t.grad = None
        (ops.selective_scan(*gargs, chunk_reset=chunk) * w).sum().backward()
    e1.record()
    torch.cuda.synchronize()
    print(f"     forward + backward: {e0.elapsed_time(e1) / 5 * 1e3:9.1f} us ({torch.cuda.max_memory_allocated() / 2**30:.1f} GB peak)", flush=True)
print("PROBE DONE")
