"""One profiled pass of each kernel family changed in round 2, for `ncu --set full --profile-from-start off`:
    enc    one eager encoder pass of the bench config (ResNet-50, 1024 x 112x112, train-mode BN)
    cfg1   one train step of the notebook small-CNN LRCN on the tensor-core trunk (B = 64)
    scan   selective scan forward + backward (B 8, L 3136, D 2048, N 16, chunk 256)
Everything before torch.cuda.profiler.start() is warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import video_classif_b200 as vc

what = sys.argv[1]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
prof = torch.cuda.profiler
if what == "enc":
    m = vc.LRCN(4, 16, 32, 8, cnn_backbone="resnet50", rnn_layers=3, dropout=0.25, precision="bf16").to(dev).train()
    x = torch.rand(64, 16, 3, 112, 112, device=dev)
    with torch.no_grad():
        m._features(x); m._features(x)
        torch.cuda.synchronize(); prof.start()
        m._features(x)
        torch.cuda.synchronize(); prof.stop()
elif what == "cfg1":
    m = vc.SmallCNNLRCN(50, 20, 32, (3, 64, 64), dropout=0.5, precision="bf16").to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
    x = torch.rand(64, 20, 3, 64, 64, device=dev) * 255
    y = torch.randint(0, 50, (64,), device=dev)
    for i in range(3):
        if i == 2:
            torch.cuda.synchronize(); prof.start()
        opt.zero_grad(set_to_none=True)
        F.cross_entropy(m(x), y).backward()
        opt.step()
    torch.cuda.synchronize(); prof.stop()
else:
    from video_classif_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, L, D, N = 8, 3136, 2048, 16
    u = torch.randn(B, L, D, generator=g).to(dev).requires_grad_(True)
    delta = F.softplus(torch.randn(B, L, D, generator=g)).to(dev).requires_grad_(True)
    A = (-torch.exp(torch.randn(D, N, generator=g))).to(dev).requires_grad_(True)
    Bm = torch.randn(B, L, N, generator=g).to(dev).requires_grad_(True)
    Cm = torch.randn(B, L, N, generator=g).to(dev).requires_grad_(True)
    w = torch.randn(B, L, D, generator=g).to(dev)
    for i in range(2):
        if i == 1:
            torch.cuda.synchronize(); prof.start()
        (ops.selective_scan(u, delta, A, Bm, Cm, chunk_reset=256) * w).sum().backward()
    torch.cuda.synchronize(); prof.stop()
print("done", what)
