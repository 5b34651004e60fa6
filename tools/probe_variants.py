"""Train-step rate of the medsos LRCN temporal variants (frozen ResNet-50 encoder replayed from a CUDA graph, 16 x 112x112, B = 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc

dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.rand(64, 16, 3, 112, 112, device=dev)
y = torch.randint(0, 4, (64,), device=dev)
for rnn_type, bidir in (("lstm", False), ("gru", False), ("gru", True), ("mamba", False), ("mamba", True)):
    m = vc.LRCN(4, 16, 32, 8, cnn_backbone="resnet50", rnn_type=rnn_type, rnn_layers=3, bidirectional=bidir, dropout=0.25).to(dev).train()
    m.enable_encoder_graph()
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward()
        opt.step()
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    n0 = vc._lib.launch_count()
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"medsos LRCN rnn_type={rnn_type} bidirectional={bidir}: {64 / ms * 1e3:8.0f} clips/s ({ms:.2f} ms/step, un-pipelined, "
          f"{(vc._lib.launch_count() - n0) // 10} b2 launches)")
    del m, opt
