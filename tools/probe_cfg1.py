"""cfg 1 (BASELINE.json configs[0]: notebook small-CNN LRCN, 20 x 64x64, 50 classes) train-step rate on one B200:
fp32 parity path vs the bf16 tensor-core path, B = 8 (the reference's batch) and B = 64 (the scaling batch).
    python tools/probe_cfg1.py [--stock]      (--stock: the same model in stock PyTorch on the same GPU, fp32 and bf16 autocast)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_classif_b200 as vc  # noqa: E402

dev = torch.device("cuda", 0)


def rate(step, B, steps=20):
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return B / ms * 1e3, ms


print(torch.cuda.get_device_name(0))
torch.manual_seed(0)
if "--profile" in sys.argv:           # 4 steps of the B = 64 bf16 path only (for an ncu launch list)
    B = 64
    x = torch.rand(B, 20, 3, 64, 64, device=dev) * 255
    y = torch.randint(0, 50, (B,), device=dev)
    m = vc.SmallCNNLRCN(50, 20, 32, (3, 64, 64), dropout=0.5, precision="bf16").to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(m(x), y).backward()
        opt.step()
    torch.cuda.synchronize()
    sys.exit(0)
for B in (8, 64):
    x = torch.rand(B, 20, 3, 64, 64, device=dev) * 255
    y = torch.randint(0, 50, (B,), device=dev)
    for prec in ("fp32", "bf16"):
        m = vc.SmallCNNLRCN(50, 20, 32, (3, 64, 64), dropout=0.5, precision=prec).to(dev).train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)

        def step():
            opt.zero_grad(set_to_none=True)
            torch.nn.functional.cross_entropy(m(x), y).backward()
            opt.step()
        r, ms = rate(step, B)
        gf = 5.0 * B / ms                      # 5.0 GFLOP per clip, train step
        print(f"cfg1 small-CNN LRCN 20x64x64 B={B} {prec}: {r:9.0f} clips/s ({ms:.3f} ms/step, {gf:.1f} TFLOP/s algorithmic)")
        # the same step replayed from one CUDA graph (GraphedTrainStep): at B = 8 the eager step is bound by the host launch path
        mg = vc.SmallCNNLRCN(50, 20, 32, (3, 64, 64), dropout=0.5, precision=prec).to(dev).train()
        og = torch.optim.Adam(mg.parameters(), lr=1e-4, fused=True, capturable=True)
        gstep = vc.GraphedTrainStep(mg, og, torch.nn.CrossEntropyLoss(), x, y)
        rg, msg = rate(lambda: gstep(x, y), B)
        print(f"      whole step as one CUDA-graph replay: {rg:9.0f} clips/s ({msg:.3f} ms/step)")
        if prec == "bf16":
            # forward only / forward + backward split
            with torch.no_grad():
                rf, msf = rate(lambda: m(x), B)
            print(f"      forward only (no grad): {msf:.3f} ms")
    if "--stock" in sys.argv:
        class Ref(torch.nn.Module):               # the notebook topology in plain torch.nn (cuDNN / cuBLAS)
            def __init__(self):
                super().__init__()
                nn = torch.nn
                self.conv1, self.conv2, self.conv3 = nn.Conv2d(3, 16, 3, padding=1), nn.Conv2d(16, 32, 3, padding=1), nn.Conv2d(32, 64, 3, padding=1)
                self.bn1, self.bn2, self.bn3 = nn.BatchNorm2d(16), nn.BatchNorm2d(32), nn.BatchNorm2d(64)
                self.pool, self.dropout = nn.MaxPool2d(2, 2), nn.Dropout(0.5)
                self.lstm = nn.LSTM(16384, 32, num_layers=2, batch_first=True)
                self.fc = nn.Linear(640, 50)

            def forward(self, x):
                b, t, c, h, w = x.shape
                x = x.view(b * t, c, h, w)
                x = torch.relu(self.bn1(self.conv1(x)))
                x = self.pool(torch.relu(self.bn2(self.conv2(x))))
                x = self.dropout(self.pool(torch.relu(self.bn3(self.conv3(x)))))
                x, _ = self.lstm(x.reshape(b, t, -1))
                return self.fc(x.contiguous().view(b, -1))
        for tag in ("fp32", "bf16 autocast channels_last"):
            r = Ref().to(dev).train()
            if tag != "fp32":
                r = r.to(memory_format=torch.channels_last)
            opt = torch.optim.Adam(r.parameters(), lr=1e-4)

            def step():
                opt.zero_grad(set_to_none=True)
                if tag == "fp32":
                    loss = torch.nn.functional.cross_entropy(r(x), y)
                else:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        loss = torch.nn.functional.cross_entropy(r(x), y)
                loss.backward()
                opt.step()
            rr, ms = rate(step, B, steps=10)
            print(f"      stock PyTorch {torch.__version__} {tag}: {rr:9.0f} clips/s ({ms:.3f} ms/step)")
