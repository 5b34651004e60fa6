"""Frozen DenseNet-121 encoder pass (crime / rgb LRCN default backbone): eager and CUDA-graph replay."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
size = int(sys.argv[2]) if len(sys.argv) > 2 else 112
arch = sys.argv[3] if len(sys.argv) > 3 else "densenet121"
GF = {"densenet121": 2 * 2.87e9, "mobilenet_v2": 2 * 0.32e9, "resnet50": 2 * 4.09e9}[arch]      # FLOP/frame at 224^2
dev = "cuda"
net, feat = vc.backbone.make_backbone(arch)
for p in net.parameters():
    p.requires_grad = False
net = net.to(dev).train()
r = vc.backbone.make_runner(net)
x = torch.rand(frames, 3, size, size, device=dev)
for mode in ("eager", "graph"):
    fn = (lambda: r(x, True)) if mode == "eager" else (lambda: r.graphed(x, True))
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        n0 = vc._lib.launch_count()
        t0 = time.time()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
    dt = (time.time() - t0) / 5
    # densenet121: 2.87 GFLOP/frame at 224^2 (torchvision), x (size/224)^2
    print("%s %s: %d frames @%d: %.2f ms -> %.0f frames/s, %.0f TFLOP/s, %d launches" %
          (arch, mode, frames, size, dt * 1e3, frames / dt, frames / dt * GF * (size / 224) ** 2 / 1e12, (vc._lib.launch_count() - n0) // 5))
