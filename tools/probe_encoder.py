"""GPU probe: the frozen encoder pass alone (CUDA-graph replay) vs the full pipelined train step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = vc.LRCN(4, 16, 32, 8, cnn_backbone="resnet50", rnn_layers=3, dropout=0.25, precision="bf16").to(dev).train()
model.enable_encoder_graph()
xs = [torch.rand(64, 16, 3, 112, 112, device=dev) for _ in range(4)]
with torch.no_grad():
    for i in range(3):
        model._features(xs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(30):
        model._features(xs[i % 4])
    e1.record(); torch.cuda.synchronize()
print(f"encoder pass alone (graph replay): {e0.elapsed_time(e1) / 30:.3f} ms per 1024 frames")
