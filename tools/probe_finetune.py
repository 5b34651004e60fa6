"""Full fine-tune step of the ResNet-50 frame encoder (forward + backward through backbone_train.py): wall time per step,
and (under ncu) the per-kernel time list."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
arch = sys.argv[3] if len(sys.argv) > 3 else "resnet50"
first = sys.argv[4] if len(sys.argv) > 4 else "conv1"
dev = "cuda"
if arch.startswith("densenet") and first == "conv1":
    first = "features"
torch.manual_seed(0)
net, feat = vc.backbone.make_backbone(arch)
on = False
for n, p in net.named_parameters():
    on = on or n.startswith(first)
    p.requires_grad_(on)
net = net.to(dev).train()
runner = vc.backbone.make_runner(net)
x = torch.rand(frames, 3, 112, 112, device=dev)
g = torch.randn(frames, feat, device=dev)
n0 = vc._lib.launch_count()
WARM = int(os.environ.get('WARM', '2'))
for it in range(steps + WARM):
    if it == WARM:
        torch.cuda.synchronize()
        t0 = time.time()
        n0 = vc._lib.launch_count()
    for p in net.parameters():
        p.grad = None
    f = runner(x, True)
    (f * g).sum().backward()
torch.cuda.synchronize()
dt = (time.time() - t0) / steps
print("%s trainable from %s, %d frames @112: fwd+bwd %.2f ms/step -> %.0f frames/s; %d b2 launches/step; mem %.1f GB" %
      (arch, first, frames, dt * 1e3, frames / dt, (vc._lib.launch_count() - n0) // steps, torch.cuda.max_memory_allocated() / 2**30))
