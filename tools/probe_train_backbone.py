"""Probe: backward kernels of the trainable encoder vs torch (fp32) -- error magnitudes and timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import video_classif_b200 as vc
from video_classif_b200 import backbone_train as BT
from video_classif_b200.ops import BF16

dev = "cuda"
torch.manual_seed(0)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


print("== wgrad")
for (N, H, W, C, Cout, R, s, p) in [(3, 10, 10, 64, 64, 3, 1, 1), (2, 9, 9, 128, 256, 1, 1, 0), (2, 14, 14, 256, 512, 3, 2, 1),
                                    (5, 7, 7, 512, 128, 1, 2, 0), (1, 1, 777, 168, 64, 1, 1, 0), (4, 28, 28, 64, 256, 1, 1, 0),
                                    (2, 8, 8, 2048, 512, 1, 1, 0), (2, 7, 7, 512, 512, 3, 1, 1)]:
    x = torch.randn(N, H, W, C, device=dev).to(BF16)
    P = (H + 2 * p - R) // s + 1
    Q = (W + 2 * p - R) // s + 1
    dy = torch.randn(N, P, Q, Cout, device=dev).to(BF16)
    dw = BT.conv_wgrad(x, dy, R, R, s, p)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wref = torch.zeros(Cout, C, R, R, device=dev, requires_grad=True)
    F.conv2d(xr, wref, stride=s, padding=p).backward(dy.float().permute(0, 3, 1, 2))
    print((N, H, W, C, Cout, R, s, p), "rel err", rel(dw.permute(0, 3, 1, 2), wref.grad))

print("== dgrad")
for (N, H, W, C, Cout, R, s, p) in [(3, 10, 10, 64, 64, 3, 1, 1), (2, 9, 9, 128, 256, 1, 1, 0), (2, 14, 14, 256, 512, 3, 2, 1),
                                    (5, 7, 7, 512, 128, 1, 2, 0), (2, 7, 7, 256, 256, 3, 2, 1)]:
    w = torch.randn(Cout, C, R, R, device=dev) * 0.05
    P = (H + 2 * p - R) // s + 1
    Q = (W + 2 * p - R) // s + 1
    dy = torch.randn(N, P, Q, Cout, device=dev).to(BF16)
    dx = BT.conv_dgrad(dy, w, (H, W), s, p)
    xr = torch.zeros(N, C, H, W, device=dev, requires_grad=True)
    F.conv2d(xr, w.to(BF16).float(), stride=s, padding=p).backward(dy.float().permute(0, 3, 1, 2))
    print((N, H, W, C, Cout, R, s, p), "rel err", rel(dx.float().permute(0, 3, 1, 2), xr.grad))

print("== single conv+bn(+res)+relu node vs torch on identical inputs")
rnd = lambda t: t.bfloat16().float()
for (N, H, W, C, Cout, R, s_, p, relu, has_res, train) in [(6, 8, 8, 64, 64, 3, 1, 1, True, False, True), (6, 8, 8, 64, 256, 1, 1, 0, True, True, True),
                                                    (6, 8, 8, 128, 128, 3, 2, 1, True, False, True), (6, 8, 8, 256, 512, 1, 2, 0, False, False, True),
                                                    (6, 4, 4, 512, 2048, 1, 1, 0, True, True, True), (6, 8, 8, 64, 256, 1, 1, 0, True, True, False)]:
    conv = torch.nn.Conv2d(C, Cout, R, s_, p, bias=False).to(dev)
    bn = torch.nn.BatchNorm2d(Cout).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.3); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 2.0)
    P = (H + 2 * p - R) // s_ + 1
    x = torch.randn(N, H, W, C, device=dev).to(BF16).requires_grad_(True)
    res = torch.randn(N, P, P, Cout, device=dev).to(BF16).requires_grad_(True) if has_res else None
    G = torch.randn(N, P, P, Cout, device=dev).to(BF16)
    bn_ref = torch.nn.BatchNorm2d(Cout).to(dev); bn_ref.load_state_dict(bn.state_dict())
    z = BT.ConvBnFn.apply(x, conv.weight, bn.weight, bn.bias, res, bn, s_, p, relu, train)
    (z.float() * G.float()).sum().backward()
    got = dict(dx=x.grad.float(), dw=conv.weight.grad, dg=bn.weight.grad, db=bn.bias.grad)
    if has_res: got["dres"] = res.grad.float()
    # reference: same graph in fp32 with bf16 storage at the same points
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    rr = res.detach().float().permute(0, 3, 1, 2).requires_grad_(True) if has_res else None
    bn_ref.train(train)
    y = rnd(F.conv2d(xr, rnd(wr), stride=s_, padding=p))
    o = bn_ref(y)
    if has_res: o = o + rr
    if relu: o = torch.relu(o)
    zr = rnd(o)
    (zr * G.float().permute(0, 3, 1, 2)).sum().backward()
    ref = dict(dx=xr.grad.permute(0, 2, 3, 1), dw=wr.grad, dg=bn_ref.weight.grad, db=bn_ref.bias.grad)
    if has_res: ref["dres"] = rr.grad.permute(0, 2, 3, 1)
    print((N, H, W, C, Cout, R, s_, p, relu, has_res, train), "z", round(rel(z.float().permute(0, 3, 1, 2), zr), 4),
          {k: round(rel(got[k], ref[k]), 4) for k in got}, "rm", round(rel(bn.running_mean, bn_ref.running_mean), 5))

print("== stem node")
conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).to(dev)
bn = torch.nn.BatchNorm2d(64).to(dev)
bn_ref = torch.nn.BatchNorm2d(64).to(dev)
x = torch.rand(6, 3, 32, 32, device=dev)
y = BT.StemFn.apply(x, conv.weight, bn.weight, bn.bias, bn, True)
G = torch.randn_like(y.float()).to(BF16)
(y.float() * G.float()).sum().backward()
wr = conv.weight.detach().clone().requires_grad_(True)
a = rnd(F.conv2d(rnd(x), rnd(wr), stride=2, padding=3))
yr = rnd(F.max_pool2d(torch.relu(bn_ref(a)), 3, 2, 1))
(yr * G.float().permute(0, 3, 1, 2)).sum().backward()
print("stem y", round(rel(y.float().permute(0, 3, 1, 2), yr), 4), "dw", round(rel(conv.weight.grad, wr.grad), 4), "dg", round(rel(bn.weight.grad, bn_ref.weight.grad), 4),
      "db", round(rel(bn.bias.grad, bn_ref.bias.grad), 4))

print("== shallow nets vs bf16-emulated oracle autograd")
import torchvision
from torchvision.models.resnet import BasicBlock, Bottleneck
from oracle import lrcn_oracle as O
for kind, nfr, size in (("basic", 12, 64), ("bottleneck", 12, 64), ("bottleneck", 6, 96)):
    for first in ("conv1", "layer1", "layer3", "layer4.0.bn2"):
        torch.manual_seed(3)
        net = torchvision.models.ResNet(BasicBlock if kind == "basic" else Bottleneck, [1, 1, 1, 1])
        net.fc = torch.nn.Identity()
        on = False
        for n, p_ in net.named_parameters():
            on = on or n.startswith(first)
            p_.requires_grad_(on)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        for k, v in sd.items():
            if v.dtype.is_floating_point and k in dict(net.named_parameters()) and dict(net.named_parameters())[k].requires_grad:
                v.requires_grad_(True)
        gen = torch.Generator().manual_seed(5)
        x = torch.nn.functional.interpolate(torch.rand(nfr, 3, size // 4, size // 4, generator=gen), size=size, mode="bilinear")
        x = (x + 0.1 * torch.rand(nfr, 3, size, size, generator=gen)).clamp(0, 1)
        g = torch.randn(nfr, 512 if kind == "basic" else 2048, generator=gen)
        arch = "resnet10" if kind == "basic" else "resnet14"
        emu, _ = O.resnet_features(sd, x, arch, True, prefix="", emulate_bf16=True)
        (emu * g).sum().backward()
        net = net.to(dev).train()
        runner = vc.backbone.ResNetRunner(net)
        f = runner(x.to(dev), True)
        (f * g.to(dev)).sum().backward()
        errs = sorted(((rel(p_.grad, sd[n].grad), n) for n, p_ in net.named_parameters() if p_.requires_grad), reverse=True)
        none = [n for n, p_ in net.named_parameters() if not p_.requires_grad and p_.grad is not None]
        print(kind, nfr, size, "trainable from", first, "feat", round(rel(f, emu), 4), "worst", [(round(e, 4), n) for e, n in errs[:3]],
              "median", round(errs[len(errs) // 2][0], 4), "frozen-with-grad", none)

print("== timing: resnet50 full fine-tune step, 256 frames 112x112")
net, feat = vc.backbone.make_backbone("resnet50")
net = net.to(dev).train()
runner = vc.backbone.ResNetRunner(net)
x = torch.rand(256, 3, 112, 112, device=dev)
g = torch.randn(256, feat, device=dev)
for it in range(5):
    if it == 2:
        torch.cuda.synchronize(); t0 = time.time()
    for p_ in net.parameters():
        p_.grad = None
    f = runner(x, True)
    (f * g).sum().backward()
torch.cuda.synchronize()
dt = (time.time() - t0) / 3
print("fwd+bwd %.2f ms -> %.0f frames/s, %.1f TFLOP/s (3x 2.152 GFLOP/frame)" % (dt * 1e3, 256 / dt, 256 * 3 * 2.152e-3 / dt))
