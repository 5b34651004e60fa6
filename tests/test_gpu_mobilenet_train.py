"""Trainable MobileNetV2 encoder (mobilenet_train.py, csrc/mobilenet_bwd.cu).  NOTE: the reference's fine-tuning scripts reject
this backbone (lrcn/lrcn.py:192-206 and rgb_lrcn.py:178-192 raise `Unsupported CNN backbone` for anything but resnet / densenet /
vgg names) and medsos models.py:144-145 always freezes it, so there is no reference golden for its gradients: the kernels are
checked one by one against torch fp32 on the same bf16 operands, and the whole trunk against torchvision's own autograd."""
import pytest
import torch

from conftest import err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("C,H,W,stride", [(32, 20, 28, 1), (96, 28, 28, 2), (144, 14, 14, 1), (24, 9, 11, 2)])
def test_depthwise_dgrad_wgrad_vs_torch(C, H, W, stride):
    from video_classif_b200._lib import call, stream_ptr
    N = 3
    torch.manual_seed(C + stride)
    x = _bf(torch.randn(N, C, H, W))
    w = torch.randn(C, 1, 3, 3) * 0.3
    P, Q = (H - 1) // stride + 1, (W - 1) // stride + 1
    dy = _bf(torch.randn(N, C, P, Q))
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    torch.nn.functional.conv2d(xr, wr, None, stride=stride, padding=1, groups=C).backward(dy)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    wd = w.reshape(C, 9).contiguous().to(DEV)
    dx = torch.empty_like(xd)
    call("b2_dwconv3x3_dgrad_nhwc_bf16", dyd.data_ptr(), wd.data_ptr(), dx.data_ptr(), N, H, W, C, stride, stream_ptr())
    assert err(dx.float().permute(0, 3, 1, 2).cpu(), xr.grad) < 6e-3
    dw = torch.zeros(C, 9, device=DEV)
    call("b2_dwconv3x3_wgrad_nhwc_bf16", xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), N, H, W, C, stride, stream_ptr())
    assert err(dw.reshape(C, 1, 3, 3).cpu(), wr.grad) < 1e-4


def test_stem_wgrad_and_relu6_bn_backward_vs_torch():
    from video_classif_b200._lib import call, stream_ptr
    torch.manual_seed(1)
    N, H, W = 3, 36, 44
    x = torch.rand(N, 3, H, W)
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dy = _bf(torch.randn(N, 32, P, Q))
    refw = torch.nn.grad.conv2d_weight(x, (32, 3, 3, 3), dy, stride=2, padding=1)
    xd = x.to(DEV)
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    dw = torch.zeros(32, 27, device=DEV)
    call("b2_mbv2_stem_wgrad", xd.data_ptr(), 0, dyd.data_ptr(), dw.data_ptr(), N, H, W, stream_ptr())
    assert err(dw.reshape(32, 3, 3, 3).cpu(), refw) < 1e-4
    # BatchNorm + ReLU6 backward
    C, M = 48, 3000
    raw = _bf(torch.randn(M, C) * 3 + 1)
    bn = torch.nn.BatchNorm1d(C)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-1, 3)
    r = raw.clone().requires_grad_(True)
    act = torch.nn.functional.relu6(bn(r))
    g = _bf(torch.randn(M, C))
    act.backward(g)
    rd, gd, ad = raw.to(DEV, torch.bfloat16), g.to(DEV, torch.bfloat16), act.detach().to(DEV, torch.bfloat16)
    sums = torch.stack([rd.float().sum(0), (rd.float() ** 2).sum(0)])
    s = torch.zeros(2, C, device=DEV)
    dyo = torch.empty_like(rd)
    gam = bn.weight.detach().to(DEV)
    call("b2_bn_bwd_relu6_nhwc_bf16", gd.data_ptr(), 0, ad.data_ptr(), rd.data_ptr(), dyo.data_ptr(), gam.data_ptr(), sums[0].data_ptr(),
         sums[1].data_ptr(), 0, 0, s[0].data_ptr(), s[1].data_ptr(), M, C, M, 1e-5, 1, stream_ptr())
    assert err(s[0].cpu(), bn.bias.grad) < 2e-3 and err(s[1].cpu(), bn.weight.grad) < 2e-3
    assert err(dyo.float().cpu(), r.grad) < 1.5e-2


def test_mobilenet_v2_trainable_trunk_vs_torch_autograd():
    """Whole trainable trunk (train-mode BN) vs torchvision's own fp32 forward / autograd on the same GPU: features within
    6e-2 of their max (bf16 activations through 53 layers), every parameter gets a finite gradient of the right shape, and the
    median relative gradient error stays below max(0.1, 1.5 x the error of torch's own bf16 autocast run)."""
    import torchvision
    from video_classif_b200.mobilenet import MobileNetRunner
    torch.manual_seed(3)
    net = torchvision.models.mobilenet_v2(weights=None)
    net.classifier = torch.nn.Identity()
    g = torch.Generator().manual_seed(5)
    x = torch.nn.functional.interpolate(torch.rand(16, 3, 16, 16, generator=g), size=64, mode="bilinear")
    x = (x + 0.1 * torch.rand(16, 3, 64, 64, generator=g)).clamp(0, 1).to(DEV)
    wgt = torch.randn(16, 1280, generator=g).to(DEV)
    ref = torchvision.models.mobilenet_v2(weights=None)
    ref.classifier = torch.nn.Identity()
    ref.load_state_dict(net.state_dict())
    ref = ref.to(DEV).train()
    fr = ref(x)
    (fr * wgt).sum().backward()
    auto = torchvision.models.mobilenet_v2(weights=None)
    auto.classifier = torch.nn.Identity()
    auto.load_state_dict(net.state_dict())
    auto = auto.to(DEV).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fa = auto(x)
    (fa.float() * wgt).sum().backward()
    net = net.to(DEV).train()
    runner = MobileNetRunner(net)
    feat = runner(x, True)
    assert feat.requires_grad
    (feat * wgt).sum().backward()
    e_feat, e_auto = err(feat, fr), err(fa.float(), fr)
    print(f"\n[mobilenet_v2 trainable] features: ours {e_feat:.3e}, torch autocast {e_auto:.3e}")
    assert e_feat < max(6e-2, 1.5 * e_auto)
    ours, yard = [], []
    for (k, p), q, a in zip(net.named_parameters(), ref.parameters(), auto.parameters()):
        assert p.grad is not None and p.grad.shape == p.shape and torch.isfinite(p.grad).all(), k
        ours.append(err(p.grad, q.grad, floor=1e-8))
        yard.append(err(a.grad, q.grad, floor=1e-8))
    ours.sort(); yard.sort()
    print(f"    median gradient error: ours {ours[len(ours) // 2]:.3e}, torch autocast {yard[len(yard) // 2]:.3e}")
    assert ours[len(ours) // 2] < max(0.1, 1.5 * yard[len(yard) // 2])
    # running statistics advanced exactly once
    assert int(net.features[0][1].num_batches_tracked) == 1
    assert err(net.features[0][1].running_mean, ref.features[0][1].running_mean) < 1e-2


def test_mobilenet_v2_trainable_trunk_eval_mode_bn_vs_torch_autograd():
    """The same comparison with eval-mode BatchNorm (running statistics = a fixed affine: nothing re-normalises and amplifies
    the bf16 rounding), which pins the whole backward chain numerically: features within 3e-2, median parameter-gradient error
    below max(5e-2, torch autocast's median), and no parameter above 3 x torch autocast's error on it (floor 0.15)."""
    import torchvision
    from video_classif_b200.mobilenet import MobileNetRunner
    torch.manual_seed(4)
    net = torchvision.models.mobilenet_v2(weights=None)
    net.classifier = torch.nn.Identity()
    with torch.no_grad():                      # non-trivial running statistics, activations kept O(1) through the depth
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(-0.1, 0.1)
                m.running_var.uniform_(0.5, 1.0)
                m.weight.uniform_(0.8, 1.2)
    g = torch.Generator().manual_seed(6)
    x = torch.nn.functional.interpolate(torch.rand(8, 3, 16, 16, generator=g), size=64, mode="bilinear")
    x = (x + 0.1 * torch.rand(8, 3, 64, 64, generator=g)).clamp(0, 1).to(DEV)
    wgt = torch.randn(8, 1280, generator=g).to(DEV)
    nets = []
    for _ in range(2):
        r = torchvision.models.mobilenet_v2(weights=None)
        r.classifier = torch.nn.Identity()
        r.load_state_dict(net.state_dict())
        nets.append(r.to(DEV).eval())
    ref, auto = nets
    fr = ref(x)
    (fr * wgt).sum().backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fa = auto(x)
    (fa.float() * wgt).sum().backward()
    net = net.to(DEV).eval()
    feat = MobileNetRunner(net)(x, False)
    (feat * wgt).sum().backward()
    e_feat, e_auto = err(feat, fr), err(fa.float(), fr)
    ours, yard, worst = [], [], 0.0
    for (k, p), q, a in zip(net.named_parameters(), ref.parameters(), auto.parameters()):
        e, ya = err(p.grad, q.grad, floor=1e-8), err(a.grad, q.grad, floor=1e-8)
        ours.append(e)
        yard.append(ya)
        assert e < max(0.15, 3.0 * ya), (k, e, ya)
    ours.sort(); yard.sort()
    print(f"\n[mobilenet_v2 trainable, eval-mode BN] features: ours {e_feat:.3e}, autocast {e_auto:.3e}; median gradient error: "
          f"ours {ours[len(ours) // 2]:.3e}, autocast {yard[len(yard) // 2]:.3e}")
    assert e_feat < 3e-2
    assert ours[len(ours) // 2] < max(5e-2, yard[len(yard) // 2])      # measured 8.1e-2 vs 1.3e-1 for torch autocast


def test_crime_lrcn_mobilenet_v2_finetune_step_runs():
    """CrimeLRCN(cnn_backbone='mobilenet_v2', finetune=True): one optimizer step moves backbone and tail parameters."""
    import video_classif_b200 as vc
    torch.manual_seed(0)
    m = vc.CrimeLRCN(3, 4, 12, 16, cnn_backbone="mobilenet_v2", rnn_layers=2, classif_mode="multiple_binary", finetune=True).to(DEV).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    x = torch.rand(2, 4, 3, 64, 64, device=DEV)
    y = (torch.rand(2, 3, device=DEV) > 0.5).float()
    w0 = m.cnn_backbone.features[3].conv[1][0].weight.detach().clone()
    losses = []
    for _ in range(4):
        opt.zero_grad()
        loss = torch.nn.functional.binary_cross_entropy_with_logits(m(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert not torch.equal(w0, m.cnn_backbone.features[3].conv[1][0].weight)
    assert min(losses[1:]) < losses[0]
