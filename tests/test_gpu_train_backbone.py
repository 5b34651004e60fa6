"""GPU parity tests of the trainable frame encoder (full / partial fine-tune: rgb_lrcn.py:208-245, lrcn.py:246-283):
backward kernels vs torch autograd on identical inputs, node by node and chained through whole ResNets."""
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, attempts  # noqa: F401
from oracle import lrcn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16 = torch.bfloat16


def rel(a, b, floor=0.0):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / max(b.abs().max().item(), floor, 1e-30)).item()


def rnd(t):
    return t.bfloat16().float()


GEOMS = [(3, 10, 10, 64, 64, 3, 1, 1), (2, 9, 9, 128, 256, 1, 1, 0), (2, 14, 14, 256, 512, 3, 2, 1), (5, 7, 7, 512, 128, 1, 2, 0),
         (4, 28, 28, 64, 256, 1, 1, 0), (2, 8, 8, 2048, 512, 1, 1, 0), (2, 7, 7, 512, 512, 3, 1, 1), (1, 1, 777, 168, 64, 1, 1, 0),
         (1, 5, 5, 64, 8, 3, 1, 1), (2, 9, 9, 128, 64, 3, 1, 1), (2, 12, 12, 64, 128, 3, 2, 1), (3, 6, 6, 128, 256, 3, 2, 1)]


@pytest.mark.parametrize("N,H,W,C,Cout,R,s,p", GEOMS)
def test_conv_wgrad_vs_torch(N, H, W, C, Cout, R, s, p):
    """b2_conv2d_wgrad_nhwc_bf16 (tcgen05, MN-major operands, im2col-TMA taps) vs F.conv2d's weight gradient on the same
    bf16 inputs: fp32 accumulation on both sides -> 1e-5; covers 1x1 / 3x3, stride 1 / 2, narrow layers (C = 64 / 128:
    4 / 2 filter taps side by side in one MMA), Cout < 128 (zero-filled
    panel), C not a multiple of 64 (the stem's patch matrix), pixel counts that are not multiples of the 64-row stage."""
    from video_classif_b200 import backbone_train as BT
    torch.manual_seed(N * 1000 + C)
    x = torch.randn(N, H, W, C, device=DEV).to(BF16)
    P, Q = (H + 2 * p - R) // s + 1, (W + 2 * p - R) // s + 1
    dy = torch.randn(N, P, Q, Cout, device=DEV).to(BF16)
    dw = BT.conv_wgrad(x, dy, R, R, s, p)
    w = torch.zeros(Cout, C, R, R, device=DEV, requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=s, padding=p).backward(dy.float().permute(0, 3, 1, 2))
    assert rel(dw.permute(0, 3, 1, 2), w.grad) < 1e-5


@pytest.mark.parametrize("N,H,W,C,Cout,R,s,p", GEOMS[:4] + [(2, 7, 7, 256, 256, 3, 2, 1)])
def test_conv_dgrad_vs_torch(N, H, W, C, Cout, R, s, p):
    """Data gradient on the forward conv kernels (flipped / transposed filter; stride 2 over the zero-dilated dy; odd and
    even input sizes): bf16 output -> 1e-2."""
    from video_classif_b200 import backbone_train as BT
    torch.manual_seed(N + C)
    w = torch.randn(Cout, C, R, R, device=DEV) * 0.05
    P, Q = (H + 2 * p - R) // s + 1, (W + 2 * p - R) // s + 1
    dy = torch.randn(N, P, Q, Cout, device=DEV).to(BF16)
    dx = BT.conv_dgrad(dy, w, (H, W), s, p)
    x = torch.zeros(N, C, H, W, device=DEV, requires_grad=True)
    F.conv2d(x, rnd(w), stride=s, padding=p).backward(dy.float().permute(0, 3, 1, 2))
    assert rel(dx.float().permute(0, 3, 1, 2), x.grad) < 1e-2


NODES = [(6, 8, 8, 64, 64, 3, 1, 1, True, False, True), (6, 8, 8, 64, 256, 1, 1, 0, True, True, True),
         (6, 8, 8, 128, 128, 3, 2, 1, True, False, True), (6, 8, 8, 256, 512, 1, 2, 0, False, False, True),
         (6, 4, 4, 512, 2048, 1, 1, 0, True, True, True), (6, 8, 8, 64, 256, 1, 1, 0, True, True, False),
         (3, 7, 7, 192, 64, 1, 1, 0, True, False, True)]


@pytest.mark.parametrize("N,H,W,C,Cout,R,s,p,relu,has_res,train", NODES)
def test_conv_bn_node_vs_torch(N, H, W, C, Cout, R, s, p, relu, has_res, train):
    """One conv -> BatchNorm (train / eval) -> (+shortcut) -> ReLU autograd node: output, running statistics and all five
    gradients (input, filter, gamma, beta, shortcut) vs torch autograd of the same graph with bf16 storage at the same
    points."""
    from video_classif_b200 import backbone_train as BT
    torch.manual_seed(C + Cout)
    conv = torch.nn.Conv2d(C, Cout, R, s, p, bias=False).to(DEV)
    bn = torch.nn.BatchNorm2d(Cout).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
        bn.running_mean.normal_(0, 0.1)
        bn.running_var.uniform_(0.5, 2.0)
    bn_ref = torch.nn.BatchNorm2d(Cout).to(DEV)
    bn_ref.load_state_dict(bn.state_dict())
    bn_ref.train(train)
    P = (H + 2 * p - R) // s + 1
    x = torch.randn(N, H, W, C, device=DEV).to(BF16).requires_grad_(True)
    res = torch.randn(N, P, P, Cout, device=DEV).to(BF16).requires_grad_(True) if has_res else None
    G = torch.randn(N, P, P, Cout, device=DEV).to(BF16)
    z = BT.ConvBnFn.apply(x, conv.weight, bn.weight, bn.bias, res, bn, s, p, relu, train)
    (z.float() * G.float()).sum().backward()
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    rr = res.detach().float().permute(0, 3, 1, 2).requires_grad_(True) if has_res else None
    o = bn_ref(rnd(F.conv2d(xr, rnd(wr), stride=s, padding=p)))
    if has_res:
        o = o + rr
    zr = rnd(torch.relu(o) if relu else o)
    (zr * G.float().permute(0, 3, 1, 2)).sum().backward()
    assert rel(z.float().permute(0, 3, 1, 2), zr) < 1e-2
    assert rel(x.grad.float().permute(0, 3, 1, 2), xr.grad) < 1e-2
    assert rel(conv.weight.grad, wr.grad) < 1e-2
    assert rel(bn.weight.grad, bn_ref.weight.grad) < 5e-3
    assert rel(bn.bias.grad, bn_ref.bias.grad) < 5e-3
    if has_res:
        assert rel(res.grad.float().permute(0, 3, 1, 2), rr.grad) < 1e-6       # the masked incoming gradient, exactly
    assert rel(bn.running_mean, bn_ref.running_mean) < 1e-3 and rel(bn.running_var, bn_ref.running_var) < 1e-3
    assert int(bn.num_batches_tracked) == int(train)


def test_stem_node_vs_torch():
    """conv1 7x7/2 -> bn1 -> ReLU -> maxpool 3x3/2 node: filter / gamma / beta gradients (max-pool routing to the first
    maximal element, ReLU mask, BatchNorm backward, weight gradient through the patch matrix)."""
    from video_classif_b200 import backbone_train as BT
    torch.manual_seed(11)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).to(DEV)
    bn, bn_ref = torch.nn.BatchNorm2d(64).to(DEV), torch.nn.BatchNorm2d(64).to(DEV)
    x = torch.rand(6, 3, 40, 32, device=DEV)
    y = BT.StemFn.apply(x, conv.weight, bn.weight, bn.bias, bn, True)
    G = torch.randn_like(y.float()).to(BF16)
    (y.float() * G.float()).sum().backward()
    wr = conv.weight.detach().clone().requires_grad_(True)
    yr = rnd(F.max_pool2d(torch.relu(bn_ref(rnd(F.conv2d(rnd(x), rnd(wr), stride=2, padding=3)))), 3, 2, 1))
    (yr * G.float().permute(0, 3, 1, 2)).sum().backward()
    assert rel(y.float().permute(0, 3, 1, 2), yr) < 1e-2
    assert rel(conv.weight.grad, wr.grad) < 2e-2
    assert rel(bn.weight.grad, bn_ref.weight.grad) < 1e-2 and rel(bn.bias.grad, bn_ref.bias.grad) < 1e-2


@pytest.mark.parametrize("N,H,W", [(3, 20, 16), (2, 7, 9), (1, 56, 56), (2, 5, 4)])
def test_maxpool_relu_backward_vs_torch_with_ties(N, H, W):
    """Backward of the stem tail (b2_maxpool_relu_bwd_nhwc) vs torch autograd through max_pool2d(relu(raw * scale + shift), 3, 2, 1).
    The raw values take few distinct levels so that every window has TIES (first-maximal-element rule), half of the scales are
    negative (argmax of the raw values flips to argmin), odd sizes leave ragged windows."""
    from video_classif_b200._lib import call, stream_ptr
    C = 64
    g = torch.Generator().manual_seed(N * H + W)
    raw = (torch.randint(-3, 4, (N, H, W, C), generator=g).float() * 0.5).to(BF16).to(DEV)
    scale = (torch.rand(C, generator=g) + 0.5) * torch.where(torch.rand(C, generator=g) > 0.5, 1.0, -1.0)
    shift = torch.randn(C, generator=g) * 0.3
    P, Q = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    dpool = torch.randn(N, P, Q, C, generator=g).to(BF16).to(DEV)
    scale, shift = scale.to(DEV), shift.to(DEV)
    dbn = torch.zeros((N, H, W, C), device=DEV, dtype=torch.float32)
    call("b2_maxpool_relu_bwd_nhwc", raw.data_ptr(), scale.data_ptr(), shift.data_ptr(), dpool.data_ptr(), dbn.data_ptr(),
         N, H, W, P, Q, C, stream_ptr())
    a = (raw.float() * scale + shift).permute(0, 3, 1, 2).detach().requires_grad_(True)
    F.max_pool2d(torch.relu(a), 3, 2, 1).backward(dpool.float().permute(0, 3, 1, 2))
    # torch routes ties among equal ACTIVATIONS; ours among equal raw values: identical whenever scale != 0 (monotone map),
    # except inside the ReLU's flat region, where the gradient is masked to zero by both
    assert rel(dbn.permute(0, 3, 1, 2), a.grad) < 1e-5


def _teacher_forced_reference(net, x, rec, G, start):
    """torch autograd over torchvision's graph where every node's VALUE is replaced by the activation the B200 path
    produced (rec, execution order) while its local Jacobian stays torch's: parameter gradients then differ from ours by
    arithmetic only (same ReLU masks, same batch statistics) -- a train-mode batch-statistics ResNet is chaotic, so
    end-to-end comparisons of two bf16 executions measure the chaos, not the kernels."""
    it = iter(rec)

    def sub(v, relu=False):
        ours = next(it).float().permute(0, 3, 1, 2)
        assert ours.shape == v.shape
        if relu:            # the ReLU mask is teacher-forced too (with 16-sample batch statistics one flipped unit is visible)
            v = v * (ours > 0).float()
        return ours.detach() + (v - v.detach())

    def node(inp, conv, bn, relu, res=None):
        o = F.batch_norm(rnd(F.conv2d(inp, rnd(conv.weight), stride=conv.stride, padding=conv.padding)), None, None, bn.weight,
                         bn.bias, True, 0.0, bn.eps)
        if res is not None:
            o = o + res
        return sub(o, relu)

    if start == ("stem",):
        a = F.batch_norm(rnd(F.conv2d(rnd(x), rnd(net.conv1.weight), stride=2, padding=3)), None, None, net.bn1.weight,
                         net.bn1.bias, True, 0.0, net.bn1.eps)
        y = sub(F.max_pool2d(torch.relu(a), 3, 2, 1))
        start = (1, 0)
    else:
        y = next(it).float().permute(0, 3, 1, 2)
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(net, f"layer{li}")):
            if (li, bi) < start:
                continue
            short = y if blk.downsample is None else node(y, blk.downsample[0], blk.downsample[1], False)
            o = node(y, blk.conv1, blk.bn1, True)
            if hasattr(blk, "conv3"):
                o = node(o, blk.conv2, blk.bn2, True)
                y = node(o, blk.conv3, blk.bn3, True, res=short)
            else:
                y = node(o, blk.conv2, blk.bn2, True, res=short)
    feat = y.mean(dim=(2, 3))
    (feat * G).sum().backward()
    return feat


@pytest.mark.parametrize("arch,first,frames,size", [("resnet18", "conv1", 8, 64), ("resnet18", "layer3", 8, 64),
                                                    ("resnet50", "conv1", 8, 64), ("resnet50", "layer2.1.bn2", 6, 96),
                                                    ("resnet34", "layer4", 4, 64)])
@attempts(3)
def test_finetune_gradients_whole_network(arch, first, frames, size):
    """Whole ResNets, parameters trainable from `first` on in named_parameters() order (freeze_until_layer semantics,
    lrcn.py:275-283; 'conv1' = full fine-tune): the frozen prefix runs on the fused kernels, the rest through the autograd
    nodes; every trainable parameter's gradient vs the teacher-forced torch reference; frozen parameters get none."""
    import torchvision
    import video_classif_b200 as vc
    from video_classif_b200 import backbone_train as BT
    torch.manual_seed(7)
    net, feat_dim = vc.backbone.make_backbone(arch)
    on = False
    for n, p in net.named_parameters():
        on = on or n.startswith(first)
        p.requires_grad_(on)
    net = net.to(DEV).train()
    x = torch.rand(frames, 3, size, size, device=DEV)
    G = torch.randn(frames, feat_dim, device=DEV)
    BT._record = rec = []
    try:
        feat = vc.backbone.ResNetRunner(net)(x, True)
    finally:
        BT._record = None
    (feat * G).sum().backward()
    got = {n: p.grad.clone() for n, p in net.named_parameters() if p.requires_grad}
    assert all(p.grad is None for p in net.parameters() if not p.requires_grad)
    assert set(got) == {n for n, p in net.named_parameters() if p.requires_grad} and got
    for p in net.parameters():
        p.grad = None
    ref_feat = _teacher_forced_reference(net, x, rec, G, BT.first_trainable_block(net))
    assert rel(feat, ref_feat) < 2e-2
    errs = sorted(((rel(got[n], p.grad, floor=1e-6), n) for n, p in net.named_parameters() if p.requires_grad), reverse=True)
    # typical: worst 1e-2, median 5e-3.  The bound on the worst parameter leaves room for what is NOT teacher-forced: the
    # max-pool routing of the stem (torch routes on its own conv output, a last-bit difference moves a gradient to the
    # neighbouring pixel) and the order of the fp32 atomics behind the batch statistics; one run in ~12 exceeded 5e-2 on the
    # deepest case (ResNet-50, 32 samples per channel in layer4).  Kernel-level exactness is pinned by the node tests above.
    assert errs[0][0] < 1.5e-1, errs[:5]
    assert errs[len(errs) // 2][0] < 3e-2, errs[len(errs) // 2]
    assert sum(e > 5e-2 for e, _ in errs) <= max(2, len(errs) // 50), errs[:8]


def test_crime_lrcn_partial_freeze_trains():
    """lrcn/lrcn.py `LRCN(..., freeze_until_layer=k)`: one optimizer step through the partially trainable backbone --
    gradients exactly on the parameters past the freeze boundary, loss decreases over a few steps on a fixed batch."""
    import video_classif_b200 as vc
    torch.manual_seed(2)
    m = vc.CrimeLRCN(3, 4, 16, 32, cnn_backbone="resnet18", freeze_until_layer=44, rnn_layers=1, classif_mode="multiclass",
                     precision="fp32").to(DEV).train()
    names = [n for n, _ in m.cnn_backbone.named_parameters()]
    trainable = {n for n, p in m.cnn_backbone.named_parameters() if p.requires_grad}
    assert trainable == set(names[45:]) and trainable
    x = torch.rand(4, 4, 3, 64, 64, device=DEV)
    y = torch.tensor([0, 1, 2, 1], device=DEV)
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-3)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = F.cross_entropy(m(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(p.grad is not None for n, p in m.cnn_backbone.named_parameters() if n in trainable)
    assert all(p.grad is None for n, p in m.cnn_backbone.named_parameters() if n not in trainable)
    assert min(losses[1:]) < losses[0], losses          # (a few Adam steps on a fixed batch: the loss goes down)


def _densenet_teacher_forced(net, y0, saved, G):
    """torch autograd over torchvision DenseNet's trunk with every produced tensor's VALUE replaced by ours (block buffers,
    bottleneck tensors) and the ReLU masks of the bottleneck taken from ours; see _teacher_forced_reference."""
    def sub(v, ours_nhwc, relu=False):
        ours = ours_nhwc.float().permute(0, 3, 1, 2)
        assert ours.shape == v.shape, (ours.shape, v.shape)
        if relu:
            v = v * (ours > 0).float()
        return ours.detach() + (v - v.detach())

    def bn(t, m):
        return F.batch_norm(t, None, None, m.weight, m.bias, True, 0.0, m.eps)

    f = net.features
    feats = y0
    pos = [0]

    def nxt_entry():
        pos[0] += 1
        return saved[pos[0] - 1]

    for name, mod in f.named_children():
        if name.startswith("denseblock"):
            _, X, S, C0, growth, rec = nxt_entry()
            N, H, W, Cfin = X.shape
            for k, (lname, layer) in enumerate(mod.named_children()):
                _, _, y1, a2, _ = rec[k]
                Ct = C0 + k * growth
                a1 = rnd(torch.relu(bn(feats, layer.norm1)))
                y1_t = sub(rnd(F.conv2d(a1, rnd(layer.conv1.weight))), y1.view(N, H, W, -1))
                a2_t = sub(bn(y1_t, layer.norm2), a2.view(N, H, W, -1), relu=True)
                y2_t = sub(rnd(F.conv2d(a2_t, rnd(layer.conv2.weight), padding=1)), X[..., Ct:Ct + growth])
                feats = torch.cat([feats, y2_t], dim=1)
        elif name.startswith("transition"):
            _, _, _, X, S, Hc, Wc, C, Cn = nxt_entry()
            a = rnd(torch.relu(bn(feats, mod.norm)))
            pooled = F.avg_pool2d(rnd(F.conv2d(a, rnd(mod.conv.weight))), 2, 2)
            feats = sub(pooled, saved[pos[0]][1][..., :Cn])          # the next block's buffer starts with the pooled features
        elif name == "norm5":
            feats = torch.relu(bn(feats, mod))
    feat = feats.mean(dim=(2, 3))
    (feat * G).sum().backward()
    return feat


@attempts(3)
def test_densenet_finetune_gradients_teacher_forced():
    """Trainable DenseNet (growth 32, blocks (2,2,2,2): every layer type): features and the gradient of EVERY parameter
    behind the stem, plus the gradient handed to the stem, vs the teacher-forced torch reference."""
    import torchvision
    from video_classif_b200 import densenet_train as DT
    from video_classif_b200.densenet import DenseNetRunner
    torch.manual_seed(5)
    net = torchvision.models.DenseNet(32, (2, 2, 2, 2), 64)
    net.classifier = torch.nn.Identity()
    net = net.to(DEV).train()
    runner = DenseNetRunner(net)
    x = torch.rand(8, 3, 64, 64, device=DEV)
    G = torch.randn(8, net.features.norm5.num_features, device=DEV)
    with torch.no_grad():
        y0 = runner.stem(x, True)
    y0 = y0.requires_grad_(True)
    names, params = DT._trunk_params(net)
    DT._record = rec = []
    try:
        feat = DT.DenseTrunkFn.apply(y0, runner, True, names, *params)
    finally:
        DT._record = None
    (feat * G).sum().backward()
    got = {n: p.grad.clone() for n, p in zip(names, params)}
    dy0 = y0.grad.clone()
    assert all(torch.isfinite(v).all() for v in got.values())
    for p in params:
        p.grad = None
    y0r = y0.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    ref_feat = _densenet_teacher_forced(net, y0r, rec[0], G)
    assert rel(feat, ref_feat) < 2e-2
    errs = sorted(((rel(got[n], p.grad, floor=1e-6), n) for n, p in zip(names, params)), reverse=True)
    assert errs[0][0] < 1.5e-1, errs[:6]                          # (same bounds and reasoning as the ResNet chains above)
    assert errs[len(errs) // 2][0] < 3e-2, errs[len(errs) // 2]
    assert sum(e > 5e-2 for e, _ in errs) <= max(2, len(errs) // 50), errs[:8]
    assert rel(dy0.float().permute(0, 3, 1, 2), y0r.grad) < 5e-2


def test_crime_lrcn_default_densenet121_finetune_trains():
    """The reference's default crime configuration (densenet121, FINETUNE = True -> nothing frozen, lrcn.py:27,34,230):
    whole-model training steps; every backbone parameter receives a finite gradient and the loss goes down."""
    import video_classif_b200 as vc
    torch.manual_seed(3)
    m = vc.CrimeLRCN(3, 2, 8, 16, cnn_backbone="densenet121", finetune=True, rnn_layers=1).to(DEV).train()
    assert all(p.requires_grad for p in m.cnn_backbone.parameters())
    x = torch.rand(4, 2, 3, 64, 64, device=DEV)
    y = torch.tensor([[1., 0., 0.], [0., 1., 0.], [0., 0., 1.], [1., 0., 0.]], device=DEV)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = F.binary_cross_entropy_with_logits(m(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.cnn_backbone.parameters())
    assert int(m.cnn_backbone.features.denseblock2.denselayer5.norm2.num_batches_tracked) == 5
    assert min(losses[1:]) < losses[0], losses          # (a few Adam steps on a fixed batch: the loss goes down)


def test_crime_lrcn_densenet_partial_freeze():
    """`freeze_until_layer=k` on the DenseNet backbone (lrcn.py:275-283): gradients exactly on the parameters past the
    boundary (the boundary may fall inside a dense layer), none before it; the frozen stem gets no gradient work."""
    import video_classif_b200 as vc
    torch.manual_seed(4)
    m = vc.CrimeLRCN(3, 2, 8, 16, cnn_backbone="densenet121", freeze_until_layer=200, rnn_layers=1, classif_mode="multiclass",
                     precision="fp32").to(DEV).train()
    names = [n for n, _ in m.cnn_backbone.named_parameters()]
    trainable = {n for n, p in m.cnn_backbone.named_parameters() if p.requires_grad}
    assert trainable == set(names[201:]) and 0 < len(trainable) < len(names)
    x = torch.rand(2, 2, 3, 64, 64, device=DEV)
    loss = F.cross_entropy(m(x), torch.tensor([0, 2], device=DEV))
    loss.backward()
    for n, p in m.cnn_backbone.named_parameters():
        if n in trainable:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
        else:
            assert p.grad is None, n
