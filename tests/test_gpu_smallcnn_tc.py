"""GPU parity of the small-CNN tensor-core kernels (csrc/smallcnn_tc.cu) through the C ABI, kernel by kernel against
plain torch fp32 on the same bf16-rounded operands, then the whole bf16 model against the reference goldens."""
import numpy as np
import pytest
import torch

from conftest import err, golden_tensors, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lib():
    from video_classif_b200 import _lib
    return _lib


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 64), (64, 32), (32, 16)])
@pytest.mark.parametrize("shape", [(3, 64, 64), (5, 32, 32), (2, 16, 16), (2, 20, 12), (1, 8, 124)])
def test_sc_conv3x3_vs_torch(cin, cout, shape):
    """b2_sc_conv3x3_bf16 (halo tile, one pixel per swizzle row, shifted tap descriptors): output within bf16 rounding
    (4e-3 of max) of torch's fp32 conv2d on the same bf16 operands; statistics = sums of the stored values (1e-3)."""
    from video_classif_b200 import ops
    N, H, W = shape
    torch.manual_seed(cin * 100 + cout + H)
    x = _bf(torch.randn(N, cin, H, W))
    w = _bf(torch.randn(cout, cin, 3, 3) * 0.2)
    b = torch.randn(cout)
    ref = torch.nn.functional.conv2d(x, w, b, padding=1)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    stats = torch.zeros(2, cout, device=DEV)
    y = ops._sc_conv(xd, ops._sc_kernel_weight(w.to(DEV)), cout, b.to(DEV), (stats[0], stats[1]))
    got = y.float().permute(0, 3, 1, 2).cpu()
    assert err(got, ref) < 6e-3
    assert err(stats[0].cpu(), got.sum((0, 2, 3)), floor=1e-3 * got.abs().sum((0, 2, 3)).max().item()) < 2e-3
    assert err(stats[1].cpu(), (got * got).sum((0, 2, 3))) < 2e-3
    # data-gradient form: the same kernel on the flipped / transposed filter == conv_transpose2d
    if (cout, cin) in ((16, 32), (32, 64)):                       # this launch IS the dgrad of the (cout -> cin) forward conv
        wf = _bf(torch.randn(cin, cout, 3, 3) * 0.2)              # forward filter [Cout_f = cin, Cin_f = cout]
        refd = torch.nn.functional.conv_transpose2d(x, wf, padding=1)
        yd = ops._sc_conv(xd, ops._sc_kernel_weight_dgrad(wf.to(DEV)), cout)
        assert err(yd.float().permute(0, 3, 1, 2).cpu(), refd) < 6e-3


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 64)])
@pytest.mark.parametrize("shape", [(3, 64, 64), (7, 32, 32), (2, 16, 16), (2, 20, 12)])
def test_sc_wgrad_vs_torch(cin, cout, shape):
    """b2_sc_conv3x3_wgrad_bf16 (both operands MN-major, pixel-shifted copies of the x tile fill M): fp32 accumulation of
    bf16 products -> 1e-3 of torch's fp32 weight gradient on the same operands."""
    from video_classif_b200._lib import call, stream_ptr
    N, H, W = shape
    torch.manual_seed(cin + cout + W)
    x = _bf(torch.randn(N, cin, H, W))
    dz = _bf(torch.randn(N, cout, H, W))
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), dz, padding=1)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    dzd = dz.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    dw = torch.zeros(cout, 3, 3, cin, device=DEV)
    call("b2_sc_conv3x3_wgrad_bf16", xd.data_ptr(), dzd.data_ptr(), N, H, W, cin, cout, dw.data_ptr(), stream_ptr())
    assert err(dw.permute(0, 3, 1, 2).cpu(), ref) < 1e-3


@pytest.mark.parametrize("shape", [(4, 64, 64), (3, 20, 28), (2, 16, 16)])
def test_sc_conv1_fwd_and_wgrad_vs_torch(shape):
    from video_classif_b200._lib import call, stream_ptr
    N, H, W = shape
    torch.manual_seed(H)
    x = torch.rand(N, 3, H, W) * 255.0                            # raw 0..255 frames (backup_ucf50.py:101)
    w = torch.randn(16, 3, 3, 3) * 0.1
    b = torch.randn(16)
    ref = torch.nn.functional.conv2d(x, w, b, padding=1)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    y = torch.empty(N, H, W, 16, device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2, 16, device=DEV)
    call("b2_sc_conv1_fwd", xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), N, H, W, stats[0].data_ptr(),
         stats[1].data_ptr(), stream_ptr())
    got = y.float().permute(0, 3, 1, 2).cpu()
    assert err(got, ref) < 5e-3
    assert err(stats[1].cpu(), (got * got).sum((0, 2, 3))) < 1e-3
    dz = _bf(torch.randn(N, 16, H, W))
    refw = torch.nn.grad.conv2d_weight(x, (16, 3, 3, 3), dz, padding=1)
    dzd = dz.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    dw = torch.zeros(16, 3, 3, 3, device=DEV)
    call("b2_sc_conv1_wgrad", xd.data_ptr(), dzd.data_ptr(), dw.data_ptr(), N, H, W, stream_ptr())
    assert err(dw.cpu(), refw) < 1e-4


@pytest.mark.parametrize("C,pool,train", [(16, 1, True), (32, 2, True), (64, 2, True), (32, 2, False)])
def test_sc_bn_act_pool_fwd_bwd_vs_torch_autograd(C, pool, train):
    """BN finalisation + BN/ReLU(/max-pool) forward and the two-pass backward against torch autograd in fp32 on the same
    bf16 raw tensor: activations 4e-3 (bf16 output), dgamma / dbeta 1e-3, dz 1e-2 of its max (bf16 output)."""
    from video_classif_b200._lib import call, stream_ptr
    N, H, W = 3, 16, 24
    torch.manual_seed(C + pool)
    raw = _bf(torch.randn(N, C, H, W) * 1.5 + 0.3)
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    rm, rv = torch.randn(C) * 0.1, torch.rand(C) + 0.5
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta); bn.running_mean.copy_(rm); bn.running_var.copy_(rv)
    bn.train(train)
    r = raw.clone().requires_grad_(True)
    yr = torch.relu(bn(r))
    if pool == 2:
        yr = torch.nn.functional.max_pool2d(yr, 2, 2)
    dy = _bf(torch.randn_like(yr))
    (yr * dy).sum().backward()
    st = stream_ptr()
    rawd = raw.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    sums = torch.stack([rawd.float().sum((0, 1, 2)), (rawd.float() ** 2).sum((0, 1, 2))])
    coef = torch.empty(4, C, device=DEV)
    g_d, b_d, rm_d, rv_d = gamma.to(DEV), beta.to(DEV), rm.to(DEV), rv.to(DEV)
    call("b2_sc_bn_finalize", sums[0].data_ptr(), sums[1].data_ptr(), 0, g_d.data_ptr(), b_d.data_ptr(), rm_d.data_ptr(),
         rv_d.data_ptr(), N * H * W, 1e-5, 0.1, int(train), coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(),
         coef[3].data_ptr(), C, st)
    if train:
        assert err(rm_d.cpu(), bn.running_mean) < 1e-4 and err(rv_d.cpu(), bn.running_var) < 1e-4
    y = torch.empty(N, H // pool, W // pool, C, device=DEV, dtype=torch.bfloat16)
    call("b2_sc_act_pool_fwd", rawd.data_ptr(), coef[0].data_ptr(), coef[1].data_ptr(), y.data_ptr(), N, H, W, C, pool, st)
    assert err(y.float().permute(0, 3, 1, 2).cpu(), yr) < 5e-3
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    s = torch.zeros(2, C, device=DEV)
    args = (rawd.data_ptr(), dyd.data_ptr(), coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(), coef[3].data_ptr(),
            s[0].data_ptr(), s[1].data_ptr())
    call("b2_sc_act_pool_bwd_reduce", *args, N, H, W, C, pool, st)
    dz = torch.empty_like(rawd)
    call("b2_sc_act_pool_bwd_apply", *args, int(train), dz.data_ptr(), N, H, W, C, pool, st)
    assert err(s[0].cpu(), bn.bias.grad) < 1e-3 and err(s[1].cpu(), bn.weight.grad) < 1e-3
    assert err(dz.float().permute(0, 3, 1, 2).cpu(), r.grad) < 1e-2


def test_sc_layout_change_round_trip_and_dropout_mask():
    from video_classif_b200._lib import call, stream_ptr
    N, HW, C = 5, 16 * 16, 64
    act = torch.randn(N, HW, C, device=DEV).to(torch.bfloat16)
    feat = torch.empty(N, C * HW, device=DEV, dtype=torch.bfloat16)
    call("b2_sc_nhwc_to_chw", act.data_ptr(), feat.data_ptr(), N, HW, C, 0.0, 0, 0, stream_ptr())
    assert torch.equal(feat.reshape(N, C, HW), act.permute(0, 2, 1))                 # c*HW + p flatten (nb:186)
    call("b2_sc_nhwc_to_chw", act.data_ptr(), feat.data_ptr(), N, HW, C, 0.5, 1234, 0, stream_ptr())
    kept = feat != 0
    assert abs(kept.float().mean().item() - 0.5) < 0.02
    assert torch.allclose(feat[kept].float(), (act.permute(0, 2, 1).reshape(N, -1)[kept].float() * 2).to(torch.bfloat16).float())
    g = torch.ones(N, C * HW, device=DEV)
    back = torch.empty(N, HW, C, device=DEV, dtype=torch.bfloat16)
    call("b2_sc_chw_to_nhwc", g.data_ptr(), 0, back.data_ptr(), N, HW, C, 0.5, 1234, 0, stream_ptr())
    assert torch.equal(back.permute(0, 2, 1).reshape(N, -1) != 0, kept)               # the backward replays the same mask
    gb = torch.randn(N, C * HW, device=DEV).to(torch.bfloat16)
    call("b2_sc_chw_to_nhwc", gb.data_ptr(), 1, back.data_ptr(), N, HW, C, 0.0, 0, 0, stream_ptr())
    assert torch.equal(back.permute(0, 2, 1).reshape(N, -1), gb)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_smallcnn_lrcn_bf16_tensor_core_path_vs_reference_golden(tag):
    """SmallCNNLRCN(precision='bf16') = the tensor-core trunk: one train step vs the notebook class's own fp32 output
    (small fixtures: 12-24 frames per BatchNorm batch).  Tolerances: logits 2e-2 of their max, loss 2e-2, gradients
    3e-1 of their max (12-24 frames per BatchNorm batch: the BASELINE-shape test carries the tight bounds: tests/test_gpu_baseline_shapes.py), running
    statistics 1e-2, argmax equal when the reference's top-2 margin exceeds the logits error."""
    import video_classif_b200 as vc
    g, meta = load_golden(f"smallcnn_lrcn_{tag}.npz")
    m = vc.SmallCNNLRCN(meta["num_classes"], meta["T"], meta["hidden"], (3, meta["size"], meta["size"]), dropout=0.0,
                        precision="bf16")
    m.load_state_dict(golden_tensors(g, "sd0/"))
    m = m.to(DEV).train()
    x, y = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["y"]).to(DEV)
    out = m(x)
    loss = torch.nn.functional.cross_entropy(out, y)
    loss.backward()
    ref = torch.from_numpy(g["logits"])
    e = err(out, ref)
    print(f"\n[smallcnn bf16 {tag}] logits rel err {e:.3e}")
    assert e < 2e-2
    assert abs(loss.item() - float(g["loss"])) < 2e-2
    worst = 0.0
    for k, v in golden_tensors(g, "grad/").items():
        got = dict(m.named_parameters())[k].grad
        assert got is not None and torch.isfinite(got).all(), k
        if k.startswith("conv") and k.endswith(".bias"):
            continue
        ek = err(got, v)
        worst = max(worst, ek)
        assert ek < 3e-1, (k, ek)
    print(f"    worst gradient rel err {worst:.3e}")
    sd1 = m.state_dict()
    for k, v in golden_tensors(g, "sd1/").items():
        if v.dtype.is_floating_point:
            assert err(sd1[k], v) < 1e-2, k
        else:
            assert int(sd1[k]) == int(v), k
    m.eval()
    with torch.no_grad():
        out_e = m(x)
    assert err(out_e, torch.from_numpy(g["logits_eval_after"])) < 3e-2


def test_sc_bias_folded_into_bn_finalize_and_input_pack():
    """The raw conv outputs are stored bias-free; b2_sc_bn_finalize folds the conv bias into the running mean (train) and
    into the effective mean (eval): BN(conv(x) + b) == scale * conv(x) + shift in both modes.  b2_sc_pack_input: NCHW fp32
    frames -> NHWC bf16 with 13 zero channels."""
    from video_classif_b200._lib import call, stream_ptr
    torch.manual_seed(4)
    C, M = 32, 4096
    z = torch.randn(M, C) * 2 + 0.5                       # bias-free conv output, rows = pixels
    b, gamma, beta = torch.randn(C), torch.rand(C) + 0.5, torch.randn(C)
    for train in (True, False):
        rm, rv = torch.randn(C) * 0.3, torch.rand(C) + 0.5
        bn = torch.nn.BatchNorm1d(C)
        with torch.no_grad():
            bn.weight.copy_(gamma); bn.bias.copy_(beta); bn.running_mean.copy_(rm); bn.running_var.copy_(rv)
        bn.train(train)
        ref = bn(z + b)
        zd = z.to(DEV)
        sums = torch.stack([zd.sum(0), (zd * zd).sum(0)])
        coef = torch.empty(4, C, device=DEV)
        d = [t.to(DEV) for t in (b, gamma, beta, rm, rv)]
        call("b2_sc_bn_finalize", sums[0].data_ptr(), sums[1].data_ptr(), d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(),
             d[3].data_ptr(), d[4].data_ptr(), M, 1e-5, 0.1, int(train), coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(),
             coef[3].data_ptr(), C, stream_ptr())
        got = zd * coef[0] + coef[1]
        assert err(got, ref) < 1e-5
        if train:
            assert err(d[3].cpu(), bn.running_mean) < 1e-5 and err(d[4].cpu(), bn.running_var) < 1e-5
    x = torch.rand(3, 3, 20, 12) * 255
    y = torch.empty(3, 20, 12, 16, device=DEV, dtype=torch.bfloat16)
    xd = x.to(DEV)
    call("b2_sc_pack_input", xd.data_ptr(), y.data_ptr(), 3, 20, 12, stream_ptr())
    assert torch.equal(y[..., :3].float().cpu(), x.permute(0, 2, 3, 1).to(torch.bfloat16).float())
    assert (y[..., 3:] == 0).all()
