"""GPU parity tests, kernel level: every b2_* kernel family through the C ABI vs the CPU oracle /
the reference-generated golden vectors.  Bit-exact for byte/index work; stated tolerances for
floating point (fp32 kernels 1e-4, bf16 tensor-core kernels 1e-2 of max|ref|)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, err, golden_tensors, load_golden
from oracle import lrcn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    import video_classif_b200 as vc
    return vc.ops


@pytest.mark.parametrize("M,N,K,bias,out_bf16,relu", [
    (128, 32, 64, False, True, False), (300, 200, 136, True, True, True), (1000, 8, 512, True, False, False),
    (160, 128, 16384, True, False, False), (1920, 224, 112, True, False, False), (4096, 50, 640, True, False, False),
    (777, 1024, 2048, True, True, True), (5, 4, 8, True, False, False)])
def test_gemm_tcgen05(ops, M, N, K, bias, out_bf16, relu):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=DEV) if bias else None
    s1, s2 = torch.zeros(N, device=DEV), torch.zeros(N, device=DEV)
    D = ops.gemm_tn(A, B, bias=b, out_dtype=torch.bfloat16 if out_bf16 else torch.float32, relu=relu, stats=(s1, s2))
    ref = A.float().cpu() @ B.float().cpu().t()          # bf16-rounded operands, fp32 CPU product
    if bias:
        ref = ref + b.cpu()
    if relu:
        ref = ref.relu()
    assert err(D.float(), ref) < (6e-3 if out_bf16 else 2e-5)
    r = ref.bfloat16().float() if out_bf16 else ref
    assert err(s1, r.sum(0), floor=1e-3 * r.abs().sum(0).max().item()) < 2e-2
    assert err(s2, (r * r).sum(0)) < 2e-2


@pytest.mark.parametrize("N,H,W,C,Cout,R,stride,pad", [
    (2, 8, 8, 64, 64, 1, 1, 0), (2, 8, 8, 64, 64, 3, 1, 1), (3, 14, 14, 128, 128, 3, 1, 1),
    (3, 28, 28, 128, 128, 3, 2, 1), (3, 7, 7, 256, 512, 1, 2, 0), (5, 7, 7, 512, 512, 3, 2, 1),
    (1, 5, 9, 64, 192, 3, 1, 1), (130, 4, 4, 64, 64, 3, 1, 1),
    # large enough for the two-CTA cluster path (weight tile multicast): odd and even row-block counts
    (40, 28, 28, 256, 256, 1, 1, 0), (64, 14, 14, 128, 512, 3, 1, 1), (97, 7, 7, 512, 512, 3, 1, 1)])
def test_conv_implicit_gemm_tma_im2col(ops, N, H, W, C, Cout, R, stride, pad):
    torch.manual_seed(N * H + C)
    x = torch.randn(N, C, H, W).bfloat16()
    w = (torch.randn(Cout, C, R, R) / (C * R * R) ** 0.5).bfloat16()
    s1, s2 = torch.zeros(Cout, device=DEV), torch.zeros(Cout, device=DEV)
    y = ops.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous().to(DEV), w.permute(0, 2, 3, 1).contiguous().to(DEV), stride,
                        pad, stats=(s1, s2))
    ref = F.conv2d(x.float(), w.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert err(y.float(), ref) < 6e-3
    r = ref.bfloat16().float().reshape(-1, Cout)
    assert err(s2, (r * r).sum(0)) < 2e-2


@pytest.mark.parametrize("N,H,W", [(2, 32, 32), (3, 112, 112), (2, 64, 96), (1, 224, 224), (5, 17, 23), (300, 16, 16)])
def test_stem_conv_direct_toeplitz(ops, N, H, W):
    """ResNet conv1 (7x7/2 pad 3) as the Toeplitz-descriptor tcgen05 kernel vs torch fp32 conv on the
    bf16-rounded operands; also the fused per-channel statistics."""
    torch.manual_seed(N + H + W)
    x = torch.rand(N, 3, H, W)
    w = (torch.randn(64, 3, 7, 7) / 147 ** 0.5)
    s1, s2 = torch.zeros(64, device=DEV), torch.zeros(64, device=DEV)
    y = ops.stem_conv(x.to(DEV), ops.pack_stem_weight(w.to(DEV)), stats=(s1, s2))
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), stride=2, padding=3).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert err(y.float(), ref) < 6e-3
    r = ref.bfloat16().float().reshape(-1, 64)
    assert err(s1, r.sum(0), floor=1e-3 * r.abs().sum(0).max().item()) < 2e-2
    assert err(s2, (r * r).sum(0)) < 2e-2
    yb = ops.stem_conv(x.to(DEV).bfloat16(), ops.pack_stem_weight(w.to(DEV)))      # bf16 input frames
    assert torch.equal(yb, y)


@pytest.mark.parametrize("N,H,W,C,Cout,R,stride,pad", [
    (3, 9, 9, 64, 64, 3, 1, 1), (2, 14, 14, 128, 256, 1, 1, 0), (5, 12, 10, 64, 128, 3, 2, 1), (70, 4, 4, 256, 64, 3, 1, 1),
    (4, 7, 7, 128, 512, 1, 2, 0), (40, 28, 28, 256, 256, 1, 1, 0), (150, 7, 7, 256, 1024, 1, 1, 0),
    (66, 14, 14, 128, 256, 3, 1, 1)])
def test_conv_bn_folded(ops, N, H, W, C, Cout, R, stride, pad):
    """b2_conv2d_bn_nhwc_bf16: input BatchNorm+ReLU applied to the A tile in shared memory (padding stays 0),
    statistics + finalisation tail, statistics-only pass, and the BN3 + shortcut(+BN) + ReLU epilogue."""
    torch.manual_seed(N * H + C + Cout)
    xr = torch.randn(N, C, H, W).bfloat16()                         # raw output of the previous conv
    a_sc, a_sh = torch.rand(C) + 0.5, torch.randn(C) * 0.3
    w = (torch.randn(Cout, C, R, R) / (C * R * R) ** 0.5).bfloat16()
    xin = torch.relu(xr.float() * a_sc.view(1, -1, 1, 1) + a_sh.view(1, -1, 1, 1)).bfloat16().float()
    raw_ref = F.conv2d(xin, w.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)        # [N,P,Q,Cout] fp32
    xg = xr.permute(0, 2, 3, 1).contiguous().to(DEV)
    wg = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    a = (a_sc.to(DEV), a_sh.to(DEV))
    # (1) raw output + statistics + finalisation
    gamma, beta = (torch.rand(Cout) + 0.5).to(DEV), torch.randn(Cout).to(DEV)
    rm, rv = torch.zeros(Cout, device=DEV), torch.ones(Cout, device=DEV)
    buf = torch.zeros(4 * Cout + 4, device=DEV)
    s1, s2, fs, fh, cnt = buf[:Cout], buf[Cout:2 * Cout], buf[2 * Cout:3 * Cout], buf[3 * Cout:4 * Cout], buf[4 * Cout:]
    y = ops.conv2d_bn_nhwc(xg, wg, stride, pad, a=a, stats=(s1, s2), fin=(gamma, beta, rm, rv, fs, fh, cnt, 1e-5, 0.1))
    assert err(y.float(), raw_ref) < 6e-3
    r = raw_ref.bfloat16().float().reshape(-1, Cout)
    cntf = r.shape[0]
    mean, var = r.mean(0), r.var(0, unbiased=False)
    assert err(s1, r.sum(0), floor=1e-3 * r.abs().sum(0).max().item()) < 2e-2
    assert err(s2, (r * r).sum(0)) < 2e-2
    sc_ref = gamma.cpu() / torch.sqrt(var + 1e-5)
    assert err(fs, sc_ref) < 2e-2 and err(fh, beta.cpu() - mean * sc_ref, floor=1.0) < 2e-2
    assert err(rm, 0.1 * mean, floor=1e-2) < 2e-2 and err(rv, 0.9 + 0.1 * var * cntf / (cntf - 1)) < 2e-2
    # (2) statistics-only pass gives the same sums and writes nothing
    buf2 = torch.zeros(4 * Cout + 4, device=DEV)
    none = ops.conv2d_bn_nhwc(xg, wg, stride, pad, a=a, stats=(buf2[:Cout], buf2[Cout:2 * Cout]),
                              fin=(gamma, beta, None, None, buf2[2 * Cout:3 * Cout], buf2[3 * Cout:4 * Cout],
                                   buf2[4 * Cout:], 1e-5, 0.1), store=False)
    assert none is None
    assert torch.allclose(buf2[:2 * Cout], buf[:2 * Cout], rtol=1e-5, atol=1e-3)
    assert torch.allclose(buf2[2 * Cout:4 * Cout], buf[2 * Cout:4 * Cout], rtol=1e-4, atol=1e-4)
    # (3) output pass: BN(out) + shortcut (identity / raw + own BN) + ReLU on the fp32 accumulators
    res = torch.randn(raw_ref.shape).bfloat16()
    r_sc, r_sh = torch.rand(Cout) + 0.5, torch.randn(Cout) * 0.3
    o = (fs, fh)
    out1 = ops.conv2d_bn_nhwc(xg, wg, stride, pad, a=a, o=o, res=res.to(DEV), relu=True)
    ref1 = torch.relu(raw_ref * fs.cpu() + fh.cpu() + res.float())
    assert err(out1.float(), ref1) < 8e-3
    out2 = ops.conv2d_bn_nhwc(xg, wg, stride, pad, a=a, o=o, res=res.to(DEV), r=(r_sc.to(DEV), r_sh.to(DEV)), relu=True)
    ref2 = torch.relu(raw_ref * fs.cpu() + fh.cpu() + res.float() * r_sc + r_sh)
    assert err(out2.float(), ref2) < 8e-3
    out3 = ops.conv2d_bn_nhwc(xg, wg, stride, pad, a=a, o=o, relu=False)
    assert err(out3.float(), raw_ref * fs.cpu() + fh.cpu()) < 8e-3
    # (4) stand-alone scale/shift apply (BasicBlock output)
    out4 = ops.scale_shift_apply(y.clone(), fs, fh, res=res.to(DEV), r=(r_sc.to(DEV), r_sh.to(DEV)), relu=True)
    ref4 = torch.relu(y.float().cpu() * fs.cpu() + fh.cpu() + res.float() * r_sc + r_sh)
    assert err(out4.float(), ref4) < 8e-3


@pytest.mark.parametrize("N,H,W", [(2, 28, 28), (3, 14, 14), (5, 7, 7), (2, 9, 13), (1, 4, 4), (40, 28, 28), (3, 56, 56),
                                   (2, 5, 60)])
@pytest.mark.parametrize("with_a", [False, True])
@pytest.mark.parametrize("C", [64, 128])
def test_conv3x3_halo(ops, N, H, W, with_a, C):
    """Halo-tile 3x3 conv (nine shifted SWIZZLE_128B descriptors over one shared-memory tile) vs torch conv2d on the
    bf16-rounded operands: output, statistics, in-kernel BatchNorm finalisation; optional input BatchNorm+ReLU with
    zero padding preserved.  C = 64: resident weights, one 128-pixel tile per stage; C = 128: streamed weight ring, two
    MMA tiles per halo stage (the wide cases that do not fit shared memory must report unsupported)."""
    from video_classif_b200._lib import lib
    Cout = C
    if C == 128 and not lib().b2_conv3x3_halo_supported(N, H, W, C, Cout):
        assert W + 2 > 32                        # only the wide maps are expected to fall back to the im2col kernel
        pytest.skip("halo stage does not fit shared memory at this width")
    torch.manual_seed(N * H + W)
    xr = torch.randn(N, C, H, W).bfloat16()
    w = (torch.randn(Cout, C, 3, 3) / (C * 9) ** 0.5).bfloat16()
    a_sc, a_sh = torch.rand(C) + 0.5, torch.randn(C) * 0.3
    xin = torch.relu(xr.float() * a_sc.view(1, -1, 1, 1) + a_sh.view(1, -1, 1, 1)).bfloat16().float() if with_a else xr.float()
    ref = F.conv2d(xin, w.float(), stride=1, padding=1).permute(0, 2, 3, 1)
    xg = xr.permute(0, 2, 3, 1).contiguous().to(DEV)
    wg = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    assert ops.conv3x3_halo_supported(xg, wg, 1, 1)
    gamma, beta = (torch.rand(Cout) + 0.5).to(DEV), torch.randn(Cout).to(DEV)
    rm, rv = torch.zeros(Cout, device=DEV), torch.ones(Cout, device=DEV)
    buf = torch.zeros(4 * Cout + 4, device=DEV)
    s1, s2, fs, fh, cnt = buf[:Cout], buf[Cout:2 * Cout], buf[2 * Cout:3 * Cout], buf[3 * Cout:4 * Cout], buf[4 * Cout:]
    y = ops.conv3x3_halo_bn(xg, wg, a=(a_sc.to(DEV), a_sh.to(DEV)) if with_a else None, stats=(s1, s2),
                            fin=(gamma, beta, rm, rv, fs, fh, cnt, 1e-5, 0.1))
    assert y.shape == ref.shape
    assert err(y.float(), ref) < 6e-3
    r = ref.bfloat16().float().reshape(-1, Cout)
    assert err(s1, r.sum(0), floor=1e-3 * r.abs().sum(0).max().item()) < 2e-2
    assert err(s2, (r * r).sum(0)) < 2e-2
    mean, var = r.mean(0), r.var(0, unbiased=False)
    sc_ref = gamma.cpu() / torch.sqrt(var + 1e-5)
    assert err(fs, sc_ref) < 2e-2 and err(fh, beta.cpu() - mean * sc_ref, floor=1.0) < 2e-2
    y2 = ops.conv2d_bn_nhwc(xg, wg, 1, 1, a=(a_sc.to(DEV), a_sh.to(DEV)) if with_a else None)     # im2col-TMA kernel
    assert err(y.float(), y2.float()) < 6e-3


@pytest.mark.parametrize("M,C,Cout", [(128, 64, 256), (256, 128, 512), (1000, 64, 64), (5000, 128, 256), (70000, 64, 256),
                                      (33333, 128, 512), (128, 256, 1024), (50176, 256, 1024), (777, 256, 64)])
def test_gram_bn_statistics(ops, M, C, Cout):
    """b2_conv1x1_gram_bnstats_bf16: sum / sum-of-squares / BN scale+shift of relu(x*a+b) @ W^T from the Gram
    matrix (MN-major tcgen05 MMA) vs the directly computed fp64 statistics of the same product."""
    torch.manual_seed(M + C)
    xr = (torch.randn(M, C) * 1.5).bfloat16()
    a_sc, a_sh = torch.rand(C) + 0.5, torch.randn(C) * 0.3 + 0.5
    w = (torch.randn(Cout, C) / C ** 0.5).bfloat16()
    t = torch.relu(xr.float() * a_sc + a_sh).bfloat16().double()
    y = t @ w.double().t()
    gamma, beta = (torch.rand(Cout) + 0.5).to(DEV), torch.randn(Cout).to(DEV)
    rm, rv = torch.zeros(Cout, device=DEV), torch.ones(Cout, device=DEV)
    s1, s2 = torch.full((Cout,), 7.0, device=DEV), torch.full((Cout,), 7.0, device=DEV)
    fs, fh = torch.empty(Cout, device=DEV), torch.empty(Cout, device=DEV)
    ops.conv1x1_gram_bnstats(xr.to(DEV), w.to(DEV), (a_sc.to(DEV), a_sh.to(DEV)),
                             (gamma, beta, rm, rv, fs, fh, None, 1e-5, 0.1), stats=(s1, s2))
    mean, var = y.mean(0), y.var(0, unbiased=False)
    assert err(s1, y.sum(0), floor=1e-3 * y.abs().sum(0).max().item()) < 1e-3
    assert err(s2, (y * y).sum(0)) < 1e-3
    sc_ref = gamma.cpu().double() / torch.sqrt(var + 1e-5)
    assert err(fs, sc_ref) < 2e-3
    assert err(fh, beta.cpu().double() - mean * sc_ref, floor=1.0) < 2e-3
    assert err(rm, 0.1 * mean, floor=1e-2) < 2e-3 and err(rv, 0.9 + 0.1 * var * M / max(M - 1, 1)) < 2e-3


def test_ingest_bit_exact_vs_cv2_golden(ops):
    g = np.load(os.path.join(GOLDEN, "resize_cv2.npz"))
    i = 0
    while f"src{i}" in g.files:
        src, rgb = g[f"src{i}"], g[f"rgb{i}"]
        out = ops.ingest_u8(torch.from_numpy(src[None]).to(DEV), rgb.shape[0], rgb.shape[1], swap_rb=True, divisor=1.0)
        got = out[0].permute(1, 2, 0).cpu().numpy()
        assert np.array_equal(got, rgb.astype(np.float32)), f"case {i}"     # uint8 result identical to cv2
        i += 1


@pytest.mark.parametrize("h,w,oh,ow", [(360, 640, 112, 112), (37, 53, 112, 112), (128, 200, 64, 100), (224, 224, 112, 112),
                                       (64, 64, 64, 64), (9, 7, 24, 24), (480, 854, 224, 224)])
def test_ingest_vs_oracle(ops, h, w, oh, ow):
    rng = np.random.default_rng(h * w)
    clip = rng.integers(0, 256, (5, h, w, 3), dtype=np.uint8)
    idx = [3, 0, -1, 4, 4, 1]
    want = O.ingest_clip(clip[[max(i, 0) for i in idx]], oh, ow, swap_rb=True, divisor=255.0)
    want[2] = 0.0
    got = ops.ingest_u8(torch.from_numpy(clip).to(DEV), oh, ow, frame_index=torch.tensor(idx, dtype=torch.int32, device=DEV))
    assert np.array_equal(got.cpu().numpy(), want)                              # fp32: bit-exact
    got16 = ops.ingest_u8(torch.from_numpy(clip).to(DEV), oh, ow, out_dtype=torch.bfloat16, swap_rb=False, divisor=1.0)
    want16 = torch.from_numpy(O.ingest_clip(clip, oh, ow, swap_rb=False, divisor=1.0)).bfloat16()
    assert torch.equal(got16.cpu(), want16)


@pytest.mark.parametrize("tag", ["uni", "bi"])
@pytest.mark.parametrize("bf16", [False])
def test_lstm_vs_reference_golden(ops, tag, bf16):
    g, meta = load_golden(f"lstm_{tag}.npz")
    rnn = torch.nn.LSTM(meta["inp"], meta["H"], num_layers=meta["layers"], bidirectional=meta["bidir"], batch_first=True)
    rnn.load_state_dict(golden_tensors(g, "p/"))
    rnn = rnn.to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    out = ops.lstm_forward(x, rnn, bf16=bf16)
    (out * torch.from_numpy(g["w"]).to(DEV)).sum().backward()
    assert err(out, torch.from_numpy(g["out"])) < 1e-5
    assert err(x.grad, torch.from_numpy(g["dx"])) < 1e-4
    for k, v in golden_tensors(g, "g/").items():
        assert err(getattr(rnn, k).grad, v) < 1e-4, k


@pytest.mark.parametrize("B,T,In,H,layers,bidir,stack", [
    (5, 7, 24, 32, 2, False, True), (5, 7, 24, 32, 2, False, False), (3, 6, 40, 56, 2, True, True),
    (9, 30, 8, 32, 3, False, True), (9, 30, 8, 32, 3, False, False), (4, 40, 64, 56, 4, False, True),
    (70, 16, 8, 32, 3, False, True), (2, 1, 3, 5, 1, False, True), (3, 64, 33, 64, 2, False, True)])
def test_lstm_vs_oracle(ops, B, T, In, H, layers, bidir, stack, monkeypatch):
    """stack=True: narrow unidirectional stacks run as ONE persistent launch per pass (lstm_stack.cu);
    stack=False forces the per-layer kernels + hoisted gate GEMM on the same shapes."""
    monkeypatch.setattr(ops, "LSTM_STACK", stack)
    torch.manual_seed(B * T)
    rnn = torch.nn.LSTM(In, H, num_layers=layers, bidirectional=bidir, batch_first=True)
    x = torch.randn(B, T, In)
    p = {"lstm." + k: v.detach().clone().requires_grad_(True) for k, v in rnn.named_parameters()}
    xo = x.clone().requires_grad_(True)
    ref = O.lstm_forward(xo, p, H, layers, bidir)
    wgt = torch.randn_like(ref)
    (ref * wgt).sum().backward()
    rnn = rnn.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    out = ops.lstm_forward(xg, rnn)
    (out * wgt.to(DEV)).sum().backward()
    assert err(out, ref) < 1e-5
    assert err(xg.grad, xo.grad) < 1e-4
    for k, v in rnn.named_parameters():
        assert err(v.grad, p["lstm." + k].grad) < 1e-4, k
    with torch.no_grad():                                     # inference: nothing saved for BPTT
        assert err(ops.lstm_forward(x.to(DEV), rnn), ref) < 1e-5


@pytest.mark.parametrize("tag", ["uni", "bi"])
def test_gru_vs_reference_golden(ops, tag):
    g, meta = load_golden(f"gru_{tag}.npz")
    rnn = torch.nn.GRU(meta["inp"], meta["H"], num_layers=meta["layers"], bidirectional=meta["bidir"], batch_first=True)
    rnn.load_state_dict(golden_tensors(g, "p/"))
    rnn = rnn.to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    out = ops.gru_forward(x, rnn)
    (out * torch.from_numpy(g["w"]).to(DEV)).sum().backward()
    assert err(out, torch.from_numpy(g["out"])) < 1e-5
    assert err(x.grad, torch.from_numpy(g["dx"])) < 1e-4
    for k, v in golden_tensors(g, "g/").items():
        assert err(getattr(rnn, k).grad, v) < 1e-4, k


@pytest.mark.parametrize("B,T,In,H,layers,bidir", [(5, 7, 24, 32, 2, False), (3, 6, 40, 56, 2, True), (9, 30, 8, 32, 3, False),
                                                  (2, 1, 3, 5, 1, True), (70, 16, 1024, 32, 1, True), (6, 20, 64, 64, 1, False)])
def test_gru_vs_oracle(ops, B, T, In, H, layers, bidir):
    torch.manual_seed(B * T + 1)
    rnn = torch.nn.GRU(In, H, num_layers=layers, bidirectional=bidir, batch_first=True)
    x = torch.randn(B, T, In) * (0.3 if In > 256 else 1.0)
    p = {"lstm." + k: v.detach().clone().requires_grad_(True) for k, v in rnn.named_parameters()}
    xo = x.clone().requires_grad_(True)
    ref = O.gru_forward(xo, p, H, layers, bidir)
    wgt = torch.randn_like(ref)
    (ref * wgt).sum().backward()
    rnn = rnn.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    out = ops.gru_forward(xg, rnn)
    (out * wgt.to(DEV)).sum().backward()
    assert err(out, ref) < 1e-5
    assert err(xg.grad, xo.grad) < 1e-4
    for k, v in rnn.named_parameters():
        assert err(v.grad, p["lstm." + k].grad) < 1e-4, k
    with torch.no_grad():
        assert err(ops.gru_forward(x.to(DEV), rnn), ref) < 1e-5


@pytest.mark.parametrize("M,N,gelu", [(37, 8, True), (64, 1024, True), (10, 960, False), (1920, 512, True)])
def test_act_layernorm_fwd_bwd(ops, M, N, gelu):
    torch.manual_seed(N)
    pre = torch.randn(M, N) * 1.5
    gam, bet = torch.rand(N) + 0.5, torch.randn(N) * 0.1
    w = torch.randn(M, N)
    po, go, bo = (t.clone().requires_grad_(True) for t in (pre, gam, bet))
    ref = O.layernorm(O.gelu_exact(po) if gelu else po, go, bo)
    (ref * w).sum().backward()
    pg, gg, bg = (t.to(DEV).requires_grad_(True) for t in (pre, gam, bet))
    out = ops.act_layernorm(pg, gg, bg, gelu, 1e-5)
    (out * w.to(DEV)).sum().backward()
    assert err(out, ref) < 1e-5
    assert err(pg.grad, po.grad) < 1e-4
    assert err(gg.grad, go.grad) < 1e-4 and err(bg.grad, bo.grad) < 1e-4


@pytest.mark.parametrize("M,K,N,bf16", [(33, 70, 19, False), (160, 640, 50, False), (1920, 2048, 1024, True), (640, 512, 8, True)])
def test_linear_fwd_bwd(ops, M, K, N, bf16):
    torch.manual_seed(K)
    x, w, b = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N)
    g = torch.randn(M, N)
    xo, wo, bo = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = xo @ wo.t() + bo
    (ref * g).sum().backward()
    xg, wg, bg = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    out = ops.linear(xg, wg, bg, bf16)
    (out * g.to(DEV)).sum().backward()
    tol = 1e-2 if bf16 else 1e-4            # bf16 operands / fp32 accumulate vs fp32
    assert err(out, ref) < tol
    assert err(xg.grad, xo.grad) < tol and err(wg.grad, wo.grad) < tol and err(bg.grad, bo.grad) < 1e-4


@pytest.mark.parametrize("N,Cin,Cout,H,W,pool", [(3, 3, 16, 16, 16, False), (4, 16, 32, 32, 32, True), (2, 32, 64, 16, 16, True),
                                                 (2, 5, 7, 10, 22, False)])
def test_conv_bn_relu_pool_fwd_bwd(ops, N, Cin, Cout, H, W, pool):
    torch.manual_seed(Cin * Cout)
    conv = torch.nn.Conv2d(Cin, Cout, 3, padding=1)
    bn = torch.nn.BatchNorm2d(Cout)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.3, 0.3)
        bn.running_mean.uniform_(-1, 1)
        bn.running_var.uniform_(0.5, 2)
    x = torch.randn(N, Cin, H, W) * 2
    xo = x.clone().requires_grad_(True)
    cw, cb, gam, bet = (t.detach().clone().requires_grad_(True) for t in (conv.weight, conv.bias, bn.weight, bn.bias))
    z = F.conv2d(xo, cw, cb, padding=1)
    yb, rm, rv = O.batchnorm2d_train(z, gam, bet, bn.running_mean.clone(), bn.running_var.clone())
    ref = torch.relu(yb)
    if pool:
        ref = F.max_pool2d(ref, 2, 2)
    wgt = torch.randn_like(ref)
    (ref * wgt).sum().backward()
    conv, bn = conv.to(DEV), bn.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    out = ops.conv_bn_relu_pool(xg, conv, bn, pool, True)
    (out * wgt.to(DEV)).sum().backward()
    assert err(out, ref) < 1e-5
    assert err(bn.running_mean, rm) < 1e-5 and err(bn.running_var, rv) < 1e-5
    assert int(bn.num_batches_tracked) == 1
    gmax = max(v.grad.abs().max().item() for v in (cw, gam, bet))
    assert err(xg.grad, xo.grad) < 1e-4
    assert err(conv.weight.grad, cw.grad) < 1e-4
    assert conv.bias.grad.abs().max().item() < 1e-4 * gmax                # analytically zero under train-mode BN
    assert err(bn.weight.grad, gam.grad) < 1e-4 and err(bn.bias.grad, bet.grad) < 1e-4
    # eval mode uses the running statistics
    out_e = ops.conv_bn_relu_pool(x.to(DEV), conv, bn, pool, False)
    ze = F.conv2d(x, conv.weight.cpu(), conv.bias.cpu(), padding=1)
    ref_e = torch.relu(O.batchnorm2d_eval(ze, bn.weight.cpu(), bn.bias.cpu(), bn.running_mean.cpu(), bn.running_var.cpu()))
    if pool:
        ref_e = F.max_pool2d(ref_e, 2, 2)
    assert err(out_e, ref_e.detach()) < 1e-5


def test_dropout_rate_scale_and_backward_mask(ops):
    x = torch.ones(1 << 20, device=DEV, requires_grad=True)
    for p in (0.25, 0.5):
        y = ops.dropout(x, p, True)
        kept = (y != 0).float().mean().item()
        assert abs(kept - (1 - p)) < 5e-3                                      # Bernoulli keep rate
        assert torch.allclose(y[y != 0], torch.tensor(1 / (1 - p), device=DEV))
        (gx,) = torch.autograd.grad(y.sum(), x)
        assert torch.equal(gx != 0, y != 0)                                    # same mask replayed
    assert ops.dropout(x, 0.5, False) is x and ops.dropout(x, 0.0, True) is x
    y1, y2 = ops.dropout(x, 0.5, True), ops.dropout(x, 0.5, True)
    assert not torch.equal(y1, y2)


def test_selective_scan_vs_reference_golden(ops):
    """Fused scan vs the outputs of the reference's own parallel_scan implementations (tests/golden/scan.npz):
    videomamba.py (state reset every 256 steps), medsos forward and 'backward' (u/delta flipped, B/C not)."""
    g = np.load(os.path.join(GOLDEN, "scan.npz"))
    args = [torch.from_numpy(g[k]).to(DEV) for k in ("u", "delta", "A", "B", "C")]
    assert err(ops.selective_scan(*args, chunk_reset=256), torch.from_numpy(g["y_videomamba"])) < 1e-4
    assert err(ops.selective_scan(*args, chunk_reset=None), torch.from_numpy(g["y_medsos_fwd"])) < 1e-4
    assert err(ops.selective_scan(*args, chunk_reset=None, reverse=True), torch.from_numpy(g["y_medsos_bwd"])) < 1e-4


@pytest.mark.parametrize("B,L,D,N,chunk,reverse", [(2, 700, 256, 16, 256, False), (3, 37, 100, 8, None, True),
                                                   (1, 1, 130, 16, 256, False), (2, 513, 2048, 16, 256, False),
                                                   (2, 16, 2048, 16, None, True), (1, 300, 64, 32, 100, False)])
def test_selective_scan_vs_oracle(ops, B, L, D, N, chunk, reverse):
    """Ragged channel counts, single-step sequences, several chunks, the BASELINE config-5 widths (D=2048, N=16)."""
    gen = torch.Generator().manual_seed(B * L + D)
    u = torch.randn(B, L, D, generator=gen)
    delta = F.softplus(torch.randn(B, L, D, generator=gen))
    A = -torch.exp(torch.randn(D, N, generator=gen))
    Bm, Cm = torch.randn(B, L, N, generator=gen), torch.randn(B, L, N, generator=gen)
    ref = O.selective_scan(u, delta, A, Bm, Cm, chunk_reset=chunk, reverse=reverse)
    got = ops.selective_scan(u.to(DEV), delta.to(DEV), A.to(DEV), Bm.to(DEV), Cm.to(DEV), chunk_reset=chunk, reverse=reverse)
    assert got.shape == ref.shape
    assert err(got, ref) < 1e-4


def test_errors_are_loud(ops):
    import video_classif_b200 as vc
    with pytest.raises(vc.B200LrcnError):
        ops.sgemm(torch.randn(4, 4), torch.randn(4, 4))                        # CPU tensors: no fallback
    with pytest.raises(vc.B200LrcnError):
        ops.conv2d_nhwc(torch.zeros(1, 4, 4, 48, device=DEV, dtype=torch.bfloat16),
                        torch.zeros(64, 3, 3, 48, device=DEV, dtype=torch.bfloat16), 1, 1)   # C % 64 != 0


# ---------------------------------------------------------------------------------------------------------------
# DenseNet / MobileNetV2 element kernels and the strided BatchNorm backward, one by one against torch
# ---------------------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("C,ld,act", [(96, 256, 1), (64, 64, 2), (200, 512, 0), (32, 40, -1)])
def test_scale_shift_apply_ld(C, ld, act):
    """Row-strided BN apply / slice copy (act -1: scale = NULL) between channel slices of concatenated buffers."""
    from video_classif_b200._lib import call, stream_ptr
    torch.manual_seed(C)
    M = 777
    X = torch.randn(M, ld, device=DEV).to(torch.bfloat16)
    Y = torch.zeros(M, ld + 8, device=DEV, dtype=torch.bfloat16)
    sc, sh = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    call("b2_scale_shift_apply_ld_bf16", X.data_ptr(), ld, Y.data_ptr() + 16, ld + 8, M, C, sc.data_ptr() if act >= 0 else 0,
         sh.data_ptr() if act >= 0 else 0, max(act, 0), stream_ptr())
    ref = X[:, :C].float()
    if act >= 0:
        ref = ref * sc + sh
        ref = ref.clamp(0, 6) if act == 2 else (ref.clamp_min(0) if act == 1 else ref)
    if act < 0:
        assert torch.equal(Y[:, 8:8 + C], X[:, :C])                              # slice copy: bit exact
    else:
        assert _rel(Y[:, 8:8 + C].float(), ref) < 8e-3                           # fused multiply-add, one bf16 rounding
    assert Y[:, :8].abs().sum() == 0 and Y[:, 8 + C:].abs().sum() == 0          # nothing outside the slice is touched


@pytest.mark.parametrize("C,ld", [(64, 256), (96, 96), (1024, 1024), (24, 32)])
def test_colstats_ld(C, ld):
    from video_classif_b200._lib import call, stream_ptr
    torch.manual_seed(ld)
    M = 5000
    X = (torch.randn(M, ld, device=DEV) * 2 + 0.5).to(torch.bfloat16)
    s = torch.zeros(2, C, device=DEV)
    call("b2_colstats_ld_bf16", X.data_ptr(), ld, M, C, s[0].data_ptr(), s[1].data_ptr(), stream_ptr())
    x = X[:, :C].double()
    assert _rel(s[0], x.sum(0)) < 1e-5 and _rel(s[1], (x * x).sum(0)) < 1e-5


@pytest.mark.parametrize("H,W", [(8, 8), (7, 7), (14, 9)])
def test_avgpool2x2_fwd_bwd(H, W):
    """AvgPool2d(2, 2) into a row-strided destination and its backward (odd sizes: the last row / column is dropped)."""
    from video_classif_b200._lib import call, stream_ptr
    torch.manual_seed(H * W)
    N, C, ldy = 3, 40, 72
    x = torch.randn(N, H, W, C, device=DEV).to(torch.bfloat16)
    P, Q = H // 2, W // 2
    y = torch.zeros(N * P * Q, ldy, device=DEV, dtype=torch.bfloat16)
    call("b2_avgpool2x2_nhwc_bf16", x.data_ptr(), y.data_ptr(), ldy, N, H, W, C, stream_ptr())
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = torch.nn.functional.avg_pool2d(xr, 2, 2)
    assert _rel(y[:, :C].float().view(N, P, Q, C).permute(0, 3, 1, 2), ref) < 1e-2
    assert y[:, C:].abs().max() == 0
    g = torch.randn(N * P * Q, ldy, device=DEV).to(torch.bfloat16)
    dx = torch.empty(N, H, W, C, device=DEV, dtype=torch.bfloat16)
    call("b2_avgpool2x2_bwd_nhwc_bf16", g.data_ptr(), ldy, dx.data_ptr(), N, H, W, C, stream_ptr())
    ref.backward(g[:, :C].float().view(N, P, Q, C).permute(0, 3, 1, 2))
    assert _rel(dx.float().permute(0, 3, 1, 2), xr.grad) < 1e-2


@pytest.mark.parametrize("C,ldx,train,relu,acc", [(96, 256, 1, 1, 1), (64, 64, 1, 0, 0), (200, 512, 0, 1, 1), (1024, 1024, 1, 1, 0)])
def test_bn_bwd_ld(C, ldx, train, relu, acc):
    """Strided BatchNorm (+ReLU mask) backward, written or accumulated into a channel slice of a gradient buffer, vs torch
    autograd of relu(batch_norm(x)) on the same values."""
    from video_classif_b200._lib import call, stream_ptr
    torch.manual_seed(C + acc)
    M = 1500
    X = (torch.randn(M, ldx, device=DEV) * 1.5 + 0.3).to(torch.bfloat16)
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.2
    rm, rv = torch.randn(C, device=DEV) * 0.1, torch.rand(C, device=DEV) + 0.5
    xr = X[:, :C].float().requires_grad_(True)
    out = torch.nn.functional.batch_norm(xr, None if train else rm, None if train else rv, gamma, beta, bool(train), 0.0, 1e-5)
    z = (out.clamp_min(0) if relu else out)
    dz = torch.randn(M, C, device=DEV).to(torch.bfloat16)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    out2 = torch.nn.functional.batch_norm(xr, None if train else rm, None if train else rv, gr, br, bool(train), 0.0, 1e-5)
    (out2.clamp_min(0) if relu else out2).backward(dz.float())
    s = torch.stack([X[:, :C].float().sum(0), (X[:, :C].float() ** 2).sum(0)]).contiguous()
    dX0 = torch.randn(M, ldx, device=DEV).to(torch.bfloat16)
    dX = dX0.clone()
    s12 = torch.zeros(2, C, device=DEV)
    zb = z.detach().to(torch.bfloat16).contiguous()
    call("b2_bn_bwd_ld_bf16", dz.data_ptr(), C, zb.data_ptr() if relu else 0, C, X.data_ptr(), ldx, dX.data_ptr(), ldx, acc,
         gamma.data_ptr(), s[0].data_ptr(), s[1].data_ptr(), rm.data_ptr(), rv.data_ptr(), s12[0].data_ptr(), s12[1].data_ptr(),
         M, C, M, 1e-5, train, stream_ptr())
    want = xr.grad + (dX0[:, :C].float() if acc else 0)
    assert _rel(dX[:, :C].float(), want) < 2e-2
    assert torch.equal(dX[:, C:], dX0[:, C:])
    assert _rel(s12[0], br.grad) < 1e-3 and _rel(s12[1], gr.grad) < 5e-3


@pytest.mark.parametrize("C,H,W,stride,act,N", [(32, 9, 9, 1, 2, 3), (96, 12, 10, 2, 2, 3), (144, 7, 7, 1, 1, 3), (960, 4, 4, 1, 2, 3),
                                                 (24, 5, 6, 2, 0, 3), (96, 56, 56, 2, 2, 3), (144, 28, 28, 1, 2, 2), (32, 56, 56, 1, 0, 2),
                                                 (192, 14, 14, 1, 2, 5), (384, 7, 7, 1, 2, 37), (576, 7, 7, 2, 1, 37), (960, 4, 4, 1, 2, 70),
                                                 (96, 13, 11, 2, 1, 4), (48, 30, 17, 1, 0, 3), (144, 28, 28, 2, 2, 3), (32, 112, 112, 1, 2, 1)])
def test_dwconv3x3_bn(C, H, W, stride, act, N):
    """Depthwise 3x3 with the previous BatchNorm + ReLU / ReLU6 on load (padding stays zero), output statistics, at the
    MobileNetV2 layer shapes of a 112x112 frame (32 @56, 96 @56 s2, 144 @28 s1 / s2, 192 @14, 384 / 576 @7, 960 @4), odd
    sizes, a channel count that is not a multiple of 16 (24) and batch sizes that leave ragged block ranges."""
    from video_classif_b200._lib import call, stream_ptr
    torch.manual_seed(C + H)
    x = torch.randn(N, H, W, C, device=DEV).to(torch.bfloat16)
    sc, sh = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    w = torch.randn(C, 1, 3, 3, device=DEV) * 0.3
    P, Q = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    y = torch.empty(N, P, Q, C, device=DEV, dtype=torch.bfloat16)
    s = torch.zeros(2, C, device=DEV)
    call("b2_dwconv3x3_bn_nhwc_bf16", x.data_ptr(), sc.data_ptr() if act else 0, sh.data_ptr() if act else 0, act,
         w.reshape(C, 9).contiguous().data_ptr(), y.data_ptr(), s[0].data_ptr(), s[1].data_ptr(), N, H, W, C, stride, stream_ptr())
    a = x.float()
    if act:
        a = a * sc + sh
        a = a.clamp(0, 6) if act == 2 else a.clamp_min(0)
        a = a.to(torch.bfloat16).float()
    ref = torch.nn.functional.conv2d(a.permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), stride=stride, padding=1, groups=C)
    assert _rel(y.float().permute(0, 3, 1, 2), ref) < 1e-2
    yf = y.float().reshape(-1, C)
    assert _rel(s[0], yf.sum(0)) < 1e-4 and _rel(s[1], (yf * yf).sum(0)) < 1e-4


def test_mbv2_stem_conv():
    from video_classif_b200._lib import call, stream_ptr
    torch.manual_seed(8)
    x = torch.rand(5, 3, 31, 40, device=DEV)
    w = torch.randn(32, 3, 3, 3, device=DEV) * 0.2
    P, Q = 16, 20
    y = torch.empty(5, P, Q, 32, device=DEV, dtype=torch.bfloat16)
    s = torch.zeros(2, 32, device=DEV)
    call("b2_mbv2_stem_conv", x.data_ptr(), 0, w.contiguous().data_ptr(), y.data_ptr(), s[0].data_ptr(), s[1].data_ptr(), 5, 31, 40,
         stream_ptr())
    ref = torch.nn.functional.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), stride=2, padding=1)
    assert _rel(y.float().permute(0, 3, 1, 2), ref) < 1e-2
    yf = y.float().reshape(-1, 32)
    assert _rel(s[0], yf.sum(0)) < 1e-4 and _rel(s[1], (yf * yf).sum(0)) < 1e-4


@pytest.mark.parametrize("M,K,N,lda", [(1000, 96, 128, 256), (4096, 224, 128, 512), (300, 512, 256, 512)])
def test_gemm_bn_fold(M, K, N, lda):
    """1x1 conv with the pre-activation BatchNorm + ReLU folded into the A-tile transform over a row-strided operand (K not a
    multiple of the 64-channel k-block: zero-padded coefficients) + output statistics."""
    from video_classif_b200._lib import call, stream_ptr
    torch.manual_seed(M)
    X = torch.randn(M, lda, device=DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) * 0.1).to(torch.bfloat16)
    Kp = (K + 63) // 64 * 64
    ss = torch.zeros(2, Kp, device=DEV)
    ss[0, :K] = torch.rand(K, device=DEV) + 0.5
    ss[1, :K] = torch.randn(K, device=DEV) * 0.5
    D = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    st = torch.zeros(2, N, device=DEV)
    call("b2_gemm_bn_bf16_tn", X.data_ptr(), lda, W.data_ptr(), K, D.data_ptr(), N, M, N, K, ss[0].data_ptr(), ss[1].data_ptr(), 1,
         st[0].data_ptr(), st[1].data_ptr(), stream_ptr())
    a = (X[:, :K].float() * ss[0, :K] + ss[1, :K]).clamp_min(0).to(torch.bfloat16).float()
    ref = a @ W.float().t()
    assert _rel(D.float(), ref) < 1e-2
    df = D.float()
    assert _rel(st[0], df.sum(0)) < 1e-3 and _rel(st[1], (df * df).sum(0)) < 1e-3


@pytest.mark.parametrize("N,H,W", [(3, 28, 28), (5, 14, 14), (4, 7, 7), (2, 3, 3), (2, 9, 20), (1024, 3, 3), (600, 2, 2)])
def test_conv3x3_halo_dense(N, H, W):
    """DenseNet dense-layer conv (128 -> 32, 3x3, pad 1) on the halo-tile kernel: output written into a channel slice of a
    wider buffer (the rest of the buffer untouched), output statistics; vs F.conv2d on the same bf16 operands.  The many-tile
    tiny-map cases exercise the shifted descriptors' reach past a small halo panel (shared-memory tail padding)."""
    from video_classif_b200._lib import call, lib, stream_ptr
    torch.manual_seed(H * W)
    assert lib().b2_conv3x3_halo_dense_supported(N, H, W, 128, 32) == 1
    x = torch.randn(N, H, W, 128, device=DEV).to(torch.bfloat16)
    w = (torch.randn(32, 128, 3, 3, device=DEV) * 0.05)
    wk = w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    ld = 256
    X0 = torch.randn(N * H * W, ld, device=DEV).to(torch.bfloat16)
    X = X0.clone()
    s = torch.zeros(2, 32, device=DEV)
    call("b2_conv3x3_halo_dense_bf16", x.data_ptr(), N, H, W, 128, wk.data_ptr(), 32, X.data_ptr() + 2 * 96, ld, s[0].data_ptr(),
         s[1].data_ptr(), stream_ptr())
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    got = X[:, 96:128].float().view(N, H, W, 32)
    assert _rel(got, ref) < 1e-2
    assert torch.equal(X[:, :96], X0[:, :96]) and torch.equal(X[:, 128:], X0[:, 128:])
    gf = X[:, 96:128].float()
    assert _rel(s[0], gf.sum(0)) < 1e-3 and _rel(s[1], (gf * gf).sum(0)) < 1e-3


@pytest.mark.parametrize("chunk,reverse", [(256, False), (None, False), (None, True), (64, False)])
def test_selective_scan_backward_vs_oracle_autograd(ops, chunk, reverse):
    """ops.selective_scan is differentiable: gradients w.r.t. u, delta, A, B, C vs torch autograd through the oracle's scan
    (videomamba chunk-reset variant with its chunks in parallel, medsos forward / reversed variants) on the golden inputs."""
    g = np.load(os.path.join(GOLDEN, "scan.npz"))
    names = ("u", "delta", "A", "B", "C")
    cpu = [torch.from_numpy(g[k]).clone().requires_grad_(True) for k in names]
    gpu = [torch.from_numpy(g[k]).to(DEV).requires_grad_(True) for k in names]
    w = torch.randn(cpu[0].shape, generator=torch.Generator().manual_seed(3))
    (O.selective_scan(*cpu, chunk_reset=chunk, reverse=reverse) * w).sum().backward()
    y = ops.selective_scan(*gpu, chunk_reset=chunk, reverse=reverse)
    (y * w.to(DEV)).sum().backward()
    for k, a, b in zip(names, gpu, cpu):
        assert err(a.grad, b.grad) < 2e-3, k


@pytest.mark.parametrize("B,L,D,N,chunk,reverse", [(3, 40, 64, 16, 16, False), (2, 33, 96, 8, None, True), (2, 20, 160, 32, None, False),
                                                   (2, 37, 128, 4, None, False), (2, 70, 128, 8, 32, False), (1, 25, 16, 64, None, True),
                                                   (2, 19, 64, 12, None, False), (1, 530, 64, 16, None, False), (2, 300, 64, 16, 256, False)])
def test_selective_scan_backward_wide(ops, B, L, D, N, chunk, reverse):
    """Both backward kernels at every compiled state width: the on-chip one (N in {4, 8, 16, 32, 64}, D a multiple of 512 / N
    channels, chunks of up to 512 steps, ragged last segment / last chunk) and the workspace one (padded N = 12, D = 96 with
    N = 8, a 530-step scan without reset)."""
    g = torch.Generator().manual_seed(L + D)
    u = torch.randn(B, L, D, generator=g)
    delta = F.softplus(torch.randn(B, L, D, generator=g))
    A = -torch.exp(torch.randn(D, N, generator=g) * 0.5)
    Bm, Cm = torch.randn(B, L, N, generator=g), torch.randn(B, L, N, generator=g)
    cpu = [t.clone().requires_grad_(True) for t in (u, delta, A, Bm, Cm)]
    gpu = [t.to(DEV).requires_grad_(True) for t in (u, delta, A, Bm, Cm)]
    w = torch.randn(B, L, D, generator=g)
    (O.selective_scan(*cpu, chunk_reset=chunk, reverse=reverse) * w).sum().backward()
    y = ops.selective_scan(*gpu, chunk_reset=chunk, reverse=reverse)
    (y * w.to(DEV)).sum().backward()
    for k, a, b in zip(("u", "delta", "A", "B", "C"), gpu, cpu):
        assert err(a.grad, b.grad) < 2e-3, k
