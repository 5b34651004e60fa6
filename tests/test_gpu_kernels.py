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
def test_conv3x3_halo(ops, N, H, W, with_a):
    """Halo-tile 3x3 conv (nine shifted SWIZZLE_128B descriptors over one shared-memory tile) vs torch conv2d on the
    bf16-rounded operands: output, statistics, in-kernel BatchNorm finalisation; optional input BatchNorm+ReLU with
    zero padding preserved."""
    C = Cout = 64
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
