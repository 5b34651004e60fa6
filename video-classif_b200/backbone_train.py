"""Trainable frame encoder: the torchvision ResNet with (some of) its parameters left trainable -- full fine-tune
(`lrcn/rgb_lrcn.py:208-245`, `CONF_FINETUNE`) and partial freezes (`freeze_cnn_layers(freeze_until_layer)`,
`lrcn/lrcn.py:246-283`: the first k entries of `named_parameters()` frozen).

The frozen PREFIX of the network (every residual block in front of the first one that owns a trainable parameter) runs on
the fused inference-style kernels of `backbone.ResNetRunner` (nothing is saved).  From the first trainable block on,
every conv -> BatchNorm -> (+shortcut) -> ReLU group is one autograd node (`ConvBnFn`) that keeps its input, the raw
conv output and the activation (bf16, NHWC) and differentiates with the kernels of `csrc/conv_bwd.cu`:

    BatchNorm (+ ReLU mask)   b2_bn_bwd_nhwc_bf16         (per-channel reductions + one elementwise pass)
    weight gradient           b2_conv2d_wgrad_nhwc_bf16   (tcgen05, MN-major operands: no transposes, no im2col matrix)
    data gradient             the forward conv kernels on the flipped / transposed filter (stride 2: zero-dilated dy)

Activations and their gradients are bf16, parameter gradients fp32 (the usual mixed-precision contract; tests compare
with the fp32 reference at bf16 tolerances)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .ops import BF16, F32, cast_bf16, conv2d_bn_nhwc, conv2d_nhwc, gemm_tn, scale_shift_apply

STEM_KP = 168      # 3 channels x 7 rows x 8 taps (one zero tap): the patch-matrix stem of backbone.py


def _momentum(bn):
    return bn.momentum if bn.momentum is not None else 0.1


def _finalize(stat, C, bn, count, train, momentum=None):
    """scale/shift (stat[2C:3C], stat[3C:4C]) from the batch statistics stat[0:C], stat[C:2C] (train) or the running ones."""
    base = stat.data_ptr()
    call("b2_bn_finalize_nhwc", base if train else 0, base + 4 * C if train else 0, bn.weight.data_ptr(), bn.bias.data_ptr(),
         bn.running_mean.data_ptr(), bn.running_var.data_ptr(), count, float(bn.eps),
         float(_momentum(bn) if momentum is None else momentum), int(train), base + 8 * C, base + 12 * C, C, stream_ptr())


def _bn_backward(dz, z, y, bn, stat, count, train, want_masked=False):
    """-> (dy bf16, dz masked by z > 0 if want_masked (else None), dgamma, dbeta).  z = None: no ReLU behind the BN."""
    C = y.shape[-1]
    M = y.numel() // C
    dy = torch.empty_like(y)
    dzm = torch.empty_like(y) if (want_masked and z is not None) else None
    s12 = torch.zeros(2 * C, device=y.device, dtype=F32)
    base = stat.data_ptr()
    call("b2_bn_bwd_nhwc_bf16", dz.data_ptr(), ptr(dzm), ptr(z), y.data_ptr(), dy.data_ptr(), bn.weight.data_ptr(),
         base if train else 0, base + 4 * C if train else 0, bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
         s12.data_ptr(), s12.data_ptr() + 4 * C, M, C, count, float(bn.eps), int(train), stream_ptr())
    if want_masked and z is None:
        dzm = dz
    return dy, dzm, s12[C:], s12[:C]


def conv_wgrad(x, dy, R, S, stride, pad):
    """x [N,H,W,C] bf16, dy [N,P,Q,Cout] bf16 -> dW [Cout,R,S,C] fp32."""
    N, H, W, C = x.shape
    Cout = dy.shape[-1]
    plain = R == 1 and S == 1 and stride == 1 and pad == 0
    PQ = dy.shape[1] * dy.shape[2]
    if not plain and (N * PQ) % 64 != 0:
        # The kernel reduces over 64-pixel tiles.  In im2col mode the box of a partial last tile runs past the last image;
        # its dy rows are zero (tiled-mode zero fill) but 0 x (whatever the x rows hold) must not meet a NaN bit pattern, so
        # such (tiny-map) calls are padded with all-zero images until the pixel count is a whole number of tiles.
        k = next(k for k in range(1, 65) if ((N + k) * PQ) % 64 == 0)
        x = torch.cat([x, x.new_zeros((k, H, W, C))])
        dy = torch.cat([dy, dy.new_zeros((k,) + tuple(dy.shape[1:]))])
        N += k
    dw = torch.zeros((Cout, R, S, C), device=x.device, dtype=F32)
    call("b2_conv2d_wgrad_nhwc_bf16", x.data_ptr(), N, H, W, C, dy.data_ptr(), Cout, R, S, stride, pad, dw.data_ptr(),
         stream_ptr())
    return dw


def weight_layouts(w, want_dgrad=True):
    """w [Cout,Cin,R,S] fp32 parameter -> (wk [Cout,R,S,Cin], wt [Cin,R,S,Cout] with flipped taps or None), bf16, one launch."""
    Cout, Cin, R, S = w.shape
    wd = w.detach()
    if not wd.is_contiguous():
        wd = wd.contiguous()
    wk = torch.empty((Cout, R, S, Cin), device=w.device, dtype=BF16)
    wt = torch.empty((Cin, R, S, Cout), device=w.device, dtype=BF16) if want_dgrad else None
    call("b2_conv_weight_layouts", wd.data_ptr(), wk.data_ptr(), ptr(wt), Cout, Cin, R, S, stream_ptr())
    return wk, wt


def conv_dgrad(dy, w, in_hw, stride, pad, wt=None, add=None):
    """dy [N,P,Q,Cout] bf16, w [Cout,Cin,R,S] fp32 parameter -> dx [N,H,W,Cin] bf16: the forward conv kernels on the
    flipped, transposed filter (stride 2: over the zero-dilated dy).  wt: that filter if the caller already has it;
    add [N,H,W,Cin] bf16: a second gradient of the same tensor (the shortcut's), added in the GEMM epilogue."""
    Cout, Cin, R, S = w.shape
    N, P, Q, _ = dy.shape
    H, W = in_hw
    if wt is None:
        wt = weight_layouts(w)[1]                                                   # [Cin, R, S, Cout]
    if stride == 1:
        src = dy
    else:
        assert stride == 2, "ResNet convolutions have stride 1 or 2"
        src = torch.zeros((N, H, W, Cout), device=dy.device, dtype=BF16)
        call("b2_dilate2_nhwc_bf16", dy.data_ptr(), src.data_ptr(), N, P, Q, H, W, Cout, stream_ptr())
    if add is not None and Cin % 32 == 0 and Cin >= 64:
        dx = conv2d_bn_nhwc(src, wt, 1, R - 1 - pad, res=add.contiguous(), relu=False)   # dgrad + add in the epilogue
    else:
        dx = conv2d_nhwc(src, wt, 1, R - 1 - pad)
        if add is not None:
            dx += add
    assert dx.shape[1:3] == (H, W), (dx.shape, H, W)
    return dx


class ConvBnFn(torch.autograd.Function):
    """z = act(bn(conv(x, w)) [+ res]) over NHWC bf16; bn in train mode (batch statistics, running-stat update) or eval."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, res, bn, stride, pad, relu, train, carry=False):
        """carry=True also returns x itself as a second output: the block hands THAT to its shortcut branch, so the
        shortcut's gradient arrives at this node (not at an autograd add) and is summed in the data-gradient GEMM's epilogue."""
        Cout, Cin, R, S = w.shape
        x = x.contiguous()
        wk, wt = weight_layouts(w, want_dgrad=x.requires_grad)
        stat = torch.zeros(4 * Cout, device=x.device, dtype=F32)         # sum | sumsq | scale | shift
        y = conv2d_nhwc(x, wk, stride, pad, stats=(stat[:Cout], stat[Cout:2 * Cout]) if train else None)
        count = y.numel() // Cout
        _finalize(stat, Cout, bn, count, train)
        z = torch.empty_like(y)
        scale_shift_apply(y, stat[2 * Cout:3 * Cout], stat[3 * Cout:], res=res, relu=relu, out=z)
        if train and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        ctx.save_for_backward(x, w, y, z if relu else None, stat, wt)
        ctx.bn, ctx.geom, ctx.train, ctx.count, ctx.has_res = bn, (stride, pad), train, count, res is not None
        if carry:
            return z, x.view_as(x)
        return z

    @staticmethod
    def backward(ctx, dz, dcarry=None):
        x, w, y, z, stat, wt = ctx.saved_tensors
        stride, pad = ctx.geom
        Cout, Cin, R, S = w.shape
        dz = dz.contiguous()
        want_res = ctx.has_res and ctx.needs_input_grad[4]
        dy, dzm, dgamma, dbeta = _bn_backward(dz, z, y, ctx.bn, stat, ctx.count, ctx.train, want_masked=want_res)
        dx = dw = None
        if ctx.needs_input_grad[1]:
            dw = conv_wgrad(x, dy, R, S, stride, pad).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[0]:
            dx = conv_dgrad(dy, w, x.shape[1:3], stride, pad, wt=wt, add=dcarry)
        return (dx, dw, dgamma if ctx.needs_input_grad[2] else None, dbeta if ctx.needs_input_grad[3] else None,
                dzm if want_res else None, None, None, None, None, None, None)


class StemFn(torch.autograd.Function):
    """conv1 7x7/2 -> bn1 -> ReLU -> maxpool 3x3/2 on the patch-matrix stem (the patch matrix is kept for the weight
    gradient); the input frames need no gradient."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, bn, train):
        x = x.contiguous()
        N, _, H, W = x.shape
        P, Q = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        st = stream_ptr()
        A = torch.empty((N * P * Q, STEM_KP), device=x.device, dtype=BF16)
        call("b2_stem_im2col", x.data_ptr(), int(x.dtype == BF16), A.data_ptr(), N, H, W, STEM_KP, st)
        wk = torch.zeros((64, 3, 7, 8), device=x.device, dtype=BF16)
        wk[:, :, :, :7] = w.detach().to(BF16)
        stat = torch.zeros(4 * 64, device=x.device, dtype=F32)
        raw = gemm_tn(A, wk.reshape(64, STEM_KP), out_dtype=BF16, stats=(stat[:64], stat[64:128]) if train else None)
        P2, Q2 = (P + 2 - 3) // 2 + 1, (Q + 2 - 3) // 2 + 1
        y = torch.empty((N, P2, Q2, 64), device=x.device, dtype=BF16)
        _finalize(stat, 64, bn, N * P * Q, train, momentum=0.0)        # scale/shift for the backward; the pool kernel
        call("b2_bn_relu_maxpool_nhwc", raw.data_ptr(), y.data_ptr(), N, P, Q, 64, ptr(stat[:64] if train else None),
             ptr(stat[64:128] if train else None), bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
             bn.running_var.data_ptr(), float(bn.eps), float(_momentum(bn)), int(train), st)   # updates the running stats
        if train and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        ctx.save_for_backward(A, raw, stat)
        ctx.bn, ctx.train, ctx.dims = bn, train, (N, P, Q, P2, Q2)
        return y

    @staticmethod
    def backward(ctx, dpool):
        A, raw, stat = ctx.saved_tensors
        N, P, Q, P2, Q2 = ctx.dims
        dpool = dpool.contiguous()
        dbn = torch.zeros((N, P, Q, 64), device=raw.device, dtype=F32)
        call("b2_maxpool_relu_bwd_nhwc", raw.data_ptr(), stat.data_ptr() + 8 * 64, stat.data_ptr() + 12 * 64, dpool.data_ptr(),
             dbn.data_ptr(), N, P, Q, P2, Q2, 64, stream_ptr())
        dzb = cast_bf16(dbn)
        del dbn
        draw, _, dgamma, dbeta = _bn_backward(dzb, None, raw, ctx.bn, stat, N * P * Q, ctx.train)
        dw = None
        if ctx.needs_input_grad[1]:
            M = N * P * Q
            d = torch.zeros((64, 1, 1, STEM_KP), device=raw.device, dtype=F32)
            call("b2_conv2d_wgrad_nhwc_bf16", A.data_ptr(), M, 1, 1, STEM_KP, draw.data_ptr(), 64, 1, 1, 1, 0, d.data_ptr(),
                 stream_ptr())
            dw = d.reshape(64, 3, 7, 8)[:, :, :, :7]
        return (None, dw, dgamma if ctx.needs_input_grad[2] else None, dbeta if ctx.needs_input_grad[3] else None, None, None)


class AvgPoolFn(torch.autograd.Function):
    """[N,H,W,C] bf16 -> [N,C] fp32 spatial mean (torchvision avgpool + flatten)."""

    @staticmethod
    def forward(ctx, y):
        Nn, Hh, Ww, C = y.shape
        feat = torch.empty((Nn, C), device=y.device, dtype=F32)
        call("b2_avgpool_nhwc", y.data_ptr(), feat.data_ptr(), 0, Nn, Hh * Ww, C, stream_ptr())
        ctx.shape = tuple(y.shape)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        Nn, Hh, Ww, C = ctx.shape
        dfeat = dfeat.contiguous().float()
        dz = torch.empty(ctx.shape, device=dfeat.device, dtype=BF16)
        call("b2_avgpool_bwd_nhwc", dfeat.data_ptr(), dz.data_ptr(), Nn, Hh * Ww, C, stream_ptr())
        return dz


_record = None      # tests: a list that collects every node's output activation in execution order


def _conv_bn(x, conv, bn, relu, train, res=None, carry=False, record=True):
    stride = conv.stride[0]
    pad = conv.padding[0]
    out = ConvBnFn.apply(x, conv.weight, bn.weight, bn.bias, res, bn, stride, pad, relu, train, carry)
    if record and _record is not None:
        _record.append(out[0] if carry else out)
    return out


def block_forward(y, blk, train):
    """One torchvision BasicBlock / Bottleneck (v1.5: the stride sits on the 3x3 conv) on ConvBnFn nodes.  The block input
    feeds conv1 AND the shortcut: conv1's node hands the input on (carry), the shortcut branch hangs off that copy, and the two
    gradients of the block input meet inside conv1's data-gradient GEMM (epilogue add) instead of in a separate add pass."""
    o, y2 = _conv_bn(y, blk.conv1, blk.bn1, True, train, carry=True, record=False)
    short = y2 if blk.downsample is None else _conv_bn(y2, blk.downsample[0], blk.downsample[1], False, train)
    if _record is not None:             # (recorded in module order: down-sample branch, conv1, conv2, conv3)
        _record.append(o)
    if hasattr(blk, "conv3"):
        o = _conv_bn(o, blk.conv2, blk.bn2, True, train)
        return _conv_bn(o, blk.conv3, blk.bn3, True, train, res=short)
    return _conv_bn(o, blk.conv2, blk.bn2, True, train, res=short)


def first_trainable_block(net):
    """('stem',) or (layer index, block index) of the first block that owns a trainable parameter, None if frozen."""
    if any(p.requires_grad for p in list(net.conv1.parameters()) + list(net.bn1.parameters())):
        return ("stem",)
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(net, f"layer{li}")):
            if any(p.requires_grad for p in blk.parameters()):
                return (li, bi)
    return None


def encode_trainable(runner, x, training):
    """Frame features [N, feat] fp32 with autograd through the trainable suffix of runner.net (x: [N,3,H,W])."""
    _lib.require_device()
    net = runner.net
    start = first_trainable_block(net)
    assert start is not None
    train = bool(training)
    if start == ("stem",):
        y = StemFn.apply(x, net.conv1.weight, net.bn1.weight, net.bn1.bias, net.bn1, train)
        if _record is not None:
            _record.append(y)
        start = (1, 0)
    else:
        with torch.no_grad():
            y = runner(x, training, stop_at=start)          # frozen prefix on the fused kernels, NHWC bf16
        if _record is not None:
            _record.append(y)
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(net, f"layer{li}")):
            if (li, bi) >= start:
                y = block_forward(y, blk, train)
    return AvgPoolFn.apply(y)
