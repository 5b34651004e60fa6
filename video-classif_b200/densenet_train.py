"""Trainable DenseNet frame encoder -- the DEFAULT of the reference's crime / rgb scripts (`CNN_BACKBONE = "densenet121"`,
`FINETUNE = True`: lrcn/lrcn.py:27,34,230, lrcn/rgb_lrcn.py:31,38,197 -- nothing is frozen).

Forward = `DenseNetRunner.stem` (autograd node `backbone_train.StemFn`: conv0 / norm0 / pool0 are the ResNet stem) +
`DenseNetRunner.trunk` with the block buffers, per-feature statistics and the per-layer bottleneck tensors kept.
Backward = ONE autograd node for the whole trunk: each dense block owns a gradient buffer dX with the layout of its
concatenated activation buffer; walking the layers in reverse, layer k
    takes d(y2) = dX[:, Ck:Ck+growth]  (complete: every later consumer has already added its share),
    conv2: weight gradient (tcgen05 wgrad kernel) + data gradient (forward conv kernel, flipped filter),
    BN2+ReLU backward, conv1: weight gradient + data gradient (GEMM), then BN1+ReLU backward ACCUMULATED into
    dX[:, :Ck] (`b2_bn_bwd_ld_bf16`) -- the transpose of "every layer reads all earlier features".
Transitions: avg-pool backward, 1x1 conv backward, BN+ReLU backward into the previous block's dX."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .backbone_train import AvgPoolFn, StemFn, _bn_backward, _momentum, conv_wgrad  # noqa: F401
from .ops import BF16, F32, conv2d_nhwc, gemm_tn


def _trunk_params(net):
    """(names, tensors) of every parameter behind the stem, in named_parameters() order."""
    items = [(n, p) for n, p in net.features.named_parameters() if not n.startswith(("conv0.", "norm0."))]
    return ["features." + n for n, _ in items], [p for _, p in items]


def _bn_bwd_ld(dz, lddz, z, ldz, X, ldx, dX, lddx, accumulate, bn, sum_ptr, sq_ptr, M, C, train):
    """-> (dgamma, dbeta); dX[:, :C] written / accumulated."""
    s12 = torch.zeros(2 * C, device=X.device, dtype=F32)
    call("b2_bn_bwd_ld_bf16", dz.data_ptr(), lddz, ptr(z), ldz, X.data_ptr(), ldx, dX.data_ptr(), lddx, int(accumulate),
         bn.weight.data_ptr(), sum_ptr if train else 0, sq_ptr if train else 0, bn.running_mean.data_ptr(),
         bn.running_var.data_ptr(), s12.data_ptr(), s12.data_ptr() + 4 * C, M, C, M, float(bn.eps), int(train), stream_ptr())
    return s12[C:], s12[:C]


def _recompute_act(runner, bn, X, ldx, S, Cfin, M, C, train):
    """relu(bn(X[:, :C])) as a contiguous [M, C] tensor (the forward's a1 / transition / norm5 activation)."""
    ss = torch.zeros((2, (C + 63) // 64 * 64), device=X.device, dtype=F32)
    call("b2_bn_finalize_nhwc", S.data_ptr() if train else 0, S.data_ptr() + 4 * Cfin if train else 0, bn.weight.data_ptr(),
         bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(), M, float(bn.eps), 0.0, int(train),
         ss.data_ptr(), ss.data_ptr() + ss.stride(0) * 4, C, stream_ptr())          # momentum 0: running stats untouched
    a = torch.empty((M, C), device=X.device, dtype=BF16)
    call("b2_scale_shift_apply_ld_bf16", X.data_ptr(), ldx, a.data_ptr(), C, M, C, ss.data_ptr(), ss.data_ptr() + ss.stride(0) * 4,
         1, stream_ptr())
    return a


_record = None      # tests: a list that receives the saved-activation records of each trunk forward


class DenseTrunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, runner, train, names, *params):
        saved = []
        feat = runner.trunk(y0, train, saved=saved)
        if _record is not None:
            _record.append(saved)
        ctx.runner, ctx.train, ctx.names, ctx.saved = runner, train, names, saved
        ctx.params = params
        ctx.y0_shape = tuple(y0.shape)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        runner, train, names, saved = ctx.runner, ctx.train, ctx.names, ctx.saved
        net = runner.net
        mods = dict(net.named_modules())
        need = {n: ctx.needs_input_grad[4 + i] for i, n in enumerate(names)}
        grads = {}
        st = stream_ptr()
        N = ctx.y0_shape[0]
        dev = dfeat.device
        dX = None              # gradient buffer of the block being processed
        dnext = None           # (tensor, ld, C): gradient w.r.t. the block's output features coming from downstream
        for entry in reversed(saved):
            kind = entry[0]
            if kind == "norm5":
                _, bn, X, S, a, Hc, Wc, C = entry
                M = N * Hc * Wc
                dz = torch.empty((N, Hc * Wc, C), device=dev, dtype=BF16)
                call("b2_avgpool_bwd_nhwc", dfeat.contiguous().float().data_ptr(), dz.data_ptr(), N, Hc * Wc, C, st)
                dX = torch.empty((M, C), device=dev, dtype=BF16)
                dg, db = _bn_bwd_ld(dz, C, a, C, X, C, dX, C, 0, bn, S.data_ptr(), S.data_ptr() + 4 * C, M, C, train)
                grads["features.norm5.weight"], grads["features.norm5.bias"] = dg, db
            elif kind == "transition":
                _, tname, mod, X, S, Hc, Wc, C, Cn = entry
                # dX currently holds the NEXT block's gradient buffer: its first Cn channels are d(pooled transition output)
                M = N * Hc * Wc
                ldn = dX.shape[-1]
                dyt = torch.empty((M, Cn), device=dev, dtype=BF16)
                call("b2_avgpool2x2_bwd_nhwc_bf16", dX.data_ptr(), ldn, dyt.data_ptr(), N, Hc, Wc, Cn, st)
                a = _recompute_act(runner, mod.norm, X, C, S, C, M, C, train)
                wname = tname + ".conv.weight"
                if need[wname]:
                    grads[wname] = conv_wgrad(a.view(M, 1, 1, C), dyt.view(M, 1, 1, Cn), 1, 1, 1, 0).permute(0, 3, 1, 2)
                wt = mod.conv.weight.detach().reshape(Cn, C).t().contiguous().to(BF16)              # [C, Cn]
                da = gemm_tn(dyt, wt, out_dtype=BF16)                                                # [M, C]
                dX = torch.empty((M, C), device=dev, dtype=BF16)
                dg, db = _bn_bwd_ld(da, C, a, C, X, C, dX, C, 0, mod.norm, S.data_ptr(), S.data_ptr() + 4 * C, M, C, train)
                grads[tname + ".norm.weight"], grads[tname + ".norm.bias"] = dg, db
                del a, da, dyt
            else:
                _, X, S, C0, growth, rec = entry
                Nn, Hc, Wc, Cfin = X.shape
                M = Nn * Hc * Wc
                assert dX.shape == (M, Cfin)
                for k in range(len(rec) - 1, -1, -1):
                    lname, layer, y1, a2, smid = rec[k]
                    Ct = C0 + k * growth
                    mid = y1.shape[1]
                    # d(y2): contiguous copy of the layer's own channel slice, zero-padded to 64 channels for the data gradient
                    gpad = torch.zeros((M, 64), device=dev, dtype=BF16) if growth < 64 else torch.empty((M, growth), device=dev, dtype=BF16)
                    call("b2_scale_shift_apply_ld_bf16", dX.data_ptr() + 2 * Ct, Cfin, gpad.data_ptr(), gpad.shape[1], M, growth, 0, 0, 0, st)
                    w2 = layer.conv2.weight
                    if need[lname + ".conv2.weight"]:
                        gnew = gpad[:, :growth].contiguous() if gpad.shape[1] != growth else gpad
                        grads[lname + ".conv2.weight"] = conv_wgrad(a2.view(Nn, Hc, Wc, mid), gnew.view(Nn, Hc, Wc, growth), 3, 3, 1,
                                                                    1).permute(0, 3, 1, 2)
                    wt2 = torch.zeros((mid, 3, 3, gpad.shape[1]), device=dev, dtype=BF16)            # [Cin, R, S, Cout(padded)]
                    wt2[..., :growth] = w2.detach().flip(2, 3).permute(1, 2, 3, 0).to(BF16)
                    da2 = conv2d_nhwc(gpad.view(Nn, Hc, Wc, gpad.shape[1]), wt2, 1, 1)               # [N,H,W,mid]
                    dy1, _, dg2, db2 = _bn_backward(da2.view(M, mid), a2, y1, layer.norm2, smid.reshape(-1), M, train)
                    grads[lname + ".norm2.weight"], grads[lname + ".norm2.bias"] = dg2, db2
                    a1 = _recompute_act(runner, layer.norm1, X, Cfin, S, Cfin, M, Ct, train)
                    if need[lname + ".conv1.weight"]:
                        grads[lname + ".conv1.weight"] = conv_wgrad(a1.view(M, 1, 1, Ct), dy1.view(M, 1, 1, mid), 1, 1, 1,
                                                                    0).permute(0, 3, 1, 2)
                    wt1 = layer.conv1.weight.detach().reshape(mid, Ct).t().contiguous().to(BF16)     # [Ct, mid]
                    da1 = gemm_tn(dy1, wt1, out_dtype=BF16)                                          # [M, Ct]
                    dg1, db1 = _bn_bwd_ld(da1, Ct, a1, Ct, X, Cfin, dX, Cfin, 1, layer.norm1, S.data_ptr(),
                                          S.data_ptr() + 4 * Cfin, M, Ct, train)
                    grads[lname + ".norm1.weight"], grads[lname + ".norm1.bias"] = dg1, db1
                    del gpad, da2, dy1, a1, da1
        # gradient of the stem output: the first 64 channels of block 1's buffer
        dy0 = None
        if ctx.needs_input_grad[0]:
            Cfin = dX.shape[-1]
            C0 = ctx.y0_shape[-1]
            dy0 = torch.empty(ctx.y0_shape, device=dev, dtype=BF16)
            call("b2_scale_shift_apply_ld_bf16", dX.data_ptr(), Cfin, dy0.data_ptr(), C0, dX.shape[0], C0, 0, 0, 0, st)
        out = [grads.get(n) if need[n] else None for n in names]
        return (dy0, None, None, None, *out)


def encode_trainable(runner, x, training):
    """Frame features [N, feat] fp32 with autograd through the whole DenseNet (x: [N,3,H,W])."""
    _lib.require_device()
    net = runner.net
    train = bool(training)
    f = net.features
    y0 = StemFn.apply(x, f.conv0.weight, f.norm0.weight, f.norm0.bias, f.norm0, train)
    names, params = _trunk_params(net)
    feat = DenseTrunkFn.apply(y0, runner, train, names, *params)
    if train:
        bns = [m for n, m in f.named_modules() if isinstance(m, torch.nn.BatchNorm2d) and n != "norm0"]
        torch._foreach_add_([b.num_batches_tracked for b in bns if b.num_batches_tracked is not None], 1)
    return feat
