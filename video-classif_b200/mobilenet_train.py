"""Trainable MobileNetV2 frame encoder: forward with saved activations + backward on the B200 kernels.

The reference's crime / rgb scripts leave the WHOLE backbone trainable when CONF_FINETUNE is set (`freeze_cnn_layers` only
un-freezes the Identity head and freezes nothing, lrcn/lrcn.py:246-258, rgb_lrcn.py:208-227) and `freeze_until_layer=k`
freezes the first k parameters (lrcn.py:275-283); with CNN_BACKBONE = "mobilenet_v2" that is this network.

One autograd node for the trunk (`MobileNetTrunkFn`).  The forward records a tape of units; every BatchNorm keeps its raw
bf16 input and its batch statistics, every ReLU6 its output:
    stem   Conv2d(3,32,3,s2)            b2_mbv2_stem_conv (+ statistics)
    bn6    BatchNorm + ReLU6            b2_scale_shift_apply_ld_bf16 (out of place)
    pw     1x1 conv                     tcgen05 GEMM (+ statistics in the epilogue)
    dw     depthwise 3x3                b2_dwconv3x3_bn_nhwc_bf16 (+ statistics)
    bnl    BatchNorm (+ shortcut)       linear bottleneck output
    pool   global average
The backward walks the tape in reverse: BatchNorm backward (two passes, ReLU6 mask 0 < z < 6: b2_bn_bwd_relu6_nhwc_bf16),
1x1 weight gradient on the tcgen05 weight-gradient kernel (MN-major operands straight from the NHWC tensors), 1x1 data
gradient on the tcgen05 GEMM, depthwise data / weight gradients and the stem weight gradient on mobilenet_bwd.cu.  The walk
stops below the first unit that owns a trainable parameter (partial freezing)."""
from __future__ import annotations

import torch

from ._lib import call, ptr, stream_ptr
from .ops import BF16, F32, gemm_tn, scale_shift_apply


def _mom(bn):
    return bn.momentum if bn.momentum is not None else 0.1


class MobileNetTrunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, x, training, *params):
        net = runner.net
        x = x.contiguous()
        N, _, H, W = x.shape
        dev = x.device
        st = stream_ptr()
        train = bool(training)
        w = runner._weights()
        bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
        maxc = max(b.num_features for b in bns)
        stat = torch.zeros((len(bns), 4, maxc), device=dev, dtype=F32)      # [sum | sumsq | scale | shift]
        row = {id(b): i for i, b in enumerate(bns)}
        tape = []

        def stats_of(bn):
            r = stat[row[id(bn)]]
            return (r[0], r[1]) if train else (None, None)

        def finalize(bn, count):
            r = stat[row[id(bn)]]
            call("b2_bn_finalize_nhwc", ptr(r[0]) if train else 0, ptr(r[1]) if train else 0, bn.weight.data_ptr(),
                 bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(), count, float(bn.eps), float(_mom(bn)),
                 int(train), r[2].data_ptr(), r[3].data_ptr(), bn.num_features, st)
            return r[2], r[3]

        def bn6(raw, bn):
            """raw [M,C] -> relu6(bn(raw)), out of place (the raw tensor feeds the BatchNorm backward)."""
            ss = finalize(bn, raw.shape[0])
            act = torch.empty_like(raw)
            M, C = raw.shape
            call("b2_scale_shift_apply_ld_bf16", raw.data_ptr(), C, act.data_ptr(), C, M, C, ptr(ss[0]), ptr(ss[1]), 2, st)
            tape.append(("bn6", raw, act, bn))
            return act

        def pw(a, conv, bn, hw):
            s = stats_of(bn)
            raw = gemm_tn(a, w[id(conv)], out_dtype=BF16, stats=s if train else None)
            tape.append(("pw", a, conv, hw))
            return raw

        def dw(a, conv, bn, Hc, Wc):
            C = conv.out_channels
            s = conv.stride[0]
            P, Q = (Hc + 2 - 3) // s + 1, (Wc + 2 - 3) // s + 1
            raw = torch.empty((N * P * Q, C), device=dev, dtype=BF16)
            so = stats_of(bn)
            call("b2_dwconv3x3_bn_nhwc_bf16", a.data_ptr(), 0, 0, 2, w[id(conv)].data_ptr(), raw.data_ptr(), ptr(so[0]), ptr(so[1]),
                 N, Hc, Wc, C, s, st)
            tape.append(("dw", a, conv, Hc, Wc))
            return raw, P, Q

        f = net.features
        conv0, bn0 = f[0][0], f[0][1]
        Hc, Wc = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        raw = torch.empty((N * Hc * Wc, 32), device=dev, dtype=BF16)
        s0 = stats_of(bn0)
        call("b2_mbv2_stem_conv", x.data_ptr(), int(x.dtype == BF16), w[id(conv0)].data_ptr(), raw.data_ptr(), ptr(s0[0]), ptr(s0[1]),
             N, H, W, st)
        tape.append(("stem", x, conv0))
        cur = bn6(raw, bn0)
        for blk in list(f)[1:-1]:
            layers = list(blk.conv.children())
            inp = cur
            if len(layers) == 4:                               # expand -> depthwise -> project
                (econv, ebn, _), (dconv, dbn, _) = list(layers[0].children()), list(layers[1].children())
                pconv, pbn = layers[2], layers[3]
                e = bn6(pw(cur, econv, ebn, (Hc, Wc)), ebn)
            else:                                              # first block: depthwise on the stem activation -> project
                (dconv, dbn, _) = list(layers[0].children())
                pconv, pbn = layers[1], layers[2]
                e = cur
            draw, P, Q = dw(e, dconv, dbn, Hc, Wc)
            d = bn6(draw, dbn)
            praw = pw(d, pconv, pbn, (P, Q))
            ss3 = finalize(pbn, praw.shape[0])
            out = torch.empty_like(praw)
            C = praw.shape[1]
            res = inp if blk.use_res_connect else None
            scale_shift_apply(praw, ss3[0], ss3[1], res=res, relu=False, out=out)
            tape.append(("bnl", praw, pbn, res is not None))
            cur, Hc, Wc = out, P, Q
        hconv, hbn = f[-1][0], f[-1][1]
        hact = bn6(pw(cur, hconv, hbn, (Hc, Wc)), hbn)
        feat = torch.empty((N, hact.shape[1]), device=dev, dtype=F32)
        call("b2_avgpool_nhwc", hact.data_ptr(), feat.data_ptr(), 0, N, Hc * Wc, hact.shape[1], st)
        tape.append(("pool", Hc * Wc, hact.shape[1]))
        if train:
            torch._foreach_add_([b.num_batches_tracked for b in bns if b.num_batches_tracked is not None], 1)
        ctx.tape, ctx.stat, ctx.row, ctx.train, ctx.N = tape, stat, row, train, N
        ctx.params = params
        ctx.runner = runner
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        tape, stat, row, train, N = ctx.tape, ctx.stat, ctx.row, ctx.train, ctx.N
        runner = ctx.runner
        w = runner._weights()
        dev = dfeat.device
        st = stream_ptr()
        grads = {}                                             # id(parameter) -> gradient
        need = {id(p) for p, flag in zip(ctx.params, ctx.needs_input_grad[3:]) if flag}

        # the walk can stop below the first unit (in forward order) that owns a trainable parameter
        def owns(unit):
            k = unit[0]
            if k in ("pw", "dw", "stem"):
                return id(unit[2].weight) in need
            if k in ("bn6", "bnl"):
                bn = unit[3] if k == "bn6" else unit[2]
                return id(bn.weight) in need or id(bn.bias) in need
            return False
        first = next((i for i, u in enumerate(tape) if owns(u)), len(tape))

        ones = {}

        def add_inplace(a, b):
            """a += b over [M, C] bf16 (the shortcut's gradient joins the block input's)."""
            C = a.shape[1]
            if C not in ones:
                ones[C] = (torch.ones(C, device=dev, dtype=F32), torch.zeros(C, device=dev, dtype=F32))
            scale_shift_apply(a, ones[C][0], ones[C][1], res=b, relu=False)

        g = None                                               # gradient w.r.t. the output of the unit being visited
        pending_res = []                                       # (gradient of a shortcut, tensor it belongs to)
        for i in range(len(tape) - 1, first - 1, -1):
            u = tape[i]
            kind = u[0]
            if kind == "pool":
                _, HW, C = u
                d = dfeat.contiguous().float()
                g = torch.empty((N * HW, C), device=dev, dtype=BF16)
                call("b2_avgpool_bwd_nhwc", d.data_ptr(), g.data_ptr(), N, HW, C, st)
            elif kind in ("bn6", "bnl"):
                if kind == "bn6":
                    _, raw, act, bn = u
                    fn, z = "b2_bn_bwd_relu6_nhwc_bf16", act
                else:
                    _, raw, bn, has_res = u
                    fn, z = "b2_bn_bwd_nhwc_bf16", None
                    if has_res:
                        pending_res.append(g)                  # the shortcut receives the same gradient as the BN output
                M, C = raw.shape
                r = stat[row[id(bn)]]
                s = torch.zeros((2, C), device=dev, dtype=F32)
                dy = torch.empty_like(raw)
                call(fn, g.data_ptr(), 0, ptr(z), raw.data_ptr(), dy.data_ptr(), bn.weight.data_ptr(), ptr(r[0]), ptr(r[1]),
                     bn.running_mean.data_ptr(), bn.running_var.data_ptr(), s[0].data_ptr(), s[1].data_ptr(), M, C, M,
                     float(bn.eps), int(train), st)
                grads[id(bn.weight)], grads[id(bn.bias)] = s[1], s[0]
                g = dy
            elif kind == "pw":
                _, a, conv, (Hc, Wc) = u
                Cout, Cin = conv.out_channels, conv.in_channels
                if id(conv.weight) in need:
                    dw = torch.zeros((Cout, 1, 1, Cin), device=dev, dtype=F32)
                    call("b2_conv2d_wgrad_nhwc_bf16", a.data_ptr(), N, Hc, Wc, Cin, g.data_ptr(), Cout, 1, 1, 1, 0, dw.data_ptr(), st)
                    grads[id(conv.weight)] = dw.reshape(Cout, Cin, 1, 1)
                if i > first:
                    wt = w[id(conv)].t().contiguous()          # [Cin, Cout] bf16 (small)
                    g = gemm_tn(g, wt, out_dtype=BF16)
                    if pending_res and tape[i - 1][0] == "bnl":   # first conv of a residual block: its input is the block input
                        add_inplace(g, pending_res.pop())
            elif kind == "dw":
                _, a, conv, Hc, Wc = u
                C = conv.out_channels
                s = conv.stride[0]
                if id(conv.weight) in need:
                    dw = torch.zeros((C, 9), device=dev, dtype=F32)
                    call("b2_dwconv3x3_wgrad_nhwc_bf16", a.data_ptr(), g.data_ptr(), dw.data_ptr(), N, Hc, Wc, C, s, st)
                    grads[id(conv.weight)] = dw.reshape(C, 1, 3, 3)
                if i > first:
                    dx = torch.empty_like(a)
                    call("b2_dwconv3x3_dgrad_nhwc_bf16", g.data_ptr(), w[id(conv)].data_ptr(), dx.data_ptr(), N, Hc, Wc, C, s, st)
                    g = dx
            elif kind == "stem":
                _, x, conv = u
                if id(conv.weight) in need:
                    dw = torch.zeros((32, 27), device=dev, dtype=F32)
                    call("b2_mbv2_stem_wgrad", x.data_ptr(), int(x.dtype == BF16), g.data_ptr(), dw.data_ptr(), N, x.shape[2],
                         x.shape[3], st)
                    grads[id(conv.weight)] = dw.reshape(32, 3, 3, 3)
        out = [grads.get(id(p)) if flag else None for p, flag in zip(ctx.params, ctx.needs_input_grad[3:])]
        return (None, None, None, *out)


def encode_trainable(runner, x, training):
    params = list(runner.net.parameters())
    return MobileNetTrunkFn.apply(runner, x, training, *params)
