"""MobileNetV2 frame encoder (torchvision mobilenet_v2 with classifier -> Identity, medsos_lrcn/src/models.py:133-143;
one of the two backbones of the reference's own search space, medsos_lrcn/src/automation.py:28) on the B200 kernels.

Every layer is memory bound, so the plan is about passes over the activations, per inverted-residual block:
  expand 1x1   tcgen05 GEMM, batch statistics of its output in the epilogue             (raw bf16 out)
  depthwise    `b2_dwconv3x3_bn_nhwc_bf16`: BN1 + ReLU6 applied to the input on load, statistics of its own output
  BN2 + ReLU6  one in-place element pass
  project 1x1  tcgen05 GEMM + statistics
  BN3 (+ x)    one element pass (linear bottleneck: no activation; the shortcut is added here)
The frozen encoder replays from a CUDA graph like the other backbones; a (partially) trainable one runs through the autograd
node of mobilenet_train.py (saved activations + backward kernels)."""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .ops import graph_epoch as _graph_epoch
from .ops import BF16, F32, gemm_tn, scale_shift_apply

SUPPORTED = ("mobilenet_v2",)


def _mom(bn):
    return bn.momentum if bn.momentum is not None else 0.1


class MobileNetRunner:
    def __init__(self, net):
        self.net = net
        self._wcache = None
        self._wkey = None
        self._convs = None
        self._params = None
        self._graphs = {}
        self.use_graph = False
        self.fuse_bn = None          # (ResNet-only switches; part of the CUDA-graph cache key of the shared graphed())
        self.stem_impl = None
        self.split_bn1 = os.environ.get("B2_MBV2_SPLIT_BN1", "1") == "1"

    def _weights(self):
        convs = self._convs
        if convs is None:
            convs = self._convs = [(n, m) for n, m in self.net.named_modules() if isinstance(m, torch.nn.Conv2d)]
        key = tuple((m.weight.data_ptr(), m.weight._version) for _, m in convs)
        if any(m.weight.requires_grad for _, m in convs):      # (a replayed train-step graph updates weights without version bumps)
            key += (_graph_epoch(),)
        if self._wkey != key:
            cache = {}
            for n, m in convs:
                w = m.weight.detach()
                if m.groups > 1 or w.shape[1] == 3:                 # depthwise / stem: fp32 [C, 9] / [32, 27] for the SIMT kernels
                    cache[id(m)] = w.reshape(w.shape[0], -1).float().contiguous()
                else:                                               # 1x1: [Cout, C] bf16, K padded to a 16-byte row
                    K = w.shape[1]
                    wk = torch.zeros((w.shape[0], (K + 7) // 8 * 8), device=w.device, dtype=BF16)
                    wk[:, :K] = w.reshape(w.shape[0], K).to(BF16)
                    cache[id(m)] = wk[:, :K]
            self._wcache, self._wkey = cache, key
        return self._wcache

    def graphed(self, x, training: bool):
        from .backbone import ResNetRunner
        return ResNetRunner.graphed(self, x, training)

    def __call__(self, x, training: bool, return_stages: bool = False):
        """x: [N,3,H,W] fp32 / bf16 NCHW -> [N, 1280] fp32."""
        _lib.require_device()
        net = self.net
        if self._params is None:
            self._params = list(net.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._params):
            # (partially) trainable encoder: lrcn.py:246-283 / rgb_lrcn.py:208-227 with CNN_BACKBONE = "mobilenet_v2"
            from .mobilenet_train import encode_trainable
            return encode_trainable(self, x, training)
        x = x.contiguous()
        N, Cin, H, W = x.shape
        assert Cin == 3, "frame encoder expects RGB frames"
        dev = x.device
        w = self._weights()
        train = bool(training)
        st = stream_ptr()
        bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
        maxc = max(b.num_features for b in bns)
        # per BatchNorm: [sum | sumsq | scale | shift] (zero-padded rows)
        stat = torch.zeros((len(bns), 4, maxc), device=dev, dtype=F32)
        bn_row = {id(b): i for i, b in enumerate(bns)}

        def stats_of(bn):
            r = stat[bn_row[id(bn)]]
            return (r[0], r[1]) if train else None

        def finalize(bn, count):
            r = stat[bn_row[id(bn)]]
            C = bn.num_features
            call("b2_bn_finalize_nhwc", r[0].data_ptr() if train else 0, r[1].data_ptr() if train else 0, bn.weight.data_ptr(),
                 bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(), count, float(bn.eps),
                 float(_mom(bn)), int(train), r[2].data_ptr(), r[3].data_ptr(), C, st)
            return r[2], r[3]

        def conv1x1(a, conv, bn):
            """a [M, K] bf16 -> raw [M, Cout] bf16 (+ statistics of bn's input)."""
            return gemm_tn(a, w[id(conv)], out_dtype=BF16, stats=stats_of(bn))

        def dwconv(xr, Hc, Wc, conv, bn_in, ss_in, bn_out):
            C = conv.out_channels
            s = conv.stride[0]
            P, Q = (Hc + 2 - 3) // s + 1, (Wc + 2 - 3) // s + 1
            y = torch.empty((N * P * Q, C), device=dev, dtype=BF16)
            so = stats_of(bn_out)
            if ss_in is not None and s == 1 and self.split_bn1:
                # stride 1: every input pixel feeds 9 outputs and the fused kernel is bound by the activation math per loaded
                # element (measured 393 us fused vs 81 + 217 us for an in-place BN + ReLU6 pass and the plain depthwise conv)
                scale_shift_apply_ld(xr, ss_in, 2)
                ss_in = None
            call("b2_dwconv3x3_bn_nhwc_bf16", xr.data_ptr(), ptr(ss_in[0]) if ss_in else 0, ptr(ss_in[1]) if ss_in else 0, 2,
                 w[id(conv)].data_ptr(), y.data_ptr(), ptr(so[0]) if so else 0, ptr(so[1]) if so else 0, N, Hc, Wc, C, s, st)
            return y, P, Q

        f = net.features
        # ---- stem: Conv2d(3, 32, 3, s2, p1) -> BN -> ReLU6 (the BN + ReLU6 is applied by the first depthwise conv on load)
        conv0, bn0 = f[0][0], f[0][1]
        Hc, Wc = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        cur = torch.empty((N * Hc * Wc, 32), device=dev, dtype=BF16)
        s0 = stats_of(bn0)
        call("b2_mbv2_stem_conv", x.data_ptr(), int(x.dtype == BF16), w[id(conv0)].data_ptr(), cur.data_ptr(), ptr(s0[0]) if s0 else 0,
             ptr(s0[1]) if s0 else 0, N, H, W, st)
        pending = (bn0, finalize(bn0, N * Hc * Wc))          # raw tensor `cur` still needs this BN + ReLU6
        stages = [] if return_stages else None
        for blk in list(f)[1:-1]:
            layers = list(blk.conv.children())
            inp = None
            if pending is not None and len(layers) == 4:
                # the block input is used by the expand GEMM (and possibly the shortcut): materialise the activation
                scale_shift_apply_ld(cur, pending[1], 2)
                pending = None
            if len(layers) == 4:                               # expand -> depthwise -> project
                inp = cur
                (econv, ebn, _), (dconv, dbn, _) = list(layers[0].children()), list(layers[1].children())
                pconv, pbn = layers[2], layers[3]
                e = conv1x1(cur, econv, ebn)
                ss1 = finalize(ebn, e.shape[0])
                d, P, Q = dwconv(e, Hc, Wc, dconv, ebn, ss1, dbn)
                del e
            else:                                              # first block: depthwise on the stem output -> project
                (dconv, dbn, _) = list(layers[0].children())
                pconv, pbn = layers[1], layers[2]
                d, P, Q = dwconv(cur, Hc, Wc, dconv, pending[0], pending[1], dbn)
                pending = None
            ss2 = finalize(dbn, d.shape[0])
            scale_shift_apply_ld(d, ss2, 2)                    # BN2 + ReLU6 in place
            pr = conv1x1(d, pconv, pbn)
            del d
            ss3 = finalize(pbn, pr.shape[0])
            C = pr.shape[1]
            if blk.use_res_connect:
                if C % 8 == 0:
                    scale_shift_apply(pr, ss3[0], ss3[1], res=inp, relu=False)
                else:
                    raise _lib.B200LrcnError("channel count not a multiple of 8")
            else:
                scale_shift_apply_ld(pr, ss3, 0)
            cur, Hc, Wc = pr, P, Q
            if return_stages:
                stages.append(cur.view(N, Hc * Wc, C).float().mean(dim=1))
        # ---- head: Conv2d(320, 1280, 1) -> BN -> ReLU6 -> global average pool
        hconv, hbn = f[-1][0], f[-1][1]
        hraw = conv1x1(cur, hconv, hbn)
        ssh = finalize(hbn, hraw.shape[0])
        scale_shift_apply_ld(hraw, ssh, 2)
        feat = torch.empty((N, hraw.shape[1]), device=dev, dtype=F32)
        call("b2_avgpool_nhwc", hraw.data_ptr(), feat.data_ptr(), 0, N, Hc * Wc, hraw.shape[1], st)
        if train:
            torch._foreach_add_([b.num_batches_tracked for b in bns if b.num_batches_tracked is not None], 1)
        if return_stages:
            return feat, stages
        return feat


def scale_shift_apply_ld(t, ss, act):
    """t [M, C] bf16 <- act(t * scale + shift) in place; act: 0 none, 1 ReLU, 2 ReLU6."""
    M, C = t.shape
    call("b2_scale_shift_apply_ld_bf16", t.data_ptr(), C, t.data_ptr(), C, M, C, ptr(ss[0]), ptr(ss[1]), act, stream_ptr())
    return t
