"""DenseNet frame encoder (torchvision densenet121 / 169 / 201 with classifier -> Identity: the default backbone of
lrcn/lrcn.py:196-209 and lrcn/rgb_lrcn.py:180-193) on the B200 kernels.

Layout: one channel-concatenated NHWC bf16 buffer per dense block (row stride = the block's final channel count);
`torch.cat` of the reference (torchvision `_DenseLayer.forward`) becomes "append `growth` channels in place".  A feature's
batch statistics are taken ONCE, in the epilogue of the conv that produces it, into a [2, C_final] table; each of the many
BatchNorms that later read the feature only finalises its own (gamma, beta, running stats) against that table.
Per dense layer: BN1+ReLU of the concatenated input (row-strided element kernel) -> 1x1 conv as a tcgen05 GEMM with
the statistics of its output in the epilogue -> BN2+ReLU in place -> 3x3 im2col-TMA conv (statistics in the epilogue)
-> slice copy into the block buffer."""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .ops import graph_epoch as _graph_epoch
from .ops import BF16, F32, conv2d_nhwc, gemm_tn, pack_stem_weight, scale_shift_apply, stem_conv

SUPPORTED = ("densenet121", "densenet169", "densenet201")


def _mom(bn):
    return bn.momentum if bn.momentum is not None else 0.1


class DenseNetRunner:
    def __init__(self, net):
        self.net = net
        self._wcache = None
        self._wkey = None
        self._convs = None
        self._params = None
        self._graphs = {}
        self.use_graph = False
        self.fold_bn1 = os.environ.get("B2_DENSE_FOLD", "1") == "1"   # BN1+ReLU in the 1x1 GEMM's A transform (tests run both)
        self.halo_conv = os.environ.get("B2_DENSE_HALO", "1") == "1"  # halo-tile 3x3 conv (tests run both)
        self.fuse_bn = None          # (ResNet-only switches; part of the CUDA-graph cache key of the shared graphed())
        self.stem_impl = None

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
                if n == "features.conv0":
                    cache[n] = pack_stem_weight(w)
                elif w.shape[2] == 1:
                    cache[n] = w.reshape(w.shape[0], w.shape[1]).to(BF16).contiguous()        # [Cout, C]
                else:
                    cache[n] = w.permute(0, 2, 3, 1).contiguous().to(BF16)                    # [Cout, R, S, C]
            self._wcache, self._wkey = cache, key
        return self._wcache

    def graphed(self, x, training: bool):
        """CUDA-graph replay of the frozen encoder pass (see ResNetRunner.graphed: ~430 launches -> one)."""
        from .backbone import ResNetRunner
        return ResNetRunner.graphed(self, x, training)

    @staticmethod
    def _finalize(bn, s_sum, s_sq, count, train, ss, C):
        """scale -> ss[0, :C], shift -> ss[1, :C] (ss rows are padded; raw addresses)."""
        call("b2_bn_finalize_nhwc", s_sum if train else 0, s_sq if train else 0, bn.weight.data_ptr(), bn.bias.data_ptr(),
             bn.running_mean.data_ptr(), bn.running_var.data_ptr(), count, float(bn.eps), float(_mom(bn)), int(train),
             ss.data_ptr(), ss.data_ptr() + ss.stride(0) * 4, C, stream_ptr())

    def __call__(self, x, training: bool, return_stages: bool = False):
        """x: [N,3,H,W] fp32 / bf16 NCHW -> [N, feat] fp32 (relu(features) -> global average pool -> flatten)."""
        _lib.require_device()
        net = self.net
        if self._params is None:
            self._params = list(net.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._params):
            # lrcn.py / rgb_lrcn.py default: FINETUNE = True leaves the whole densenet121 trainable
            from .densenet_train import encode_trainable
            return encode_trainable(self, x, training)
        y = self.stem(x, bool(training))
        stages = [y.float().mean(dim=(1, 2))] if return_stages else None
        feat = self.trunk(y, bool(training), stages=stages)
        if training:
            bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
            torch._foreach_add_([b.num_batches_tracked for b in bns if b.num_batches_tracked is not None], 1)
        if return_stages:
            return feat, stages
        return feat

    def stem(self, x, train):
        """conv0 7x7/2 -> norm0 -> relu -> maxpool 3x3/2 (the ResNet stem kernels) -> [N, H/4, W/4, 64] bf16."""
        x = x.contiguous()
        N, Cin, H, W = x.shape
        assert Cin == 3, "frame encoder expects RGB frames"
        dev = x.device
        w = self._weights()
        st = stream_ptr()
        s0 = torch.zeros(2 * 64, device=dev, dtype=F32)
        raw = stem_conv(x, w["features.conv0"], stats=(s0[:64], s0[64:]) if train else None)
        P, Q = raw.shape[1], raw.shape[2]
        Hc, Wc = (P + 2 - 3) // 2 + 1, (Q + 2 - 3) // 2 + 1
        y = torch.empty((N, Hc, Wc, 64), device=dev, dtype=BF16)
        bn0 = self.net.features.norm0
        call("b2_bn_relu_maxpool_nhwc", raw.data_ptr(), y.data_ptr(), N, P, Q, 64, ptr(s0[:64] if train else None),
             ptr(s0[64:] if train else None), bn0.weight.data_ptr(), bn0.bias.data_ptr(), bn0.running_mean.data_ptr(),
             bn0.running_var.data_ptr(), float(bn0.eps), float(_mom(bn0)), int(train), st)
        return y

    def trunk(self, y, train, stages=None, saved=None):
        """Dense blocks, transitions, norm5 + ReLU + global average pool on the stem output y [N,H,W,64] bf16.
        saved: a list that receives what the backward needs (densenet_train.py); nothing is kept otherwise."""
        dev = y.device
        w = self._weights()
        f = self.net.features
        st = stream_ptr()
        N, Hc, Wc, C = y.shape
        src, src_ld = y, C                                # the tensor (and row stride) holding the block's input features
        mods = [(n, m) for n, m in f.named_children() if n.startswith(("denseblock", "transition", "norm5"))]
        X = S = ss = None
        for mi, (name, mod) in enumerate(mods):
            if name.startswith("denseblock"):
                layers = list(mod.children())
                growth = layers[0].conv2.out_channels
                Cfin = C + len(layers) * growth
                M = N * Hc * Wc
                if X is None or X.shape[-1] != Cfin or X.shape[1] != Hc:      # the transition may have pre-allocated it
                    X = torch.empty((N, Hc, Wc, Cfin), device=dev, dtype=BF16)
                    call("b2_scale_shift_apply_ld_bf16", src.data_ptr(), src_ld, X.data_ptr(), Cfin, M, C, 0, 0, 0, st)
                S = torch.zeros((2, Cfin), device=dev, dtype=F32)            # per-feature sum | sumsq
                if train:
                    call("b2_colstats_ld_bf16", X.data_ptr(), Cfin, M, C, S.data_ptr(), S.data_ptr() + 4 * Cfin, st)
                ss = torch.zeros((2, (Cfin + 63) // 64 * 64), device=dev, dtype=F32)
                mid = layers[0].conv1.out_channels
                smid = torch.zeros((len(layers), 2, mid), device=dev, dtype=F32)
                ssmid = torch.empty((2, mid), device=dev, dtype=F32)
                pfx = "features." + name + "."
                rec = []
                for k, (lname, layer) in enumerate(mod.named_children()):
                    Ct = C + k * growth
                    self._finalize(layer.norm1, S.data_ptr(), S.data_ptr() + 4 * Cfin, M, train, ss, Ct)
                    w1 = w[pfx + lname + ".conv1"]
                    if self.fold_bn1:
                        # BN1 + ReLU of the concatenated input folded into the GEMM's A-tile transform: X is read once,
                        # the normalised copy is never written
                        y1 = torch.empty((M, mid), device=dev, dtype=BF16)
                        call("b2_gemm_bn_bf16_tn", X.data_ptr(), Cfin, w1.data_ptr(), w1.stride(0), y1.data_ptr(), mid, M, mid, Ct,
                             ss.data_ptr(), ss.data_ptr() + ss.stride(0) * 4, 1, smid[k, 0].data_ptr() if train else 0,
                             smid[k, 1].data_ptr() if train else 0, st)
                    else:
                        a1 = torch.empty((M, Ct), device=dev, dtype=BF16)
                        call("b2_scale_shift_apply_ld_bf16", X.data_ptr(), Cfin, a1.data_ptr(), Ct, M, Ct, ss.data_ptr(),
                             ss.data_ptr() + ss.stride(0) * 4, 1, st)
                        y1 = gemm_tn(a1, w1, out_dtype=BF16, stats=(smid[k, 0], smid[k, 1]) if train else None)
                        del a1
                    self._finalize(layer.norm2, smid[k, 0].data_ptr(), smid[k, 1].data_ptr(), M, train, ssmid, mid)
                    a2 = torch.empty_like(y1) if saved is not None else y1
                    scale_shift_apply(y1, ssmid[0], ssmid[1], relu=True, out=a2)
                    w2 = w[pfx + lname + ".conv2"]
                    if self.halo_conv and _lib.lib().b2_conv3x3_halo_dense_supported(N, Hc, Wc, mid, growth):
                        # halo-tile conv: the input halo is fetched once for all 9 taps (the im2col path re-reads it 9
                        # times and is L2 -> SM bound at 32 output channels); the 32 new channels go straight into X
                        y2 = None
                        call("b2_conv3x3_halo_dense_bf16", a2.data_ptr(), N, Hc, Wc, mid, w2.data_ptr(), growth,
                             X.data_ptr() + 2 * Ct, Cfin, S.data_ptr() + 4 * Ct if train else 0,
                             S.data_ptr() + 4 * (Cfin + Ct) if train else 0, st)
                    else:
                        y2 = conv2d_nhwc(a2.view(N, Hc, Wc, mid), w2, 1, 1,
                                         stats=(S.data_ptr() + 4 * Ct, S.data_ptr() + 4 * (Cfin + Ct)) if train else None)
                        call("b2_scale_shift_apply_ld_bf16", y2.data_ptr(), growth, X.data_ptr() + 2 * Ct, Cfin, M, growth, 0, 0, 0, st)
                    if saved is not None:
                        rec.append((pfx + lname, layer, y1, a2, smid[k]))
                    del y1, y2, a2
                if saved is not None:
                    saved.append(("block", X, S, C, growth, rec))
                C = Cfin
                if stages is not None:
                    stages.append(X.float().mean(dim=(1, 2)))
            elif name.startswith("transition"):
                M = N * Hc * Wc
                self._finalize(mod.norm, S.data_ptr(), S.data_ptr() + 4 * C, M, train, ss, C)
                a = torch.empty((M, C), device=dev, dtype=BF16)
                call("b2_scale_shift_apply_ld_bf16", X.data_ptr(), C, a.data_ptr(), C, M, C, ss.data_ptr(),
                     ss.data_ptr() + ss.stride(0) * 4, 1, st)
                yt = gemm_tn(a, w["features." + name + ".conv"], out_dtype=BF16)
                Cn = yt.shape[1]
                del a
                # AvgPool2d(2, 2) straight into the next block's concatenated buffer
                nl = list(mods[mi + 1][1].children())
                Cfin = Cn + len(nl) * nl[0].conv2.out_channels
                Hn, Wn = Hc // 2, Wc // 2
                Xn = torch.empty((N, Hn, Wn, Cfin), device=dev, dtype=BF16)
                call("b2_avgpool2x2_nhwc_bf16", yt.data_ptr(), Xn.data_ptr(), Cfin, N, Hc, Wc, Cn, st)
                del yt
                if saved is not None:
                    saved.append(("transition", "features." + name, mod, X, S, Hc, Wc, C, Cn))
                X, Hc, Wc, C = Xn, Hn, Wn, Cn
                src, src_ld = X, Cfin
            else:                                           # norm5 -> F.relu -> adaptive_avg_pool2d(1) -> flatten
                M = N * Hc * Wc
                self._finalize(mod, S.data_ptr(), S.data_ptr() + 4 * C, M, train, ss, C)
                a = torch.empty((N, Hc * Wc, C), device=dev, dtype=BF16)
                call("b2_scale_shift_apply_ld_bf16", X.data_ptr(), C, a.data_ptr(), C, M, C, ss.data_ptr(),
                     ss.data_ptr() + ss.stride(0) * 4, 1, st)
                feat = torch.empty((N, C), device=dev, dtype=F32)
                call("b2_avgpool_nhwc", a.data_ptr(), feat.data_ptr(), 0, N, Hc * Wc, C, st)
                if saved is not None:
                    saved.append(("norm5", mod, X, S, a, Hc, Wc, C))
        return feat
