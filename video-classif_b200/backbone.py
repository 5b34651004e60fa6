"""Frozen torchvision-ResNet frame encoder executed on the sm_100a kernels.

The reference builds `getattr(torchvision.models, name)(pretrained=True)`, sets fc=Identity and
freezes every parameter (medsos_lrcn/src/models.py:133-145, lrcn/ucf50-lrcn.py:263-272) but keeps
the module in .train() mode (train_eval.py:12), so every BatchNorm uses BATCH statistics over all
B*T frames and updates its running stats.  This runner reproduces exactly that, NHWC bf16:

  stem   : 7x7/2 patches (b2_stem_im2col) -> tcgen05 GEMM (+column stats) -> BN+ReLU+maxpool
  blocks : every conv is the tcgen05 implicit GEMM (1x1 = plain TMA GEMM, 3x3 / strided =
           im2col-mode TMA) whose epilogue already emits the BN batch statistics; one
           bn_apply pass normalises, adds the shortcut (identity or down-sample BN) and ReLUs
  head   : global average pool -> [N, C] fp32

The torchvision module is only a PARAMETER CONTAINER (state_dict keys / shapes / default init are
the reference's); its forward is never called."""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .ops import graph_epoch as _graph_epoch
from .ops import (BF16, F32, GRAM_CHANNELS, conv1x1_gram_bnstats, conv2d_bn_nhwc, conv3x3_halo_bn, conv3x3_halo_supported, conv2d_nhwc, gemm_tn, pack_stem_weight, stem_pack,
                  scale_shift_apply, stem_conv)

SUPPORTED = ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152")
STEM_KP = 168   # stem patch columns: (c*7 + r)*8 + s, filter rows padded from 7 to 8 taps


def make_backbone(name: str, pretrained: bool = False):
    """torchvision constructor + the reference's head surgery (fc/classifier -> Identity).
    Returns (module, feature_size)."""
    import torchvision.models as tvm
    from .densenet import SUPPORTED as DENSE
    from .mobilenet import SUPPORTED as MOBILE
    if name in MOBILE:
        net = getattr(tvm, name)(weights="IMAGENET1K_V1" if pretrained else None)
        feat = net.classifier[-1].in_features                  # models.py:138-140: Sequential classifier
        net.classifier = torch.nn.Identity()
        return net, feat
    if name in DENSE:
        net = getattr(tvm, name)(weights="IMAGENET1K_V1" if pretrained else None)
        feat = net.classifier.in_features
        net.classifier = torch.nn.Identity()
        return net, feat
    if name not in SUPPORTED:
        raise NotImplementedError(
            f"cnn_backbone={name!r}: the B200 kernels run the torchvision backbones {SUPPORTED}, {DENSE} and {MOBILE} "
            "(other torchvision families are listed under 'next' in DESIGN.md)")
    net = getattr(tvm, name)(weights="IMAGENET1K_V1" if pretrained else None)
    feat = net.fc.in_features
    net.fc = torch.nn.Identity()
    return net, feat


def make_runner(net):
    """The kernel-side executor of a backbone module built by make_backbone()."""
    if type(net).__name__ == "MobileNetV2":
        from .mobilenet import MobileNetRunner
        return MobileNetRunner(net)
    if hasattr(net, "features"):
        from .densenet import DenseNetRunner
        return DenseNetRunner(net)
    return ResNetRunner(net)


class ResNetRunner:
    FUSE_BN = os.environ.get("B2_FUSE_BN", "1") == "1"   # class default; tests run both settings
    def __init__(self, net):
        self.net = net
        self._wcache = None
        self._wkey = None
        self._convs = None
        self._bns = None
        self._frozen = None
        self._graphs = {}
        self.use_graph = False        # models route through graphed() when set (enable_encoder_graph())
        self.stem_impl = "direct"     # "im2col": patch matrix + plain GEMM (A/B parity tests)
        self.fuse_bn = ResNetRunner.FUSE_BN   # BatchNorms folded into the conv kernels vs stand-alone bn_apply passes
        self.gram_wide = os.environ.get("B2_GRAM_WIDE", "0") == "1"   # opt-in: Gram-form BN3 statistics for 256-channel conv3 inputs (measured: 50 us vs 35 us for the statistics-only conv pass at layer3 -- the reduce + finalise kernels dominate)
        self._ident = None

    def _identity(self, dev):
        """(scale, shift) = (1, 0) over 256 channels: the Gram statistics kernel always applies its A transform."""
        if self._ident is None or self._ident[0].device != dev:
            self._ident = (torch.ones(256, device=dev, dtype=F32), torch.zeros(256, device=dev, dtype=F32))
        return self._ident

    # ---- weights in kernel layout: [Cout, R, S, C] bf16 (K-major), stem [64, STEM_KP] ----
    def _weights(self):
        convs = self._convs          # module list is fixed after construction: walk the tree once, not per call
        if convs is None:
            convs = self._convs = [(n, m) for n, m in self.net.named_modules() if isinstance(m, torch.nn.Conv2d)]
        key = tuple((m.weight.data_ptr(), m.weight._version) for _, m in convs)
        if any(m.weight.requires_grad for _, m in convs):      # (a replayed train-step graph updates weights without version bumps)
            key += (_graph_epoch(),)
        if self._wkey != key:
            cache = {}
            for n, m in convs:
                w = m.weight.detach()
                if n == "conv1":
                    wk = torch.zeros((w.shape[0], 3, 7, 8), device=w.device, dtype=BF16)
                    wk[:, :, :, :7] = w.to(BF16)                  # [Cout, c, r, s] with a zero 8th tap
                    cache["conv1.im2col"] = wk.reshape(w.shape[0], STEM_KP)
                    # direct stem kernel: k = r*32 + s*4 + c (s = 7, c = 3 zero), stored [k/8][Cout][8]
                    wk = pack_stem_weight(w)
                else:
                    wk = w.permute(0, 2, 3, 1).contiguous().to(BF16)
                cache[n] = wk
            self._wcache, self._wkey = cache, key
        return self._wcache

    # ---- CUDA-graph replay of the whole encoder pass (static shapes, frozen weights) ---------------------------
    def graphed(self, x, training: bool):
        """Same result as self(x, training), replayed from a CUDA graph captured on first use for this input shape /
        mode / weight version: one graph launch instead of ~110 kernel launches from Python per pass.  The captured
        pass reads a static copy of x and writes static activations; the returned features live in one of 4 rotating
        buffers (valid until the 4th following call)."""
        _lib.require_device()
        if torch.cuda.is_current_stream_capturing():       # inside a captured train step the pass is part of THAT graph
            return self(x, training)
        self._weights()
        key = (tuple(x.shape), x.dtype, x.device.index, bool(training), self._wkey, self.fuse_bn, self.stem_impl)
        entry = self._graphs.get(key)
        if entry is None:
            bufs = [b for b in self.net.buffers()]
            saved = [b.detach().clone() for b in bufs]          # the warm-up pass must not advance the running stats
            static_x = x.detach().clone()
            # ResNet direct stem: its operand packing (a pure re-layout of the frames) stays outside the graph and reads the
            # caller's tensor directly, so no copy of the frames into a static buffer is needed per call
            packs = (bool(getattr(self, "PACKED_STEM", False)) and self.stem_impl == "direct"
                     and os.environ.get("B2_NO_PACKED_STEM") is None)
            static_xp = stem_pack(static_x) if packs else None
            kw = {"packed": static_xp} if packs else {}
            self(static_x, training, **kw)                       # warm-up: function attributes, allocator pools
            for b, sv in zip(bufs, saved):
                b.copy_(sv)
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            # kernel nodes inherit the capture stream's priority: B2_ENC_PRIORITY=-1 lets the encoder's CTAs go first and
            # the small tail kernels of the previous batch fill the gaps of its partial waves
            prio = int(os.environ.get("B2_ENC_PRIORITY", "0"))
            with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=x.device, priority=prio)):
                static_feat = self(static_x, training, **kw)
            # results rotate through a small ring of persistent buffers: no per-call allocation (a tensor handed from
            # the encoder's stream to the consumer's stream every step made the caching allocator hold blocks back)
            ring = [torch.empty_like(static_feat) for _ in range(4)]
            entry = [graph, static_x, static_feat, _lib.launch_count() - n0, ring, 0, static_xp]
            self._graphs[key] = entry
        graph, static_x, static_feat, n_launch, ring, idx, static_xp = entry
        if static_xp is not None:
            stem_pack(x, out=static_xp)                          # one kernel: caller's frames -> the graph's packed stem operand
        elif x.data_ptr() != static_x.data_ptr():
            static_x.copy_(x)
        graph.replay()
        call("b2_add_launch_count", n_launch)
        out = ring[idx]
        entry[5] = (idx + 1) % len(ring)
        out.copy_(static_feat)
        return out

    def static_input(self, shape, dtype, device, training: bool):
        """The static input tensor of the graph captured for this frame-batch shape / mode (None before its first replay):
        a producer that writes its frames THERE (ingest_batch(..., out=...)) saves the per-step copy into the graph."""
        for key, entry in self._graphs.items():
            if key[0] == tuple(shape) and key[1] == dtype and key[2] == torch.device(device).index and key[3] == bool(training) \
                    and key[4] == self._wkey:
                return entry[1]
        return None

    def _bn(self, x, bn, stats, count, train, relu=True, res_mode=0, res=None, rbn=None, rstats=None):
        rows = x.numel() // x.shape[-1]
        C = x.shape[-1]
        rs1 = rs2 = rg = rb = rrm = rrv = None
        if res_mode == 2:
            rs1, rs2 = (rstats if train else (None, None))
            rg, rb, rrm, rrv = rbn.weight, rbn.bias, rbn.running_mean, rbn.running_var
        s1, s2 = stats if train else (None, None)
        mom = bn.momentum if bn.momentum is not None else 0.1
        call("b2_bn_apply_nhwc", x.data_ptr(), x.data_ptr(), rows, C, ptr(s1), ptr(s2), bn.weight.data_ptr(),
             bn.bias.data_ptr(), ptr(bn.running_mean), ptr(bn.running_var), res_mode, ptr(res), ptr(rs1), ptr(rs2),
             ptr(rg), ptr(rb), ptr(rrm), ptr(rrv), count, float(bn.eps), float(mom), int(train), int(relu),
             stream_ptr())
        return x   # in place

    def _pool(self, y):
        Nn, Hh, Ww, C = y.shape
        feat = torch.empty((Nn, C), device=y.device, dtype=F32)
        call("b2_avgpool_nhwc", y.data_ptr(), feat.data_ptr(), 0, Nn, Hh * Ww, C, stream_ptr())
        return feat

    PACKED_STEM = True      # graphed(): the stem's operand packing runs OUTSIDE the graph, straight from the caller's frames

    def __call__(self, x, training: bool, return_stages: bool = False, stop_at=None, packed=None):
        """x: [N,3,H,W] fp32 or bf16 (NCHW, the reference's frame tensor) -> [N, feat] fp32.
        packed: the frames already in the direct stem's packed layout (ops.stem_pack; x then only gives the shape).
        return_stages=True also returns the spatial means after the stem and each stage (tests).
        stop_at=(layer, block): return the NHWC bf16 activation entering that block (the frozen prefix of a partially
        trainable encoder, backbone_train.py)."""
        _lib.require_device()
        net = self.net
        if self._frozen is None:
            self._frozen = list(net.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._frozen):
            # (partially) trainable encoder: rgb_lrcn.py:208-245 / lrcn.py:246-283 -- autograd path with backward kernels
            from .backbone_train import encode_trainable
            return encode_trainable(self, x, training)
        x = x.contiguous()
        N, Cin, H, W = x.shape
        assert Cin == 3, "frame encoder expects RGB frames"
        dev = x.device
        w = self._weights()
        if self._bns is None:
            bl = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
            self._bns = (bl, max(b.num_features for b in bl), {id(b): i for i, b in enumerate(bl)})
        bns, maxc, bn_index = self._bns
        train = bool(training)
        # per BatchNorm: [sum | sumsq | scale | shift] (maxc floats each) + a finalisation counter word.  The slots are
        # handed to the kernels as raw device addresses (no per-launch tensor views on the host)
        stat_buf = torch.zeros((len(bns), 4 * maxc + 4), device=dev, dtype=F32)
        row_bytes = (4 * maxc + 4) * 4
        sb_base = stat_buf.data_ptr()

        def slot(bn, k):
            return sb_base + bn_index[id(bn)] * row_bytes + k * maxc * 4

        def stats_of(bn):
            if not train:
                return None
            return (slot(bn, 0), slot(bn, 1))

        def ss_of(bn):                         # (scale, shift) written by the producer's finalisation tail
            return (slot(bn, 2), slot(bn, 3))

        def fin_of(bn):
            if not train:
                return None
            return (bn.weight, bn.bias, bn.running_mean, bn.running_var, slot(bn, 2), slot(bn, 3),
                    slot(bn, 4), bn.eps, bn.momentum if bn.momentum is not None else 0.1)

        fuse = self.fuse_bn
        if fuse and not train:                 # eval: scale/shift straight from the running statistics
            for b in bns:
                sc, sh = ss_of(b)
                call("b2_bn_finalize_nhwc", 0, 0, b.weight.data_ptr(), b.bias.data_ptr(), b.running_mean.data_ptr(),
                     b.running_var.data_ptr(), 1, float(b.eps), 0.1, 0, sc, sh, b.num_features,
                     stream_ptr())

        st = stream_ptr()
        # ---- stem ----
        P, Q = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        s_stem = stats_of(net.bn1)
        if self.stem_impl == "direct":
            raw = stem_conv(x, w["conv1"], stats=s_stem, xp=packed)
        else:                                   # patch-matrix path (kept for A/B parity tests)
            A = torch.empty((N * P * Q, STEM_KP), device=dev, dtype=BF16)
            call("b2_stem_im2col", x.data_ptr(), int(x.dtype == BF16), A.data_ptr(), N, H, W, STEM_KP, st)
            raw = gemm_tn(A, w["conv1.im2col"], out_dtype=BF16, stats=s_stem)
            del A
        P2, Q2 = (P + 2 - 3) // 2 + 1, (Q + 2 - 3) // 2 + 1
        y = torch.empty((N, P2, Q2, 64), device=dev, dtype=BF16)
        s = stats_of(net.bn1)
        mom = net.bn1.momentum if net.bn1.momentum is not None else 0.1
        call("b2_bn_relu_maxpool_nhwc", raw.data_ptr(), y.data_ptr(), N, P, Q, 64, ptr(s[0] if s else None),
             ptr(s[1] if s else None), net.bn1.weight.data_ptr(), net.bn1.bias.data_ptr(),
             net.bn1.running_mean.data_ptr(), net.bn1.running_var.data_ptr(), float(net.bn1.eps), float(mom),
             int(train), st)
        del raw
        stages = [self._pool(y)] if return_stages else None
        # ---- residual stages ----
        for li in range(1, 5):
            layer = getattr(net, f"layer{li}")
            for bi, blk in enumerate(layer):
                pfx = f"layer{li}.{bi}"
                if stop_at == (li, bi):
                    if train:
                        done = [net.bn1] + [m for lj in range(1, 5) for bj, b in enumerate(getattr(net, f"layer{lj}"))
                                            if (lj, bj) < stop_at for m in b.modules() if isinstance(m, torch.nn.BatchNorm2d)]
                        torch._foreach_add_([b.num_batches_tracked for b in done if b.num_batches_tracked is not None], 1)
                    return y
                stride = blk.stride if isinstance(blk.stride, int) else blk.stride[0]
                bottleneck = hasattr(blk, "conv3")
                if fuse:
                    # BatchNorms folded into the conv kernels: statistics + finalisation in the producer,
                    # normalise+ReLU in the consumer's A-tile transform, BN3 + shortcut + ReLU in the epilogue
                    # of a second conv3 pass (gemm_tc.cu); no stand-alone BN pass over the activations
                    ds = blk.downsample is not None
                    if bottleneck:
                        r1 = conv2d_bn_nhwc(y, w[pfx + ".conv1"], 1, 0, stats=stats_of(blk.bn1), fin=fin_of(blk.bn1))
                        # a 3x3 consumer would re-normalise every input pixel 9 times (once per tap) and
                        # doubles the shared-memory traffic per k-block: one in-place pass over this small
                        # tensor is cheaper; the 1x1 consumer (conv3) takes the A transform
                        if conv3x3_halo_supported(r1, w[pfx + ".conv2"], stride, 1):
                            # 64-channel stage: halo-tile conv, BN1+ReLU applied to the halo once per tile
                            r2 = conv3x3_halo_bn(r1, w[pfx + ".conv2"], a=ss_of(blk.bn1), stats=stats_of(blk.bn2),
                                                 fin=fin_of(blk.bn2))
                        else:
                            sc1, sh1 = ss_of(blk.bn1)
                            scale_shift_apply(r1, sc1, sh1, relu=True)
                            r2 = conv2d_bn_nhwc(r1, w[pfx + ".conv2"], stride, 1, stats=stats_of(blk.bn2),
                                                fin=fin_of(blk.bn2))
                        del r1
                        if ds:
                            dbn = blk.downsample[1]
                            rd = conv2d_bn_nhwc(y, w[pfx + ".downsample.0"], stride, 0, stats=stats_of(dbn),
                                                fin=fin_of(dbn))
                        # BN2 + ReLU on conv3's input: folded into conv3's A-tile transform for the narrow stages; at
                        # 256+ channels the transform warps are the serial stage of a short-K pipeline (measured:
                        # 62 -> 39 us statistics pass, 67 -> 46 us output pass at layer4), so the small tensor is
                        # normalised in place once instead
                        a2 = ss_of(blk.bn2)
                        if r2.shape[-1] >= 256:
                            scale_shift_apply(r2, a2[0], a2[1], relu=True)
                            a2 = None
                        if train and r2.shape[-1] in GRAM_CHANNELS and (a2 is not None or self.gram_wide):
                            # BN3 statistics from the Gram matrix of conv3's (transformed) input: one HBM-bound
                            # pass over the small tensor instead of a full conv3 whose output is thrown away
                            # (256 channels: the input is already normalised in place -> identity transform)
                            conv1x1_gram_bnstats(r2, w[pfx + ".conv3"], a2 if a2 is not None else self._identity(dev),
                                                 fin_of(blk.bn3), a_relu=a2 is not None)
                        elif train:            # statistics-only pass of conv3
                            conv2d_bn_nhwc(r2, w[pfx + ".conv3"], 1, 0, a=a2, stats=stats_of(blk.bn3),
                                           fin=fin_of(blk.bn3), store=False)
                        y = conv2d_bn_nhwc(r2, w[pfx + ".conv3"], 1, 0, a=a2, o=ss_of(blk.bn3),
                                           res=rd if ds else y, r=ss_of(dbn) if ds else None, relu=True)
                        del r2
                    else:
                        if conv3x3_halo_supported(y, w[pfx + ".conv1"], stride, 1):
                            r1 = conv3x3_halo_bn(y, w[pfx + ".conv1"], stats=stats_of(blk.bn1), fin=fin_of(blk.bn1))
                        else:
                            r1 = conv2d_bn_nhwc(y, w[pfx + ".conv1"], stride, 1, stats=stats_of(blk.bn1),
                                                fin=fin_of(blk.bn1))
                        if conv3x3_halo_supported(r1, w[pfx + ".conv2"], 1, 1):
                            r2 = conv3x3_halo_bn(r1, w[pfx + ".conv2"], a=ss_of(blk.bn1), stats=stats_of(blk.bn2),
                                                 fin=fin_of(blk.bn2))
                        else:
                            sc1, sh1 = ss_of(blk.bn1)
                            scale_shift_apply(r1, sc1, sh1, relu=True)
                            r2 = conv2d_bn_nhwc(r1, w[pfx + ".conv2"], 1, 1, stats=stats_of(blk.bn2), fin=fin_of(blk.bn2))
                        del r1
                        if ds:
                            dbn = blk.downsample[1]
                            rd = conv2d_bn_nhwc(y, w[pfx + ".downsample.0"], stride, 0, stats=stats_of(dbn),
                                                fin=fin_of(dbn))
                        sc, sh = ss_of(blk.bn2)
                        y = scale_shift_apply(r2, sc, sh, res=rd if ds else y, r=ss_of(dbn) if ds else None, relu=True)
                    continue
                if bottleneck:
                    o = conv2d_nhwc(y, w[pfx + ".conv1"], 1, 0, stats_of(blk.bn1))
                    o = self._bn(o, blk.bn1, stats_of(blk.bn1), o.numel() // o.shape[-1], train)
                    o = conv2d_nhwc(o, w[pfx + ".conv2"], stride, 1, stats_of(blk.bn2))
                    o = self._bn(o, blk.bn2, stats_of(blk.bn2), o.numel() // o.shape[-1], train)
                    o = conv2d_nhwc(o, w[pfx + ".conv3"], 1, 0, stats_of(blk.bn3))
                    last_bn = blk.bn3
                else:
                    o = conv2d_nhwc(y, w[pfx + ".conv1"], stride, 1, stats_of(blk.bn1))
                    o = self._bn(o, blk.bn1, stats_of(blk.bn1), o.numel() // o.shape[-1], train)
                    o = conv2d_nhwc(o, w[pfx + ".conv2"], 1, 1, stats_of(blk.bn2))
                    last_bn = blk.bn2
                cnt = o.numel() // o.shape[-1]
                if blk.downsample is not None:
                    dbn = blk.downsample[1]
                    d = conv2d_nhwc(y, w[pfx + ".downsample.0"], stride, 0, stats_of(dbn))
                    y = self._bn(o, last_bn, stats_of(last_bn), cnt, train, res_mode=2, res=d, rbn=dbn,
                                 rstats=stats_of(dbn))
                else:
                    y = self._bn(o, last_bn, stats_of(last_bn), cnt, train, res_mode=1, res=y)
            if return_stages:
                stages.append(self._pool(y))
        # ---- head ----
        feat = self._pool(y)
        if train:
            torch._foreach_add_([b.num_batches_tracked for b in bns if b.num_batches_tracked is not None], 1)
        if return_stages:
            return feat, stages
        return feat
