"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the LRCN clip-classification hot path.

A plain CPU restatement (numpy for integer/byte work, torch-CPU fp32/fp64 primitives for the
floating-point algebra) of what the reference computes on the path
    frames(u8) -> resize -> /255 -> [B,T,C,H,W] -> per-frame CNN -> LSTM over T -> head -> logits
forward and (through torch autograd on these explicit formulas) backward.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product package (video-classif_b200/) never does.

Pinned dependency whose published algorithms are restated here: the reference does all of its
arithmetic through torch==2.4.1 / torchvision==0.19.1 (pins: medsos_lrcn/build/worker.dockerfile:36-38)
and OpenCV (`cv2.resize`, loader_data.py:162).  The reference has NO tests / golden vectors for
this path, so the oracle is pinned instead against outputs of the reference's own
classes run in the authoring container: tests/golden/*.npz, produced by
tests/golden/make_golden.py (committed) and checked by tests/test_oracle_golden.py.

Reference call sites each function follows are cited as file:line relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# 1. Frame-index sampling (integer, bit-exact)
# --------------------------------------------------------------------------------------


def uniform_sampling_indices(n: int, T: int) -> List[int]:
    """medsos_lrcn/src/loader_data.py:35-41 (same code: lrcn/ucf50-lrcn.py:84-90).

    n <= T returns every frame; otherwise interval = n // T and the first T of range(0,n,interval).
    """
    if n <= T:
        return list(range(n))
    interval = n // T
    return list(range(0, n, interval))[:T]


def duplicate_indices(idx: Sequence[int], T: int) -> List[int]:
    """loader_data.py:43-51: short clips are cycled until T frames are reached."""
    idx = list(idx)
    if len(idx) >= T:
        return idx[:T]
    if len(idx) == 0:
        raise ValueError("duplicate_frames on an empty clip never terminates in the reference")
    out: List[int] = []
    while len(out) < T:
        out.extend(idx)
    return out[:T]


def medsos_indices(n: int, T: int) -> List[int]:
    """loader_data.py:171-178: uniform_sampling then duplicate_frames when short."""
    idx = uniform_sampling_indices(n, T)
    if len(idx) < T:
        idx = duplicate_indices(idx, T)
    return idx


def seek_indices(n: int, T: int) -> Optional[List[int]]:
    """lrcn/backup_ucf50.py:52-62 / notebook cell 2: clips with n < T are skipped (None);
    otherwise frame i*(n//T) for i in range(T)."""
    if n < T:
        return None
    interval = n // T
    return [i * interval for i in range(T)]


def crime_indices(n: int, T: int) -> List[int]:
    """lrcn/lrcn.py:151-155: n >= T -> range(0,n,n//T)[:T]; short clips are padded with all-zero
    frames, encoded here as index -1."""
    if n >= T:
        interval = n // T
        return list(range(0, n, interval))[:T]
    return list(range(n)) + [-1] * (T - n)


# --------------------------------------------------------------------------------------
# 2. Frame ingest: cv2.resize(INTER_LINEAR) on uint8, channel swap, /255, HWC -> CHW
# --------------------------------------------------------------------------------------

_COEF_BITS = 11
_COEF_ONE = 1 << _COEF_BITS


def _linear_coeffs(src: int, dst: int, vertical: bool) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """OpenCV imgproc resize.cpp, INTER_LINEAR coefficient tables for 8-bit images:
    f = (d+0.5)*scale-0.5 (float32), s = floor(f), weights rounded to 11-bit fixed point
    (saturate_cast<short>(w * 2048), round-half-even).  Horizontally a tap that falls off the
    image gets weight 0 and the index is clamped; vertically only the ROW INDICES are clamped
    (both taps keep their weights), which rounds differently by up to 1 LSB."""
    scale = float(src) / float(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if vertical:
        s1 = np.clip(s + 1, 0, src - 1)
        s = np.clip(s, 0, src - 1)
    else:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0.0
        s[hi] = src - 1
        s1 = np.minimum(s + 1, src - 1)
    w1 = np.rint(f.astype(np.float32) * np.float32(_COEF_ONE)).astype(np.int64)
    w0 = np.rint((np.float32(1.0) - f) * np.float32(_COEF_ONE)).astype(np.int64)
    return np.stack([s, s1], 0), w0, w1


def resize_bilinear_u8(src: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2.resize(frame, (out_w, out_h)) with the default INTER_LINEAR on uint8 HxWxC
    (call site: loader_data.py:162, ucf50-lrcn.py:372, backup_ucf50.py:64).

    Restates OpenCV's fixed-point path: horizontal pass to 11-bit-scaled ints, vertical pass
    ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2.  The exact 2x2 decimation case is routed by
    OpenCV to INTER_AREA ((a+b+c+d+2)>>2)."""
    assert src.dtype == np.uint8 and src.ndim == 3
    h, w, _ = src.shape
    if h == out_h and w == out_w:
        return src.copy()
    if h == 2 * out_h and w == 2 * out_w:
        s = src.astype(np.int64)
        acc = s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2]
        return ((acc + 2) >> 2).astype(np.uint8)
    xs, xw0, xw1 = _linear_coeffs(w, out_w, False)
    ys, yw0, yw1 = _linear_coeffs(h, out_h, True)
    s = src.astype(np.int64)
    hor = s[:, xs[0], :] * xw0[None, :, None] + s[:, xs[1], :] * xw1[None, :, None]
    r0 = hor[ys[0]]
    r1 = hor[ys[1]]
    out = (((yw0[:, None, None] * (r0 >> 4)) >> 16) + ((yw1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def ingest_clip(frames_u8: np.ndarray, out_h: int, out_w: int, swap_rb: bool = True,
                divisor: float = 255.0) -> np.ndarray:
    """[T,H0,W0,3] uint8 (decoder order, BGR) -> float32 [T,3,out_h,out_w].

    loader_data.py:162-163 (resize, BGR->RGB), :182 (`np.array(frames)/255.0` in float64, then
    float32 at :201), :112 (permute to CHW).  swap_rb=False / CHW is the crime path
    (lrcn/lrcn.py:136-142); divisor=1.0 is the UCF50 small-CNN path which never divides
    (backup_ucf50.py:66-68,101)."""
    T = frames_u8.shape[0]
    out = np.empty((T, 3, out_h, out_w), dtype=np.float32)
    for t in range(T):
        r = resize_bilinear_u8(frames_u8[t], out_h, out_w)
        if swap_rb:
            r = r[:, :, ::-1]
        v = r.astype(np.float32) if divisor == 1.0 else (r.astype(np.float64) / divisor).astype(np.float32)
        out[t] = v.transpose(2, 0, 1)
    return out


# --------------------------------------------------------------------------------------
# 3. Layer algebra (explicit formulas; torch-CPU conv/matmul as the arithmetic primitive)
# --------------------------------------------------------------------------------------


def batchnorm2d_train(x, weight, bias, running_mean=None, running_var=None, momentum=0.1,
                      eps=1e-5):
    """nn.BatchNorm2d in train mode (SURVEY a.1): biased variance normalises, unbiased variance
    feeds running_var; returns (y, new_running_mean, new_running_var)."""
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var_b = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
    y = (x - mean[None, :, None, None]) * torch.rsqrt(var_b + eps)[None, :, None, None]
    y = y * weight[None, :, None, None] + bias[None, :, None, None]
    new_rm = new_rv = None
    if running_mean is not None:
        var_u = var_b * (n / max(n - 1, 1))
        new_rm = (1 - momentum) * running_mean + momentum * mean.detach()
        new_rv = (1 - momentum) * running_var + momentum * var_u.detach()
    return y, new_rm, new_rv


def batchnorm2d_eval(x, weight, bias, running_mean, running_var, eps=1e-5):
    y = (x - running_mean[None, :, None, None]) * torch.rsqrt(running_var + eps)[None, :, None, None]
    return y * weight[None, :, None, None] + bias[None, :, None, None]


def layernorm(x, weight, bias, eps=1e-5):
    """nn.LayerNorm over the last dim, biased variance (models.py:148-152 'bn*' are LayerNorms)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * weight + bias


def gelu_exact(x):
    """F.gelu default (approximate='none'): 0.5*x*(1+erf(x/sqrt(2)))  (models.py:200-202)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def lstm_forward(x, params: Dict[str, torch.Tensor], hidden: int, num_layers: int,
                 bidirectional: bool, prefix: str = "lstm."):
    """nn.LSTM(batch_first=True, zero initial state, no dropout/proj) -- nb:169, models.py:156-158.

    Gate order i,f,g,o along dim 0 of weight_ih_l{k} [4H,in] / weight_hh_l{k} [4H,H];
    g = x W_ih^T + b_ih + h W_hh^T + b_hh; c' = sig(f) c + sig(i) tanh(g~); h' = sig(o) tanh(c').
    `_reverse` parameters run t = T-1..0; outputs are cat([fwd, bwd], -1) per step."""
    B, T, _ = x.shape
    inp = x
    for layer in range(num_layers):
        outs = []
        for d in range(2 if bidirectional else 1):
            sfx = f"l{layer}" + ("_reverse" if d == 1 else "")
            w_ih = params[f"{prefix}weight_ih_{sfx}"]
            w_hh = params[f"{prefix}weight_hh_{sfx}"]
            b = params[f"{prefix}bias_ih_{sfx}"] + params[f"{prefix}bias_hh_{sfx}"]
            h = x.new_zeros(B, hidden)
            c = x.new_zeros(B, hidden)
            hs = [None] * T
            order = range(T - 1, -1, -1) if d == 1 else range(T)
            for t in order:
                g = inp[:, t] @ w_ih.t() + h @ w_hh.t() + b
                i, f, gg, o = g.split(hidden, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                hs[t] = h
            outs.append(torch.stack(hs, dim=1))
        inp = torch.cat(outs, dim=-1) if bidirectional else outs[0]
    return inp


def gru_forward(x, params: Dict[str, torch.Tensor], hidden: int, num_layers: int, bidirectional: bool,
                prefix: str = "lstm."):
    """nn.GRU(batch_first=True, zero initial state) -- lrcn/backup_ucf50.py:126 (attribute `lstm`), medsos
    models.py:160-170.  Gate order r,z,n; r = sig(W_ir x + b_ir + W_hr h + b_hr), z likewise,
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)), h' = (1 - z) n + z h; `_reverse` runs t = T-1..0."""
    B, T, _ = x.shape
    inp = x
    for layer in range(num_layers):
        outs = []
        for d in range(2 if bidirectional else 1):
            sfx = f"l{layer}" + ("_reverse" if d == 1 else "")
            w_ih, w_hh = params[f"{prefix}weight_ih_{sfx}"], params[f"{prefix}weight_hh_{sfx}"]
            b_ih, b_hh = params[f"{prefix}bias_ih_{sfx}"], params[f"{prefix}bias_hh_{sfx}"]
            h = x.new_zeros(B, hidden)
            hs = [None] * T
            for t in (range(T - 1, -1, -1) if d == 1 else range(T)):
                gi = inp[:, t] @ w_ih.t() + b_ih
                gh = h @ w_hh.t() + b_hh
                ir, iz, i_n = gi.split(hidden, dim=1)
                hr, hz, hn = gh.split(hidden, dim=1)
                r = torch.sigmoid(ir + hr)
                z = torch.sigmoid(iz + hz)
                n = torch.tanh(i_n + r * hn)
                h = (1.0 - z) * n + z * h
                hs[t] = h
            outs.append(torch.stack(hs, dim=1))
        inp = torch.cat(outs, dim=-1) if bidirectional else outs[0]
    return inp


def small_cnn_lrcn_forward(params: Dict[str, torch.Tensor], x: torch.Tensor, hidden: int,
                           train: bool = True, lstm_layers: int = 2, gru: bool = False):
    """Notebook `LRCN.forward` (nb:174-193): conv1-bn1-relu, conv2-bn2-relu-pool,
    conv3-bn3-relu-pool, (dropout p=0 for parity), reshape(B,T,C*H*W) [channel-major],
    2-layer LSTM, flatten all T, fc.  Returns (logits, dict of new running stats)."""
    B, T, C, H, W = x.shape
    y = x.reshape(B * T, C, H, W)
    new_stats = {}
    for k in (1, 2, 3):
        y = F.conv2d(y, params[f"conv{k}.weight"], params[f"conv{k}.bias"], padding=1)
        if train:
            y, rm, rv = batchnorm2d_train(y, params[f"bn{k}.weight"], params[f"bn{k}.bias"],
                                          params[f"bn{k}.running_mean"], params[f"bn{k}.running_var"])
            new_stats[f"bn{k}.running_mean"] = rm
            new_stats[f"bn{k}.running_var"] = rv
        else:
            y = batchnorm2d_eval(y, params[f"bn{k}.weight"], params[f"bn{k}.bias"],
                                 params[f"bn{k}.running_mean"], params[f"bn{k}.running_var"])
        y = torch.relu(y)
        if k >= 2:
            y = F.max_pool2d(y, 2, 2)
    feat = y.reshape(B, T, -1)
    if gru:        # LRCN2 (lrcn/backup_ucf50.py:126,146): one bidirectional GRU layer stored as `lstm`
        out = gru_forward(feat, params, hidden, 1, True, prefix="lstm.")
    else:
        out = lstm_forward(feat, params, hidden, lstm_layers, False, prefix="lstm.")
    logits = out.reshape(B, -1) @ params["fc.weight"].t() + params["fc.bias"]
    return logits, new_stats


# ---- torchvision ResNet, restated functionally from a state_dict (train-mode BN) ----

_RESNET_CFG = {
    "resnet18": ("basic", [2, 2, 2, 2]),
    "resnet34": ("basic", [3, 4, 6, 3]),
    "resnet50": ("bottleneck", [3, 4, 6, 3]),
    "resnet101": ("bottleneck", [3, 4, 23, 3]),
    # test-only shallow variants (one block per stage: every block geometry, little chaos)
    "resnet10": ("basic", [1, 1, 1, 1]),
    "resnet14": ("bottleneck", [1, 1, 1, 1]),
}


def _bn(x, sd, name, train, new_stats):
    if train:
        y, rm, rv = batchnorm2d_train(x, sd[name + ".weight"], sd[name + ".bias"],
                                      sd[name + ".running_mean"], sd[name + ".running_var"])
        new_stats[name + ".running_mean"] = rm
        new_stats[name + ".running_var"] = rv
        return y
    return batchnorm2d_eval(x, sd[name + ".weight"], sd[name + ".bias"],
                            sd[name + ".running_mean"], sd[name + ".running_var"])


def _bf16(t):
    return t.bfloat16().float()


def resnet_features(sd: Dict[str, torch.Tensor], x: torch.Tensor, arch: str, train: bool = True,
                    prefix: str = "cnn_backbone.", emulate_bf16: bool = False, return_stages: bool = False):
    """torchvision.models.resnet{18,34,50,101} with fc=Identity (models.py:133-137): conv7x7/2,
    BN, ReLU, maxpool3x3/2, 4 stages of basic / bottleneck (v1.5: stride on the 3x3) blocks,
    global average pool.  Train-mode BN even when frozen (train_eval.py:12).

    emulate_bf16=True restates the SAME graph with bf16 storage at the points the B200 path rounds
    (conv operands, raw conv outputs, activation outputs; fp32 accumulation, fp32 BN statistics)
    -- the yardstick for how far any bf16 execution of this (chaotic, randomly initialised,
    batch-statistics) network may drift from the fp32 reference.
    return_stages=True additionally returns the spatial mean of the stem and of each stage."""
    kind, depths = _RESNET_CFG[arch]
    g = lambda k: sd[prefix + k]
    sdp = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    ns: Dict[str, torch.Tensor] = {}
    rnd = _bf16 if emulate_bf16 else (lambda t: t)

    def conv(inp, w, **kw):
        return rnd(F.conv2d(rnd(inp), rnd(w), None, **kw))

    stages = []
    y = conv(x, g("conv1.weight"), stride=2, padding=3)
    y = torch.relu(_bn(y, sdp, "bn1", train, ns))
    y = rnd(F.max_pool2d(y, 3, 2, 1))
    stages.append(y.mean(dim=(2, 3)))
    for si, depth in enumerate(depths):
        for bi in range(depth):
            p = f"layer{si + 1}.{bi}"
            stride = 2 if (si > 0 and bi == 0) else 1
            idt = y
            if kind == "basic":
                o = conv(y, sdp[p + ".conv1.weight"], stride=stride, padding=1)
                o = rnd(torch.relu(_bn(o, sdp, p + ".bn1", train, ns)))
                o = conv(o, sdp[p + ".conv2.weight"], padding=1)
                o = _bn(o, sdp, p + ".bn2", train, ns)
            else:
                o = conv(y, sdp[p + ".conv1.weight"])
                o = rnd(torch.relu(_bn(o, sdp, p + ".bn1", train, ns)))
                o = conv(o, sdp[p + ".conv2.weight"], stride=stride, padding=1)
                o = rnd(torch.relu(_bn(o, sdp, p + ".bn2", train, ns)))
                o = conv(o, sdp[p + ".conv3.weight"])
                o = _bn(o, sdp, p + ".bn3", train, ns)
            if (p + ".downsample.0.weight") in sdp:
                idt = conv(y, sdp[p + ".downsample.0.weight"], stride=stride)
                idt = _bn(idt, sdp, p + ".downsample.1", train, ns)
            y = rnd(torch.relu(o + idt))
        stages.append(y.mean(dim=(2, 3)))
    feat = y.mean(dim=(2, 3))
    ns = {prefix + k: v for k, v in ns.items()}
    if return_stages:
        return feat, ns, stages
    return feat, ns


_DENSENET_CFG = {"densenet121": (32, (6, 12, 24, 16), 64), "densenet169": (32, (6, 12, 32, 32), 64),
                 "densenet201": (32, (6, 12, 48, 32), 64),
                 "densenet_tiny": (32, (2, 2, 2, 2), 64)}     # test-only: every layer type, little depth


def densenet_features(sd: Dict[str, torch.Tensor], x: torch.Tensor, arch: str, train: bool = True,
                      prefix: str = "cnn_backbone.", emulate_bf16: bool = False, return_stages: bool = False):
    """torchvision.models.densenet{121,169,201} with classifier=Identity (lrcn/lrcn.py:196-209): conv0 7x7/2, norm0,
    relu, maxpool 3x3/2; dense blocks of layers [norm1, relu, conv1 1x1 -> 4*growth, norm2, relu, conv2 3x3 -> growth,
    concatenated to the input]; transitions [norm, relu, conv 1x1 -> C/2, avgpool 2x2]; norm5, relu, global average pool.
    Train-mode BN uses batch statistics.  emulate_bf16: bf16 storage at the points the B200 path rounds.
    Returns (features, new running stats) [+ per-stage spatial means]."""
    growth, blocks, c0 = _DENSENET_CFG[arch]
    sdp = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    ns: Dict[str, torch.Tensor] = {}
    rnd = _bf16 if emulate_bf16 else (lambda t: t)

    def conv(inp, w, **kw):
        return rnd(F.conv2d(rnd(inp), rnd(w), None, **kw))

    y = conv(x, sdp["features.conv0.weight"], stride=2, padding=3)
    y = rnd(F.max_pool2d(torch.relu(_bn(y, sdp, "features.norm0", train, ns)), 3, 2, 1))
    stages = [y.mean(dim=(2, 3))]
    for bi, depth in enumerate(blocks):
        for li in range(depth):
            p = f"features.denseblock{bi + 1}.denselayer{li + 1}"
            o = rnd(torch.relu(_bn(y, sdp, p + ".norm1", train, ns)))
            o = conv(o, sdp[p + ".conv1.weight"])
            o = rnd(torch.relu(_bn(o, sdp, p + ".norm2", train, ns)))
            o = conv(o, sdp[p + ".conv2.weight"], padding=1)
            y = torch.cat([y, o], dim=1)
        stages.append(y.mean(dim=(2, 3)))
        if bi + 1 < len(blocks):
            p = f"features.transition{bi + 1}"
            o = rnd(torch.relu(_bn(y, sdp, p + ".norm", train, ns)))
            y = rnd(F.avg_pool2d(conv(o, sdp[p + ".conv.weight"]), 2, 2))
    y = rnd(torch.relu(_bn(y, sdp, "features.norm5", train, ns)))
    feat = y.mean(dim=(2, 3))
    ns = {prefix + k: v for k, v in ns.items()}
    if return_stages:
        return feat, ns, stages
    return feat, ns


_MBV2_CFG = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]


def mobilenetv2_features(sd: Dict[str, torch.Tensor], x: torch.Tensor, arch: str = "mobilenet_v2", train: bool = True,
                         prefix: str = "cnn_backbone.", emulate_bf16: bool = False, return_stages: bool = False):
    """torchvision.models.mobilenet_v2 with classifier=Identity (medsos models.py:133-143): Conv 3x3/2 -> BN -> ReLU6; 17
    inverted residual blocks [1x1 expand (t > 1) -> BN -> ReLU6 -> depthwise 3x3 (stride s) -> BN -> ReLU6 -> 1x1 project ->
    BN (+ input when stride 1 and equal widths)]; Conv 1x1 -> 1280 -> BN -> ReLU6; global average pool."""
    sdp = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    ns: Dict[str, torch.Tensor] = {}
    rnd = _bf16 if emulate_bf16 else (lambda t: t)
    relu6 = lambda t: torch.clamp(t, 0.0, 6.0)

    def conv(inp, w, **kw):
        return rnd(F.conv2d(rnd(inp), rnd(w), None, **kw))

    y = conv(x, sdp["features.0.0.weight"], stride=2, padding=1)
    y = rnd(relu6(_bn(y, sdp, "features.0.1", train, ns)))
    stages = []
    idx, cin = 1, 32
    for t, c, n, s in _MBV2_CFG:
        for i in range(n):
            stride = s if i == 0 else 1
            p = f"features.{idx}.conv"
            o, j = y, 0
            if t != 1:
                o = conv(o, sdp[f"{p}.0.0.weight"])
                o = rnd(relu6(_bn(o, sdp, f"{p}.0.1", train, ns)))
                j = 1
            wd = sdp[f"{p}.{j}.0.weight"]
            o = conv(o, wd, stride=stride, padding=1, groups=wd.shape[0])
            o = rnd(relu6(_bn(o, sdp, f"{p}.{j}.1", train, ns)))
            o = conv(o, sdp[f"{p}.{j + 1}.weight"])
            o = _bn(o, sdp, f"{p}.{j + 2}", train, ns)
            if stride == 1 and cin == c:
                o = o + y
            y = rnd(o)
            stages.append(y.mean(dim=(2, 3)))
            cin = c
            idx += 1
    y = conv(y, sdp["features.18.0.weight"])
    y = rnd(relu6(_bn(y, sdp, "features.18.1", train, ns)))
    feat = y.mean(dim=(2, 3))
    ns = {prefix + k: v for k, v in ns.items()}
    if return_stages:
        return feat, ns, stages
    return feat, ns


def medsos_lrcn_forward(sd, x, arch, hidden, rnn_layers, bidirectional, rnn_out="all",
                        train=True, emulate_bf16_backbone=False):
    """medsos_lrcn/src/models.py:188-234 with rnn_type='lstm', multiclass head, dropout p=0."""
    B, T, C, H, W = x.shape
    backbone = (mobilenetv2_features if arch.startswith("mobilenet") else
                densenet_features if arch.startswith("densenet") else resnet_features)
    feat, ns = backbone(sd, x.reshape(B * T, C, H, W), arch, train, emulate_bf16=emulate_bf16_backbone)
    y = feat.reshape(B, T, -1)
    for k in (1, 2, 3):
        y = y @ sd[f"adapt{k}.weight"].t() + sd[f"adapt{k}.bias"]
        y = layernorm(gelu_exact(y), sd[f"bn{k}.weight"], sd[f"bn{k}.bias"])
    r = lstm_forward(y, sd, hidden, rnn_layers, bidirectional, prefix="rnn.")
    r = r.reshape(B, -1) if rnn_out == "all" else r[:, -1, :]
    o = layernorm(r, sd["bn0.weight"], sd["bn0.bias"])
    o = layernorm(gelu_exact(o @ sd["fc.weight"].t() + sd["fc.bias"]), sd["bna.weight"], sd["bna.bias"])
    o = layernorm(gelu_exact(o @ sd["fca.weight"].t() + sd["fca.bias"]), sd["bnb.weight"], sd["bnb.bias"])
    return o @ sd["fcb.weight"].t() + sd["fcb.bias"], ns


def simple_lrcn_forward(sd, x, arch, hidden, rnn_layers, adapt_names=("adapt1", "adapt2", "adapt3"),
                        rnn_prefix="rnn.", rnn_out="all", num_heads: Optional[int] = None,
                        train=True, emulate_bf16_backbone=False):
    """lrcn/ucf50-lrcn.py:304-336 (three plain Linear adapts, attribute `rnn`) and
    lrcn/lrcn.py:285-305 / rgb_lrcn.py:247-263 (one `adapt`, attribute `lstm`); always
    bidirectional.  num_heads!=None -> per-class binary heads fc.{i} concatenated (lrcn.py:303)."""
    B, T, C, H, W = x.shape
    backbone = densenet_features if arch.startswith("densenet") else resnet_features
    feat, ns = backbone(sd, x.reshape(B * T, C, H, W), arch, train, emulate_bf16=emulate_bf16_backbone)
    y = feat.reshape(B, T, -1)
    for a in adapt_names:
        y = y @ sd[a + ".weight"].t() + sd[a + ".bias"]
    r = lstm_forward(y, sd, hidden, rnn_layers, True, prefix=rnn_prefix)
    r = r.reshape(B, -1) if rnn_out == "all" else r[:, -1, :]
    if num_heads is None:
        return r @ sd["fc.weight"].t() + sd["fc.bias"], ns
    outs = [r @ sd[f"fc.{i}.weight"].t() + sd[f"fc.{i}.bias"] for i in range(num_heads)]
    return torch.cat(outs, dim=1), ns


# --------------------------------------------------------------------------------------
# 4. Selective scan (config 5)
# --------------------------------------------------------------------------------------


def selective_scan(u, delta, A, Bm, Cm, chunk_reset: Optional[int] = 256, reverse: bool = False):
    """lrcn/videomamba.py:242-284: x_t = exp(delta_t A) * x_{t-1} + delta_t B_t u_t ; y_t = <x_t, C_t>,
    with the state RESET to zero every `chunk_reset` steps (videomamba.py:260,269).
    chunk_reset=None, reverse=True/False restates medsos models.py:47-71 (u and delta flipped in
    time, B and C NOT flipped, output flipped back)."""
    Bsz, L, D = u.shape
    N = A.shape[1]
    if reverse:
        u = torch.flip(u, dims=[1])
        delta = torch.flip(delta, dims=[1])
    x = u.new_zeros(Bsz, D, N)
    ys = []
    for t in range(L):
        if chunk_reset is not None and t % chunk_reset == 0:
            x = u.new_zeros(Bsz, D, N)
        dA = torch.exp(delta[:, t, :, None] * A[None])
        dBu = delta[:, t, :, None] * Bm[:, t, None, :] * u[:, t, :, None]
        x = dA * x + dBu
        ys.append((x * Cm[:, t, None, :]).sum(-1))
    y = torch.stack(ys, dim=1)
    if reverse:
        y = torch.flip(y, dims=[1])
    return y


def mamba_block_forward(sd, x, prefix, bidirectional=False, eps=1e-5):
    """medsos_lrcn/src/models.py:19-117 `ResidualBlock.forward`: mixer(norm(x)) + x, restated functionally.
    norm: x * rsqrt(mean(x^2) + eps) * w; mixer: in_proj -> split (x, res) -> causal depthwise conv1d (padding k-1,
    trimmed to L) -> SiLU -> x_proj -> split (delta_raw, B, C) -> softplus(dt_proj) -> A = -exp(A_log) -> scan forward
    [and reversed: u / delta flipped, B / C not] -> y * silu(res [repeated when bidirectional]) -> out_proj."""
    g = lambda k: sd[prefix + k]
    Bsz, L, dm = x.shape
    xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * g("norm.weight")
    xr = xn @ g("mixer.in_proj.weight").t() + g("mixer.in_proj.bias")
    di = g("mixer.A_log").shape[0]
    n = g("mixer.A_log").shape[1]
    xi, res = xr[..., :di], xr[..., di:]
    w = g("mixer.conv1d.weight")
    K = w.shape[-1]
    xc = F.conv1d(xi.transpose(1, 2), w, g("mixer.conv1d.bias"), padding=K - 1, groups=di)[:, :, :L].transpose(1, 2)
    xc = xc * torch.sigmoid(xc)
    xp = xc @ g("mixer.x_proj.weight").t()
    dt_rank = g("mixer.dt_proj.weight").shape[1]
    delta = F.softplus(xp[..., :dt_rank] @ g("mixer.dt_proj.weight").t() + g("mixer.dt_proj.bias"))
    Bm, Cm = xp[..., dt_rank:dt_rank + n], xp[..., dt_rank + n:]
    A = -torch.exp(g("mixer.A_log"))
    y = selective_scan(xc, delta, A, Bm, Cm, chunk_reset=None)
    if bidirectional:
        y = torch.cat([y, selective_scan(xc, delta, A, Bm, Cm, chunk_reset=None, reverse=True)], dim=-1)
        res = torch.cat([res, res], dim=-1)
    y = y * (res * torch.sigmoid(res))
    return y @ g("mixer.out_proj.weight").t() + g("mixer.out_proj.bias") + x


# --------------------------------------------------------------------------------------
# 5. Train step + prediction (train_eval.py:20-43)
# --------------------------------------------------------------------------------------


def cross_entropy_mean(logits, labels):
    """nn.CrossEntropyLoss() default (main.py:147): mean over the batch of -log softmax[label]."""
    lse = torch.logsumexp(logits, dim=1)
    return (lse - logits.gather(1, labels[:, None]).squeeze(1)).mean()


def predict(logits):
    """torch.max(outputs, 1) -> first maximal index (train_eval.py:27)."""
    return torch.from_numpy(np.argmax(logits.detach().cpu().numpy(), axis=1))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| -- the tolerance statistic used by every parity test."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)
