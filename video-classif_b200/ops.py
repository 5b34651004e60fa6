"""Host-side operator layer: thin wrappers over the C ABI (include/b200lrcn.h) and the
torch.autograd.Functions that give the hand-written kernels a backward.

torch is used here for device memory (torch.empty / zeros), streams and autograd bookkeeping only;
every arithmetic step of the forward and backward pass is a b2_* kernel.  No CPU fallback: all
entry points raise B200LrcnError when the tensors are not on an sm_100 device."""
from __future__ import annotations

import os
import weakref

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

F32, BF16 = torch.float32, torch.bfloat16
# GEMMs smaller than this many MACs stay on the fp32 SIMT kernel even in bf16 mode
TC_MIN_MACS = 1 << 22


def _chk(*ts):
    _lib.require_device()
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.B200LrcnError("b200-lrcn operators need CUDA tensors (no CPU fallback)")


def _pad8(n):
    return (n + 7) // 8 * 8


# ----------------------------------------------------------------------------------------
# raw kernel wrappers
# ----------------------------------------------------------------------------------------

def gemm_tn(A, B, bias=None, out_dtype=F32, relu=False, stats=None, M=None, N=None, K=None, bias2=None):
    """D[M,N] = A[M,K] @ B[N,K]^T (+bias).  A, B bf16 2-D (row stride multiple of 8)."""
    _chk(A, B)
    assert A.dtype == BF16 and B.dtype == BF16
    M = A.shape[0] if M is None else M
    K = A.shape[1] if K is None else K
    N = B.shape[0] if N is None else N
    D = torch.empty((M, N), device=A.device, dtype=out_dtype)
    s1, s2 = stats if stats is not None else (None, None)
    call("b2_gemm_bf16_tn", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), D.data_ptr(), N, M, N, K,
         ptr(bias), ptr(bias2), int(out_dtype == BF16), int(relu), ptr(s1), ptr(s2), stream_ptr())
    return D


def conv2d_nhwc(x, w, stride, pad, stats=None):
    """x [N,H,W,C] bf16, w [Cout,R,S,C] bf16 -> raw conv output [N,P,Q,Cout] bf16 (+ column stats)."""
    _chk(x, w)
    N, H, W, C = x.shape
    Cout, R, S, _ = w.shape
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - S) // stride + 1
    y = torch.empty((N, P, Q, Cout), device=x.device, dtype=BF16)
    s1, s2 = stats if stats is not None else (None, None)
    call("b2_conv2d_nhwc_bf16", x.data_ptr(), N, H, W, C, w.data_ptr(), Cout, R, S, stride, pad, y.data_ptr(), 0, 1,
         0, ptr(s1), ptr(s2), stream_ptr())
    return y


def conv2d_bn_nhwc(x, w, stride, pad, a=None, a_relu=True, o=None, res=None, r=None, relu=False, stats=None,
                   fin=None, store=True):
    """b2_conv2d_bn_nhwc_bf16: convolution with the BatchNorms around it folded in (see include/b200lrcn.h).
    a / o / r = (scale, shift) fp32 vectors for the input / output / shortcut BatchNorm; stats = (sum, sumsq);
    fin = (gamma, beta, running_mean, running_var, scale_out, shift_out, counter_u32, eps, momentum);
    store=False: statistics-only pass (returns None)."""
    _chk(x, w)
    N, H, W, C = x.shape
    Cout, R, S, _ = w.shape
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - S) // stride + 1
    y = torch.empty((N, P, Q, Cout), device=x.device, dtype=BF16) if store else None
    a0, a1 = a if a is not None else (None, None)
    o0, o1 = o if o is not None else (None, None)
    r0, r1 = r if r is not None else (None, None)
    s1, s2 = stats if stats is not None else (None, None)
    if fin is not None:
        g, b, rm, rv, fs, fh, cnt, eps, mom = fin
    else:
        g = b = rm = rv = fs = fh = cnt = None
        eps, mom = 1e-5, 0.1
    call("b2_conv2d_bn_nhwc_bf16", x.data_ptr(), N, H, W, C, w.data_ptr(), Cout, R, S, stride, pad, ptr(y), ptr(a0),
         ptr(a1), int(a_relu), ptr(o0), ptr(o1), ptr(res), ptr(r0), ptr(r1), int(relu), ptr(s1), ptr(s2), ptr(g), ptr(b),
         ptr(rm), ptr(rv), ptr(fs), ptr(fh), ptr(cnt), float(eps), float(mom), stream_ptr())
    return y


def conv3x3_halo_supported(x, w, stride, pad):
    N, H, W, C = x.shape
    return (stride == 1 and pad == 1 and tuple(w.shape[1:3]) == (3, 3)
            and bool(_lib.lib().b2_conv3x3_halo_supported(N, H, W, C, w.shape[0])))


def conv3x3_halo_bn(x, w, a=None, a_relu=True, stats=None, fin=None):
    """b2_conv3x3_halo_bn_nhwc_bf16: 3x3/1/1 conv of a 64-channel stage from one halo tile per TH output rows
    (nine shifted descriptors, resident weights); a = (scale, shift) of the input BatchNorm(+ReLU), applied to the
    halo once per tile; stats / fin as conv2d_bn_nhwc.  Returns the raw bf16 output [N,H,W,64]."""
    _chk(x, w)
    N, H, W, C = x.shape
    Cout = w.shape[0]
    y = torch.empty((N, H, W, Cout), device=x.device, dtype=BF16)
    a0, a1 = a if a is not None else (None, None)
    s1, s2 = stats if stats is not None else (None, None)
    if fin is not None:
        g, b, rm, rv, fs, fh, cnt, eps, mom = fin
    else:
        g = b = rm = rv = fs = fh = cnt = None
        eps, mom = 1e-5, 0.1
    call("b2_conv3x3_halo_bn_nhwc_bf16", x.data_ptr(), N, H, W, C, w.data_ptr(), Cout, y.data_ptr(), ptr(a0), ptr(a1),
         int(a_relu), ptr(s1), ptr(s2), ptr(g), ptr(b), ptr(rm), ptr(rv), ptr(fs), ptr(fh), ptr(cnt), float(eps),
         float(mom), stream_ptr())
    return y


def conv1x1_gram_bnstats(x, w, a, fin, stats=None, a_relu=True):
    """b2_conv1x1_gram_bnstats_bf16: train-mode BatchNorm statistics + finalisation of the 1x1 convolution
    relu?(x*a_scale+a_shift) @ w^T without computing its output (Gram-matrix form, one pass over x).
    x [..., C] bf16 (C in GRAM_CHANNELS), w [Cout, (1, 1,) C] bf16; a = (scale, shift); fin as in conv2d_bn_nhwc;
    stats = optional (sum, sumsq) outputs (overwritten)."""
    _chk(x, w)
    C = x.shape[-1]
    Cout = w.shape[0]
    M = x.numel() // C
    ws = torch.empty(_lib.lib().b2_gram_workspace_floats(C), device=x.device, dtype=F32)
    s1, s2 = stats if stats is not None else (None, None)
    g, b, rm, rv, fs, fh, _cnt, eps, mom = fin
    call("b2_conv1x1_gram_bnstats_bf16", x.data_ptr(), M, C, w.data_ptr(), Cout, ptr(a[0]), ptr(a[1]),
         int(a_relu), ws.data_ptr(), ptr(s1), ptr(s2), g.data_ptr(), b.data_ptr(), ptr(rm), ptr(rv), ptr(fs),
         ptr(fh), float(eps), float(mom), stream_ptr())


GRAM_CHANNELS = (64, 128, 256)


def scale_shift_apply(x, scale, shift, res=None, r=None, relu=True, out=None):
    """y = act(x*scale[c] + shift[c] [+ res | + res*rscale[c] + rshift[c]]) over NHWC bf16 (in place by default)."""
    _chk(x)
    C = x.shape[-1]
    out = x if out is None else out
    r0, r1 = r if r is not None else (None, None)
    call("b2_scale_shift_apply_nhwc", x.data_ptr(), out.data_ptr(), x.numel() // C, C, ptr(scale), ptr(shift),
         ptr(res), ptr(r0), ptr(r1), int(relu), stream_ptr())
    return out


def pack_stem_weight(w):
    """torch conv1 weight [64,3,7,7] -> bf16 [28][64][8] for b2_stem_conv_bf16 (k = r*32 + s*4 + c)."""
    assert tuple(w.shape[1:]) == (3, 7, 7) and w.shape[0] == 64
    w4 = torch.zeros((64, 7, 8, 4), device=w.device, dtype=BF16)
    w4[:, :, :7, :3] = w.detach().permute(0, 2, 3, 1).to(BF16)
    return w4.reshape(64, 28, 8).permute(1, 0, 2).contiguous()


def stem_pack(x, out=None):
    """x [N,3,H,W] fp32/bf16 NCHW -> the packed bf16 operand of the direct stem conv (uint8 tensor of b2_stem_packed_bytes)."""
    _chk(x)
    x = x.contiguous()
    N, C, H, W = x.shape
    assert C == 3 and x.dtype in (F32, BF16)
    nbytes = _lib.lib().b2_stem_packed_bytes(N, H, W)
    xp = out if out is not None else torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    assert xp.numel() == nbytes and xp.dtype == torch.uint8 and xp.is_contiguous()
    call("b2_stem_pack", x.data_ptr(), int(x.dtype == BF16), xp.data_ptr(), N, H, W, stream_ptr())
    return xp


def stem_conv(x, wk, stats=None, xp=None):
    """x [N,3,H,W] fp32/bf16 NCHW, wk = pack_stem_weight(conv1.weight) -> raw conv output [N,P,Q,64] bf16
    (Conv2d 7x7 stride 2 pad 3) + optional per-channel sum / sum-of-squares.  xp: the frames already packed by stem_pack
    (x is then read for its shape only)."""
    _chk(wk)
    N, C, H, W = x.shape
    assert C == 3 and x.dtype in (F32, BF16)
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if xp is None:
        xp = stem_pack(x)
    st = stream_ptr()
    y = torch.empty((N, P, Q, 64), device=x.device, dtype=BF16)
    s1, s2 = stats if stats is not None else (None, None)
    call("b2_stem_conv_bf16", xp.data_ptr(), wk.data_ptr(), y.data_ptr(), N, H, W, ptr(s1), ptr(s2), st)
    return y


def sgemm(A, B, trans_a=False, trans_b=False, out=None, alpha=1.0, beta=0.0, M=None, N=None, K=None, bias=None,
          bias2=None):
    """fp32 row-major C = alpha op(A) op(B) + beta C on the SIMT kernel."""
    _chk(A, B)
    assert A.dtype == F32 and B.dtype == F32 and A.stride(-1) == 1 and B.stride(-1) == 1
    if M is None:
        M = A.shape[1] if trans_a else A.shape[0]
    if K is None:
        K = A.shape[0] if trans_a else A.shape[1]
    if N is None:
        N = B.shape[0] if trans_b else B.shape[1]
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=F32)
        beta = 0.0
    call("b2_sgemm", int(trans_a), int(trans_b), M, N, K, float(alpha), A.data_ptr(), A.stride(0), B.data_ptr(),
         B.stride(0), float(beta), out.data_ptr(), out.stride(0), ptr(bias), ptr(bias2), stream_ptr())
    return out


def colsum(X, out=None, accumulate=False):
    _chk(X)
    M, N = X.shape
    if out is None:
        out = torch.empty(N, device=X.device, dtype=F32)
        accumulate = False
    call("b2_colsum_f32", X.data_ptr(), X.stride(0), M, N, out.data_ptr(), int(accumulate), stream_ptr())
    return out


def cast_bf16(x):
    _chk(x)
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    call("b2_cast_f32_bf16", x.data_ptr(), y.data_ptr(), x.numel(), stream_ptr())
    return y


_weight_cache = {}      # id(parameter) -> (weakref, {kind: (data_ptr, _version, bf16 copy)}); entries die with their parameter

# A replayed CUDA graph updates parameters WITHOUT bumping their autograd version counters, so everything that caches a
# derived copy of a trainable weight also keys on this epoch; graph_step.GraphedTrainStep bumps it after every replay.
_graph_epoch = [0]
# Device-side counter added to every dropout seed inside the kernels (None: eager mode, the host-side counter alone varies the
# masks).  A captured launch has its scalar seed baked in; with this pointer set during capture it draws a new mask per replay.
_seed_offset = [None]


def graph_epoch():
    return _graph_epoch[0]


def bump_graph_epoch():
    _graph_epoch[0] += 1


def set_seed_offset(t):
    """t: a 1-element int64 CUDA tensor (or None).  Returns the previous value."""
    prev = _seed_offset[0]
    _seed_offset[0] = t
    return prev


def _seed_ptr():
    t = _seed_offset[0]
    return 0 if t is None else t.data_ptr()


def _cached_weight(w, kind):
    """bf16 ("cast") or transposed bf16 ("tcast") copy of a weight, reused while the parameter is unchanged (same storage,
    same version counter): eval / serving and frozen layers pay the conversion once, a training step once per optimizer
    update instead of once per use (forward + backward)."""
    if not isinstance(w, torch.nn.Parameter):
        return cast_bf16(w) if kind == "cast" else transpose_cast_bf16(w)      # a view / stacked heads: no cache
    key = id(w)
    ent = _weight_cache.get(key)
    if ent is None or ent[0]() is not w:
        ent = _weight_cache[key] = (weakref.ref(w, lambda _r, k=key: _weight_cache.pop(k, None)), {})
    slot = ent[1]
    hit = slot.get(kind)
    ver = (w._version, _graph_epoch[0] if w.requires_grad else 0)
    if hit is None or hit[0] != w.data_ptr() or hit[1] != ver:
        copy = cast_bf16(w.detach()) if kind == "cast" else transpose_cast_bf16(w.detach())
        hit = slot[kind] = (w.data_ptr(), ver, copy)
    return hit[2]


def transpose_cast_bf16(x):
    """fp32 [R,C] (any row stride) -> bf16 [C, pad8(R)] view [:, :R] (row stride padded for TMA)."""
    _chk(x)
    assert x.stride(1) == 1
    R, C = x.shape
    buf = torch.empty((C, _pad8(R)), device=x.device, dtype=BF16)
    call("b2_transpose_cast_f32_bf16", x.data_ptr(), x.stride(0), buf.data_ptr(), buf.stride(0), R, C, stream_ptr())
    return buf[:, :R]


def transpose_bf16(x):
    """bf16 [R,C] (any row stride) -> bf16 [C, pad8(R)] view [:, :R]: the K-major form of a bf16 operand for a weight-gradient GEMM."""
    _chk(x)
    assert x.dtype == BF16 and x.stride(1) == 1
    R, C = x.shape
    buf = torch.empty((C, _pad8(R)), device=x.device, dtype=BF16)
    call("b2_transpose_bf16", x.data_ptr(), x.stride(0), buf.data_ptr(), buf.stride(0), R, C, stream_ptr())
    return buf[:, :R]


_WGRAD_LINEAR = os.environ.get("B2_LINEAR_WGRAD", "1") == "1"


def linear_wgrad_tc(x2, dy2):
    """dW [N, K] fp32 = dy2^T @ x2 for x2 [M, K], dy2 [M, N] (nn.Linear / the hoisted RNN gate GEMM) on the tcgen05
    weight-gradient kernel: both operands are read MN-major straight from their row-major tensors (pixels = rows of the
    batch are the reduction), so neither transposed copy of the old gemm_tn form is materialised.  Needs K % 8 == N % 8 == 0."""
    M, K = x2.shape
    N = dy2.shape[1]
    xa = x2 if x2.dtype == BF16 else cast_bf16(x2)
    da = dy2 if dy2.dtype == BF16 else cast_bf16(dy2)
    dw = torch.zeros((N, K), device=x2.device, dtype=F32)
    call("b2_conv2d_wgrad_nhwc_bf16", xa.data_ptr(), M, 1, 1, K, da.data_ptr(), N, 1, 1, 1, 0, dw.data_ptr(), stream_ptr())
    return dw


def _wgrad_ok(x2, dy2):
    return (_WGRAD_LINEAR and x2.shape[1] % 8 == 0 and dy2.shape[1] % 8 == 0 and x2.is_contiguous() and dy2.is_contiguous())


def _kmajor_bf16(x):
    """[R,C] fp32 or bf16 -> bf16 [C,R] (transposed copy) for gemm_tn."""
    return transpose_bf16(x) if x.dtype == BF16 else transpose_cast_bf16(x)


def ingest_u8(frames_u8, out_h, out_w, frame_index=None, out_dtype=F32, swap_rb=True, divisor=255.0, out=None):
    """uint8 [F,H0,W0,3] device frames -> [n_out,3,out_h,out_w]; see b2_ingest_u8.  out: a preallocated contiguous result
    (e.g. the static input buffer of the encoder's CUDA graph, so that no copy sits between ingest and the encoder pass)."""
    _chk(frames_u8)
    assert frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 and frames_u8.shape[-1] == 3
    frames_u8 = frames_u8.contiguous()
    Fr, H0, W0, _ = frames_u8.shape
    n_out = Fr if frame_index is None else frame_index.numel()
    if frame_index is not None:
        assert frame_index.dtype == torch.int32 and frame_index.is_cuda
    if out is None:
        out = torch.empty((n_out, 3, out_h, out_w), device=frames_u8.device, dtype=out_dtype)
    else:
        if (out.numel() != n_out * 3 * out_h * out_w or not out.is_contiguous() or out.device != frames_u8.device
                or out.dtype not in (F32, BF16)):
            raise ValueError("ingest_u8: out must be a contiguous fp32 / bf16 tensor of n_out * 3 * out_h * out_w elements on the frames' device")
        out_dtype = out.dtype
        out = out.view(n_out, 3, out_h, out_w)
    call("b2_ingest_u8", frames_u8.data_ptr(), Fr, H0, W0, H0 * W0 * 3, ptr(frame_index), n_out, out.data_ptr(), out_h,
         out_w, int(out_dtype == BF16), int(swap_rb), float(divisor), stream_ptr())
    return out


def _scan_fwd(u, delta, A, Bm, Cm, chunk_reset, reverse):
    Bsz, L, D = u.shape
    N = A.shape[1]
    y = torch.empty_like(u)
    call("b2_selective_scan_fwd", u.data_ptr(), delta.data_ptr(), A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), y.data_ptr(),
         Bsz, L, D, N, int(chunk_reset or 0), int(reverse), stream_ptr())
    return y


def _scan_workspace(batch, L, D, N, chunk, device):
    """State workspace of b2_selective_scan_bwd: None when the chunk's states stay on chip (<= 512 steps)."""
    n = int(_lib.lib().b2_scan_bwd_workspace_floats(batch, L, D, N, int(chunk or 0)))
    return torch.empty(n, device=device, dtype=F32) if n else None


class SelectiveScanFn(torch.autograd.Function):
    """y = selective scan(u, delta, A, B, C): forward b2_selective_scan_fwd, backward b2_selective_scan_bwd (BPTT over the
    recomputed states; chunks of the chunk-reset variant in parallel)."""

    @staticmethod
    def forward(ctx, u, delta, A, Bm, Cm, chunk_reset, reverse):
        ctx.save_for_backward(u, delta, A, Bm, Cm)
        ctx.cfg = (int(chunk_reset or 0), int(reverse))
        return _scan_fwd(u, delta, A, Bm, Cm, chunk_reset, reverse)

    @staticmethod
    def backward(ctx, dy):
        u, delta, A, Bm, Cm = ctx.saved_tensors
        chunk, rev = ctx.cfg
        Bsz, L, D = u.shape
        N = A.shape[1]
        dy = dy.contiguous().float()
        ws = _scan_workspace(Bsz, L, D, N, chunk, u.device)
        du, dd = torch.empty_like(u), torch.empty_like(u)
        dA = torch.zeros_like(A)
        dB, dC = torch.zeros_like(Bm), torch.zeros_like(Cm)
        call("b2_selective_scan_bwd", u.data_ptr(), delta.data_ptr(), A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), dy.data_ptr(),
             ptr(ws), du.data_ptr(), dd.data_ptr(), dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), Bsz, L, D, N, chunk, rev, 0,
             stream_ptr())
        return du, dd, dA, dB, dC, None, None


def selective_scan(u, delta, A, Bm, Cm, chunk_reset=256, reverse=False):
    """y[B,L,D] of the VideoMamba scan (videomamba.py:242-284: state reset every `chunk_reset` steps; chunk_reset=None +
    reverse for medsos models.py:47-71).  Differentiable w.r.t. u, delta, A, B, C (BASELINE config 5)."""
    _chk(u, delta, A, Bm, Cm)
    u, delta, A, Bm, Cm = (t.contiguous().float() for t in (u, delta, A, Bm, Cm))
    Bsz, L, D = u.shape
    N = A.shape[1]
    assert delta.shape == u.shape and A.shape[0] == D and Bm.shape == (Bsz, L, N) and Cm.shape == (Bsz, L, N)
    if torch.is_grad_enabled() and any(t.requires_grad for t in (u, delta, A, Bm, Cm)):
        return SelectiveScanFn.apply(u, delta, A, Bm, Cm, chunk_reset, reverse)
    return _scan_fwd(u, delta, A, Bm, Cm, chunk_reset, reverse)


def rmsnorm(x, weight, eps=1e-5):
    _chk(x, weight)
    xc = x.contiguous()
    y = torch.empty_like(xc)
    D = xc.shape[-1]
    call("b2_rmsnorm_f32", xc.data_ptr(), weight.data_ptr(), y.data_ptr(), xc.numel() // D, D, float(eps), stream_ptr())
    return y


def _mamba_fwd(x, norm_w, eps, A_log, in_w, in_b, conv_w, conv_b, xproj_w, dt_w, dt_b, out_w, out_b, bidirectional):
    """Forward chain of the Mamba block; returns (out, saved intermediates)."""
    B, L, dm = x.shape
    di, n = A_log.shape
    dt_rank = dt_w.shape[1]
    K = conv_w.shape[-1]
    xc_in = x.contiguous()
    xn = rmsnorm(xc_in, norm_w, eps)
    xr = sgemm(xn.reshape(B * L, dm), in_w, trans_b=True, bias=in_b)                                   # [B*L, 2*di]
    xc = torch.empty((B, L, di), device=x.device, dtype=F32)
    call("b2_dwconv1d_silu_f32", xr.data_ptr(), 2 * di, conv_w.data_ptr(), ptr(conv_b), xc.data_ptr(), B, L, di, K,
         stream_ptr())
    xp = sgemm(xc.reshape(B * L, di), xproj_w, trans_b=True)                                           # [B*L, dt_rank + 2n]
    dpre = sgemm(xp[:, :dt_rank], dt_w, trans_b=True, bias=dt_b)                                       # [B*L, di]
    delta = torch.empty_like(dpre)
    call("b2_softplus_f32", dpre.data_ptr(), delta.data_ptr(), dpre.numel(), stream_ptr())
    Bm = xp[:, dt_rank:dt_rank + n].contiguous().reshape(B, L, n)
    Cm = xp[:, dt_rank + n:].contiguous().reshape(B, L, n)
    A = -torch.exp(A_log.detach())                        # parameter transform (models.py:94), like the weight re-layouts
    delta3 = delta.reshape(B, L, di)
    y = selective_scan(xc, delta3, A, Bm, Cm, chunk_reset=None)
    if bidirectional:
        y = torch.cat([y, selective_scan(xc, delta3, A, Bm, Cm, chunk_reset=None, reverse=True)], dim=-1)
    cols = y.shape[-1]
    g = torch.empty_like(y)
    call("b2_mul_silu_f32", y.data_ptr(), xr.data_ptr() + di * 4, 2 * di, di, g.data_ptr(), B * L, cols, stream_ptr())
    out = xc_in.reshape(B * L, dm).clone()                # residual: out = y W^T + b + 1 * x
    sgemm(g.reshape(B * L, cols), out_w, trans_b=True, out=out, beta=1.0, bias=out_b)
    return out.reshape(B, L, dm), (xc_in, xn, xr, xc, xp, dpre, delta3, Bm, Cm, A, y, g)


class MambaBlockFn(torch.autograd.Function):
    """The reference's Mamba `ResidualBlock` (medsos_lrcn/src/models.py:19-117) as one autograd node: forward chain of
    `_mamba_fwd`, backward through b2_selective_scan_bwd (BPTT with recomputed states) and the b2_*_bwd element kernels;
    all weight / input gradients of the projections on the fp32 SIMT GEMM."""

    @staticmethod
    def forward(ctx, x, norm_w, A_log, in_w, in_b, conv_w, conv_b, xproj_w, dt_w, dt_b, out_w, out_b, eps, bidirectional):
        _chk(x)
        out, saved = _mamba_fwd(x, norm_w, eps, A_log, in_w, in_b, conv_w, conv_b, xproj_w, dt_w, dt_b, out_w, out_b,
                                bidirectional)
        ctx.save_for_backward(norm_w, in_w, conv_w, conv_b if conv_b is not None else norm_w.new_empty(0), xproj_w, dt_w,
                              out_w, *saved)
        ctx.eps = eps
        ctx.bidirectional = bidirectional
        ctx.has = (in_b is not None, conv_b is not None, dt_b is not None, out_b is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        (norm_w, in_w, conv_w, conv_b, xproj_w, dt_w, out_w, x, xn, xr, xc, xp, dpre, delta3, Bm, Cm, A, y, g) = \
            ctx.saved_tensors
        has_in_b, has_conv_b, has_dt_b, has_out_b = ctx.has
        B, L, dm = x.shape
        di, n = A.shape
        dt_rank = dt_w.shape[1]
        K = conv_w.shape[-1]
        R = B * L
        cols = y.shape[-1]
        dev = x.device
        st = stream_ptr()
        do2 = dout.reshape(R, dm)
        if not do2.is_contiguous():
            do2 = do2.contiguous()
        if do2.dtype != F32:
            do2 = do2.float()
        g2, y2 = g.reshape(R, cols), y.reshape(R, cols)
        # out_proj
        d_out_w = sgemm(do2, g2, trans_a=True)                                   # [dm, cols]
        d_out_b = colsum(do2) if has_out_b else None
        dg = sgemm(do2, out_w)                                                   # [R, cols]
        # gate: g = y * silu(res)
        dxr = torch.zeros((R, 2 * di), device=dev, dtype=F32)
        dy = torch.empty((R, cols), device=dev, dtype=F32)
        call("b2_mul_silu_bwd_f32", dg.data_ptr(), y2.data_ptr(), xr.data_ptr() + di * 4, 2 * di, di, dy.data_ptr(),
             dxr.data_ptr() + di * 4, 2 * di, R, cols, st)
        # scan(s)
        ws = _scan_workspace(B, L, di, n, 0, dev)
        dA_log = torch.zeros((di, n), device=dev, dtype=F32)
        dBC = torch.zeros((2, B, L, n), device=dev, dtype=F32)
        dxc = ddelta = None
        for rev in range(2 if ctx.bidirectional else 1):
            dyd = dy if cols == di else dy[:, rev * di:(rev + 1) * di].contiguous()
            du = torch.empty((B, L, di), device=dev, dtype=F32)
            dd = torch.empty((B, L, di), device=dev, dtype=F32)
            call("b2_selective_scan_bwd", xc.data_ptr(), delta3.data_ptr(), A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(),
                 dyd.data_ptr(), ptr(ws), du.data_ptr(), dd.data_ptr(), dA_log.data_ptr(), dBC[0].data_ptr(),
                 dBC[1].data_ptr(), B, L, di, n, 0, rev, 1, st)
            if rev == 0:
                dxc, ddelta = du, dd
            else:
                dxc += du
                ddelta += dd
        # delta = softplus(dt_proj(xp[:, :dt_rank]))
        ddpre = torch.empty((R, di), device=dev, dtype=F32)
        call("b2_softplus_bwd_f32", ddelta.data_ptr(), dpre.data_ptr(), ddpre.data_ptr(), ddpre.numel(), st)
        d_dt_w = sgemm(ddpre, xp[:, :dt_rank], trans_a=True)                     # [di, dt_rank]
        d_dt_b = colsum(ddpre) if has_dt_b else None
        dxp = torch.empty_like(xp)
        sgemm(ddpre, dt_w, out=dxp[:, :dt_rank])
        dxp[:, dt_rank:dt_rank + n].copy_(dBC[0].reshape(R, n))
        dxp[:, dt_rank + n:].copy_(dBC[1].reshape(R, n))
        # x_proj
        d_xproj_w = sgemm(dxp, xc.reshape(R, di), trans_a=True)                  # [dt_rank + 2n, di]
        sgemm(dxp, xproj_w, out=dxc.reshape(R, di), beta=1.0)
        # causal depthwise conv + SiLU over xr[:, :di]
        d_conv_w = torch.zeros((di, K), device=dev, dtype=F32)
        d_conv_b = torch.zeros(di, device=dev, dtype=F32) if has_conv_b else None
        call("b2_dwconv1d_silu_bwd_f32", dxc.data_ptr(), xr.data_ptr(), 2 * di, conv_w.data_ptr(),
             conv_b.data_ptr() if has_conv_b else 0, dxr.data_ptr(), 2 * di, d_conv_w.data_ptr(), ptr(d_conv_b), B, L, di, K, st)
        # in_proj
        d_in_w = sgemm(dxr, xn.reshape(R, dm), trans_a=True)                     # [2 di, dm]
        d_in_b = colsum(dxr) if has_in_b else None
        dxn = sgemm(dxr, in_w)                                                   # [R, dm]
        # RMSNorm, + the residual branch
        dx = torch.empty((R, dm), device=dev, dtype=F32)
        d_norm_w = torch.zeros(dm, device=dev, dtype=F32)
        call("b2_rmsnorm_bwd_f32", dxn.data_ptr(), x.data_ptr(), norm_w.data_ptr(), dx.data_ptr(), d_norm_w.data_ptr(), R, dm,
             float(ctx.eps), st)
        dx += do2
        return (dx.reshape(B, L, dm), d_norm_w, dA_log, d_in_w, d_in_b, d_conv_w.reshape(conv_w.shape), d_conv_b, d_xproj_w,
                d_dt_w, d_dt_b, d_out_w, d_out_b, None, None)


def mamba_block_forward(x, blk, bidirectional=False):
    """The reference's Mamba `ResidualBlock` (medsos_lrcn/src/models.py:19-117: RMSNorm -> in_proj -> causal
    depthwise conv + SiLU -> x_proj / dt_proj + softplus -> selective scan (forward [+ reversed]) -> * silu(res) ->
    out_proj, + x) on the b2_* kernels.  `blk` is a parameter container with the reference's attribute names
    (norm.weight, mixer.{A_log, in_proj, conv1d, x_proj, dt_proj, out_proj}).  Differentiable (MambaBlockFn)."""
    _chk(x)
    mx = blk.mixer
    args = (blk.norm.weight, mx.A_log, mx.in_proj.weight, mx.in_proj.bias, mx.conv1d.weight, mx.conv1d.bias,
            mx.x_proj.weight, mx.dt_proj.weight, mx.dt_proj.bias, mx.out_proj.weight, mx.out_proj.bias)
    if torch.is_grad_enabled() and (x.requires_grad or any(a is not None and a.requires_grad for a in args)):
        return MambaBlockFn.apply(x, *args, float(blk.norm.eps), bool(bidirectional))
    return _mamba_fwd(x, args[0], float(blk.norm.eps), *args[1:], bidirectional)[0]


# ----------------------------------------------------------------------------------------
# autograd: Linear
# ----------------------------------------------------------------------------------------

def _linear_fwd(x2, weight, bias, bf16, bias2=None):
    M, K = x2.shape
    N = weight.shape[0]
    if bf16 and M * N * K >= TC_MIN_MACS and K % 8 == 0:
        xa = x2 if x2.dtype == BF16 else cast_bf16(x2)
        return gemm_tn(xa, _cached_weight(weight, "cast"), bias=bias, bias2=bias2, out_dtype=F32)
    xf = x2 if x2.dtype == F32 else x2.float()
    return sgemm(xf, weight, trans_b=True, bias=bias, bias2=bias2)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b  (nn.Linear).  bf16=True routes large products to the tcgen05 GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias, bf16):
        _chk(x, weight)
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        y = _linear_fwd(x2, weight, bias, bf16)
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        ctx.bf16 = bf16
        ctx.x_shape = x.shape
        return y.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        N, K = weight.shape
        dy2 = dy.reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        M = dy2.shape[0]
        dx = dw = db = None
        big = ctx.bf16 and M * N * K >= TC_MIN_MACS
        if ctx.needs_input_grad[0]:
            if big and N % 8 == 0:
                dx = gemm_tn(cast_bf16(dy2), _cached_weight(weight, "tcast"), out_dtype=F32)   # [M,N]x[K,N]^T
            else:
                dx = sgemm(dy2, weight)
            dx = dx.reshape(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            if big and _wgrad_ok(x2, dy2):
                dw = linear_wgrad_tc(x2, dy2)
            elif big:
                xf = x2 if x2.dtype == F32 else x2.float()
                dw = gemm_tn(transpose_cast_bf16(dy2), transpose_cast_bf16(xf), out_dtype=F32)  # [N,M]x[K,M]^T
            else:
                xf = x2 if x2.dtype == F32 else x2.float()
                dw = sgemm(dy2, xf, trans_a=True)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)
        return dx, dw, db, None


def linear(x, weight, bias=None, bf16=False):
    return LinearFn.apply(x, weight, bias, bf16)


# ----------------------------------------------------------------------------------------
# autograd: (GELU +) LayerNorm
# ----------------------------------------------------------------------------------------

class ActLayerNormFn(torch.autograd.Function):
    """LayerNorm(gelu(pre)) (apply_gelu) or LayerNorm(pre): models.py:200-202,222-224."""

    @staticmethod
    def forward(ctx, pre, gamma, beta, apply_gelu, eps):
        _chk(pre, gamma, beta)
        N = pre.shape[-1]
        p2 = pre.reshape(-1, N).contiguous()
        M = p2.shape[0]
        out = torch.empty_like(p2)
        mean = torch.empty(M, device=pre.device, dtype=F32)
        rstd = torch.empty(M, device=pre.device, dtype=F32)
        call("b2_act_ln_fwd", p2.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), 0, mean.data_ptr(),
             rstd.data_ptr(), M, N, float(eps), int(apply_gelu), stream_ptr())
        ctx.save_for_backward(p2, gamma, mean, rstd)
        ctx.apply_gelu = apply_gelu
        return out.reshape(pre.shape)

    @staticmethod
    def backward(ctx, dout):
        p2, gamma, mean, rstd = ctx.saved_tensors
        M, N = p2.shape
        d2 = dout.reshape(M, N).contiguous()
        dpre = torch.empty_like(p2)
        dgamma = torch.zeros(N, device=p2.device, dtype=F32)
        dbeta = torch.zeros(N, device=p2.device, dtype=F32)
        call("b2_act_ln_bwd", d2.data_ptr(), p2.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
             dpre.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), M, N, int(ctx.apply_gelu), stream_ptr())
        return dpre.reshape(dout.shape), dgamma, dbeta, None, None


def act_layernorm(pre, gamma, beta, apply_gelu=True, eps=1e-5):
    return ActLayerNormFn.apply(pre, gamma, beta, apply_gelu, eps)


# ----------------------------------------------------------------------------------------
# autograd: stand-alone activation (string-programmed Adapt stacks, SiLU head of models_bidir.py)
# ----------------------------------------------------------------------------------------

ACT_KINDS = {"relu": 0, "gelu": 1, "silu": 2}


class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind):
        _chk(x)
        xc = x.contiguous()
        if xc.dtype != F32:
            xc = xc.float()
        y = torch.empty_like(xc)
        call("b2_act_fwd_f32", xc.data_ptr(), y.data_ptr(), xc.numel(), kind, stream_ptr())
        ctx.save_for_backward(xc)
        ctx.kind = kind
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        dc = dy.contiguous()
        dx = torch.empty_like(xc)
        call("b2_act_bwd_f32", dc.data_ptr(), xc.data_ptr(), dx.data_ptr(), xc.numel(), ctx.kind, stream_ptr())
        return dx, None


def act(x, kind):
    """relu / gelu (erf form) / silu as one kernel each way."""
    return ActFn.apply(x, ACT_KINDS[kind])


# ----------------------------------------------------------------------------------------
# autograd: dropout
# ----------------------------------------------------------------------------------------

class DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        _chk(x)
        xc = x.contiguous()
        y = torch.empty_like(xc)
        call("b2_dropout_f32", xc.data_ptr(), y.data_ptr(), xc.numel(), float(p), int(seed), _seed_ptr(), stream_ptr())
        ctx.p, ctx.seed = p, seed
        return y

    @staticmethod
    def backward(ctx, dy):
        dc = dy.contiguous()
        dx = torch.empty_like(dc)
        call("b2_dropout_f32", dc.data_ptr(), dx.data_ptr(), dc.numel(), float(ctx.p), int(ctx.seed), _seed_ptr(), stream_ptr())
        return dx, None, None


_dropout_counter = [0]


def dropout(x, p, training):
    """nn.Dropout(p): identity in eval mode or when p == 0."""
    if not training or p <= 0.0:
        return x
    _dropout_counter[0] += 1
    seed = (torch.initial_seed() * 1000003 + _dropout_counter[0]) & 0x7FFFFFFFFFFFFFFF
    return DropoutFn.apply(x, float(p), seed)


# ----------------------------------------------------------------------------------------
# autograd: one LSTM layer (all directions)
# ----------------------------------------------------------------------------------------

class LSTMLayerFn(torch.autograd.Function):
    """One nn.LSTM layer, batch_first, zero initial state; params = (w_ih, w_hh, b_ih, b_hh) per
    direction (forward first, then `_reverse`).  Output [B,T,dirs*H] = cat(fwd, bwd)."""

    @staticmethod
    def forward(ctx, x, hidden, bf16, need_grad, *params):
        _chk(x, *params)
        B, T, In = x.shape
        dirs = len(params) // 4
        H = hidden
        xc = x.contiguous()
        x2 = xc.reshape(B * T, In)
        out = torch.empty((B, T, dirs * H), device=x.device, dtype=F32)
        saved = []
        for d in range(dirs):
            w_ih, w_hh, b_ih, b_hh = params[4 * d:4 * d + 4]
            G = _linear_fwd(x2, w_ih, b_ih, bf16, bias2=b_hh)   # hoisted gate GEMM for all T steps
            if need_grad:
                gates = torch.empty((B, T, 4 * H), device=x.device, dtype=F32)
                cst = torch.empty((B, T, H), device=x.device, dtype=F32)
            else:
                gates = cst = None
            o = out[:, :, d * H:]
            call("b2_lstm_seq_fwd", G.data_ptr(), w_hh.data_ptr(), o.data_ptr(), dirs * H, ptr(gates), ptr(cst), B, T,
                 H, int(d == 1), stream_ptr())
            saved += [gates, cst]
        if need_grad:
            ctx.save_for_backward(x2, out, *saved, *params)
        ctx.dims = (B, T, In, H, dirs)
        ctx.bf16 = bf16
        return out

    @staticmethod
    def backward(ctx, dout):
        B, T, In, H, dirs = ctx.dims
        x2, out = ctx.saved_tensors[:2]
        saved = ctx.saved_tensors[2:2 + 2 * dirs]
        params = ctx.saved_tensors[2 + 2 * dirs:]
        dout = dout.contiguous()
        grads = []
        big = ctx.bf16 and B * T * 4 * H * In >= TC_MIN_MACS and H % 2 == 0
        H4 = 4 * H
        dG_all = torch.empty((B * T, dirs * H4), device=dout.device, dtype=F32)   # [dG_fwd | dG_bwd]
        xf = x2 if (x2.dtype == F32 or big) else x2.float()     # big: the tcgen05 GEMM takes the bf16 operand as is
        use_wg = big and _WGRAD_LINEAR and In % 8 == 0 and xf.is_contiguous()   # MN-major weight-gradient kernel: no transposes
        xt = _kmajor_bf16(xf) if (big and not use_wg) else None  # (old form) one transposed copy of x for all directions
        xa = (xf if xf.dtype == BF16 else cast_bf16(xf)) if use_wg else None
        for d in range(dirs):
            w_ih, w_hh, b_ih, b_hh = params[4 * d:4 * d + 4]
            gates, cst = saved[2 * d:2 * d + 2]
            dG = dG_all[:, d * H4:(d + 1) * H4]
            dWhh = torch.zeros_like(w_hh)
            do = dout[:, :, d * H:]
            oo = out[:, :, d * H:]
            call("b2_lstm_seq_bwd", do.data_ptr(), dirs * H, oo.data_ptr(), dirs * H, gates.data_ptr(), cst.data_ptr(),
                 w_hh.data_ptr(), dG.data_ptr(), dirs * H4, dWhh.data_ptr(), B, T, H, int(d == 1), stream_ptr())
            if use_wg:
                dWih = linear_wgrad_tc(xa, dG if dG.is_contiguous() else dG.contiguous())
            elif big:
                dWih = gemm_tn(transpose_cast_bf16(dG), xt, out_dtype=F32)
            else:
                dWih = sgemm(dG, xf, trans_a=True)
            db = colsum(dG)
            grads += [dWih, dWhh, db, db]
        dx = None
        if ctx.needs_input_grad[0]:
            # dx = [dG_fwd | dG_bwd] @ [W_ih ; W_ih_reverse]  -- one GEMM over K = dirs*4H
            w_cat = params[0] if dirs == 1 else torch.cat([params[0], params[4]], dim=0)
            if big:      # a bf16 input (the tensor-core CNN's feature) takes its gradient in bf16: half the bytes of the widest tensor
                w_t = _cached_weight(w_cat, "tcast") if dirs == 1 else transpose_cast_bf16(w_cat)
                dx = gemm_tn(cast_bf16(dG_all), w_t, out_dtype=BF16 if x2.dtype == BF16 else F32)
            else:
                dx = sgemm(dG_all, w_cat)
            dx = dx.reshape(B, T, In)
        return (dx, None, None, None, *grads)


class GRULayerFn(torch.autograd.Function):
    """One nn.GRU layer, batch_first, zero initial state; params = (w_ih, w_hh, b_ih, b_hh) per direction (forward
    first, then `_reverse`).  Output [B,T,dirs*H] = cat(fwd, bwd).  The input-to-gate product for all T steps is one
    hoisted GEMM (tcgen05 when bf16 and large enough), the recurrence one persistent kernel per direction."""

    @staticmethod
    def forward(ctx, x, hidden, bf16, need_grad, *params):
        _chk(x, *params)
        B, T, In = x.shape
        dirs = len(params) // 4
        H = hidden
        xc = x.contiguous()
        x2 = xc.reshape(B * T, In)
        out = torch.empty((B, T, dirs * H), device=x.device, dtype=F32)
        saved = []
        for d in range(dirs):
            w_ih, w_hh, b_ih, b_hh = params[4 * d:4 * d + 4]
            G = _linear_fwd(x2, w_ih, b_ih, bf16)
            sv = torch.empty((B, T, 4 * H), device=x.device, dtype=F32) if need_grad else None
            o = out[:, :, d * H:]
            call("b2_gru_seq_fwd", G.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr(), o.data_ptr(), dirs * H, ptr(sv), B, T, H,
                 int(d == 1), stream_ptr())
            saved.append(sv)
        if need_grad:
            ctx.save_for_backward(x2, out, *saved, *params)
        ctx.dims = (B, T, In, H, dirs)
        ctx.bf16 = bf16
        return out

    @staticmethod
    def backward(ctx, dout):
        B, T, In, H, dirs = ctx.dims
        x2, out = ctx.saved_tensors[:2]
        saved = ctx.saved_tensors[2:2 + dirs]
        params = ctx.saved_tensors[2 + dirs:]
        dout = dout.contiguous()
        H3 = 3 * H
        big = ctx.bf16 and B * T * H3 * In >= TC_MIN_MACS and H % 8 == 0
        dG_all = torch.empty((B * T, dirs * H3), device=dout.device, dtype=F32)
        xf = x2 if (x2.dtype == F32 or big) else x2.float()
        use_wg = big and _WGRAD_LINEAR and In % 8 == 0 and xf.is_contiguous()
        xt = _kmajor_bf16(xf) if (big and not use_wg) else None
        xa = (xf if xf.dtype == BF16 else cast_bf16(xf)) if use_wg else None
        grads = []
        for d in range(dirs):
            w_ih, w_hh, b_ih, b_hh = params[4 * d:4 * d + 4]
            dG = dG_all[:, d * H3:(d + 1) * H3]
            dWhh = torch.zeros_like(w_hh)
            dbhh = torch.zeros_like(b_hh)
            call("b2_gru_seq_bwd", dout[:, :, d * H:].data_ptr(), dirs * H, out[:, :, d * H:].data_ptr(), dirs * H,
                 saved[d].data_ptr(), w_hh.data_ptr(), dG.data_ptr(), dirs * H3, dWhh.data_ptr(), dbhh.data_ptr(), B, T, H,
                 int(d == 1), stream_ptr())
            if use_wg:
                dWih = linear_wgrad_tc(xa, dG if dG.is_contiguous() else dG.contiguous())
            elif big:
                dWih = gemm_tn(transpose_cast_bf16(dG), xt, out_dtype=F32)
            else:
                dWih = sgemm(dG, xf, trans_a=True)
            grads += [dWih, dWhh, colsum(dG), dbhh]
        dx = None
        if ctx.needs_input_grad[0]:
            w_cat = params[0] if dirs == 1 else torch.cat([params[0], params[4]], dim=0)
            if big:
                dx = gemm_tn(cast_bf16(dG_all), transpose_cast_bf16(w_cat), out_dtype=F32)
            else:
                dx = sgemm(dG_all, w_cat)
            dx = dx.reshape(B, T, In)
        return (dx, None, None, None, *grads)


def gru_forward(x, gru_module, bf16=False):
    """Runs a torch.nn.GRU *parameter container* (batch_first, dropout 0) on the persistent kernels."""
    assert gru_module.batch_first
    dirs = 2 if gru_module.bidirectional else 1
    H = gru_module.hidden_size
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in gru_module.parameters()))
    y = x
    for layer in range(gru_module.num_layers):
        params = []
        for d in range(dirs):
            sfx = f"_l{layer}" + ("_reverse" if d == 1 else "")
            params += [getattr(gru_module, "weight_ih" + sfx), getattr(gru_module, "weight_hh" + sfx),
                       getattr(gru_module, "bias_ih" + sfx), getattr(gru_module, "bias_hh" + sfx)]
        y = GRULayerFn.apply(y, H, bf16, need_grad, *params)
    return y


def rnn_forward(x, module, bf16=False):
    """nn.LSTM or nn.GRU parameter container -> [B,T,dirs*H]."""
    return gru_forward(x, module, bf16) if isinstance(module, torch.nn.GRU) else lstm_forward(x, module, bf16)


def _ptr_array(tensors):
    import ctypes
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


class LSTMStackFn(torch.autograd.Function):
    """A whole unidirectional nn.LSTM stack in one launch per direction of time (b2_lstm_stack_fwd / _bwd):
    params = (w_ih, w_hh, b_ih, b_hh) per layer.  Output = the top layer's [B,T,H] sequence."""

    @staticmethod
    def forward(ctx, x, hidden, need_grad, *params):
        _chk(x, *params)
        B, T, In = x.shape
        L = len(params) // 4
        H = hidden
        xc = x.contiguous()
        dev = x.device
        out = torch.empty((L, B, T, H), device=dev, dtype=F32)
        gates = torch.empty((L, B, T, 4 * H), device=dev, dtype=F32) if need_grad else None
        cst = torch.empty((L, B, T, H), device=dev, dtype=F32) if need_grad else None
        keep = [_ptr_array([params[4 * l + i] for l in range(L)]) for i in range(4)]
        call("b2_lstm_stack_fwd", xc.data_ptr(), In, keep[0][1], keep[1][1], keep[2][1], keep[3][1], L, out.data_ptr(),
             ptr(gates), ptr(cst), B, T, H, stream_ptr())
        if need_grad:
            ctx.save_for_backward(xc, out, gates, cst, *params)
        ctx.dims = (B, T, In, H, L)
        return out[L - 1]

    @staticmethod
    def backward(ctx, dout):
        B, T, In, H, L = ctx.dims
        xc, out, gates, cst = ctx.saved_tensors[:4]
        params = ctx.saved_tensors[4:]
        dev = dout.device
        dout = dout.contiguous()
        # one zeroed slab for every parameter gradient of the stack (accumulated with atomics in the kernel)
        sizes = []
        for l in range(L):
            sizes += [params[4 * l].numel(), params[4 * l + 1].numel(), 4 * H]
        slab = torch.zeros(sum(sizes), device=dev, dtype=F32)
        views, off = [], 0
        for n in sizes:
            views.append(slab[off:off + n])
            off += n
        dwih = [views[3 * l].view_as(params[4 * l]) for l in range(L)]
        dwhh = [views[3 * l + 1].view_as(params[4 * l + 1]) for l in range(L)]
        db = [views[3 * l + 2] for l in range(L)]
        dx = torch.empty((B, T, In), device=dev, dtype=F32) if ctx.needs_input_grad[0] else None
        k_wih = _ptr_array([params[4 * l] for l in range(L)])
        k_whh = _ptr_array([params[4 * l + 1] for l in range(L)])
        k_dwih, k_dwhh, k_db = _ptr_array(dwih), _ptr_array(dwhh), _ptr_array(db)
        call("b2_lstm_stack_bwd", dout.data_ptr(), xc.data_ptr(), In, k_wih[1], k_whh[1], L, out.data_ptr(),
             gates.data_ptr(), cst.data_ptr(), ptr(dx), k_dwih[1], k_dwhh[1], k_db[1], B, T, H, stream_ptr())
        grads = []
        for l in range(L):
            grads += [dwih[l], dwhh[l], db[l], db[l]]
        return (dx, None, None, *grads)


STACK_MAX_H, STACK_MAX_IN, STACK_MAX_T, STACK_MAX_LAYERS = 64, 64, 64, 8
LSTM_STACK = True      # False: always the per-layer kernels + hoisted gate GEMM (A/B parity tests)


def lstm_forward(x, lstm_module, bf16=False):
    """Runs a torch.nn.LSTM *parameter container* (batch_first, no proj, dropout 0) on the
    persistent kernels.  Returns the [B,T,dirs*H] output of the last layer."""
    assert lstm_module.batch_first and lstm_module.proj_size == 0
    dirs = 2 if lstm_module.bidirectional else 1
    H = lstm_module.hidden_size
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in lstm_module.parameters()))
    B, T, In = x.shape
    if (LSTM_STACK and dirs == 1 and H <= STACK_MAX_H and In <= STACK_MAX_IN and T <= STACK_MAX_T
            and lstm_module.num_layers <= STACK_MAX_LAYERS and x.dtype == F32):
        # narrow unidirectional stack: all layers and timesteps in one persistent launch
        params = []
        for layer in range(lstm_module.num_layers):
            sfx = f"_l{layer}"
            params += [getattr(lstm_module, "weight_ih" + sfx), getattr(lstm_module, "weight_hh" + sfx),
                       getattr(lstm_module, "bias_ih" + sfx), getattr(lstm_module, "bias_hh" + sfx)]
        return LSTMStackFn.apply(x, H, need_grad, *params)
    y = x
    for layer in range(lstm_module.num_layers):
        params = []
        for d in range(dirs):
            sfx = f"_l{layer}" + ("_reverse" if d == 1 else "")
            params += [getattr(lstm_module, "weight_ih" + sfx), getattr(lstm_module, "weight_hh" + sfx),
                       getattr(lstm_module, "bias_ih" + sfx), getattr(lstm_module, "bias_hh" + sfx)]
        y = LSTMLayerFn.apply(y, H, bf16, need_grad, *params)
    return y


# ----------------------------------------------------------------------------------------
# autograd: small-CNN block  conv3x3 -> BatchNorm2d -> ReLU [-> MaxPool2d(2,2)]   (fp32 NCHW)
# ----------------------------------------------------------------------------------------

class ConvBnReluPoolFn(torch.autograd.Function):
    """`pool(F.relu(bn(conv(x))))` of the notebook LRCN (nb:181-183) as conv + statistics +
    fused BN/ReLU/pool kernels, with the matching backward (pool/ReLU mask recomputed, two-pass BN
    backward, conv data- and weight-gradient kernels)."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, running_mean, running_var, train, pool, momentum, eps):
        _chk(x, w, gamma, beta)
        x = x.contiguous()
        N, Cin, H, W = x.shape
        Cout = w.shape[0]
        dev = x.device
        st = stream_ptr()
        z = torch.empty((N, Cout, H, W), device=dev, dtype=F32)
        call("b2_conv3x3_f32", x.data_ptr(), w.data_ptr(), ptr(b), z.data_ptr(), N, Cin, Cout, H, W, 0, st)
        scale = torch.empty(Cout, device=dev, dtype=F32)
        shift = torch.empty_like(scale)
        mean = torch.empty_like(scale)
        rstd = torch.empty_like(scale)
        count = N * H * W
        if train:
            sums = torch.zeros((2, Cout), device=dev, dtype=torch.float64)
            call("b2_bn2d_stats_f32", z.data_ptr(), N, Cout, H * W, sums[0].data_ptr(), sums[1].data_ptr(), st)
            call("b2_bn2d_finalize", sums[0].data_ptr(), sums[1].data_ptr(), count, gamma.data_ptr(), beta.data_ptr(),
                 ptr(running_mean), ptr(running_var), float(momentum), float(eps), 1, scale.data_ptr(),
                 shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), Cout, st)
        else:
            call("b2_bn2d_finalize", 0, 0, count, gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
                 running_var.data_ptr(), float(momentum), float(eps), 0, scale.data_ptr(), shift.data_ptr(),
                 mean.data_ptr(), rstd.data_ptr(), Cout, st)
        Ho, Wo = (H // 2, W // 2) if pool else (H, W)
        y = torch.empty((N, Cout, Ho, Wo), device=dev, dtype=F32)
        call("b2_bn2d_act_pool_fwd_f32", z.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(), 0, N, Cout, H,
             W, int(pool), st)
        ctx.save_for_backward(x, w, z, scale, shift, mean, rstd, gamma)
        ctx.cfg = (train, pool, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, z, scale, shift, mean, rstd, gamma = ctx.saved_tensors
        train, pool, has_bias = ctx.cfg
        N, Cin, H, W = x.shape
        Cout = w.shape[0]
        dev = x.device
        st = stream_ptr()
        dy = dy.contiguous()
        s = torch.zeros((2, Cout), device=dev, dtype=torch.float64)
        call("b2_bn2d_act_pool_bwd_reduce_f32", z.data_ptr(), dy.data_ptr(), scale.data_ptr(), shift.data_ptr(),
             mean.data_ptr(), rstd.data_ptr(), N, Cout, H, W, int(pool), s[0].data_ptr(), s[1].data_ptr(), st)
        dz = torch.empty_like(z)
        call("b2_bn2d_act_pool_bwd_apply_f32", z.data_ptr(), dy.data_ptr(), scale.data_ptr(), shift.data_ptr(),
             mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), s[0].data_ptr(), s[1].data_ptr(), N * H * W,
             int(train), dz.data_ptr(), N, Cout, H, W, int(pool), st)
        dbeta = torch.empty(Cout, device=dev, dtype=F32)
        dgamma = torch.empty(Cout, device=dev, dtype=F32)
        call("b2_f64_to_f32", s[0].data_ptr(), dbeta.data_ptr(), Cout, 0, st)
        call("b2_f64_to_f32", s[1].data_ptr(), dgamma.data_ptr(), Cout, 0, st)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            call("b2_conv3x3_f32", dz.data_ptr(), w.data_ptr(), 0, dx.data_ptr(), N, Cout, Cin, H, W, 1, st)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros_like(w)
            call("b2_conv3x3_wgrad_f32", x.data_ptr(), dz.data_ptr(), dw.data_ptr(), N, Cin, Cout, H, W, st)
        if has_bias and ctx.needs_input_grad[2]:
            sb = torch.zeros(Cout, device=dev, dtype=torch.float64)
            call("b2_bn2d_stats_f32", dz.data_ptr(), N, Cout, H * W, sb.data_ptr(), 0, st)
            db = torch.empty(Cout, device=dev, dtype=F32)
            call("b2_f64_to_f32", sb.data_ptr(), db.data_ptr(), Cout, 0, st)
        return dx, dw, db, dgamma, dbeta, None, None, None, None, None, None


def conv_bn_relu_pool(x, conv, bn, pool, training):
    """conv: nn.Conv2d(k=3,pad=1) container, bn: nn.BatchNorm2d container."""
    train = training or not bn.track_running_stats
    y = ConvBnReluPoolFn.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, train,
                               pool, bn.momentum if bn.momentum is not None else 0.1, bn.eps)
    if training and bn.track_running_stats:
        bn.num_batches_tracked += 1
    return y


# ----------------------------------------------------------------------------------------
# autograd: the whole small-CNN trunk on the bf16 tensor-core path (csrc/smallcnn_tc.cu)
# ----------------------------------------------------------------------------------------

def _sc_conv(x, w_k, cout, bias=None, stats=None):
    """b2_sc_conv3x3_bf16: x [N,H,W,Cin] bf16, w_k [Cout,3,3,Cin] bf16 -> raw [N,H,W,Cout] bf16 (+ bias, + statistics)."""
    N, H, W, Cin = x.shape
    y = torch.empty((N, H, W, cout), device=x.device, dtype=BF16)
    s1, s2 = stats if stats is not None else (None, None)
    call("b2_sc_conv3x3_bf16", x.data_ptr(), N, H, W, Cin, w_k.data_ptr(), cout, y.data_ptr(), ptr(bias), ptr(s1), ptr(s2),
         stream_ptr())
    return y


def _sc_kernel_weight(w):
    """torch [Cout,Cin,3,3] fp32 -> [Cout,3,3,Cin] bf16 (forward B operand)."""
    return w.detach().permute(0, 2, 3, 1).contiguous().to(BF16)


def _sc_kernel_weight_dgrad(w):
    """data-gradient filter: Wd[ci][r][s][co] = w[co][ci][2-r][2-s] as [Cin,3,3,Cout] bf16."""
    return w.detach().flip(2, 3).permute(1, 2, 3, 0).contiguous().to(BF16)


class SmallCNNTrunkFn(torch.autograd.Function):
    """conv1 -> BN -> ReLU -> conv2 -> BN -> ReLU -> pool -> conv3 -> BN -> ReLU -> pool -> dropout -> channel-major
    flatten (nb:178-186) as ONE autograd node on the bf16 NHWC tensor-core kernels: raw conv outputs + their batch
    statistics come from the conv epilogues, BN + ReLU (+ pool) is one element pass per layer, the backward walks
    BN backward (two passes) -> weight gradient (tcgen05, both operands MN-major) -> data gradient (the forward conv
    kernel on the flipped filter).  Returns the feature [N, 64 * H/4 * W/4] bf16 = the gate GEMM's A operand."""

    @staticmethod
    def forward(ctx, x, w1, b1, g1, be1, w2, b2, g2, be2, w3, b3, g3, be3, running, train, momentum, eps, p_drop, seed):
        _chk(x, w1, w2, w3)
        x = x.contiguous()
        if x.dtype != F32:
            x = x.float()
        N, _, H, W = x.shape
        dev = x.device
        st = stream_ptr()
        stats = torch.zeros((3, 2, 64), device=dev, dtype=F32)
        coef = torch.empty((3, 4, 64), device=dev, dtype=F32)          # scale, shift, mean, rstd per layer

        def finalize(k, bias, gamma, beta, count, C):
            rm, rv = running[2 * k], running[2 * k + 1]
            call("b2_sc_bn_finalize", stats[k, 0].data_ptr(), stats[k, 1].data_ptr(), ptr(bias), gamma.data_ptr(), beta.data_ptr(),
                 ptr(rm), ptr(rv), count, float(eps), float(momentum[k]), int(train), coef[k, 0].data_ptr(), coef[k, 1].data_ptr(),
                 coef[k, 2].data_ptr(), coef[k, 3].data_ptr(), C, st)

        def act(raw, k, pool):
            n, h, w, c = raw.shape
            y = torch.empty((n, h // pool, w // pool, c), device=dev, dtype=BF16)
            call("b2_sc_act_pool_fwd", raw.data_ptr(), coef[k, 0].data_ptr(), coef[k, 1].data_ptr(), y.data_ptr(), n, h, w, c,
                 pool, st)
            return y

        def layer_stats(k):
            return (stats[k, 0], stats[k, 1]) if train else None

        # the conv bias is folded into the BatchNorm finalisation (raw tensors are stored bias-free)
        x16 = torch.empty((N, H, W, 16), device=dev, dtype=BF16)        # frames as NHWC bf16, channels 3..15 zero
        call("b2_sc_pack_input", x.data_ptr(), x16.data_ptr(), N, H, W, st)
        w1k = torch.zeros((16, 3, 3, 16), device=dev, dtype=BF16)
        w1k[..., :3] = w1.detach().permute(0, 2, 3, 1)
        raw1 = _sc_conv(x16, w1k, 16, None, layer_stats(0))
        finalize(0, b1, g1, be1, N * H * W, 16)
        a1 = act(raw1, 0, 1)
        raw2 = _sc_conv(a1, _sc_kernel_weight(w2), 32, None, layer_stats(1))
        finalize(1, b2, g2, be2, N * H * W, 32)
        a2 = act(raw2, 1, 2)
        raw3 = _sc_conv(a2, _sc_kernel_weight(w3), 64, None, layer_stats(2))
        finalize(2, b3, g3, be3, N * (H // 2) * (W // 2), 64)
        a3 = act(raw3, 2, 2)
        HW = (H // 4) * (W // 4)
        feat = torch.empty((N, 64 * HW), device=dev, dtype=BF16)
        call("b2_sc_nhwc_to_chw", a3.data_ptr(), feat.data_ptr(), N, HW, 64, float(p_drop), int(seed), _seed_ptr(), st)
        ctx.save_for_backward(x16, raw1, a1, raw2, a2, raw3, coef, w2, w3)
        ctx.cfg = (bool(train), float(p_drop), int(seed), b1 is not None, b2 is not None, b3 is not None)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        x16, raw1, a1, raw2, a2, raw3, coef, w2, w3 = ctx.saved_tensors
        train, p_drop, seed, hb1, hb2, hb3 = ctx.cfg
        N, H, W, _ = x16.shape
        x = x16
        dev = x.device
        st = stream_ptr()
        dfeat = dfeat.contiguous()
        if dfeat.dtype not in (F32, BF16):
            dfeat = dfeat.float()
        HW = (H // 4) * (W // 4)
        s = torch.zeros((3, 2, 64), device=dev, dtype=F32)               # [layer][dbeta | dgamma][channel]

        def bn_bwd(raw, dy, k, pool):
            n, h, w, c = raw.shape
            args = (raw.data_ptr(), dy.data_ptr(), coef[k, 0].data_ptr(), coef[k, 1].data_ptr(), coef[k, 2].data_ptr(),
                    coef[k, 3].data_ptr(), s[k, 0].data_ptr(), s[k, 1].data_ptr())
            call("b2_sc_act_pool_bwd_reduce", *args, n, h, w, c, pool, st)
            dz = torch.empty_like(raw)
            call("b2_sc_act_pool_bwd_apply", *args, int(train), dz.data_ptr(), n, h, w, c, pool, st)
            return dz

        def wgrad(xa, dz, cin, cout):
            n, h, w, _ = xa.shape
            dw = torch.zeros((cout, 3, 3, cin), device=dev, dtype=F32)
            call("b2_sc_conv3x3_wgrad_bf16", xa.data_ptr(), dz.data_ptr(), n, h, w, cin, cout, dw.data_ptr(), st)
            return dw.permute(0, 3, 1, 2)                                # torch layout [Cout,Cin,3,3] (a view)

        d3 = torch.empty((N, H // 4, W // 4, 64), device=dev, dtype=BF16)
        call("b2_sc_chw_to_nhwc", dfeat.data_ptr(), int(dfeat.dtype == BF16), d3.data_ptr(), N, HW, 64, p_drop, seed, _seed_ptr(), st)
        dz3 = bn_bwd(raw3, d3, 2, 2)
        dw3 = wgrad(a2, dz3, 32, 64)
        da2 = _sc_conv(dz3, _sc_kernel_weight_dgrad(w3), 32)
        dz2 = bn_bwd(raw2, da2, 1, 2)
        dw2 = wgrad(a1, dz2, 16, 32)
        da1 = _sc_conv(dz2, _sc_kernel_weight_dgrad(w2), 16)
        dz1 = bn_bwd(raw1, da1, 0, 1)
        dw1 = wgrad(x16, dz1, 16, 16)[:, :3]                              # the zero-padded input channels carry no gradient

        def dbias(k, C, has):
            # a bias ahead of train-mode BatchNorm has a zero gradient; eval mode: sum dz = scale * sum dpre
            if not has:
                return None
            return torch.zeros(C, device=dev, dtype=F32) if train else (coef[k, 0, :C] * s[k, 0, :C])
        return (None, dw1, dbias(0, 16, hb1), s[0, 1, :16], s[0, 0, :16], dw2, dbias(1, 32, hb2), s[1, 1, :32], s[1, 0, :32],
                dw3, dbias(2, 64, hb3), s[2, 1, :64], s[2, 0, :64], None, None, None, None, None, None)


def smallcnn_trunk_supported(x):
    """The tensor-core trunk needs two exact 2x2 pools and a padded row that fits one TMA box."""
    return x.dim() == 4 and x.shape[1] == 3 and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0 and x.shape[3] <= 254


def smallcnn_trunk(x, m, training):
    """x [N,3,H,W] -> feature [N, 64*H/4*W/4] bf16 through SmallCNNTrunkFn; `m` carries conv1-3 / bn1-3 / dropout containers."""
    bns = (m.bn1, m.bn2, m.bn3)
    train = training or not all(b.track_running_stats for b in bns)
    running = [t for b in bns for t in (b.running_mean, b.running_var)]
    if training:
        running_arg = running
    else:
        running_arg = running           # eval: read only
    mom = [b.momentum if b.momentum is not None else 0.1 for b in bns]
    p = m.dropout.p if training else 0.0
    seed = 0
    if p > 0.0:
        _dropout_counter[0] += 1
        seed = (torch.initial_seed() * 1000003 + _dropout_counter[0]) & 0x7FFFFFFFFFFFFFFF
    feat = SmallCNNTrunkFn.apply(x, m.conv1.weight, m.conv1.bias, m.bn1.weight, m.bn1.bias, m.conv2.weight, m.conv2.bias,
                                 m.bn2.weight, m.bn2.bias, m.conv3.weight, m.conv3.bias, m.bn3.weight, m.bn3.bias, running_arg,
                                 train, mom, m.bn1.eps, p, seed)
    if training:
        for b in bns:
            if b.track_running_stats:
                b.num_batches_tracked += 1
    return feat
