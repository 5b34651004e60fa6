"""Data-parallel training by clip: one process per GPU, replicated weights, bucketed gradient
all-reduce over NCCL (NVLink 5 / NVSwitch) overlapped with the rest of backward.

The reference is single-process; this is the new multi-GPU capability of
BASELINE config 4.  The path shards by independent clips, so the ONLY exchange is the gradient
all-reduce (mean over ranks).  BatchNorm statistics stay per replica, exactly like N independent
reference processes (no SyncBN exists in the reference); the parity oracle for DP is therefore the
mean of per-shard gradients.

Gradients are tiny (0.75-35 MB) -> the all-reduce is latency bound, so what matters is WHEN each
collective starts, not its size:
  * parameters are packed into flat buckets of at most `bucket_bytes` (default: 1/8 of the payload,
    clamped to 1..8 MB; a larger parameter gets a bucket of its own) in reverse registration order = the order backward
    produces them, so the head / LSTM / adapt3 / adapt2 gradients are on the wire while the
    weight-gradient GEMM of the first trainable layer is still running;
  * a bucket's all-reduce (NCCL `AVG`: no separate scaling pass) is launched from the autograd hook
    of its last-arriving gradient on a high-priority side stream;
  * after the all-reduce the parameters' `.grad` simply POINT INTO the flat bucket (no copy back);
    the only staging pass is one fused `_foreach_copy_` per bucket on the side stream.
`finish()` joins the side stream before `optimizer.step()`.  `timeline()` returns CUDA-event times
of every bucket (ready / reduced) relative to the first hook of the step, for profiles/."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        self.flat = None
        self.views = None
        self.pending = 0
        self.work = None
        self.ev_ready = None
        self.ev_done = None


class GradBucketAllReduce:
    """Usage (per step):  loss.backward(); dp.finish(); optimizer.step()"""

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = None, process_group=None, record_timeline: bool = False):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU, NCCL backend)")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.record_timeline = record_timeline
        self._avg = dist.get_backend(process_group) == "nccl"          # gloo has no AVG: SUM + one scaling pass
        params = [p for p in module.parameters() if p.requires_grad]
        if bucket_bytes is None:
            # ~8 collectives per step: each bucket costs one staging launch + one NCCL launch on the host thread, and the
            # trainable-backbone steps are launch bound (DenseNet-121 fine-tune, 364 parameters: 34 buckets of 1 MB cost
            # 5 ms of a 27 ms step at N = 2, profiles/r02_probe_dp_n2.log); never below 1 MB, never above 8 MB
            total = sum(p.numel() * p.element_size() for p in params)
            bucket_bytes = min(max(total // 8, 1 << 20), 8 << 20)
        self.buckets: List[_Bucket] = []
        cur, cur_bytes = [], 0
        for p in reversed(params):
            nbytes = p.numel() * p.element_size()
            if cur and cur_bytes + nbytes > bucket_bytes:              # close the bucket BEFORE it overflows
                self.buckets.append(_Bucket(cur))
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(_Bucket(cur))
        self._owner = {}
        for b in self.buckets:
            for p in b.params:
                self._owner[p] = b
                p.register_post_accumulate_grad_hook(self._hook)
        self._stream = None
        self._t0 = None
        self._last_timeline = None
        self._reset()

    def _reset(self):
        for b in self.buckets:
            b.pending = len(b.params)
            b.work = None

    def _hook(self, p):
        if self.record_timeline and self._t0 is None and p.is_cuda:
            self._t0 = torch.cuda.Event(enable_timing=True)
            self._t0.record(torch.cuda.current_stream(p.device))
        b = self._owner[p]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket):
        dev = b.params[0].device
        if b.flat is None or b.flat.device != dev:
            b.flat = torch.empty(b.numel, device=dev, dtype=b.params[0].dtype)
            b.views, off = [], 0
            for p in b.params:
                b.views.append(b.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        grads = [p.grad for p in b.params]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        if dev.type == "cuda":
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=dev, priority=-1)
            timed = self.record_timeline
            b.ev_ready = torch.cuda.Event(enable_timing=timed)
            b.ev_ready.record(torch.cuda.current_stream(dev))          # gradients of this bucket are final
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(b.ev_ready)
                torch._foreach_copy_(b.views, grads)                    # (a grad that already IS its view copies onto itself)
                b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
                if timed:
                    b.work.wait()                                       # stream-level wait: orders ev_done after the collective
                    b.ev_done = torch.cuda.Event(enable_timing=True)
                    b.ev_done.record(self._stream)
            for g in grads:                                             # the side stream reads them: keep the allocator honest
                g.record_stream(self._stream)
        else:
            torch._foreach_copy_(b.views, grads)
            b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def finish(self):
        """Wait for all buckets and point every `.grad` at its averaged slice of the flat bucket (params whose grad never
        arrived this step -- unused branches -- contribute zeros so every rank reduces the same shape)."""
        for b in self.buckets:
            if b.pending != 0:                                        # some grads missing: reduce what exists
                for p in b.params:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                b.pending = 0
                self._launch(b)
        for b in self.buckets:
            dev = b.params[0].device
            if dev.type == "cuda":
                with torch.cuda.stream(self._stream):
                    b.work.wait()
                    if not self._avg:
                        b.flat.mul_(1.0 / self.world)
            else:
                b.work.wait()
                if not self._avg:
                    b.flat.mul_(1.0 / self.world)
            for p, v in zip(b.params, b.views):
                p.grad = v                                             # no copy back: the optimizer reads the bucket
        if self._stream is not None:
            cur = torch.cuda.current_stream()
            cur.wait_stream(self._stream)
            if self.record_timeline and self._t0 is not None:
                end = torch.cuda.Event(enable_timing=True)
                end.record(cur)
                self._last_timeline = (self._t0, [(b.numel * b.params[0].element_size(), b.ev_ready, b.ev_done)
                                                  for b in self.buckets], end)
                self._t0 = None
        self._reset()

    def timeline(self):
        """[{bytes, ready_ms, reduced_ms}] per bucket + joined_ms for the last finished step (record_timeline=True),
        times relative to the first gradient hook of that step.  Synchronises the device."""
        if self._last_timeline is None:
            return None
        torch.cuda.synchronize()
        t0, rows, end = self._last_timeline
        return {"buckets": [{"bytes": n, "ready_ms": t0.elapsed_time(r), "reduced_ms": t0.elapsed_time(d)} for n, r, d in rows],
                "joined_ms": t0.elapsed_time(end)}

    @property
    def payload_bytes(self) -> int:
        return sum(b.numel * b.params[0].element_size() for b in self.buckets)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, process_group=None):
    """Replicate rank-`src` weights and buffers to every rank (start of training)."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)
