"""Data-parallel training by clip: one process per GPU, replicated weights, bucketed gradient
all-reduce over NCCL (NVLink 5 / NVSwitch) overlapped with the rest of backward.

The reference is single-process; this is the new multi-GPU capability of
BASELINE config 4.  The path shards by independent clips, so the ONLY exchange is the gradient
all-reduce (sum / world).  BatchNorm statistics stay per replica, exactly like N independent
reference processes (no SyncBN exists in the reference); the parity oracle for DP is therefore the
mean of per-shard gradients.

Gradients are tiny (0.75-35 MB) -> the all-reduce is latency bound: parameters are packed into a
few flat buckets in reverse registration order (the order backward produces them); a bucket's
all-reduce is launched from the autograd hook of its last-arriving gradient on a side stream and
overlaps the remaining backward kernels.  `finish()` waits and scatters the averaged values back."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        self.flat = None
        self.pending = 0
        self.work = None
        self.event = None


class GradBucketAllReduce:
    """Usage (per step):  loss.backward(); dp.finish(); optimizer.step()"""

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 8 << 20, process_group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU, NCCL backend)")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        params = [p for p in module.parameters() if p.requires_grad]
        self.buckets: List[_Bucket] = []
        cur, cur_bytes = [], 0
        for p in reversed(params):
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= bucket_bytes:
                self.buckets.append(_Bucket(cur))
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(_Bucket(cur))
        self._owner = {}
        for b in self.buckets:
            for p in b.params:
                self._owner[p] = b
                p.register_post_accumulate_grad_hook(self._hook)
        self._stream = None
        self._reset()

    def _reset(self):
        for b in self.buckets:
            b.pending = len(b.params)
            b.work = None
            b.event = None

    def _hook(self, p):
        b = self._owner[p]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket):
        dev = b.params[0].device
        if b.flat is None or b.flat.device != dev:
            b.flat = torch.empty(b.numel, device=dev, dtype=b.params[0].dtype)
        views = []
        off = 0
        for p in b.params:
            views.append(b.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        b.views = views
        if dev.type == "cuda":
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=dev, priority=-1)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))          # gradients of this bucket are final
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ready)
                torch._foreach_copy_(views, [p.grad for p in b.params])
                b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            torch._foreach_copy_(views, [p.grad for p in b.params])
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for all buckets, write the averaged gradients back (params whose grad never
        arrived this step -- unused branches -- contribute zeros so every rank reduces the same shape)."""
        for b in self.buckets:
            if b.pending != 0:                                        # some grads missing: reduce what exists
                for p in b.params:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                b.pending = 0
                self._launch(b)
        inv = 1.0 / self.world
        for b in self.buckets:
            dev = b.params[0].device
            if dev.type == "cuda":
                with torch.cuda.stream(self._stream):
                    b.work.wait()
                    b.flat.mul_(inv)
                    torch._foreach_copy_([p.grad for p in b.params], b.views)
            else:
                b.work.wait()
                b.flat.mul_(inv)
                torch._foreach_copy_([p.grad for p in b.params], b.views)
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        self._reset()

    @property
    def payload_bytes(self) -> int:
        return sum(b.numel * b.params[0].element_size() for b in self.buckets)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, process_group=None):
    """Replicate rank-`src` weights and buffers to every rank (start of training)."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)
