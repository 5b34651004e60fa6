"""One whole train step -- zero_grad, forward, loss, backward, optimizer.step (medsos_lrcn/src/train_eval.py:20-43,
lrcn/lrcn.py:310-347) -- captured once into a CUDA graph and replayed per batch.

The drop-in modules launch ~100-800 small kernels per step from Python; at the reference's own batch sizes (8 clips of the
notebook model, 2 clips at 224x224) the host launch path, not the GPU, bounds the step (measured: the trainable tail of the
medsos LRCN takes 2.1 ms of host time for 0.53 ms of GPU time).  A graph replay removes the host from the loop:

    step = GraphedTrainStep(model, optimizer, criterion, example_inputs, example_labels)
    for inputs, labels in loader:
        loss = step(inputs, labels)          # same numbers as the eager loop of train_eval.py:20-43

What makes a replay equal to an eager step:
  * inputs / labels are copied into static buffers; outputs and loss are static tensors (clone them to keep them);
  * dropout masks: the kernels add a device-side step counter to their seed (ops.set_seed_offset), bumped before every replay;
  * BatchNorm running statistics / num_batches_tracked are device tensors updated by captured kernels;
  * the optimizer must keep its step count on the device: torch.optim.Adam / AdamW(capturable=True), or SGD;
  * every cache of a derived copy of a trainable weight (bf16 / transposed / kernel-layout copies) is invalidated after a
    replay (ops.bump_graph_epoch): a replay changes parameters without touching their autograd version counters, so an eager
    call that follows (model.eval()(x), a checkpoint conversion) re-derives what it needs.
The warm-up steps that PyTorch's capture recipe needs run on the example batch and are UNDONE before the capture: parameters,
buffers and optimizer state are restored in place, so constructing the object does not train the model.
Shapes are fixed at construction; a batch of another shape needs its own GraphedTrainStep (or the eager path).
Single process: the data-parallel gradient buckets (dp.GradBucketAllReduce) launch their collectives from autograd hooks on a
side stream and are not captured -- use the eager step under DP."""
from __future__ import annotations

import copy

import torch

from . import _lib, ops
from ._lib import call


class GraphedTrainStep:
    def __init__(self, model, optimizer, criterion, example_inputs, example_labels, warmup: int = 3):
        _lib.require_device()
        if not example_inputs.is_cuda:
            raise ValueError("GraphedTrainStep: example_inputs must live on the GPU the model is on")
        for grp in optimizer.param_groups:
            if "capturable" in grp and not grp["capturable"]:
                raise ValueError("GraphedTrainStep: construct the optimizer with capturable=True (its step count must live on "
                                 "the device to be advanced by a graph replay)")
        self.model, self.optimizer, self.criterion = model, optimizer, criterion
        dev = example_inputs.device
        self.static_inputs = example_inputs.detach().clone()
        self.static_labels = example_labels.detach().to(dev).clone()
        self._seed = torch.zeros(1, dtype=torch.int64, device=dev)
        params = [p for g in optimizer.param_groups for p in g["params"]]
        # ---- snapshot (restored in place after the warm-up so that the captured graph keeps pointing at the same tensors)
        saved = [t.detach().clone() for t in list(model.parameters()) + list(model.buffers())]
        had_state = {p: copy.deepcopy(optimizer.state[p]) for p in params if p in optimizer.state and optimizer.state[p]}
        prev_offset = ops.set_seed_offset(self._seed)
        try:
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):        # allocator pools, lazy optimizer state, function attributes
                    self._eager_step()
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            with torch.no_grad():
                for t, s in zip(list(model.parameters()) + list(model.buffers()), saved):
                    t.copy_(s)
                for p in params:
                    st = optimizer.state.get(p, {})
                    old = had_state.get(p)
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if old is not None and torch.is_tensor(old.get(k)):
                                v.copy_(old[k])
                            else:
                                v.zero_()                # state created by the warm-up: back to its initial value
            self._seed.zero_()
            ops.bump_graph_epoch()
            optimizer.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self.outputs = model(self.static_inputs)
                self.loss = criterion(self.outputs, self.static_labels)
                self.loss.backward()
                optimizer.step()
            self._launches = _lib.launch_count() - n0
        finally:
            ops.set_seed_offset(prev_offset)
        ops.bump_graph_epoch()

    def _eager_step(self):
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(self.static_inputs)
        loss = self.criterion(out, self.static_labels)
        loss.backward()
        self.optimizer.step()
        return loss

    def __call__(self, inputs, labels):
        """One train step on (inputs, labels); returns the (static) loss tensor.  self.outputs holds the logits."""
        if tuple(inputs.shape) != tuple(self.static_inputs.shape) or tuple(labels.shape) != tuple(self.static_labels.shape):
            raise ValueError(f"GraphedTrainStep was captured for inputs {tuple(self.static_inputs.shape)} / labels "
                             f"{tuple(self.static_labels.shape)}, got {tuple(inputs.shape)} / {tuple(labels.shape)}")
        self.static_inputs.copy_(inputs, non_blocking=True)
        self.static_labels.copy_(labels, non_blocking=True)
        self._seed += 1
        self.graph.replay()
        call("b2_add_launch_count", self._launches)
        ops.bump_graph_epoch()
        return self.loss
