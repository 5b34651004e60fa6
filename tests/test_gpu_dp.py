"""Multi-GPU parity: NCCL gradient buckets of a real model == mean of the per-shard gradients.
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r"""
import copy, os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import video_classif_b200 as vc
from video_classif_b200.dp import GradBucketAllReduce, broadcast_parameters
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
model = vc.SmallCNNLRCN(10, 6, 16, (3, 32, 32), dropout=0.0).to(dev).train()
broadcast_parameters(model)
ref = copy.deepcopy(model)                                  # same weights, no hooks: per-shard gradients
dp = GradBucketAllReduce(model, bucket_bytes=64 << 10, record_timeline=True)
assert len(dp.buckets) >= 3, len(dp.buckets)
g = torch.Generator().manual_seed(5)
X = torch.rand(4 * world, 6, 3, 32, 32, generator=g)
Y = torch.randint(0, 10, (4 * world,), generator=g)
xs, ys = X.chunk(world)[rank].to(dev), Y.chunk(world)[rank].to(dev)
for step in range(2):                                       # step 2: .grad starts as a view of the flat bucket
    model.zero_grad(set_to_none=(step == 0))
    torch.nn.functional.cross_entropy(model(xs), ys).backward()
    dp.finish()
ref.zero_grad()
torch.nn.functional.cross_entropy(ref(xs), ys).backward()
worst = 0.0
for (k, p), q in zip(model.named_parameters(), ref.parameters()):
    gq = q.grad.clone()
    dist.all_reduce(gq)                                     # oracle: mean over ranks of the per-shard gradients
    gq /= world
    den = gq.abs().max().item() or 1.0
    e = (p.grad - gq).abs().max().item() / den
    if k.startswith("conv") and k.endswith(".bias"):        # zero true gradient ahead of train-mode BN
        continue
    worst = max(worst, e)
    assert e < 1e-4, (rank, k, e)
tl = dp.timeline()
assert tl is not None and len(tl["buckets"]) == len(dp.buckets)
# the first bucket's all-reduce must have STARTED before backward produced the last gradient (overlap)
assert tl["buckets"][0]["ready_ms"] < tl["buckets"][-1]["ready_ms"]
print("nccl dp ok", rank, f"{worst:.2e}", len(dp.buckets), dp.payload_bytes)
dist.destroy_process_group()
"""


def test_nccl_gradient_buckets_match_mean_of_shard_gradients(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = tmp_path / "dp_nccl_worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29741", str(script), ROOT],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("nccl dp ok") == 2
