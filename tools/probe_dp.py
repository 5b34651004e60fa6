"""Multi-GPU probe (run under torchrun): train-step time and the gradient-bucket timeline of the data-parallel path
for the bench config (frozen ResNet-50, 2.8 M trainable parameters) and the TRAINABLE configs where the overlap matters
(cfg-1 small-CNN LRCN 8.7 MB of gradients; crime LRCN densenet121 full fine-tune ~29 MB).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe_dp.py [models...]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_classif_b200 as vc  # noqa: E402
from video_classif_b200.dp import GradBucketAllReduce, broadcast_parameters  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank() if world > 1 else 0
which = sys.argv[1:] or ["medsos", "smallcnn", "crime_densenet"]


def build(name):
    torch.manual_seed(0)
    if name == "medsos":
        m = vc.LRCN(4, 16, 32, 8, cnn_backbone="resnet50", rnn_layers=3, dropout=0.25)
        x = torch.rand(64, 16, 3, 112, 112, device=dev)
        y = torch.randint(0, 4, (64,), device=dev)
        loss = torch.nn.functional.cross_entropy
    elif name == "smallcnn":
        m = vc.SmallCNNLRCN(50, 20, 32, (3, 64, 64), dropout=0.5, precision="bf16")
        x = torch.rand(64, 20, 3, 64, 64, device=dev)
        y = torch.randint(0, 50, (64,), device=dev)
        loss = torch.nn.functional.cross_entropy
    else:
        m = vc.CrimeLRCN(3, 16, 56, 512, cnn_backbone="densenet121", finetune=True, rnn_layers=4, classif_mode="multiple_binary")
        x = torch.rand(32, 16, 3, 112, 112, device=dev)
        y = (torch.rand(32, 3, device=dev) > 0.5).float()
        loss = torch.nn.functional.binary_cross_entropy_with_logits
    return m.to(dev).train(), x, y, loss


for name in which:
    m, x, y, loss_fn = build(name)
    dp = None
    if world > 1:
        broadcast_parameters(m)
        dp = GradBucketAllReduce(m, record_timeline=True)
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        loss_fn(m(x), y).backward()
        if dp is not None:
            dp.finish()
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 8
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        row = {"model": name, "world": world, "ms_per_step": ms.item(), "clips_per_s": world * x.shape[0] / ms.item() * 1e3,
               "trainable_MB": sum(p.numel() for p in m.parameters() if p.requires_grad) * 4 / 1e6}
        if dp is not None:
            tl = dp.timeline()
            last_ready = max(b["ready_ms"] for b in tl["buckets"])
            row["buckets"] = len(tl["buckets"])
            row["backward_ms_first_to_last_grad"] = last_ready
            row["joined_ms"] = tl["joined_ms"]
            row["exposed_after_last_grad_ms"] = tl["joined_ms"] - last_ready
            row["timeline"] = tl["buckets"] if len(tl["buckets"]) <= 24 else tl["buckets"][:6] + tl["buckets"][-6:]
        print(json.dumps(row), flush=True)
    del m, opt, dp
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
