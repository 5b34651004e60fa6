"""CPU tests: the C-ABI library loads and exports exactly what include/b200lrcn.h declares; host
logic (sampling, module surface / state_dict layout, loud failures); world_size-2 gloo test of the
gradient-bucket all-reduce.  No kernel is launched here."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, golden_tensors, load_golden
from oracle import lrcn_oracle as O


def test_library_exports_every_declared_symbol():
    import video_classif_b200 as vc
    sigs = vc._lib.parse_header()
    assert len(sigs) >= 28
    lib = vc._lib.lib()                                   # binds every prototype (AttributeError if missing)
    nm = subprocess.run(["nm", "-D", "--defined-only", vc._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in nm.splitlines() if " T " in l}
    assert set(sigs) <= exported, set(sigs) - exported
    assert {e for e in exported if e.startswith("b2_")} <= set(sigs), "exported but undeclared symbols"
    assert lib.b2_abi_version() == 1
    assert lib.b2_launch_count() >= 0


def test_no_gpu_means_loud_failure_not_fallback():
    import video_classif_b200 as vc
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vc.B200LrcnError):
        vc.ops.sgemm(torch.randn(4, 4), torch.randn(4, 4))
    m = vc.SmallCNNLRCN(5, 2, 4, (3, 8, 8))
    with pytest.raises(vc.B200LrcnError):
        m(torch.rand(1, 2, 3, 8, 8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "video-classif_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("parity oracle", ""), f


def test_sampling_bit_exact():
    import video_classif_b200 as vc
    S = vc.sampling
    t = json.load(open(os.path.join(GOLDEN, "sampling.json")))
    for key, want in t["medsos"].items():
        n, T = map(int, key.split(","))
        assert S.medsos_indices(n, T) == want, key
    for key, want in t["crime"].items():
        n, T = map(int, key.split(","))
        assert S.crime_indices(n, T) == want, key
    for key, want in t["seek"].items():
        n, T = map(int, key.split(","))
        assert S.seek_indices(n, T) == want, key
    for n in range(1, 300):                                # exhaustive vs the oracle restatement
        for T in (1, 7, 16, 20, 40, 60):
            assert S.medsos_indices(n, T) == O.medsos_indices(n, T)
            assert S.crime_indices(n, T) == O.crime_indices(n, T)
            assert S.seek_indices(n, T) == O.seek_indices(n, T)
            assert len(S.medsos_indices(n, T)) == T
    with pytest.raises(ValueError):
        S.duplicate_frames([], 4)


def test_state_dict_layout_matches_reference_checkpoints():
    import video_classif_b200 as vc
    g, meta = load_golden("smallcnn_lrcn_b.npz")
    ref_sd = golden_tensors(g, "sd0/")
    m = vc.SmallCNNLRCN(meta["num_classes"], meta["T"], meta["hidden"], (3, meta["size"], meta["size"]))
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    assert all(sd[k].shape == ref_sd[k].shape and sd[k].dtype == ref_sd[k].dtype for k in sd)
    m.load_state_dict(ref_sd)                              # a reference checkpoint loads unchanged
    # notebook model at its real size: 2,165,810 parameters
    big = vc.SmallCNNLRCN(50, 20, 32)
    assert sum(p.numel() for p in big.parameters()) == 2165810
    # medsos topology parameter counts printed by the reference's count_parameters
    # (dumps/new_medsos_log_bayesian.txt:27 -> frozen 21,284,672 for resnet34; 42,500,160 for resnet101)
    for arch, frozen in (("resnet34", 21284672), ("resnet101", 42500160)):
        mm = vc.LRCN(4, 4, 8, 8, cnn_backbone=arch)
        assert vc.count_parameters(mm)[1] == frozen
    med = vc.LRCN(4, 3, 32, 8, cnn_backbone="resnet18")
    keys = set(med.state_dict().keys())
    for k in ("adapt1.weight", "bn1.weight", "adapt3.bias", "rnn.weight_ih_l0", "rnn.weight_hh_l2", "rnn.bias_hh_l1",
              "fc.weight", "fca.weight", "fcb.bias", "bn0.weight", "bna.bias", "bnb.weight",
              "cnn_backbone.layer4.1.bn2.running_var", "cnn_backbone.conv1.weight"):
        assert k in keys, k
    crime = vc.CrimeLRCN(3, 4, 56, 512, cnn_backbone="resnet18")
    ck = crime.state_dict()
    assert ck["lstm.weight_ih_l0_reverse"].shape == (224, 512) and ck["lstm.weight_ih_l3"].shape == (224, 112)
    assert ck["fc.2.weight"].shape == (1, 2 * 56 * 4) and "adapt.weight" in ck
    with pytest.raises(NotImplementedError):
        vc.LRCN(4, 3, 32, 8, cnn_backbone="densenet121")
    mam = vc.LRCN(4, 3, 32, 8, cnn_backbone="resnet18", rnn_type="mamba", rnn_layers=2).state_dict()   # models.py:159-164
    assert mam["rnn.1.mixer.A_log"].shape == (16, 32) and mam["rnn.0.mixer.in_proj.weight"].shape == (32, 8)
    assert mam["rnn.0.mixer.conv1d.weight"].shape == (16, 1, 3) and mam["rnn.0.mixer.x_proj.weight"].shape == (96, 16)
    assert mam["rnn.0.norm.weight"].shape == (8,) and mam["fc.weight"].shape == (12, 24)
    gru = vc.LRCN(4, 3, 32, 8, cnn_backbone="resnet18", rnn_type="gru", rnn_layers=2, bidirectional=True).state_dict()
    assert gru["rnn.weight_ih_l1_reverse"].shape == (96, 64)
    with pytest.raises(ValueError):
        vc.LRCN(4, 3, 32, 8, cnn_backbone="resnet18", rnn_type="transformer")


def test_ingest_divisor_rounding_is_exact_in_fp32():
    # float32(v)/255f equals float32(double(v)/255.0) for every byte value -> the kernel's fp32 divide is exact
    v = np.arange(256)
    assert np.array_equal((v / 255.0).astype(np.float32), v.astype(np.float32) / np.float32(255))


_DP_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from video_classif_b200.dp import GradBucketAllReduce, broadcast_parameters
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(100 + rank)
model = torch.nn.Sequential(torch.nn.Linear(12, 20), torch.nn.Tanh(), torch.nn.Linear(20, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
model[2].weight.requires_grad_(True)
broadcast_parameters(model)
dp = GradBucketAllReduce(model, bucket_bytes=600)          # several buckets
assert len(dp.buckets) >= 2
torch.manual_seed(7)
X = torch.randn(8, 12); Y = torch.randint(0, 3, (8,))
xs, ys = X.chunk(world)[rank], Y.chunk(world)[rank]
for step in range(2):
    model.zero_grad()
    torch.nn.functional.cross_entropy(model(xs), ys).backward()
    dp.finish()
# oracle for DP (SURVEY 8e): mean over ranks of the per-shard gradients of the single-process model
ref = torch.nn.Sequential(torch.nn.Linear(12, 20), torch.nn.Tanh(), torch.nn.Linear(20, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
ref.load_state_dict(model.state_dict())
acc = [torch.zeros_like(p) for p in ref.parameters()]
for r in range(world):
    ref.zero_grad()
    torch.nn.functional.cross_entropy(ref(X.chunk(world)[r]), Y.chunk(world)[r]).backward()
    for a, p in zip(acc, ref.parameters()):
        a += p.grad / world
for a, p in zip(acc, model.parameters()):
    assert torch.allclose(a, p.grad, atol=1e-6), (rank, (a - p.grad).abs().max())
print("dp ok", rank, dp.payload_bytes)
dist.destroy_process_group()
"""


def test_gradient_bucket_allreduce_world2_gloo(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), ROOT],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("dp ok") == 2
