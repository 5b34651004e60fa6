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
    with pytest.raises(vc.B200LrcnError):         # the captured train step has no CPU form either
        vc.GraphedTrainStep(m, torch.optim.SGD(m.parameters(), lr=0.1), torch.nn.CrossEntropyLoss(),
                            torch.rand(1, 2, 3, 8, 8), torch.zeros(1, dtype=torch.long))


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
        vc.LRCN(4, 3, 32, 8, cnn_backbone="efficientnet_b1")
    # the other torchvision families the reference configures: densenet121 (lrcn.py:27 default; classifier -> Identity) and
    # mobilenet_v2 (automation.py:28; Sequential classifier, models.py:138-140)
    dn = vc.CrimeLRCN(3, 2, 8, 16, cnn_backbone="densenet121", finetune=True, rnn_layers=1)
    assert dn.adapt.in_features == 1024 and all(p.requires_grad for p in dn.cnn_backbone.parameters())   # FINETUNE: nothing frozen
    assert "cnn_backbone.features.denseblock4.denselayer16.conv2.weight" in dn.state_dict()
    assert type(dn._runner).__name__ == "DenseNetRunner"
    mb = vc.LRCN(4, 3, 32, 8, cnn_backbone="mobilenet_v2")
    assert mb.adapt1.in_features == 1280 and not any(p.requires_grad for p in mb.cnn_backbone.parameters())
    assert type(mb._runner).__name__ == "MobileNetRunner" and "cnn_backbone.features.18.0.weight" in mb.state_dict()
    assert vc.count_parameters(mb)[1] == sum(p.numel() for p in mb.cnn_backbone.parameters())
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


def _same_state(a, b):
    assert set(a) == set(b), set(a) ^ set(b)
    for k in a:
        assert torch.equal(a[k].cpu(), b[k].cpu()), k


def test_reference_whole_module_checkpoints_load_without_the_reference_source():
    """train_eval.py:53 / ucf50-lrcn.py:468 `torch.save(model)` files (committed tiny fixtures made by the reference's
    own classes, pickled as `__main__.LRCN` / `__main__.LRCN2`): plain torch.load cannot resolve the class here; the
    shim rebuilds the equivalent module with identical weights and inferred hyper-parameters."""
    import video_classif_b200 as vc
    from video_classif_b200 import checkpoint as C
    path = os.path.join(GOLDEN, "ckpt_smallcnn_lstm.pt")
    with pytest.raises(Exception):
        torch.load(path, weights_only=False)
    raw = torch.load(path, pickle_module=C._pickle_module(), weights_only=False)
    assert isinstance(raw, C.ReferencePlaceholder) and raw._ref_name == "LRCN"
    m = vc.load_reference_checkpoint(path, precision="fp32")
    assert type(m).__name__ == "SmallCNNLRCN" and not m.training
    assert (m.sequence_length, m.hidden_size, m.num_classes, m.lstm.num_layers, m.dropout.p) == (4, 8, 5, 2, 0.5)
    _same_state(m.state_dict(), raw.state_dict())
    g = vc.load_reference_checkpoint(os.path.join(GOLDEN, "ckpt_smallcnn_gru.pt"), precision="fp32")
    assert type(g).__name__ == "SmallCNNGRU" and isinstance(g.lstm, torch.nn.GRU) and g.lstm.bidirectional
    assert g.fc.weight.shape == (3, 64) and g.dropout.p == 0.3


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="builds the reference's backbone classes on the fly")
def test_reference_backbone_checkpoints_roundtrip(tmp_path):
    """medsos `models.LRCN` (lstm and mamba variants), ucf50-lrcn `__main__.LRCN` and crime `LRCN` pickles written by the
    reference classes -> load_reference_checkpoint: topology, hyper-parameters, requires_grad flags and every tensor."""
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    from oracle import refload
    import video_classif_b200 as vc
    mod = refload.medsos_models(CONF_RNN_LAYER=2, CONF_CLASSIF_MODE="multiclass", CONF_DROPOUT=0.25)
    for rnn_type, bidir in (("lstm", False), ("gru", True), ("mamba", True)):
        torch.manual_seed(5)
        ref = mod.LRCN(4, 6, 16, 8, cnn_backbone="resnet18", rnn_type=rnn_type, rnn_out="all", bidirectional=bidir)
        p = str(tmp_path / f"medsos_{rnn_type}.pt")
        MG.save_as(ref, p, "models", extra=(mod.ResidualBlock, mod.ParallelMamba, mod.RMSNorm))
        m = vc.load_reference_checkpoint(p)
        assert type(m).__name__ == "LRCN" and m.rnn_type == rnn_type and m.bidirectional == bidir and m.training
        assert (m.sequence_length, m.hidden_size) == (6, 16)
        _same_state(m.state_dict(), ref.state_dict())
        assert {n for n, q in m.named_parameters() if q.requires_grad} == {n for n, q in ref.named_parameters() if q.requires_grad}
    U, _ = refload.ucf50_lrcn(CONF_RNN_LAYER=2, CONF_CNN_BACKBONE="resnet18")
    ref = U(5, 4, 12, 16, cnn_backbone="resnet18").eval()
    p = str(tmp_path / "ucf50.pt")
    MG.save_as(ref, p, "__main__")
    m = vc.load_reference_checkpoint(p)
    assert type(m).__name__ == "UCF50LRCN" and not m.training and m.rnn.num_layers == 2 and m.rnn.hidden_size == 12
    _same_state(m.state_dict(), ref.state_dict())
    Cr, _ = refload.crime_lrcn(CONF_RNN_LAYER=2, CONF_CNN_BACKBONE="resnet18")
    ref = Cr(3, 4, 12, 16, cnn_backbone="resnet18")
    p = str(tmp_path / "crime.pt")
    MG.save_as(ref, p, "__main__")
    m = vc.load_reference_checkpoint(p)
    assert type(m).__name__ == "CrimeLRCN" and len(m.fc) == 3
    _same_state(m.state_dict(), ref.state_dict())
    # the crime script's own default: densenet121 with FINETUNE = True (nothing frozen) -- the trainable flags survive
    Cd, _ = refload.crime_lrcn(CONF_RNN_LAYER=1, CONF_CNN_BACKBONE="densenet121", CONF_FINETUNE=True)
    ref = Cd(3, 2, 8, 16, cnn_backbone="densenet121")
    p = str(tmp_path / "crime_dn.pt")
    MG.save_as(ref, p, "__main__")
    m = vc.load_reference_checkpoint(p)
    assert type(m).__name__ == "CrimeLRCN" and m.backbone == "densenet121" and type(m._runner).__name__ == "DenseNetRunner"
    _same_state(m.state_dict(), ref.state_dict())
    assert all(q.requires_grad for q in m.cnn_backbone.parameters()) == all(q.requires_grad for q in ref.cnn_backbone.parameters())
    mm = refload.medsos_models(CONF_RNN_LAYER=2, CONF_CLASSIF_MODE="multiclass", CONF_DROPOUT=0.25)
    ref = mm.LRCN(4, 3, 16, 8, cnn_backbone="mobilenet_v2", rnn_type="lstm", rnn_out="all", bidirectional=False)
    p = str(tmp_path / "medsos_mb.pt")
    MG.save_as(ref, p, "models", extra=(mm.ResidualBlock, mm.ParallelMamba, mm.RMSNorm))
    m = vc.load_reference_checkpoint(p)
    assert type(m).__name__ == "LRCN" and m.backbone == "mobilenet_v2" and type(m._runner).__name__ == "MobileNetRunner"
    _same_state(m.state_dict(), ref.state_dict())


def test_trainable_prefix_boundaries_host_logic():
    """freeze_until_layer semantics (lrcn.py:275-283: the first k+1 entries of named_parameters() frozen) -> which block the
    autograd path starts at; the DenseNet trunk's parameter list excludes the stem (it has its own autograd node)."""
    import video_classif_b200 as vc
    from video_classif_b200 import backbone_train as BT, densenet_train as DT
    m = vc.CrimeLRCN(3, 2, 8, 16, cnn_backbone="resnet18", freeze_until_layer=None, rnn_layers=1)
    assert BT.first_trainable_block(m.cnn_backbone) is None                       # everything frozen
    m = vc.CrimeLRCN(3, 2, 8, 16, cnn_backbone="resnet18", finetune=True, rnn_layers=1)
    assert BT.first_trainable_block(m.cnn_backbone) == ("stem",)                  # FINETUNE: nothing frozen
    names = [n for n, _ in m.cnn_backbone.named_parameters()]
    k = names.index("layer3.0.conv1.weight")
    m = vc.CrimeLRCN(3, 2, 8, 16, cnn_backbone="resnet18", freeze_until_layer=k - 1, rnn_layers=1)
    assert BT.first_trainable_block(m.cnn_backbone) == (3, 0)
    k = names.index("layer2.1.bn2.weight")                                        # boundary inside a block
    m = vc.CrimeLRCN(3, 2, 8, 16, cnn_backbone="resnet18", freeze_until_layer=k - 1, rnn_layers=1)
    assert BT.first_trainable_block(m.cnn_backbone) == (2, 1)
    assert not m.cnn_backbone.layer2[1].conv2.weight.requires_grad and m.cnn_backbone.layer2[1].bn2.weight.requires_grad
    dn = vc.CrimeLRCN(3, 2, 8, 16, cnn_backbone="densenet121", finetune=True, rnn_layers=1)
    tn, tp = DT._trunk_params(dn.cnn_backbone)
    assert len(tn) == len(tp) == len(list(dn.cnn_backbone.parameters())) - 3     # conv0.weight, norm0.{weight,bias}
    assert tn[0] == "features.denseblock1.denselayer1.norm1.weight" and tn[-1] == "features.norm5.bias"


def test_header_is_plain_c(tmp_path):
    """include/b200lrcn.h is the drop-in boundary: it must compile as C99 (extern "C" guards, no C++ types), so that any
    language with a C FFI can bind it."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "b200lrcn.h"\nint main(void) { return b2_abi_version() == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-Wno-comment", "-fsyntax-only", "-I",
                        os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_checkpoint_unpickler_never_resolves_foreign_globals(tmp_path):
    """`os.system`, `builtins.eval` ... must not be importable through load_reference_checkpoint: they become
    inert placeholders (REDUCE only constructs an empty nn.Module), nothing is executed."""
    import pickle

    import video_classif_b200.checkpoint as ck

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("echo pwned > %s" % (tmp_path / "pwned"),))

    for name, obj in (("evil.pkl", Evil()),):
        p = tmp_path / name
        with open(p, "wb") as f:
            pickle.dump(obj, f)
        with open(p, "rb") as f:
            try:
                out = ck._Unpickler(f).load()
            except Exception:
                out = None
        assert not (tmp_path / "pwned").exists()
        assert out is None or isinstance(out, ck.ReferencePlaceholder)
    assert not ck._allowed_global("builtins", "eval") and not ck._allowed_global("builtins", "getattr")
    assert not ck._allowed_global("os", "system") and not ck._allowed_global("subprocess", "Popen")
    assert ck._allowed_global("collections", "OrderedDict") and ck._allowed_global("torch._utils", "_rebuild_tensor_v2")


def test_variant_modules_convert_from_reference_layout():
    """convert_reference_module() recognises the new layouts from their state_dict keys / module types (CPU side)."""
    import video_classif_b200 as vc
    for tag in ("ucf50_mamba", "dump_gru", "adapt_gru_bi"):
        from conftest import load_golden
        g, meta = load_golden(f"variant_{tag}.npz")
        build = meta["build"]
        m = getattr(vc, build["cls"])(**build["kw"])
        m2 = vc.convert_reference_module(m)
        assert type(m2) is type(m)
        assert set(m2.state_dict()) == set(m.state_dict())
        for attr in ("rnn_type", "rnn_attr"):
            if hasattr(m, attr):
                assert getattr(m2, attr) == getattr(m, attr)
        if hasattr(m, "adapt") and hasattr(m.adapt, "mode"):
            assert m2.adapt.mode == m.adapt.mode


def test_host_side_shape_contracts_need_no_gpu():
    """Pure host functions of the C ABI that the Python layer uses to pick a kernel: which 3x3 shapes the halo-tile convs cover
    (64 channels: resident weights, padded width <= 64; 128 channels: streamed weights, the halo stages must fit shared memory)
    and when the selective-scan backward keeps its states on chip (no workspace)."""
    import video_classif_b200 as vc
    lib = vc._lib.lib()
    sup = lib.b2_conv3x3_halo_supported
    assert sup(1024, 14, 14, 128, 128) == 1 and sup(32, 28, 28, 128, 128) == 1 and sup(2, 4, 4, 128, 128) == 1
    assert sup(8, 56, 56, 128, 128) == 0 and sup(2, 5, 60, 128, 128) == 0          # halo stages larger than shared memory
    assert sup(1024, 28, 28, 64, 64) == 1 and sup(4, 62, 62, 64, 64) == 1 and sup(4, 63, 63, 64, 64) == 0
    assert sup(4, 14, 14, 256, 256) == 0 and sup(4, 14, 14, 128, 64) == 0 and sup(0, 14, 14, 64, 64) == 0
    ws = lib.b2_scan_bwd_workspace_floats
    assert ws(8, 3136, 2048, 16, 256) == 0                                           # config 5: 256-step chunks stay on chip
    assert ws(8, 3136, 2048, 16, 0) == 8 * 2048 * 3136 * 16                          # one 3136-step scan: workspace kernel
    assert ws(2, 19, 64, 12, 0) == 2 * 64 * 19 * 16 and lib.b2_scan_padded_states(12) == 16   # padded state width
    assert ws(2, 19, 96, 8, 0) == 2 * 96 * 19 * 8                                    # 96 channels are not a multiple of 512 / 8
