"""CPU: pins oracle/lrcn_oracle.py against golden vectors produced by the reference itself
(tests/golden/make_golden.py), and -- when /root/reference is present -- against the live reference."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, build_backbone_model, err, golden_tensors, load_golden
from oracle import lrcn_oracle as O
from oracle import refload


def test_sampling_matches_reference_vectors():
    t = json.load(open(os.path.join(GOLDEN, "sampling.json")))
    for key, want in t["medsos"].items():
        n, T = map(int, key.split(","))
        assert O.medsos_indices(n, T) == want, key
    for key, want in t["crime"].items():
        n, T = map(int, key.split(","))
        assert O.crime_indices(n, T) == want, key
    for key, want in t["seek"].items():
        n, T = map(int, key.split(","))
        assert O.seek_indices(n, T) == want, key


def test_resize_bit_exact_vs_cv2_golden():
    g = np.load(os.path.join(GOLDEN, "resize_cv2.npz"))
    i = 0
    while f"src{i}" in g.files:
        src, dst = g[f"src{i}"], g[f"dst{i}"]
        got = O.resize_bilinear_u8(src, dst.shape[0], dst.shape[1])
        assert np.array_equal(got, dst), f"case {i}"
        clip = O.ingest_clip(src[None], dst.shape[0], dst.shape[1], swap_rb=True)
        want = (g[f"rgb{i}"].astype(np.float64) / 255.0).astype(np.float32).transpose(2, 0, 1)
        assert np.array_equal(clip[0], want)
        i += 1
    assert i >= 5


def test_resize_bit_exact_vs_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for _ in range(25):
        h, w = rng.integers(2, 200, 2)
        oh, ow = rng.integers(1, 130, 2)
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(O.resize_bilinear_u8(src, int(oh), int(ow)), cv2.resize(src, (int(ow), int(oh))))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_smallcnn_oracle_vs_golden(tag):
    g, meta = load_golden(f"smallcnn_lrcn_{tag}.npz")
    sd = golden_tensors(g, "sd0/")
    p = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
         for k, v in sd.items()}
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    logits, ns = O.small_cnn_lrcn_forward(p, x, meta["hidden"])
    loss = O.cross_entropy_mean(logits, y)
    loss.backward()
    assert err(logits, torch.from_numpy(g["logits"])) < 2e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    gmax = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith("grad/"))
    for k, v in golden_tensors(g, "grad/").items():
        if k.startswith("conv") and k.endswith(".bias"):
            # analytically zero (train-mode BN removes the bias): both sides are rounding noise
            assert p[k].grad.abs().max().item() < 1e-4 * gmax and v.abs().max().item() < 1e-4 * gmax, k
        else:
            assert err(p[k].grad, v) < 2e-4, k
    for k, v in golden_tensors(g, "sd1/").items():
        if "running" in k:
            assert err(ns[k], v) < 1e-5, k
    assert torch.equal(O.predict(logits), O.predict(torch.from_numpy(g["logits"])))


@pytest.mark.parametrize("tag", ["uni", "bi"])
def test_lstm_oracle_vs_golden(tag):
    g, meta = load_golden(f"lstm_{tag}.npz")
    p = {"lstm." + k: v.clone().requires_grad_(True) for k, v in golden_tensors(g, "p/").items()}
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    out = O.lstm_forward(x, p, meta["H"], meta["layers"], meta["bidir"])
    (out * torch.from_numpy(g["w"])).sum().backward()
    assert err(out, torch.from_numpy(g["out"])) < 1e-5
    assert err(x.grad, torch.from_numpy(g["dx"])) < 1e-5
    for k, v in golden_tensors(g, "g/").items():
        assert err(p["lstm." + k].grad, v) < 1e-5, k


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_medsos_oracle_vs_golden(arch):
    m, g, meta = build_backbone_model(f"medsos_lrcn_{arch}.npz")
    sd = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point and not k.startswith("cnn_backbone.")
              else v.detach().clone()) for k, v in m.state_dict().items()}
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    B, T = x.shape[:2]
    feat, ns = O.resnet_features(sd, x.reshape(B * T, *x.shape[2:]), arch, train=True)
    assert err(feat, torch.from_numpy(g["features"])) < 1e-4
    logits, ns = O.medsos_lrcn_forward(sd, x, arch, meta["hidden"], meta["rnn_layers"], False)
    O.cross_entropy_mean(logits, y).backward()
    assert err(logits, torch.from_numpy(g["logits"])) < 1e-4
    for k, v in golden_tensors(g, "grad/").items():
        assert err(sd[k].grad, v, floor=1e-7) < 2e-3, k
    for k, v in golden_tensors(g, "gradsub16/").items():
        assert err(sd[k].grad[::16, ::16], v, floor=1e-7) < 2e-3, k
    for k, v in golden_tensors(g, "sd1/").items():
        assert err(ns[k], v) < 1e-4, k


def test_simple_and_crime_oracle_vs_golden():
    m, g, meta = build_backbone_model("ucf50_lrcn_resnet18.npz")
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.from_numpy(g["x"])
    logits, _ = O.simple_lrcn_forward(sd, x, "resnet18", meta["hidden"], meta["rnn_layers"])
    assert err(logits, torch.from_numpy(g["logits"])) < 1e-4
    m, g, meta = build_backbone_model("crime_lrcn_resnet18.npz")
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    logits, _ = O.simple_lrcn_forward(sd, x, "resnet18", meta["hidden"], meta["rnn_layers"], adapt_names=("adapt",),
                                      rnn_prefix="lstm.", num_heads=meta["num_classes"])
    assert err(logits, torch.from_numpy(g["logits"])) < 1e-4


def test_mobilenet_oracle_vs_golden():
    """MobileNetV2 restated by the oracle vs the reference medsos LRCN built on it (models.py:133-143): pooled features in
    train- and eval-mode BN, running statistics, logits."""
    m, g, meta = build_backbone_model("medsos_lrcn_mobilenet_v2.npz")
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.from_numpy(g["x"])
    frames = x.reshape(-1, *x.shape[2:])
    feat, ns = O.mobilenetv2_features(sd, frames, train=True)
    assert err(feat, torch.from_numpy(g["features"])) < 1e-4
    for k, v in golden_tensors(g, "sd1/").items():
        assert err(ns[k], v) < 1e-4, k
    feat_e, _ = O.mobilenetv2_features(sd, frames, train=False)
    assert err(feat_e, torch.from_numpy(g["features_eval"])) < 1e-4
    logits, _ = O.medsos_lrcn_forward(sd, x, "mobilenet_v2", meta["hidden"], meta["rnn_layers"], False)
    assert err(logits, torch.from_numpy(g["logits"])) < 1e-4


def test_densenet_oracle_vs_golden():
    """DenseNet-121 restated by the oracle vs the reference crime LRCN's own default backbone (lrcn.py:196-209): pooled
    features in train- and eval-mode BN, running statistics after the step, logits through the crime tail."""
    m, g, meta = build_backbone_model("crime_lrcn_densenet121.npz")
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.from_numpy(g["x"])
    frames = x.reshape(-1, *x.shape[2:])
    feat, ns = O.densenet_features(sd, frames, "densenet121", train=True)
    assert err(feat, torch.from_numpy(g["features"])) < 1e-4
    for k, v in golden_tensors(g, "sd1/").items():
        assert err(ns[k], v) < 1e-4, k
    feat_e, _ = O.densenet_features(sd, frames, "densenet121", train=False)
    assert err(feat_e, torch.from_numpy(g["features_eval"])) < 1e-4
    logits, _ = O.simple_lrcn_forward(sd, x, "densenet121", meta["hidden"], meta["rnn_layers"], adapt_names=("adapt",),
                                      rnn_prefix="lstm.", num_heads=meta["num_classes"])
    assert err(logits, torch.from_numpy(g["logits"])) < 1e-4


@pytest.mark.parametrize("tag", ["uni", "bi"])
def test_gru_oracle_vs_golden(tag):
    """Manual GRU cell of the oracle vs torch.nn.GRU outputs and gradients recorded from the reference's engine."""
    g, meta = load_golden(f"gru_{tag}.npz")
    p = {"lstm." + k: v.clone().requires_grad_(True) for k, v in golden_tensors(g, "p/").items()}
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    out = O.gru_forward(x, p, meta["H"], meta["layers"], meta["bidir"])
    (out * torch.from_numpy(g["w"])).sum().backward()
    assert err(out, torch.from_numpy(g["out"])) < 1e-5
    assert err(x.grad, torch.from_numpy(g["dx"])) < 1e-4
    for k, v in golden_tensors(g, "g/").items():
        assert err(p["lstm." + k].grad, v) < 1e-4, k


def test_smallcnn_gru_oracle_vs_golden():
    """LRCN2 (lrcn/backup_ucf50.py:105-151) restated by the oracle vs the reference class's own train step."""
    g, meta = load_golden("smallcnn_gru.npz")
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in golden_tensors(g, "sd0/").items()}
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    logits, _ = O.small_cnn_lrcn_forward(sd, x, meta["hidden"], gru=True)
    O.cross_entropy_mean(logits, y).backward()
    assert err(logits, torch.from_numpy(g["logits"])) < 1e-4
    gmax = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith("grad/"))
    for k, v in golden_tensors(g, "grad/").items():
        if k.startswith("conv") and k.endswith(".bias"):       # cancelled by the BatchNorm that follows: pure rounding noise
            assert sd[k].grad.abs().max().item() < 1e-4 * gmax, k
        else:
            assert err(sd[k].grad, v, floor=1e-7) < 2e-3, k


@pytest.mark.parametrize("tag", ["uni", "bi"])
def test_mamba_block_oracle_vs_golden(tag):
    """Mamba ResidualBlock restated by the oracle vs the output of the reference's own class (medsos models.py:107-117)."""
    g, meta = load_golden(f"mamba_block_{tag}.npz")
    sd = golden_tensors(g, "p/")
    for v in sd.values():
        v.requires_grad_(True)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    out = O.mamba_block_forward(sd, x, "", bidirectional=meta["bidir"])
    assert err(out, torch.from_numpy(g["out"])) < 1e-5
    (out * torch.from_numpy(g["r"])).sum().backward()            # the reference's autograd gradients of <out, r>
    assert err(x.grad, torch.from_numpy(g["dx"])) < 1e-4
    for k, v in golden_tensors(g, "g/").items():
        assert err(sd[k].grad, v, floor=1e-7) < 1e-4, k


def test_scan_oracle_vs_golden():
    g = np.load(os.path.join(GOLDEN, "scan.npz"))
    t = lambda k: torch.from_numpy(g[k])
    args = (t("u"), t("delta"), t("A"), t("B"), t("C"))
    assert err(O.selective_scan(*args, chunk_reset=256), t("y_videomamba")) < 1e-5
    assert err(O.selective_scan(*args, chunk_reset=None), t("y_medsos_fwd")) < 1e-5
    assert err(O.selective_scan(*args, chunk_reset=None, reverse=True), t("y_medsos_bwd")) < 1e-5


@pytest.mark.skipif(not refload.available(), reason="/root/reference only exists in the authoring container")
def test_oracle_vs_live_reference_smallcnn():
    torch.manual_seed(3)
    LRCN = refload.notebook_lrcn()
    m = LRCN(7, 3, 6, (3, 8, 8))
    m.dropout.p = 0.0
    m.train()
    x = torch.rand(2, 3, 3, 8, 8)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    ref = m(x)
    got, _ = O.small_cnn_lrcn_forward(sd, x, 6)
    assert err(got, ref) < 1e-5


@pytest.mark.parametrize("arch", ["resnet34", "resnet101", "densenet169", "densenet201", "mobilenet_v2"])
def test_oracle_backbone_configs_vs_torchvision(arch):
    """The oracle's depth / width tables for the backbones without a committed fixture, against torchvision's own modules
    (the arithmetic engine under the reference) with the reference's head surgery: train-mode and eval-mode
    features of a seeded random-init network."""
    import torchvision
    torch.manual_seed(23)
    net = getattr(torchvision.models, arch)(weights=None)
    if hasattr(net, "fc"):
        net.fc = torch.nn.Identity()
    else:
        net.classifier = torch.nn.Identity()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.rand(3, 3, 64, 64)
    fn = (O.mobilenetv2_features if arch.startswith("mobilenet") else
          O.densenet_features if arch.startswith("densenet") else O.resnet_features)
    for train in (True, False):
        net.train(train)
        net.load_state_dict(sd)
        with torch.no_grad():
            want = net(x)
            got, _ = fn(sd, x, arch, train, prefix="")
        # train mode: 12 samples per channel in the last stage of a 100-layer random-init network amplify fp32 summation-order
        # differences (3.8e-4 for resnet101); eval mode (fixed affine) pins the tables tightly
        assert err(got, want) < (2e-3 if train else 1e-4), (arch, train)


@pytest.mark.skipif(not refload.available(), reason="/root/reference only exists in the authoring container")
def test_oracle_and_module_surface_vs_live_rgb_lrcn():
    """lrcn/rgb_lrcn.py:168-263 (the multiclass sibling of the crime model: one `adapt`, biLSTM stored as `lstm`, one `fc`): the
    oracle's forward against the live reference class, and the replacement module's constructor / state_dict against it."""
    import video_classif_b200 as vc
    Rgb, g0 = refload.rgb_lrcn(CONF_CNN_BACKBONE="resnet18", CONF_RNN_LAYER=2, CONF_FINETUNE=False)
    torch.manual_seed(29)
    ref = Rgb(5, 3, 12, 16, cnn_backbone="resnet18").train()
    x = torch.rand(2, 3, 3, 32, 32)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    want = ref(x)
    got, _ = O.simple_lrcn_forward(sd, x, "resnet18", 12, 2, adapt_names=("adapt",), rnn_prefix="lstm.")
    assert err(got, want) < 1e-4
    torch.manual_seed(29)
    mine = vc.CrimeLRCN(5, 3, 12, 16, cnn_backbone="resnet18", rnn_layers=2, classif_mode="multiclass")
    mine_sd = mine.state_dict()
    assert set(mine_sd) == set(sd)
    for k, v in sd.items():
        assert mine_sd[k].shape == v.shape and torch.equal(mine_sd[k], v), k        # same construction order -> same seeded init
    assert [n for n, p in mine.named_parameters() if p.requires_grad] == [n for n, p in ref.named_parameters() if p.requires_grad]
