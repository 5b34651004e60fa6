"""GPU parity at the BASELINE.json config shapes: the drop-in modules vs golden outputs of the
reference's own classes run at those shapes (tests/golden/make_golden.py: gold_cfg1 / gold_cfg2 / gold_cfg3 /
gold_crime_trainable).  The fixtures hold seeds + outputs: weights are the seeded default init (bit-identical between
the reference class and the drop-in module, checked by checksum), clips are regenerated from their seed.

Every test asserts logits AND parameter gradients with the tolerance written next to the assert, and prints our
error beside torch's OWN bf16-autocast error on the reference module (`yard/*` in the fixture) where one exists.
A JSON summary of the measured errors goes to gpurun_out/parity_baseline_shapes.json."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, condition_backbone, err, load_golden, state_checksum

pytestmark = pytest.mark.gpu
DEV = "cuda"
_REPORT = {}


def _clips(meta, classes):
    g = torch.Generator().manual_seed(meta["clip_seed"])
    x = torch.randint(0, 256, (meta["B"], meta["T"], 3, meta["size"], meta["size"]), generator=g).float() / 255.0
    y = torch.randint(0, classes, (meta["B"],), generator=g)
    return x, y


def _grad_errors(g, params, skip=()):
    """{param: (our rel-to-max error, torch-autocast yardstick or None)} over every gradient the fixture holds."""
    out = {}
    for k in g.files:
        if k.startswith("grad/"):
            name, ref = k[5:], torch.from_numpy(g[k])
            got = params[name].grad
        elif k.startswith("gradsub/"):
            _, r, c, name = k.split("/", 3)
            ref = torch.from_numpy(g[k])
            got = params[name].grad
            got = got.reshape(got.shape[0], -1)[::int(r), ::int(c)]
        else:
            continue
        if name in skip:
            continue
        assert got is not None, name
        den = float(g["gradabs/" + name])
        e = (got.detach().double().cpu() - ref.double()).abs().max().item() / (den if den > 0 else 1.0)
        yk, sk = "yard/grad/" + name, "sens/grad/" + name
        out[name] = (e, float(g[yk]) if yk in g.files else None, float(g[sk]) if sk in g.files else None)
    return out


def _bound(floor, yard, factor=1.25):
    """tolerance = max(floor, factor x torch's own bf16-autocast error on the reference module)"""
    return max(floor, factor * (yard or 0.0))


def _report(tag, logits_err, yard_logits, gerrs):
    worst = sorted(gerrs.items(), key=lambda kv: -kv[1][0])[:6]
    print(f"\n[{tag}] logits rel err {logits_err:.3e}" + (f" (torch bf16 autocast on the reference: {yard_logits:.3e})" if yard_logits else ""))
    for k, (e, yd, sn) in worst:
        print(f"    grad {k:48s} {e:.3e}" + (f"   autocast {yd:.3e}" if yd is not None else "")
              + (f"   reference under a 1e-7 input perturbation {sn:.3e}" if sn is not None else ""))
    _REPORT[tag] = {"logits": logits_err, "logits_autocast": yard_logits,
                    "grads": {k: {"ours": e, "autocast": yd, "ref_sensitivity": sn} for k, (e, yd, sn) in gerrs.items()}}
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(_REPORT, open(os.path.join(ROOT, "gpurun_out", "parity_baseline_shapes.json"), "w"), indent=1)
    except OSError:
        pass


# ------------------------------------------------------------------------------------------------ cfg 1
def _cfg1(precision):
    import video_classif_b200 as vc
    g, meta = load_golden("cfg1_smallcnn.npz")
    torch.manual_seed(meta["seed"])
    m = vc.SmallCNNLRCN(meta["num_classes"], meta["T"], meta["hidden"], (3, meta["size"], meta["size"]), dropout=0.0,
                        precision=precision)
    cs = state_checksum(m.state_dict())
    if abs(cs - float(g["state_checksum"])) > 1e-6 * cs:
        pytest.skip("torch RNG stream differs from the authoring container")
    x, y = _clips(meta, meta["num_classes"])
    m = m.to(DEV).train()
    out = m(x.to(DEV))
    loss = torch.nn.functional.cross_entropy(out, y.to(DEV))
    loss.backward()
    return m, g, out, loss


def test_cfg1_smallcnn_fp32_exact_shape():
    """BASELINE.json configs[0] exactly (B=8, T=20, 3x64x64, 50 classes, H=32), fp32 path vs the notebook class
    (nb:148-193): logits <= 1e-4, loss 1e-4, running statistics <= 1e-4, argmax bit-equal; LSTM / fc gradients
    <= 1e-3 of their max.  CNN gradients: <= max(1e-3, 2 x the reference's own movement under a 1e-7 perturbation of
    the clips) -- at 31 M ReLU / max-pool decisions per step a handful flip under ANY last-bit difference, and one
    flip moves a conv gradient by ~1e-3 of its max (the reference moves by 1.6e-3 .. 6.2e-3: `sens/*` in the fixture;
    kernel by kernel the path is at 1e-6 of torch fp64, tools/probe_cfg1_parity.py)."""
    m, g, out, loss = _cfg1("fp32")
    ref = torch.from_numpy(g["logits"])
    e = err(out, ref)
    bias = tuple(f"conv{i}.bias" for i in (1, 2, 3))      # a bias ahead of train-mode BN has a zero true gradient
    ge = _grad_errors(g, dict(m.named_parameters()), skip=bias)
    _report("cfg1_fp32", e, None, ge)
    assert e < 1e-4
    assert abs(loss.item() - float(g["loss"])) < 1e-4
    assert torch.equal(out.argmax(1).cpu(), ref.argmax(1))
    for k, (v, _, sens) in ge.items():
        cnn = k.startswith(("conv", "bn"))
        assert v < (max(1e-3, 2.0 * sens) if cnn else 1e-3), (k, v, sens)
    gmax = max(float(g[k]) for k in g.files if k.startswith("gradabs/") and "bias" not in k)
    for k in bias:
        assert dict(m.named_parameters())[k].grad.abs().max().item() < 1e-4 * gmax, k
    sd1 = m.state_dict()
    for k in g.files:
        if k.startswith("sd1/"):
            v = torch.from_numpy(g[k])
            if v.dtype.is_floating_point:
                assert err(sd1[k[4:]], v) < 1e-4, k
            else:
                assert int(sd1[k[4:]]) == int(v), k


def test_cfg1_smallcnn_bf16_exact_shape():
    """Same shape through precision='bf16' (NHWC bf16 tensor-core trunk).  Every activation is stored in bf16, as under
    torch autocast, and train-mode BatchNorm re-normalises the rounding noise of each layer: torch's own CPU bf16 autocast
    on the notebook class is off by 2.0e-2 on the logits and 0.14-0.37 on the CNN gradients at this shape, so the north_star's
    1e-2 is out of reach of any bf16 execution of this model.  Bounds: logits <= max(1e-2, 2 x autocast), loss 2e-2, each
    gradient <= max(5e-2, 3 x autocast's error on that tensor), median gradient error <= 1.25 x autocast's median
    (measured: 0.174 vs 0.159)."""
    m, g, out, loss = _cfg1("bf16")
    ref = torch.from_numpy(g["logits"])
    e = err(out, ref)
    bias = tuple(f"conv{i}.bias" for i in (1, 2, 3))
    ge = _grad_errors(g, dict(m.named_parameters()), skip=bias)
    _report("cfg1_bf16", e, float(g["yard/logits"]), ge)
    assert e < _bound(1e-2, float(g["yard/logits"]), 2.0), e
    assert abs(loss.item() - float(g["loss"])) < 2e-2
    for k, (v, yd, _) in ge.items():
        assert v < _bound(5e-2, yd, 3.0), (k, v, yd)
    ours = sorted(v[0] for v in ge.values())
    auto = sorted(v[1] for v in ge.values())
    print(f"    median gradient error: ours {ours[len(ours) // 2]:.3e}, torch autocast {auto[len(auto) // 2]:.3e}")
    assert ours[len(ours) // 2] <= 1.25 * auto[len(auto) // 2]


# ------------------------------------------------------------------------------------------------ cfg 2 / cfg 3
# bf16 tolerance, as measured (profiles/README.md "parity at the BASELINE shapes"): a randomly initialised 50-layer
# ResNet in train-mode BatchNorm amplifies ANY bf16 rounding -- torch's own bf16 autocast on the reference module is off
# by 0.41 of the feature range and 3.6e-2 on the logits at these shapes (fp32 vs fp64: 7e-5), so the north_star's 1e-2 is
# not reachable by any bf16 execution of this model with its default init.  Each assert below is therefore
#     ours <= max(floor, factor x torch-autocast's error on the same inputs and weights)     (`yard/*` in the fixture)
# with factor 1.25 for the aggregates (features, median gradient error) and 2-3 for single quantities, which scatter from
# run to run (see _frozen_backbone_case),
# and the *_cond fixtures repeat the comparison with 'trained-like' conditioning (gamma 0.25 on every block's last
# BatchNorm; the reference loads ImageNet weights, which are unreachable here), where the floors are what binds.

def _backbone_model(fixture, cls, **kw):
    import video_classif_b200 as vc
    g, meta = load_golden(fixture)
    torch.manual_seed(meta["seed"])
    m = getattr(vc, cls)(meta["num_classes"], meta["T"], meta["hidden"], meta["rnn_input"], cnn_backbone=meta["arch"],
                         rnn_layers=meta["rnn_layers"], precision="bf16", **kw)
    if meta.get("conditioned"):
        condition_backbone(m.cnn_backbone)
    sd = m.state_dict()
    key = "backbone_checksum" if "backbone_checksum" in g.files else "state_checksum"
    cs = state_checksum(sd, "cnn_backbone." if key == "backbone_checksum" else "")
    if abs(cs - float(g[key])) > 1e-6 * cs:
        pytest.skip("torch RNG stream differs from the authoring container")
    return m, g, meta


def _yard(g, key):
    return float(g[key]) if key in g.files else None


def _frozen_backbone_case(fixture, cls, tag, floors, **kw):
    """One train step of a frozen-backbone LRCN at a BASELINE shape vs the reference class's fp32 golden."""
    f_feat, f_logit, f_grad = floors
    m, g, meta = _backbone_model(fixture, cls, **kw)
    x, y = _clips(meta, meta["num_classes"])
    m = m.to(DEV).train()
    xd = x.to(DEV)
    with torch.no_grad():
        feat = m._runner(xd.reshape(-1, 3, meta["size"], meta["size"]), True)
    fe = (feat[:, ::8].double().cpu() - torch.from_numpy(g["features_sub8"]).double()).abs().max().item() / float(g["features_absmax"])
    for mod in m.cnn_backbone.modules():          # the feature probe advanced the running statistics: back to 0 / 1
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.reset_running_stats()
    out = m(xd)
    loss = torch.nn.functional.cross_entropy(out, y.to(DEV))
    loss.backward()
    ref = torch.from_numpy(g["logits"])
    e = err(out, ref)
    ge = _grad_errors(g, dict(m.named_parameters()))
    _report(tag, e, _yard(g, "yard/logits"), ge)
    _REPORT[tag]["features"] = fe
    _REPORT[tag]["features_autocast"] = _yard(g, "yard/features")
    print(f"    pooled features rel err {fe:.3e} (torch bf16 autocast: {_yard(g, 'yard/features')})")
    assert fe < _bound(f_feat, _yard(g, "yard/features")), fe
    # Single quantities scatter from run to run: the order of the fp32 statistics atomics differs, and the default-init
    # network amplifies that last-bit difference like any other (12 runs on one box: logits 0.047-0.067 against the
    # yardstick's single draw of 0.047, one tensor's gradient up to 1.7 x its yardstick).  So: 2 x for the logits, 3 x per
    # gradient tensor (same order of magnitude), and the stable aggregates -- the features (1.25 x) and the MEDIAN over
    # the gradient tensors (<= 1.25 x the yardstick's median) -- carry the comparison.
    assert e < _bound(f_logit, _yard(g, "yard/logits"), 2.0), e
    assert abs(loss.item() - float(g["loss"])) < _bound(f_logit, _yard(g, "yard/logits"), 2.0) * max(1.0, abs(float(g["loss"])))
    for k, (v, yd, _) in ge.items():
        assert v < _bound(f_grad, yd, 3.0), (k, v, yd)
    ours = sorted(v[0] for v in ge.values())
    auto = sorted(v[1] for v in ge.values() if v[1] is not None)
    if auto:
        print(f"    median gradient error: ours {ours[len(ours) // 2]:.3e}, torch autocast {auto[len(auto) // 2]:.3e}")
        _REPORT[tag]["median_grad"] = {"ours": ours[len(ours) // 2], "autocast": auto[len(auto) // 2]}
        assert ours[len(ours) // 2] <= max(f_grad, 1.25 * auto[len(auto) // 2])
    sd1 = m.state_dict()
    for k in g.files:
        if k.startswith("sd1/"):                   # BatchNorm running statistics after one step (first layers: 1e-2)
            assert err(sd1[k[4:]], torch.from_numpy(g[k])) < (1e-2 if ".bn1.running" in k and "layer" not in k else 1e-1), k
    return m, g, out, ref


@pytest.mark.parametrize("tag", ["b8", "b64", "b8_cond"])
def test_cfg2_medsos_resnet50_bench_shape(tag):
    """BASELINE.json configs[1]: medsos LRCN (models.py:121-234), frozen ResNet-50 in train-mode BN, 16 x 112x112:
    an 8-clip slice (128 frames per BatchNorm batch), the bench's exact 64-clip batch, and the 8-clip slice with
    trained-like conditioning.  bf16 tcgen05 path vs the reference class's fp32 output; floors: pooled features 2e-2,
    logits 1e-2, tail gradients 5e-2 of their max; bound = max(floor, factor x torch autocast) with the factors of
    _frozen_backbone_case.  On the conditioned slice the north_star's own 1e-2 holds for the logits (measured 2.7e-3)."""
    m, g, out, ref = _frozen_backbone_case(f"cfg2_medsos_{tag}.npz", "LRCN", f"cfg2_{tag}", (2e-2, 1e-2, 5e-2), dropout=0.0)
    if tag == "b8_cond":
        assert err(out, ref) < 1e-2                                     # north_star bf16 tolerance, no yardstick needed
        assert torch.equal(out.argmax(1).cpu(), ref.argmax(1))


def test_cfg3_ucf50_resnet50_224():
    """BASELINE.json configs[2]: frozen ResNet-50 at 224x224 x 16 frames + 4-layer biLSTM H=56 (ucf50-lrcn.py:252-336),
    B=2; same floors and yardstick rule as cfg 2."""
    _frozen_backbone_case("cfg3_ucf50_224.npz", "UCF50LRCN", "cfg3", (2e-2, 1e-2, 5e-2))


# ------------------------------------------------------------------------------------------------ trainable backbone
@pytest.mark.parametrize("cond", [False, True], ids=["default_init", "conditioned"])
def test_crime_trainable_resnet18_gradients_vs_reference(cond):
    """crime LRCN with the whole ResNet-18 trainable (lrcn.py:181-305, CONF_FINETUNE=True), 8 clips x 8 frames x 64x64:
    logits and EVERY parameter gradient (backbone included) against the reference class's own autograd.
    Tolerance: logits <= max(1e-2, 2 x autocast); each gradient <= max(5e-2, 3 x torch's own bf16-autocast error
    on that tensor) and the MEDIAN gradient error <= 1.25 x the median autocast error (default init: autocast errors are
    0.2-0.6 on the backbone tensors -- bf16 gradients of a randomly initialised batch-statistics ResNet are noise for
    any implementation; the conditioned fixture is the meaningful one)."""
    fixture = "crime_trainable_resnet18_cond.npz" if cond else "crime_trainable_resnet18.npz"
    m, g, meta = _backbone_model(fixture, "CrimeLRCN", classif_mode="multiple_binary", finetune=True)
    x, _ = _clips(meta, meta["num_classes"])
    y = torch.from_numpy(g["y"])
    m = m.to(DEV).train()
    out = m(x.to(DEV))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y.to(DEV), reduction="mean")
    loss.backward()
    ref = torch.from_numpy(g["logits"])
    e = err(out, ref)
    ge = _grad_errors(g, dict(m.named_parameters()))
    tag = "crime_trainable_resnet18" + ("_cond" if cond else "")
    _report(tag, e, _yard(g, "yard/logits"), ge)
    ours = sorted(v[0] for v in ge.values())
    auto = sorted(v[1] for v in ge.values() if v[1] is not None)
    print(f"    median gradient error: ours {ours[len(ours) // 2]:.3e}, torch autocast {auto[len(auto) // 2]:.3e}")
    _REPORT[tag]["median_grad"] = {"ours": ours[len(ours) // 2], "autocast": auto[len(auto) // 2]}
    assert e < _bound(1e-2, _yard(g, "yard/logits"), 2.0), e
    assert abs(loss.item() - float(g["loss"])) < 1e-2
    bad = {k: v for k, v in ge.items() if v[0] > _bound(5e-2, v[1], 3.0)}
    assert not bad, bad
    assert ours[len(ours) // 2] <= max(2e-2, 1.25 * auto[len(auto) // 2])
