import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def attempts(n=3):
    """Decorator for tests whose asserted statistic is a random variable of THIS implementation's run: the batch statistics and
    split weight gradients are reduced with fp32 atomics whose order differs from run to run, and a randomly initialised
    batch-statistics network turns that last-bit difference into flipped ReLU masks, so a per-parameter gradient error at the
    edge of its bound lands on either side (observed: about one run in thirty).  The test body is repeated on AssertionError, at
    most n times; a defect in a kernel fails every attempt.  Every failed attempt is printed."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            last = None
            for i in range(n):
                try:
                    return fn(*args, **kwargs)
                except AssertionError as e:
                    last = e
                    print(f"[attempt {i + 1}/{n} of {fn.__name__} failed] {str(e)[:300]}")
            raise last
        return wrapper
    return deco


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    meta = json.loads(str(g["meta"])) if "meta" in g.files else {}
    return g, meta


def golden_tensors(g, prefix):
    import torch
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


def state_checksum(sd, prefix=""):
    return sum(float(v.double().abs().sum()) for k, v in sorted(sd.items())
               if k.startswith(prefix) and v.dtype.is_floating_point)


def err(a, b, floor=0.0):
    """max|a-b| / max(max|b|, floor)"""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = max(b.abs().max().item(), floor)
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


def build_backbone_model(fixture):
    """Re-creates the seeded model a backbone fixture was generated from (weights are
    seed-reproducible: checked against the stored checksum)."""
    import torch
    import video_classif_b200 as vc
    g, meta = load_golden(fixture)
    torch.manual_seed(meta["seed"])
    src = meta["source"]
    if "models.py" in src:
        m = vc.LRCN(meta["num_classes"], meta["T"], meta["hidden"], meta["rnn_input"], cnn_backbone=meta["arch"],
                    rnn_layers=meta["rnn_layers"], dropout=0.0, precision="bf16")
    elif "ucf50" in src:
        m = vc.UCF50LRCN(meta["num_classes"], meta["T"], meta["hidden"], meta["rnn_input"], cnn_backbone=meta["arch"],
                         rnn_layers=meta["rnn_layers"], precision="bf16")
    else:
        m = vc.CrimeLRCN(meta["num_classes"], meta["T"], meta["hidden"], meta["rnn_input"], cnn_backbone=meta["arch"],
                         rnn_layers=meta["rnn_layers"], classif_mode="multiple_binary", precision="bf16")
    sd = m.state_dict()
    cs = state_checksum(sd, "cnn_backbone.")
    if abs(cs - float(g["backbone_checksum"])) > 1e-6 * cs:
        pytest.skip("torch RNG stream differs from the authoring container: seeded backbone not reproducible")
    for k in g.files:
        if k.startswith("sd0/"):
            assert np.array_equal(g[k], sd[k[4:]].numpy()), k
    return m, g, meta


def condition_backbone(net, gamma=0.25):
    """Same 'trained-like' conditioning tests/golden/make_golden.py applied before generating the *_cond fixtures:
    gamma = 0.25 on the last BatchNorm of every residual block (bn3 of a Bottleneck, bn2 of a BasicBlock)."""
    import torch
    mods = dict(net.named_modules())
    with torch.no_grad():
        for name, mod in mods.items():
            if not isinstance(mod, torch.nn.BatchNorm2d) or "." not in name:
                continue
            parent = mods[name.rsplit(".", 1)[0]]
            if name.endswith(".bn3") or (name.endswith(".bn2") and not hasattr(parent, "bn3")):
                mod.weight.fill_(gamma)
