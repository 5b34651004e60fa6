"""Temporal-layer / adapt-stack variants of the backbone LRCNs (ucf50-lrcn.py gru / mamba, dump_lrcn.py lstm / gru under
`rnn`, models_bidir.py string-programmed Adapt + LN->SiLU head) against one train step of the reference's own classes
(tests/golden/make_golden.py::gold_variants).  The trainable tail runs in fp32 on the reference's pooled backbone features,
so the bounds are the fp32 ones: logits 1e-4, loss 1e-4, every tail gradient 2e-3 of its max; checkpoint keys load strictly."""
import pytest
import torch

from conftest import err, golden_tensors, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
CASES = ["ucf50_gru", "ucf50_mamba", "dump_gru", "dump_lstm", "adapt_lstm", "adapt_gru_bi", "adapt_mamba"]


@pytest.mark.parametrize("tag", CASES)
def test_variant_tail_vs_reference_golden(tag):
    import video_classif_b200 as vc
    g, meta = load_golden(f"variant_{tag}.npz")
    build = meta["build"]
    torch.manual_seed(0)
    m = getattr(vc, build["cls"])(precision="fp32", **build["kw"])
    tail = golden_tensors(g, "sd0/")
    missing, unexpected = m.load_state_dict(tail, strict=False)
    assert not unexpected and all(k.startswith("cnn_backbone.") for k in missing), (missing[:4], unexpected[:4])
    if hasattr(m, "adapt") and hasattr(m.adapt, "precision"):
        m.adapt.precision = "fp32"
    for p in m.cnn_backbone.parameters():
        p.requires_grad = False
    m = m.to(DEV).train()
    x = torch.from_numpy(g["x"])
    B, T = x.shape[:2]
    feat = torch.from_numpy(g["features"]).reshape(B, T, -1).to(DEV)
    object.__setattr__(m, "_features", lambda _x: feat)
    out = m(x.to(DEV))
    y = torch.from_numpy(g["y"]).to(DEV)
    if meta["loss"] == "ce":
        loss = torch.nn.functional.cross_entropy(out, y)
    else:
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y, reduction="mean")
    loss.backward()
    ref = torch.from_numpy(g["logits"])
    assert err(out, ref) < 1e-4, err(out, ref)
    assert abs(loss.item() - float(g["loss"])) < 1e-4
    params = dict(m.named_parameters())
    worst = 0.0
    for k, v in golden_tensors(g, "grad/").items():
        if k.startswith("cnn_backbone."):
            continue
        e = err(params[k].grad, v, floor=1e-6)
        worst = max(worst, e)
        assert e < 2e-3, (k, e)
    print(f"\n[variant {tag}] logits {err(out, ref):.2e}, worst tail gradient {worst:.2e}")
