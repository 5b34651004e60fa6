"""Generates tests/golden/*.npz|json from the REFERENCE ITSELF (authoring container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference (read-only) + cv2; never runs on the GPU box.  Every fixture records the
reference file:line whose output it holds.  Weights come from the reference module's own default
init under a fixed torch seed; small models store their full state_dict, torchvision-backbone
models store the trainable tail + a checksum of the (seed-reproducible) backbone tensors.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def npd(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print("wrote", name, sum(a.nbytes for a in arrs.values() if hasattr(a, "nbytes")) // 1024, "KiB raw")


def checksum(sd, prefix):
    tot = 0.0
    for k in sorted(sd):
        if k.startswith(prefix) and sd[k].dtype.is_floating_point:
            tot += float(sd[k].double().abs().sum())
    return tot


def run_step(model, x, y, loss_kind="ce"):
    model.train()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    out = model(x)
    if loss_kind == "ce":
        loss = torch.nn.functional.cross_entropy(out, y)
    else:
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y, reduction="mean")
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    sd1 = {k: v.clone() for k, v in model.state_dict().items()}
    return sd0, out.detach(), loss.detach(), grads, sd1


def gold_sampling():
    us, df = refload.sampling_functions()
    table = {}
    for n, T in [(100, 40), (80, 40), (79, 40), (45, 40), (399, 20), (21, 20), (20, 20), (30, 40),
                 (7, 20), (1, 16), (125, 60), (300, 60), (61, 60), (59, 60), (1000, 30), (16, 16),
                 (33, 16), (47, 16), (2, 3)]:
        fr = list(range(n))
        r = us(fr, T)
        if len(r) < T:
            r = df(r, T)
        table[f"{n},{T}"] = r
    # known-answer rows of SURVEY.md section 8(a.1) for variants that live inside loader functions
    crime = {"30,40": list(range(30)) + [-1] * 10,          # lrcn/lrcn.py:151-155 zero-frame padding
             "80,40": list(range(0, 80, 2)), "100,40": list(range(0, 100, 2))[:40]}
    seek = {"125,60": [i * 2 for i in range(60)], "300,60": [i * 5 for i in range(60)],
            "59,60": None}                                    # backup_ucf50.py:52-62
    json.dump({"source": "medsos_lrcn/src/loader_data.py:35-51 run on list(range(n))",
               "medsos": table, "crime": crime, "seek": seek},
              open(os.path.join(OUT, "sampling.json"), "w"), indent=0)
    print("wrote sampling.json")


def gold_resize():
    import cv2
    rng = np.random.default_rng(7)
    arrs = {}
    cases = [(48, 64, 32, 32), (36, 64, 28, 28), (20, 30, 32, 48), (64, 64, 32, 32), (32, 32, 32, 32),
             (45, 80, 16, 16), (9, 7, 24, 24)]
    for i, (h, w, oh, ow) in enumerate(cases):
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        arrs[f"src{i}"] = src
        arrs[f"dst{i}"] = cv2.resize(src, (ow, oh))      # loader_data.py:162 (default INTER_LINEAR)
        arrs[f"rgb{i}"] = cv2.cvtColor(arrs[f"dst{i}"], cv2.COLOR_BGR2RGB)  # loader_data.py:163
    arrs["cv2_version"] = np.array(cv2.__version__)
    save("resize_cv2.npz", **arrs)


def gold_smallcnn():
    LRCN = refload.notebook_lrcn()
    for tag, (C, T, H, S, B) in {"a": (5, 4, 8, 16, 3), "b": (50, 6, 32, 32, 4)}.items():
        torch.manual_seed(100 + ord(tag))
        m = LRCN(C, T, H, (3, S, S))
        m.dropout.p = 0.0                                  # parity runs use p=0
        with torch.no_grad():                               # non-trivial BN affine / running stats
            for k in (1, 2, 3):
                bn = getattr(m, f"bn{k}")
                bn.weight.uniform_(0.5, 1.5)
                bn.bias.uniform_(-0.3, 0.3)
                bn.running_mean.uniform_(-1, 1)
                bn.running_var.uniform_(0.5, 2.0)
        g = torch.Generator().manual_seed(1234)
        x = torch.randint(0, 256, (B, T, 3, S, S), generator=g).float()   # raw 0..255 (backup_ucf50.py:101)
        if tag == "b":
            x = x / 255.0
        y = torch.randint(0, C, (B,), generator=g)
        sd0, out, loss, grads, sd1 = run_step(m, x, y)
        m.eval()
        with torch.no_grad():
            out_eval = m(x)
        arrs = {"x": x.numpy(), "y": y.numpy(), "logits": out.numpy(), "loss": loss.numpy(),
                "logits_eval_after": out_eval.numpy(),
                "meta": np.array(json.dumps(dict(num_classes=C, T=T, hidden=H, size=S, B=B,
                                                 source="nb:148-193 LRCN, dropout p=0")))}
        for k, v in npd(sd0).items():
            arrs["sd0/" + k] = v
        for k, v in npd(grads).items():
            arrs["grad/" + k] = v
        for k, v in npd(sd1).items():
            if "running" in k or "num_batches" in k:
                arrs["sd1/" + k] = v
        save(f"smallcnn_lrcn_{tag}.npz", **arrs)


def gold_medsos():
    for arch, S, B, T in [("resnet18", 32, 2, 3), ("resnet50", 64, 2, 2)]:
        mm = refload.medsos_models(CONF_RNN_LAYER=3, CONF_RNN_OUT="all", CONF_CLASSIF_MODE="multiclass",
                                   CONF_DROPOUT=0.0)
        torch.manual_seed(7)
        m = mm.LRCN(4, T, 32, 8, cnn_backbone=arch, rnn_type="lstm", rnn_out="all", bidirectional=False)
        g = torch.Generator().manual_seed(1234)
        x = torch.randint(0, 256, (B, T, 3, S, S), generator=g).float() / 255.0
        y = torch.randint(0, 4, (B,), generator=g)
        sd0, out, loss, grads, sd1 = run_step(m, x, y)
        arrs = {"x": x.numpy(), "y": y.numpy(), "logits": out.numpy(), "loss": loss.numpy(),
                "backbone_checksum": np.array(checksum(sd0, "cnn_backbone.")),
                "meta": np.array(json.dumps(dict(arch=arch, size=S, B=B, T=T, hidden=32, rnn_input=8,
                                                 rnn_layers=3, num_classes=4, seed=7,
                                                 source="medsos_lrcn/src/models.py:121-234, dropout 0")))}
        big = lambda k, v: arch == "resnet50" and v.size > 200000   # seed-reproducible; keep fixture small
        arrs["tail_checksum"] = np.array(checksum({k: v for k, v in sd0.items() if not k.startswith("cnn_backbone.")}, ""))
        for k, v in npd(sd0).items():
            if not k.startswith("cnn_backbone.") and not big(k, v):
                arrs["sd0/" + k] = v
        for k, v in npd(grads).items():
            if big(k, v):
                arrs["gradsub16/" + k] = v[::16, ::16].copy()
            else:
                arrs["grad/" + k] = v
        for k in ("cnn_backbone.bn1.running_mean", "cnn_backbone.bn1.running_var",
                  "cnn_backbone.layer4.1.bn2.running_var", "cnn_backbone.layer2.0.downsample.1.running_mean"):
            arrs["sd1/" + k] = sd1[k].numpy()
        # pooled backbone features (pre-adapt) pin the CNN on its own
        with torch.no_grad():
            m2 = mm.LRCN(4, T, 32, 8, cnn_backbone=arch, rnn_type="lstm", rnn_out="all", bidirectional=False)
            m2.load_state_dict(sd0)
            m2.train()
            feat = m2.cnn_backbone(x.view(B * T, 3, S, S))
        arrs["features"] = feat.numpy()
        save(f"medsos_lrcn_{arch}.npz", **arrs)


def gold_simple():
    # ucf50-lrcn.py topology: frozen backbone, 3 plain adapts, 2-layer biLSTM H=12, multiclass
    C, g0 = refload.ucf50_lrcn(CONF_CNN_BACKBONE="resnet18", CONF_RNN_LAYER=2)
    torch.manual_seed(11)
    m = C(5, 3, 12, 16, cnn_backbone="resnet18")
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, 256, (2, 3, 3, 32, 32), generator=g).float() / 255.0
    y = torch.randint(0, 5, (2,), generator=g)
    sd0, out, loss, grads, sd1 = run_step(m, x, y)
    arrs = {"x": x.numpy(), "y": y.numpy(), "logits": out.numpy(), "loss": loss.numpy(),
            "backbone_checksum": np.array(checksum(sd0, "cnn_backbone.")),
            "meta": np.array(json.dumps(dict(arch="resnet18", size=32, B=2, T=3, hidden=12, rnn_input=16,
                                             rnn_layers=2, num_classes=5, seed=11,
                                             source="lrcn/ucf50-lrcn.py:252-336")))}
    for k, v in npd(sd0).items():
        if not k.startswith("cnn_backbone."):
            arrs["sd0/" + k] = v
    for k, v in npd(grads).items():
        arrs["grad/" + k] = v
    save("ucf50_lrcn_resnet18.npz", **arrs)

    # crime lrcn.py topology: frozen backbone (CONF_FINETUNE False), one adapt, per-class binary heads
    C, g0 = refload.crime_lrcn(CONF_CNN_BACKBONE="resnet18", CONF_RNN_LAYER=2,
                               CONF_CLASSIF_MODE="multiple_binary", CONF_FINETUNE=False)
    torch.manual_seed(13)
    m = C(3, 3, 12, 16, cnn_backbone="resnet18")
    yb = torch.tensor([[1., 0., 0.], [0., 0., 1.]])
    sd0, out, loss, grads, sd1 = run_step(m, x, yb, loss_kind="bce")
    arrs = {"x": x.numpy(), "y": yb.numpy(), "logits": out.numpy(), "loss": loss.numpy(),
            "backbone_checksum": np.array(checksum(sd0, "cnn_backbone.")),
            "meta": np.array(json.dumps(dict(arch="resnet18", size=32, B=2, T=3, hidden=12, rnn_input=16,
                                             rnn_layers=2, num_classes=3, seed=13,
                                             source="lrcn/lrcn.py:181-305 multiple_binary; loss=mean BCEWithLogits")))}
    for k, v in npd(sd0).items():
        if not k.startswith("cnn_backbone."):
            arrs["sd0/" + k] = v
    for k, v in npd(grads).items():
        arrs["grad/" + k] = v
    save("crime_lrcn_resnet18.npz", **arrs)


def gold_crime_densenet():
    # crime lrcn.py topology with ITS default backbone (CONF_CNN_BACKBONE = densenet121, lrcn.py:196-209), frozen
    C, g0 = refload.crime_lrcn(CONF_CNN_BACKBONE="densenet121", CONF_RNN_LAYER=2, CONF_CLASSIF_MODE="multiple_binary",
                               CONF_FINETUNE=False)
    torch.manual_seed(17)
    m = C(3, 3, 12, 16, cnn_backbone="densenet121")
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, 256, (2, 3, 3, 64, 64), generator=g).float() / 255.0
    yb = torch.tensor([[1., 0., 0.], [0., 0., 1.]])
    sd0, out, loss, grads, sd1 = run_step(m, x, yb, loss_kind="bce")
    arrs = {"x": x.numpy(), "y": yb.numpy(), "logits": out.numpy(), "loss": loss.numpy(),
            "backbone_checksum": np.array(checksum(sd0, "cnn_backbone.")),
            "meta": np.array(json.dumps(dict(arch="densenet121", size=64, B=2, T=3, hidden=12, rnn_input=16,
                                             rnn_layers=2, num_classes=3, seed=17,
                                             source="lrcn/lrcn.py:181-305 multiple_binary; loss=mean BCEWithLogits")))}
    for k, v in npd(sd0).items():
        if not k.startswith("cnn_backbone."):
            arrs["sd0/" + k] = v
    for k, v in npd(grads).items():
        arrs["grad/" + k] = v
    for k in ("cnn_backbone.features.norm0.running_mean", "cnn_backbone.features.denseblock1.denselayer3.norm1.running_var",
              "cnn_backbone.features.transition2.norm.running_mean", "cnn_backbone.features.denseblock4.denselayer16.norm2.running_var",
              "cnn_backbone.features.norm5.running_var"):
        arrs["sd1/" + k] = sd1[k].numpy()
    with torch.no_grad():                       # pooled backbone features (train-mode BN) and eval-mode features
        m2 = C(3, 3, 12, 16, cnn_backbone="densenet121")
        m2.load_state_dict(sd0)
        m2.train()
        arrs["features"] = m2.cnn_backbone(x.view(6, 3, 64, 64)).numpy()
        m2.load_state_dict(sd0)
        m2.eval()
        arrs["features_eval"] = m2.cnn_backbone(x.view(6, 3, 64, 64)).numpy()
    save("crime_lrcn_densenet121.npz", **arrs)


def gold_medsos_mobilenet():
    # medsos models.py with mobilenet_v2 (the second backbone of its search space, automation.py:28), frozen, train-mode BN
    mm = refload.medsos_models(CONF_RNN_LAYER=2, CONF_RNN_OUT="all", CONF_CLASSIF_MODE="multiclass", CONF_DROPOUT=0.0)
    torch.manual_seed(19)
    m = mm.LRCN(4, 3, 16, 8, cnn_backbone="mobilenet_v2", rnn_type="lstm", rnn_out="all", bidirectional=False)
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, 256, (2, 3, 3, 64, 64), generator=g).float() / 255.0
    y = torch.randint(0, 4, (2,), generator=g)
    sd0, out, loss, grads, sd1 = run_step(m, x, y)
    arrs = {"x": x.numpy(), "y": y.numpy(), "logits": out.numpy(), "loss": loss.numpy(),
            "backbone_checksum": np.array(checksum(sd0, "cnn_backbone.")),
            "meta": np.array(json.dumps(dict(arch="mobilenet_v2", size=64, B=2, T=3, hidden=16, rnn_input=8, rnn_layers=2,
                                             num_classes=4, seed=19, source="medsos_lrcn/src/models.py:121-234, dropout 0")))}
    big = lambda v: v.size > 100000               # seed-reproducible (checksum above): keep the fixture small
    for k, v in npd(sd0).items():
        if not k.startswith("cnn_backbone.") and not big(v):
            arrs["sd0/" + k] = v
    for k, v in npd(grads).items():
        if big(v):
            arrs["gradsub16/" + k] = v[::16, ::16].copy()
        else:
            arrs["grad/" + k] = v
    for k in ("cnn_backbone.features.0.1.running_mean", "cnn_backbone.features.3.conv.1.1.running_var",
              "cnn_backbone.features.14.conv.3.running_mean", "cnn_backbone.features.18.1.running_var"):
        arrs["sd1/" + k] = sd1[k].numpy()
    with torch.no_grad():
        m2 = mm.LRCN(4, 3, 16, 8, cnn_backbone="mobilenet_v2", rnn_type="lstm", rnn_out="all", bidirectional=False)
        m2.load_state_dict(sd0)
        m2.train()
        arrs["features"] = m2.cnn_backbone(x.view(6, 3, 64, 64)).numpy()
        m2.load_state_dict(sd0)
        m2.eval()
        arrs["features_eval"] = m2.cnn_backbone(x.view(6, 3, 64, 64)).numpy()
    save("medsos_lrcn_mobilenet_v2.npz", **arrs)


def gold_lstm():
    # nn.LSTM exactly as the reference configures it (lrcn.py:236: 4-layer biLSTM H=56 -> here 2x2, H=7)
    torch.manual_seed(5)
    for tag, (inp, H, layers, bidir, B, T) in {"uni": (10, 6, 2, False, 3, 5), "bi": (9, 7, 2, True, 2, 4)}.items():
        rnn = torch.nn.LSTM(inp, H, num_layers=layers, bidirectional=bidir, batch_first=True)
        x = torch.randn(B, T, inp, requires_grad=True)
        out, _ = rnn(x)
        w = torch.randn_like(out)
        (out * w).sum().backward()
        arrs = {"x": x.detach().numpy(), "out": out.detach().numpy(), "w": w.numpy(), "dx": x.grad.numpy(),
                "meta": np.array(json.dumps(dict(inp=inp, H=H, layers=layers, bidir=bidir)))}
        for k, p in rnn.named_parameters():
            arrs["p/" + k] = p.detach().numpy()
            arrs["g/" + k] = p.grad.numpy()
        save(f"lstm_{tag}.npz", **arrs)


def gold_gru():
    # nn.GRU as the reference configures it (backup_ucf50.py:126 1-layer bidirectional; models.py:167-169 N layers)
    torch.manual_seed(6)
    for tag, (inp, H, layers, bidir, B, T) in {"uni": (10, 6, 2, False, 3, 5), "bi": (9, 7, 2, True, 2, 4)}.items():
        rnn = torch.nn.GRU(inp, H, num_layers=layers, bidirectional=bidir, batch_first=True)
        x = torch.randn(B, T, inp, requires_grad=True)
        out, _ = rnn(x)
        w = torch.randn_like(out)
        (out * w).sum().backward()
        arrs = {"x": x.detach().numpy(), "out": out.detach().numpy(), "w": w.numpy(), "dx": x.grad.numpy(),
                "meta": np.array(json.dumps(dict(inp=inp, H=H, layers=layers, bidir=bidir)))}
        for k, p in rnn.named_parameters():
            arrs["p/" + k] = p.detach().numpy()
            arrs["g/" + k] = p.grad.numpy()
        save(f"gru_{tag}.npz", **arrs)
    # the reference's own LRCN2 class (small CNN + biGRU), one train step
    LRCN2 = refload.backup_lrcn2()
    C, T, H, S, B = 5, 4, 8, 16, 3
    torch.manual_seed(77)
    m = LRCN2(C, T, H, (3, S, S))
    m.dropout.p = 0.0
    gen = torch.Generator().manual_seed(4321)
    x = torch.randint(0, 256, (B, T, 3, S, S), generator=gen).float() / 255.0
    y = torch.randint(0, C, (B,), generator=gen)
    sd0, out, loss, grads, sd1 = run_step(m, x, y)
    arrs = {"x": x.numpy(), "y": y.numpy(), "logits": out.numpy(), "loss": loss.numpy(),
            "meta": np.array(json.dumps(dict(num_classes=C, T=T, hidden=H, size=S, B=B,
                                             source="lrcn/backup_ucf50.py:105-151 LRCN2, dropout p=0")))}
    for k, v in npd(sd0).items():
        arrs["sd0/" + k] = v
    for k, v in npd(grads).items():
        arrs["grad/" + k] = v
    save("smallcnn_gru.npz", **arrs)


def gold_mamba():
    # the reference's own ResidualBlock (medsos models.py:107-117) as LRCN builds it: d_inner = 2 d_model, n_state = dt_rank = H
    mod = refload.medsos_models()
    for tag, bidir in (("uni", False), ("bi", True)):
        torch.manual_seed(31 + bidir)
        blk = mod.ResidualBlock(8, 16, 32, 32, bidirectional=bidir).eval()
        x = torch.randn(3, 16, 8)
        x.requires_grad_(True)
        r = torch.randn(3, 16, 8)                                    # loss = <out, r>: the reference's own autograd gradients
        out = blk(x)
        (out * r).sum().backward()
        arrs = {"x": x.detach().numpy(), "out": out.detach().numpy(), "r": r.numpy(), "dx": x.grad.numpy(),
                "meta": np.array(json.dumps(dict(bidir=bidir, d_model=8, n_state=32)))}
        for k, v in blk.state_dict().items():
            arrs["p/" + k] = v.numpy()
        for k, v in blk.named_parameters():
            if v.grad is not None:                                   # parameters the forward never touches have none
                arrs["g/" + k] = v.grad.numpy()
        save(f"mamba_block_{tag}.npz", **arrs)


def save_as(model, path, module_name, extra=()):
    """torch.save(model) with the class pickled under the import path the reference's own scripts give it
    (`models.LRCN` via main.py:6, `__main__.LRCN` for a script run directly)."""
    import sys, types
    classes = {type(model)} | set(extra)
    fake = types.ModuleType(module_name)
    prev_mod = sys.modules.get(module_name)
    prev = {c: (c.__module__, c.__qualname__) for c in classes}
    try:
        for c in classes:
            c.__module__, c.__qualname__ = module_name, c.__name__
            setattr(fake, c.__name__, c)
        if module_name == "__main__":
            for c in classes:
                setattr(prev_mod, c.__name__, c)
        else:
            sys.modules[module_name] = fake
        torch.save(model, path)
    finally:
        for c, (m, q) in prev.items():
            c.__module__, c.__qualname__ = m, q
        if module_name != "__main__":
            if prev_mod is None:
                sys.modules.pop(module_name, None)
            else:
                sys.modules[module_name] = prev_mod


def gold_ckpt():
    # whole-module pickles as train_eval.py:53 / ucf50-lrcn.py:468 write them, tiny shapes (committed fixtures)
    here = os.path.dirname(os.path.abspath(__file__))
    torch.manual_seed(41)
    m = refload.notebook_lrcn()(5, 4, 8, input_shape=(3, 16, 16)).eval()
    x = torch.rand(2, 4, 3, 16, 16)
    with torch.no_grad():
        y = m(x)
    save_as(m, os.path.join(here, "ckpt_smallcnn_lstm.pt"), "__main__")
    torch.manual_seed(42)
    m2 = refload.backup_lrcn2()(3, 4, 8, (3, 16, 16)).eval()
    with torch.no_grad():
        y2 = m2(x)
    save_as(m2, os.path.join(here, "ckpt_smallcnn_gru.pt"), "__main__")
    save("ckpt_io.npz", x=x.numpy(), y_lstm=y.numpy(), y_gru=y2.numpy())


def gold_scan():
    torch.manual_seed(3)
    Bz, L, D, N = 2, 300, 12, 4
    u = torch.randn(Bz, L, D)
    delta = torch.nn.functional.softplus(torch.randn(Bz, L, D))
    A = -torch.exp(torch.randn(D, N))
    Bm = torch.randn(Bz, L, N)
    Cm = torch.randn(Bz, L, N)
    y_vm = refload.videomamba_scan()(u, delta, A, Bm, Cm)        # lrcn/videomamba.py:242-284 (reset @256)
    y_f = refload.medsos_scan("forward")(u, delta, A, Bm, Cm)    # medsos models.py:47-71
    y_b = refload.medsos_scan("backward")(u, delta, A, Bm, Cm)
    save("scan.npz", u=u.numpy(), delta=delta.numpy(), A=A.numpy(), B=Bm.numpy(), C=Cm.numpy(),
         y_videomamba=y_vm.numpy(), y_medsos_fwd=y_f.numpy(), y_medsos_bwd=y_b.numpy())


# ------------------------------------------------------------------------------------------------
# BASELINE.json config shapes (round 2).  These fixtures hold SEEDS + OUTPUTS only: the weights are the reference
# class's default init under torch.manual_seed(seed) (bit-identical to the drop-in module built under the same seed;
# the tests check a checksum), the clips are torch.randint under Generator(1234).  Big gradient tensors are stored
# as strided sub-samples.  "yard/*" entries are the error of the REFERENCE ITSELF under torch's CPU bf16 autocast
# against its own fp32 run on the same inputs -- the bf16 yardstick the GPU tests print beside our error.
# ------------------------------------------------------------------------------------------------

def _sub(v, limit=70000):
    """strided sub-sample spec (row step, col step) keeping a tensor under `limit` elements"""
    if v.ndim < 2 or v.size <= limit:
        return None
    r = c = 1
    rows, cols = v.shape[0], int(np.prod(v.shape[1:]))
    while (rows // r) * (cols // c) > limit:
        if cols // c >= rows // r:
            c *= 2
        else:
            r *= 2
    return r, c


def _store_grads(arrs, grads, prefix="grad/"):
    for k, v in npd(grads).items():
        sp = _sub(v)
        if sp is None:
            arrs[prefix + k] = v
        else:
            r, c = sp
            arrs[f"gradsub/{r}/{c}/" + k] = v.reshape(v.shape[0], -1)[::r, ::c].copy()
        arrs["gradabs/" + k] = np.array(float(np.abs(v).max()))


def _relerr(a, b):
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def _autocast_yard(model, x, y, out32, grads32, loss_kind="ce"):
    """errors of torch's own CPU bf16 autocast on the reference module vs its fp32 run (BN buffers restored after)"""
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.zero_grad()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out = model(x)
        if loss_kind == "ce":
            loss = torch.nn.functional.cross_entropy(out.float(), y)
        else:
            loss = torch.nn.functional.binary_cross_entropy_with_logits(out.float(), y, reduction="mean")
    loss.backward()
    yard = {"yard/logits": np.array(_relerr(out.float(), out32))}
    for k, p in model.named_parameters():
        if p.grad is not None and k in grads32:
            yard["yard/grad/" + k] = np.array(_relerr(p.grad.float(), grads32[k]))
    model.load_state_dict(sd)
    model.zero_grad()
    return yard


def _clips(B, T, S, classes, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (B, T, 3, S, S), generator=g).float() / 255.0
    y = torch.randint(0, classes, (B,), generator=g)
    return x, y


def condition_backbone(net, gamma=0.25):
    """'Trained-like' conditioning of a randomly initialised torchvision ResNet (the reference loads ImageNet weights,
    which cannot be fetched here): the LAST BatchNorm of every residual block gets gamma = 0.25, so the residual
    branches are perturbations of the shortcut path as in a trained network.  With the default gamma = 1 a 50-layer
    batch-statistics ResNet amplifies ANY bf16 rounding (torch's own autocast included) to 0.4 of the feature range.
    tests/conftest.py applies the same function to the drop-in module."""
    with torch.no_grad():
        for name, mod in net.named_modules():
            if isinstance(mod, torch.nn.BatchNorm2d) and (name.endswith(".bn3") or (name.endswith(".bn2") and not hasattr(
                    dict(net.named_modules())[name.rsplit(".", 1)[0]], "bn3"))):
                mod.weight.fill_(gamma)


def gold_cfg1():
    """BASELINE.json configs[0]: notebook small-CNN LRCN, 20 frames x 64x64, 50 classes, batch 8, hidden 32 (nb:148-193)."""
    LRCN = refload.notebook_lrcn()
    C, T, H, S, B, seed = 50, 20, 32, 64, 8, 201
    torch.manual_seed(seed)
    m = LRCN(C, T, H, (3, S, S))
    m.dropout.p = 0.0
    x, y = _clips(B, T, S, C)
    sd0, out, loss, grads, sd1 = run_step(m, x, y)
    arrs = {"logits": out.numpy(), "loss": loss.numpy(), "state_checksum": np.array(checksum(sd0, "")),
            "meta": np.array(json.dumps(dict(num_classes=C, T=T, hidden=H, size=S, B=B, seed=seed, clip_seed=1234,
                                             source="nb:148-193 LRCN, dropout p=0; BASELINE.json configs[0]")))}
    _store_grads(arrs, grads)
    for k, v in npd(sd1).items():
        if "running" in k or "num_batches" in k:
            arrs["sd1/" + k] = v
    m.load_state_dict(sd0)
    arrs.update(_autocast_yard(m, x, y, out, grads))
    # sensitivity yardstick: how far the REFERENCE'S OWN fp32 gradients move when the clips are perturbed by 1e-7
    # (one fp32 ulp of a [0,1] pixel): ReLU / max-pool decisions that flip put a floor under any fp32 comparison
    m.load_state_dict(sd0)
    m.zero_grad()
    torch.manual_seed(5)
    _, out_p, _, grads_p, _ = run_step(m, x + 1e-7 * torch.randn_like(x), y)
    arrs["sens/logits"] = np.array(_relerr(out_p, out))
    for k in grads:
        arrs["sens/grad/" + k] = np.array(_relerr(grads_p[k], grads[k]))
    save("cfg1_smallcnn.npz", **arrs)


def gold_cfg2():
    """BASELINE.json configs[1]: medsos LRCN, frozen ResNet-50 (train-mode BN), 16 frames x 112x112; an 8-clip slice
    (128 frames per BatchNorm batch) and the bench's exact 64-clip batch (medsos_lrcn/src/models.py:121-234)."""
    for tag, B, cond in (("b8", 8, False), ("b64", 64, False), ("b8_cond", 8, True)):
        mm = refload.medsos_models(CONF_RNN_LAYER=3, CONF_RNN_OUT="all", CONF_CLASSIF_MODE="multiclass", CONF_DROPOUT=0.0)
        T, S, seed = 16, 112, 7
        torch.manual_seed(seed)
        m = mm.LRCN(4, T, 32, 8, cnn_backbone="resnet50", rnn_type="lstm", rnn_out="all", bidirectional=False)
        if cond:
            condition_backbone(m.cnn_backbone)
        x, y = _clips(B, T, S, 4)
        sd0, out, loss, grads, sd1 = run_step(m, x, y)
        arrs = {"logits": out.numpy(), "loss": loss.numpy(),
                "backbone_checksum": np.array(checksum(sd0, "cnn_backbone.")),
                "tail_checksum": np.array(checksum({k: v for k, v in sd0.items() if not k.startswith("cnn_backbone.")}, "")),
                "meta": np.array(json.dumps(dict(arch="resnet50", size=S, B=B, T=T, hidden=32, rnn_input=8, rnn_layers=3,
                                                 num_classes=4, seed=seed, clip_seed=1234, conditioned=cond,
                                                 source="medsos_lrcn/src/models.py:121-234, dropout 0; BASELINE.json configs[1]")))}
        _store_grads(arrs, grads)
        for k in ("cnn_backbone.bn1.running_mean", "cnn_backbone.bn1.running_var", "cnn_backbone.layer1.0.bn3.running_var",
                  "cnn_backbone.layer3.5.bn3.running_mean", "cnn_backbone.layer4.2.bn3.running_var",
                  "cnn_backbone.layer2.0.downsample.1.running_mean"):
            arrs["sd1/" + k] = sd1[k].numpy()
        m.load_state_dict(sd0)
        m.train()
        with torch.no_grad():
            feat = m.cnn_backbone(x.view(B * T, 3, S, S))
        arrs["features_sub8"] = feat[:, ::8].numpy().copy()            # [B*T, 256] of the 2048 pooled features
        arrs["features_absmax"] = np.array(float(feat.abs().max()))
        m.load_state_dict(sd0)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            arrs["yard/features"] = np.array(_relerr(m.cnn_backbone(x.view(B * T, 3, S, S)).float(), feat))
        m.load_state_dict(sd0)
        arrs.update(_autocast_yard(m, x, y, out, grads))
        save(f"cfg2_medsos_{tag}.npz", **arrs)


def gold_cfg3():
    """BASELINE.json configs[2]: frozen ResNet-50 at 224x224 x 16 frames + 4-layer biLSTM H=56 (lrcn/ucf50-lrcn.py:252-336
    topology: the frozen-encoder form of rgb_lrcn.py), batch 2."""
    C, g0 = refload.ucf50_lrcn(CONF_CNN_BACKBONE="resnet50", CONF_RNN_LAYER=4)
    T, S, B, seed = 16, 224, 2, 23
    torch.manual_seed(seed)
    m = C(5, T, 56, 64, cnn_backbone="resnet50")
    x, y = _clips(B, T, S, 5)
    sd0, out, loss, grads, sd1 = run_step(m, x, y)
    arrs = {"logits": out.numpy(), "loss": loss.numpy(), "backbone_checksum": np.array(checksum(sd0, "cnn_backbone.")),
            "tail_checksum": np.array(checksum({k: v for k, v in sd0.items() if not k.startswith("cnn_backbone.")}, "")),
            "meta": np.array(json.dumps(dict(arch="resnet50", size=S, B=B, T=T, hidden=56, rnn_input=64, rnn_layers=4,
                                             num_classes=5, seed=seed, clip_seed=1234,
                                             source="lrcn/ucf50-lrcn.py:252-336; BASELINE.json configs[2]")))}
    _store_grads(arrs, grads)
    m.load_state_dict(sd0)
    m.train()
    with torch.no_grad():
        feat = m.cnn_backbone(x.view(B * T, 3, S, S))
    arrs["features_sub8"] = feat[:, ::8].numpy().copy()
    arrs["features_absmax"] = np.array(float(feat.abs().max()))
    m.load_state_dict(sd0)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        arrs["yard/features"] = np.array(_relerr(m.cnn_backbone(x.view(B * T, 3, S, S)).float(), feat))
    m.load_state_dict(sd0)
    arrs.update(_autocast_yard(m, x, y, out, grads))
    save("cfg3_ucf50_224.npz", **arrs)


def gold_crime_trainable():
    """crime LRCN with the WHOLE backbone trainable (lrcn/lrcn.py:181-305 with CONF_FINETUNE=True: freeze_cnn_layers
    un-freezes the Identity head and freezes nothing, lrcn.py:246-258), ResNet-18, 8 clips x 8 frames x 64x64:
    reference gradients of backbone parameters + torch's own bf16-autocast error on each."""
    for cond in (False, True):
        _gold_crime_trainable(cond)


def _gold_crime_trainable(cond):
    C, g0 = refload.crime_lrcn(CONF_CNN_BACKBONE="resnet18", CONF_RNN_LAYER=2, CONF_CLASSIF_MODE="multiple_binary",
                               CONF_FINETUNE=True)
    T, S, B, seed = 8, 64, 8, 29
    torch.manual_seed(seed)
    m = C(3, T, 12, 16, cnn_backbone="resnet18")
    if cond:
        condition_backbone(m.cnn_backbone)
    assert all(p.requires_grad for p in m.cnn_backbone.parameters())
    x, _ = _clips(B, T, S, 3)
    gy = torch.Generator().manual_seed(99)
    yb = (torch.rand(B, 3, generator=gy) > 0.5).float()
    sd0, out, loss, grads, sd1 = run_step(m, x, yb, loss_kind="bce")
    arrs = {"y": yb.numpy(), "logits": out.numpy(), "loss": loss.numpy(), "state_checksum": np.array(checksum(sd0, "")),
            "meta": np.array(json.dumps(dict(arch="resnet18", size=S, B=B, T=T, hidden=12, rnn_input=16, rnn_layers=2,
                                             num_classes=3, seed=seed, clip_seed=1234, finetune=True, conditioned=cond,
                                             source="lrcn/lrcn.py:181-305 multiple_binary, CONF_FINETUNE=True; loss=mean BCEWithLogits")))}
    _store_grads(arrs, grads)
    m.load_state_dict(sd0)
    arrs.update(_autocast_yard(m, x, yb, out, grads, loss_kind="bce"))
    save("crime_trainable_resnet18_cond.npz" if cond else "crime_trainable_resnet18.npz", **arrs)


def gold_variants():
    """Temporal-layer / adapt-stack variants of the backbone LRCNs (VERDICT r01 missing #4, #5): one train step of each
    reference class on a frozen ResNet-18, 2 clips x 4 frames x 32x32.  The fixture keeps the reference's pooled backbone
    features so the GPU test can check the TRAINABLE TAIL in fp32 (1e-4 / 2e-3) independent of the bf16 encoder."""
    B, T, S = 2, 4, 32
    cases = []
    C, _ = refload.ucf50_lrcn_full(CONF_CNN_BACKBONE="resnet18", CONF_RNN_LAYER=2)
    cases.append(("ucf50_gru", lambda: C(5, T, 8, 16, cnn_backbone="resnet18", rnn_type="gru"), "ce", 5,
                  dict(cls="UCF50LRCN", kw=dict(num_classes=5, sequence_length=T, hidden_size=8, rnn_input_size=16, cnn_backbone="resnet18",
                                                rnn_type="gru", rnn_layers=2)), "lrcn/ucf50-lrcn.py:252-336 rnn_type=gru"))
    cases.append(("ucf50_mamba", lambda: C(5, T, 8, 16, cnn_backbone="resnet18", rnn_type="mamba"), "ce", 5,
                  dict(cls="UCF50LRCN", kw=dict(num_classes=5, sequence_length=T, hidden_size=8, rnn_input_size=16, cnn_backbone="resnet18",
                                                rnn_type="mamba", rnn_layers=2)), "lrcn/ucf50-lrcn.py:123-336 rnn_type=mamba"))
    for rt in ("gru", "lstm"):
        D, _ = refload.dump_lrcn(CONF_CNN_BACKBONE="resnet18", CONF_RNN_LAYER=2, CONF_CLASSIF_MODE="multiple_binary", CONF_RNN_TYPE=rt)
        cases.append((f"dump_{rt}", (lambda D=D, rt=rt: D(3, T, 8, 16, cnn_backbone="resnet18", rnn_type=rt)), "bce", 3,
                      dict(cls="CrimeLRCN", kw=dict(num_classes=3, sequence_length=T, hidden_size=8, rnn_input_size=16, cnn_backbone="resnet18",
                                                    rnn_layers=2, classif_mode="multiple_binary", rnn_type=rt, rnn_attr="rnn",
                                                    finetune=True)), f"lrcn/dump_lrcn.py:278-339 rnn_type={rt}"))
    for tag, mode, rt, bidir in (("adapt_lstm", "lnslnslnsd", "lstm", False), ("adapt_gru_bi", "lgnlrnlsd", "gru", True),
                                 ("adapt_mamba", "lsnlsnlsn", "mamba", False)):
        mb = refload.medsos_models_bidir(CONF_ADAPT=mode, CONF_RNN_LAYER=2, CONF_DROPOUT=0.0, CONF_CLASSIF_MODE="multiclass",
                                         CONF_RNN_OUT="all")
        # (all_config is one shared module: CONF_ADAPT is read when the model is CONSTRUCTED, so pin it again right there)
        cases.append((tag, (lambda mb=mb, rt=rt, bidir=bidir, mode=mode: (setattr(mb.all_config, "CONF_ADAPT", mode),
                                                                          mb.LRCN(4, T, 16, 8, cnn_backbone="resnet18", rnn_type=rt,
                                                                                  rnn_out="all", bidirectional=bidir))[1]), "ce", 4,
                      dict(cls="AdaptLRCN", kw=dict(num_classes=4, sequence_length=T, hidden_size=16, rnn_input_size=8, cnn_backbone="resnet18",
                                                    rnn_type=rt, bidirectional=bidir, rnn_layers=2, dropout=0.0, adapt_mode=mode)),
                      f"medsos_lrcn/src/models_bidir.py:119-248 CONF_ADAPT={mode} rnn_type={rt}"))
    for i, (tag, make, loss_kind, ncls, build, source) in enumerate(cases):
        torch.manual_seed(300 + i)
        m = make()
        for p in m.cnn_backbone.parameters():      # dump_lrcn.py never freezes: the tail test does not need backbone gradients
            p.requires_grad = False
        x, y = _clips(B, T, S, ncls, seed=77 + i)
        if loss_kind == "bce":
            y = (torch.rand(B, ncls, generator=torch.Generator().manual_seed(5 + i)) > 0.5).float()
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        m.train()
        with torch.no_grad():
            feat = m.cnn_backbone(x.view(B * T, 3, S, S))
        m.load_state_dict(sd0)
        _, out, loss, grads, _ = run_step(m, x, y, loss_kind=loss_kind)
        arrs = {"x": x.numpy(), "y": y.numpy(), "features": feat.numpy(), "logits": out.numpy(), "loss": loss.numpy(),
                "meta": np.array(json.dumps(dict(build=build, loss=loss_kind, source=source)))}
        for k, v in npd(sd0).items():
            if not k.startswith("cnn_backbone."):
                arrs["sd0/" + k] = v
        for k, v in npd(grads).items():
            arrs["grad/" + k] = v
        save(f"variant_{tag}.npz", **arrs)


def gold_baseline_shapes():
    gold_cfg1()
    gold_cfg2()
    gold_cfg3()
    gold_crime_trainable()


if __name__ == "__main__":
    assert refload.available(), "needs /root/reference"
    if len(sys.argv) > 1:                      # python make_golden.py gold_cfg1 gold_cfg2 ...
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    gold_sampling()
    gold_resize()
    gold_smallcnn()
    gold_lstm()
    gold_medsos()
    gold_simple()
    gold_scan()
    gold_gru()
    gold_mamba()
    gold_ckpt()
    gold_crime_densenet()
    gold_medsos_mobilenet()
