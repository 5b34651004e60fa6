"""Loading the reference's pickled whole-module checkpoints into the B200 modules.

The reference saves the best model with `torch.save(model, path)` (medsos_lrcn/src/train_eval.py:53,
lrcn/ucf50-lrcn.py:468) and serves it with `torch.load(path)` (medsos_lrcn/src/deployment.py:165, worker.py:114).
Such a file pickles the module OBJECT: it names the class by import path (`models.LRCN` when trained through
main.py, `__main__.LRCN` when the defining script was run directly) and carries the instance `__dict__`
(`_parameters`, `_modules`, plain attributes).  `torch.load` therefore needs the reference's source tree on the
path and yields a module that runs on torch's kernels.

`load_reference_checkpoint(path)` unpickles the same file WITHOUT the reference's code: classes that cannot be
imported (or that carry the reference's class names) are materialised as inert `nn.Module` placeholders, the
topology and hyper-parameters are read back from the placeholder's attributes and tensor shapes, the matching
video_classif_b200 module is constructed and the weights go in through `load_state_dict` (same keys, same shapes).
`state_dict` checkpoints (lrcn/lrcn.py:347, rgb_lrcn.py:302) need none of this: `model.load_state_dict(torch.load(p))`.

Host-side only; nothing here touches the GPU.  Trust model: the unpickler resolves ONLY an allowlist of globals --
torch tensor / storage rebuild helpers, classes under torch.nn / torchvision.models (module objects), OrderedDict, numpy
array reconstruction and a handful of inert builtins (set, frozenset, slice, range, complex, int, float, bool, ...).
Every other global (`os.system`, `builtins.eval`, `subprocess.Popen`, the reference's own classes, ...) is NEVER imported
or called: it becomes an inert `ReferencePlaceholder` class, so a REDUCE on it only constructs an empty nn.Module.
That is stricter than the reference's own `torch.load(path)`; it is still a pickle -- load checkpoints you trust.
"""
from __future__ import annotations

import math
import pickle
import re
import types

import torch
import torch.nn as nn

REFERENCE_CLASS_NAMES = ("LRCN", "LRCN2", "ResidualBlock", "ParallelMamba", "RMSNorm")
# globals a torch module pickle legitimately needs (checked by (module, name); prefixes end with a dot)
_ALLOWED_EXACT = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"), ("_codecs", "encode"), ("copyreg", "_reconstructor"),
    ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "slice"), ("builtins", "range"), ("builtins", "complex"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"),
    ("builtins", "bytearray"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "dict"), ("builtins", "object"),
    ("torch", "Size"), ("torch", "device"), ("torch", "dtype"), ("torch", "Tensor"), ("torch", "layout"),
    ("torch", "memory_format"), ("torch.serialization", "_get_layout"), ("torch._tensor", "_rebuild_from_type_v2"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
}
_ALLOWED_PREFIX = ("torch._utils._rebuild_", "torch.nn.", "torch.storage.", "torchvision.models.", "torchvision.ops.")
_ALLOWED_TORCH_ATTRS = re.compile(r"^(Float|Double|Half|BFloat16|Long|Int|Short|Char|Byte|Bool)Storage$")


def _allowed_global(mod_name, name):
    if mod_name == "__builtin__":          # protocol-2 spelling of `builtins` (pickle's fix_imports maps it on lookup)
        mod_name = "builtins"
    if (mod_name, name) in _ALLOWED_EXACT:
        return True
    if mod_name == "torch" and _ALLOWED_TORCH_ATTRS.match(name):
        return True
    full = mod_name + "." + name
    return any(full.startswith(p) for p in _ALLOWED_PREFIX)


class ReferencePlaceholder(nn.Module):
    """Stands in for a reference class while unpickling: keeps `__dict__` (parameters, sub-modules, attributes)."""
    _ref_module = "?"
    _ref_name = "?"

    def forward(self, *a, **k):
        raise RuntimeError(f"{self._ref_module}.{self._ref_name} placeholder: convert it with convert_reference_module()")


class _Unpickler(pickle.Unpickler):
    _made: dict = {}

    def find_class(self, mod_name, name):
        if _allowed_global(mod_name, name):
            obj = super().find_class(mod_name, name)
            # torch.nn.* / torchvision.* entries must be classes (module types, parameter containers), never functions
            if mod_name.startswith(("torch.nn", "torchvision")) and not isinstance(obj, type) \
                    and not name.startswith("_rebuild_"):
                raise pickle.UnpicklingError(f"refusing non-class global {mod_name}.{name}")
            return obj
        key = (mod_name, name)
        if key not in _Unpickler._made:
            _Unpickler._made[key] = type(name, (ReferencePlaceholder,), {"_ref_module": mod_name, "_ref_name": name})
        return _Unpickler._made[key]


def _pickle_module():
    m = types.ModuleType("b2_reference_pickle")
    m.__dict__.update({k: v for k, v in pickle.__dict__.items() if not k.startswith("__")})
    m.Unpickler = _Unpickler
    m.load = lambda f, **kw: _Unpickler(f, **kw).load()
    return m


def _count_layers(sd, prefix):
    ks = [int(m.group(1)) for k in sd for m in [re.match(re.escape(prefix) + r"weight_ih_l(\d+)$", k)] if m]
    return max(ks) + 1 if ks else 0


def infer_backbone(sd, prefix="cnn_backbone."):
    """torchvision backbone name (ResNet / DenseNet / MobileNetV2) from the block structure of the state_dict keys."""
    if (prefix + "features.denseblock1.denselayer1.conv1.weight") in sd:
        depth = [len({k.split(".")[len(prefix.split(".")) + 1] for k in sd if k.startswith(f"{prefix}features.denseblock{b}.")})
                 for b in range(1, 5)]
        growth = sd[prefix + "features.denseblock1.denselayer1.conv2.weight"].shape[0]
        table = {(32, (6, 12, 24, 16)): "densenet121", (32, (6, 12, 32, 32)): "densenet169", (32, (6, 12, 48, 32)): "densenet201"}
        if (growth, tuple(depth)) not in table:
            raise NotImplementedError(f"unrecognised DenseNet layout growth={growth} blocks={depth}")
        return table[(growth, tuple(depth))]
    if (prefix + "features.18.0.weight") in sd and (prefix + "features.1.conv.0.0.weight") in sd:
        return "mobilenet_v2"
    blocks = []
    for layer in range(1, 5):
        ids = {int(m.group(1)) for k in sd for m in [re.match(re.escape(prefix) + rf"layer{layer}\.(\d+)\.", k)] if m}
        blocks.append(len(ids))
    if not all(blocks):
        raise NotImplementedError("checkpoint backbone is not a torchvision ResNet (only ResNet-class backbones are built)")
    bottleneck = (prefix + "layer1.0.conv3.weight") in sd
    table = {(False, (2, 2, 2, 2)): "resnet18", (False, (3, 4, 6, 3)): "resnet34", (True, (3, 4, 6, 3)): "resnet50",
             (True, (3, 4, 23, 3)): "resnet101", (True, (3, 8, 36, 3)): "resnet152"}
    key = (bottleneck, tuple(blocks))
    if key not in table:
        raise NotImplementedError(f"unrecognised ResNet layout {key}")
    return table[key]


def _heads(sd):
    """(num_classes, classif_mode, fc_in) from `fc.weight` / `fc.{i}.weight` / `fcb.weight`."""
    per_class = [k for k in sd if re.match(r"fc\.\d+\.weight$", k)]
    if per_class:
        return len(per_class), "multiple_binary", sd["fc.0.weight"].shape[1]
    if "fcb.weight" in sd:
        return sd["fcb.weight"].shape[0], "multiclass", sd["fc.weight"].shape[1]
    return sd["fc.weight"].shape[0], "multiclass", sd["fc.weight"].shape[1]


def convert_reference_module(ref, precision="bf16", **overrides):
    """Build the video_classif_b200 module equivalent to `ref` (a reference module or its unpickled placeholder) and
    load its weights.  `overrides` replace inferred constructor arguments (e.g. input_shape for a non-square small CNN)."""
    from . import models as M
    sd = ref.state_dict()
    attr = lambda k, d=None: getattr(ref, k, d)

    if "conv1.weight" in sd and "lstm.weight_ih_l0" in sd:                                 # small frame CNN (nb:148-193)
        H = sd["lstm.weight_hh_l0"].shape[1]
        gates = sd["lstm.weight_ih_l0"].shape[0] // H
        dirs = 2 if "lstm.weight_ih_l0_reverse" in sd else 1
        ncls, _, fc_in = _heads(sd)
        T = fc_in // (H * dirs)
        side = int(round(math.sqrt(sd["lstm.weight_ih_l0"].shape[1] / 64))) * 4
        kw = dict(num_classes=ncls, sequence_length=T, hidden_size=H, input_shape=(sd["conv1.weight"].shape[1], side, side),
                  dropout=float(getattr(attr("dropout"), "p", 0.5)), precision=precision)
        if gates == 3:                                                                     # LRCN2: biGRU stored as `lstm`
            kw.update(overrides)
            model = M.SmallCNNGRU(**kw)
        else:
            kw["lstm_layers"] = _count_layers(sd, "lstm.")
            kw.update(overrides)
            model = M.SmallCNNLRCN(**kw)
    elif "adapt1.weight" in sd and "bn1.weight" in sd:                                     # medsos models.py:121-234
        rnn_type = attr("rnn_type", "lstm")
        bidir = bool(attr("bidirectional", "rnn.weight_ih_l0_reverse" in sd))
        H = attr("hidden_size")
        rnn_in = sd["adapt3.weight"].shape[0]
        ncls, mode, fc_in = _heads(sd)
        if rnn_type == "mamba":
            layers = len({k.split(".")[1] for k in sd if k.startswith("rnn.")})
            out_w = rnn_in
            H = H if H is not None else sd["rnn.0.mixer.A_log"].shape[1]
        else:
            layers = _count_layers(sd, "rnn.")
            H = H if H is not None else sd["rnn.weight_hh_l0"].shape[1]
            out_w = H * (2 if bidir else 1)
        T = attr("sequence_length") or fc_in // out_w
        kw = dict(num_classes=ncls, sequence_length=T, hidden_size=H, rnn_input_size=rnn_in,
                  cnn_backbone=infer_backbone(sd), rnn_type=rnn_type, rnn_out="all" if fc_in == out_w * T and T > 1 else "last",
                  bidirectional=bidir, rnn_layers=layers, dropout=float(getattr(attr("drop1"), "p", 0.25)),
                  classif_mode=mode, precision=precision)
        kw.update(overrides)
        model = M.LRCN(**kw)
    elif "adapt1.weight" in sd:                                                            # lrcn/ucf50-lrcn.py:252-336
        ncls, mode, fc_in = _heads(sd)
        if "rnn.0.mixer.A_log" in sd:                                                      # Mamba blocks (:285-289)
            rnn_type, H = "mamba", sd["rnn.0.mixer.A_log"].shape[1]
            layers = len({k.split(".")[1] for k in sd if k.startswith("rnn.")})
        else:
            H = sd["rnn.weight_hh_l0"].shape[1]
            rnn_type = "gru" if sd["rnn.weight_ih_l0"].shape[0] == 3 * H else "lstm"
            layers = _count_layers(sd, "rnn.")
        T = attr("sequence_length") or fc_in // (2 * H)
        kw = dict(num_classes=ncls, sequence_length=T, hidden_size=H, rnn_input_size=sd["adapt3.weight"].shape[0],
                  cnn_backbone=infer_backbone(sd), rnn_type=rnn_type,
                  rnn_out="all" if fc_in == 2 * H * T and T > 1 else "last", rnn_layers=layers,
                  classif_mode=mode, precision=precision)
        kw.update(overrides)
        model = M.UCF50LRCN(**kw)
    elif "adapt.weight" in sd and ("lstm.weight_ih_l0" in sd or "rnn.weight_ih_l0" in sd):
        # lrcn/lrcn.py:181-305, rgb_lrcn.py (`lstm`); lrcn/dump_lrcn.py:278-339 (`rnn`, lstm / gru switch)
        rnn_attr = "lstm" if "lstm.weight_ih_l0" in sd else "rnn"
        H = sd[rnn_attr + ".weight_hh_l0"].shape[1]
        ncls, mode, fc_in = _heads(sd)
        T = attr("sequence_length") or fc_in // (2 * H)
        kw = dict(num_classes=ncls, sequence_length=T, hidden_size=H, rnn_input_size=sd["adapt.weight"].shape[0],
                  cnn_backbone=infer_backbone(sd), rnn_out="all" if fc_in == 2 * H * T and T > 1 else "last",
                  rnn_layers=_count_layers(sd, rnn_attr + "."), classif_mode=mode,
                  finetune=any(p.requires_grad for n, p in ref.named_parameters() if n.startswith("cnn_backbone.")),
                  rnn_type="gru" if sd[rnn_attr + ".weight_ih_l0"].shape[0] == 3 * H else "lstm", rnn_attr=rnn_attr,
                  precision=precision)
        kw.update(overrides)
        model = M.CrimeLRCN(**kw)
    elif any(k.startswith("adapt.adapt.") for k in sd):                                    # models_bidir.py:158-248
        seq = getattr(getattr(ref, "adapt", None), "adapt", None)
        if seq is None and "adapt_mode" not in overrides:
            raise NotImplementedError("a models_bidir state_dict does not record its activation layers: pass adapt_mode=")
        mode_str = overrides.pop("adapt_mode", None) or "".join(
            {"Linear": "l", "LayerNorm": "n", "SiLU": "s", "GELU": "g", "ReLU": "r", "Dropout": "d"}[type(m).__name__] for m in seq)
        lin = sorted((int(k.split(".")[2]), v) for k, v in sd.items() if k.startswith("adapt.adapt.") and k.endswith(".weight")
                     and v.dim() == 2)
        rnn_in = lin[-1][1].shape[0]
        ncls, mode, fc_in = _heads(sd)
        if "rnn.0.mixer.A_log" in sd:
            rnn_type, H = "mamba", sd["rnn.0.mixer.A_log"].shape[1]
            layers = len({k.split(".")[1] for k in sd if k.startswith("rnn.")})
            bidir = sd["rnn.0.mixer.out_proj.weight"].shape[1] == 2 * sd["rnn.0.mixer.A_log"].shape[0]
            out_w = rnn_in
        else:
            H = sd["rnn.weight_hh_l0"].shape[1]
            rnn_type = "gru" if sd["rnn.weight_ih_l0"].shape[0] == 3 * H else "lstm"
            layers = _count_layers(sd, "rnn.")
            bidir = "rnn.weight_ih_l0_reverse" in sd
            out_w = H * (2 if bidir else 1)
        T = attr("sequence_length") or fc_in // out_w
        drops = [m.p for m in (seq or []) if type(m).__name__ == "Dropout"]
        kw = dict(num_classes=ncls, sequence_length=T, hidden_size=H, rnn_input_size=rnn_in, cnn_backbone=infer_backbone(sd),
                  rnn_type=rnn_type, rnn_out="all" if fc_in == out_w * T and T > 1 else "last", bidirectional=bidir,
                  rnn_layers=layers, dropout=float(drops[0]) if drops else 0.25, classif_mode=mode, adapt_mode=mode_str,
                  adapt_depth=len(lin), precision=precision)
        kw.update(overrides)
        model = M.AdaptLRCN(**kw)
    else:
        raise NotImplementedError("unrecognised reference checkpoint layout: " + ", ".join(sorted(sd)[:8]) + " ...")

    missing, unexpected = model.load_state_dict(sd, strict=False)
    # `mixer.D` is created but never used by the reference block; everything else must match key for key
    bad = [k for k in list(missing) + list(unexpected) if not k.endswith("mixer.D")]
    if bad:
        raise RuntimeError(f"reference checkpoint does not match the rebuilt {type(model).__name__}: {bad[:6]}")
    flags = {n: p.requires_grad for n, p in ref.named_parameters()}
    for n, p in model.named_parameters():
        if n in flags:
            p.requires_grad_(flags[n])
    model.train(bool(getattr(ref, "training", False)))
    return model


def load_reference_checkpoint(path, precision="bf16", map_location="cpu", **overrides):
    """`torch.load(path)` replacement for the reference's `torch.save(model)` files (train_eval.py:53; consumers
    deployment.py:165, worker.py:114): returns the equivalent video_classif_b200 module with the checkpoint's weights,
    on `map_location` (move it with `.to("cuda")` like the reference does).  A file holding a plain state_dict is
    returned as the dict."""
    obj = torch.load(path, map_location=map_location, pickle_module=_pickle_module(), weights_only=False)
    if isinstance(obj, dict):
        return obj
    if not isinstance(obj, nn.Module):
        raise TypeError(f"{path}: expected a pickled nn.Module or a state_dict, found {type(obj).__name__}")
    return convert_reference_module(obj, precision=precision, **overrides)
