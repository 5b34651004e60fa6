"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference classes.

Imports the reference's own model classes / sampling functions from /root/reference
(read-only, exists only in the authoring container, never on the GPU box) so that
  * tests/golden/make_golden.py can generate golden vectors from the reference itself,
  * tests can cross-check oracle/lrcn_oracle.py against the reference when it is present.
Nothing in the product package may import this module.

Shims:
  1. matplotlib / skimage / h5py are absent -> stubbed in sys.modules before import.
  2. torchvision constructors are wrapped so `pretrained=True` does not hit the network.
  3. scripts that run dataset loading at import time are not imported; the class source is
     cut out by line range and exec'd with the CONF_* globals defined first.
"""
import importlib.util
import json
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py (GPU box)
REF_ROOT = os.environ.get("B200LRCN_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "lrcn")) and os.path.isdir(os.path.join(_STAGED, "lrcn")):
    REF_ROOT = _STAGED


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lrcn"))


def is_staged_copy() -> bool:
    """True when the classes come from oracle/_ref (the unmodified files staged by oracle/build_ref.py)."""
    return os.path.abspath(REF_ROOT) == os.path.abspath(_STAGED)


def _stub_modules():
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.metrics", "h5py",
                 "seaborn"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                m.__dict__.setdefault("structural_similarity", lambda *a, **k: 0.0)
                sys.modules[name] = m
    if "skimage" in sys.modules and "skimage.metrics" in sys.modules:
        setattr(sys.modules["skimage"], "metrics", sys.modules["skimage.metrics"])
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])


class _NoDownloadModels:
    """Proxy for `torchvision.models` that ignores pretrained=/weights= (no network)."""

    def __init__(self):
        import torchvision.models as tvm
        self._tvm = tvm

    def __getattr__(self, name):
        obj = getattr(self._tvm, name)
        if callable(obj) and name[0].islower():
            def ctor(*a, pretrained=False, weights=None, **k):
                return obj(*a, weights=None, **k)
            return ctor
        return obj


def _exec_lines(path, first, last, glob):
    with open(path) as f:
        lines = f.readlines()
    src = "".join(lines[first - 1:last])
    exec(compile(src, path, "exec"), glob)
    return glob


def _base_globals(**conf):
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from einops import rearrange
    g = dict(torch=torch, nn=nn, F=F, rearrange=rearrange, models=_NoDownloadModels())
    g.update(conf)
    return g


def notebook_lrcn():
    """class LRCN of lrcn/.ipynb_checkpoints/LRCN-ucf50-checkpoint.ipynb cell 4 (nb:148-193)."""
    _stub_modules()
    nb = json.load(open(os.path.join(REF_ROOT, "lrcn/.ipynb_checkpoints/LRCN-ucf50-checkpoint.ipynb")))
    src = "".join(nb["cells"][4]["source"])
    g = _base_globals()
    exec(compile(src, "nb-cell4", "exec"), g)
    return g["LRCN"]


def backup_lrcn2():
    """class LRCN2 of lrcn/backup_ucf50.py:105-151 (small CNN + biGRU)."""
    g = _base_globals()
    _exec_lines(os.path.join(REF_ROOT, "lrcn/backup_ucf50.py"), 105, 151, g)
    return g["LRCN2"]


def medsos_models(**conf):
    """module medsos_lrcn/src/models.py with all_config patched (models.py:121-234)."""
    _stub_modules()
    src_dir = os.path.join(REF_ROOT, "medsos_lrcn/src")
    if src_dir not in sys.path:
        sys.path.insert(0, src_dir)
    import all_config
    for k, v in conf.items():
        setattr(all_config, k, v)
    import torchvision
    spec = importlib.util.spec_from_file_location("ref_medsos_models", os.path.join(src_dir, "models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.models = _NoDownloadModels()
    mod.all_config = all_config
    return mod


def ucf50_lrcn(**conf):
    """class LRCN of lrcn/ucf50-lrcn.py:252-336 (frozen backbone, 3 plain adapts, biLSTM)."""
    base = dict(CONF_CNN_BACKBONE="resnet50", CONF_RNN_TYPE="lstm", CONF_RNN_OUT="all",
                CONF_RNN_LAYER=4, CONF_CLASSIF_MODE="multiclass")
    base.update(conf)
    g = _base_globals(**base)
    g["ResidualBlock"] = None
    _exec_lines(os.path.join(REF_ROOT, "lrcn/ucf50-lrcn.py"), 252, 336, g)
    return g["LRCN"], g


def ucf50_lrcn_full(**conf):
    """class LRCN of lrcn/ucf50-lrcn.py:252-336 together with that file's own RMSNorm / ParallelMamba / ResidualBlock
    (:123-250), so that rnn_type='mamba' constructs."""
    base = dict(CONF_CNN_BACKBONE="resnet50", CONF_RNN_TYPE="lstm", CONF_RNN_OUT="all",
                CONF_RNN_LAYER=4, CONF_CLASSIF_MODE="multiclass")
    base.update(conf)
    g = _base_globals(**base)
    _exec_lines(os.path.join(REF_ROOT, "lrcn/ucf50-lrcn.py"), 123, 336, g)
    return g["LRCN"], g


def dump_lrcn(**conf):
    """class LRCN of lrcn/dump_lrcn.py:278-339 (one adapt, temporal layer stored as `rnn`, lstm / gru switch)."""
    base = dict(CONF_CNN_BACKBONE="resnet50", CONF_RNN_TYPE="lstm", CONF_RNN_OUT="all", CONF_RNN_LAYER=4,
                CONF_CLASSIF_MODE="multiple_binary", CONF_FINETUNE=False)
    base.update(conf)
    g = _base_globals(**base)
    _exec_lines(os.path.join(REF_ROOT, "lrcn/dump_lrcn.py"), 278, 339, g)
    return g["LRCN"], g


def medsos_models_bidir(**conf):
    """module medsos_lrcn/src/models_bidir.py (string-programmed Adapt :119-155, LRCN :158-248) with all_config patched."""
    _stub_modules()
    src_dir = os.path.join(REF_ROOT, "medsos_lrcn/src")
    if src_dir not in sys.path:
        sys.path.insert(0, src_dir)
    import all_config
    for k, v in conf.items():
        setattr(all_config, k, v)
    spec = importlib.util.spec_from_file_location("ref_medsos_models_bidir", os.path.join(src_dir, "models_bidir.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.models = _NoDownloadModels()
    mod.all_config = all_config
    return mod


def crime_lrcn(**conf):
    """class LRCN of lrcn/lrcn.py:181-305 (trainable/frozen backbone, one adapt, biLSTM)."""
    base = dict(CONF_CNN_BACKBONE="densenet121", CONF_RNN_OUT="all", CONF_RNN_LAYER=4,
                CONF_CLASSIF_MODE="multiple_binary", CONF_FINETUNE=False)
    base.update(conf)
    g = _base_globals(**base)
    _exec_lines(os.path.join(REF_ROOT, "lrcn/lrcn.py"), 181, 305, g)
    return g["LRCN"], g


def rgb_lrcn(**conf):
    """class LRCN of lrcn/rgb_lrcn.py:168-263."""
    base = dict(CONF_CNN_BACKBONE="resnet50", CONF_RNN_OUT="all", CONF_RNN_LAYER=4,
                CONF_FINETUNE=True)
    base.update(conf)
    g = _base_globals(**base)
    _exec_lines(os.path.join(REF_ROOT, "lrcn/rgb_lrcn.py"), 168, 263, g)
    return g["LRCN"], g


def sampling_functions():
    """uniform_sampling / duplicate_frames of medsos_lrcn/src/loader_data.py:35-51."""
    g = {}
    _exec_lines(os.path.join(REF_ROOT, "medsos_lrcn/src/loader_data.py"), 35, 51, g)
    return g["uniform_sampling"], g["duplicate_frames"]


def videomamba_scan():
    """ParallelMamba.parallel_scan of lrcn/videomamba.py:242-284 as a free function."""
    import torch
    g = dict(torch=torch)
    with open(os.path.join(REF_ROOT, "lrcn/videomamba.py")) as f:
        lines = f.readlines()
    import textwrap
    src = textwrap.dedent("".join(lines[241:284]))
    exec(compile(src, "videomamba.py:242-284", "exec"), g)
    fn = g["parallel_scan"]
    return lambda u, delta, A, B, C: fn(None, u, delta, A, B, C)


def medsos_scan(direction="forward"):
    """ParallelMamba.parallel_scan of medsos_lrcn/src/models.py:47-71 (unchunked, bidirectional)."""
    import torch
    g = dict(torch=torch)
    with open(os.path.join(REF_ROOT, "medsos_lrcn/src/models.py")) as f:
        lines = f.readlines()
    import textwrap
    src = textwrap.dedent("".join(lines[46:71]))
    exec(compile(src, "models.py:47-71", "exec"), g)
    fn = g["parallel_scan"]
    return lambda u, delta, A, B, C: fn(None, u, delta, A, B, C, direction=direction)
