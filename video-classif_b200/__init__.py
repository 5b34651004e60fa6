"""b200-lrcn: B200-native LRCN clip-classification hot path.

Importing the package never needs a GPU; running any operator does (sm_100a, no fallback)."""
from . import _lib, sampling  # noqa: F401
from ._lib import B200LrcnError  # noqa: F401


def __getattr__(name):   # lazy: models/ops import torch + torchvision
    if name in ("SmallCNNLRCN", "SmallCNNGRU", "LRCN", "UCF50LRCN", "CrimeLRCN", "AdaptLRCN", "Adapt", "GraphedInference",
                "count_parameters"):
        from . import models
        return getattr(models, name)
    if name in ("load_reference_checkpoint", "convert_reference_module"):
        from . import checkpoint
        return getattr(checkpoint, name)
    if name == "GraphedTrainStep":
        from . import graph_step
        return graph_step.GraphedTrainStep
    if name in ("ops", "models", "ingest", "backbone", "dp", "scan", "checkpoint", "graph_step"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
