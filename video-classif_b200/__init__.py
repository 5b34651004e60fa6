"""b200-lrcn: B200-native LRCN clip-classification hot path."""
from . import _lib  # noqa: F401
from ._lib import B200LrcnError  # noqa: F401

__all__ = ["B200LrcnError"]
