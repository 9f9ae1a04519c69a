"""e3dge_b200 — B200-native StyleSDF generator hot path for E3DGE."""
from . import _lib  # noqa: F401
from .options import Opt, model_options, rendering_options  # noqa: F401

__all__ = ["Opt", "model_options", "rendering_options"]
