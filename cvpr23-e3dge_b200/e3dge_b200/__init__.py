"""e3dge_b200 — B200-native StyleSDF generator hot path for E3DGE."""
from . import _lib  # noqa: F401
from ._lib import invalidate_packed  # noqa: F401
from .options import Opt, model_options, rendering_options  # noqa: F401

__all__ = ["Opt", "model_options", "rendering_options", "invalidate_packed", "accumulate"]


def accumulate(model1, model2, decay=0.999):
    """EMA of generator weights, the reference's `accumulate` (project/utils/training_utils.py:40-45):
    the same in-place `.data` update, followed by the packed-weight invalidation it requires here."""
    par1, par2 = dict(model1.named_parameters()), dict(model2.named_parameters())
    for k in par1.keys():
        par1[k].data.mul_(decay).add_(par2[k].data, alpha=1 - decay)
    invalidate_packed()
