from . import helpers, resnetfc, sft  # noqa: F401
