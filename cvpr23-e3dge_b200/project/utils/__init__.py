from . import volume_renderer  # noqa: F401
