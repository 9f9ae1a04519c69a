from . import camera_utils, misc_utils, volume_renderer  # noqa: F401
