from . import camera_utils, mesh_utils, misc_utils, volume_renderer  # noqa: F401
from .mesh_utils import align_volume  # noqa: F401  (project/utils/__init__.py re-exports it, volume_renderer.py:13)
