from .fpn_encoders import HybridGradualStyleEncoder_V2  # noqa: F401
