from e3dge_b200.frontend import HybridGradualStyleEncoder_V2  # noqa: F401
