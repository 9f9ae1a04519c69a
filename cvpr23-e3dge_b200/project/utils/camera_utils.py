from e3dge_b200.frontend import generate_camera_params  # noqa: F401
