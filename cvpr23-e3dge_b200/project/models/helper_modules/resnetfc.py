from e3dge_b200.local_branch import ResnetBlockFC  # noqa: F401
