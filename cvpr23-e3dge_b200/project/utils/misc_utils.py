from e3dge_b200.local_branch import PosEncoding  # noqa: F401
