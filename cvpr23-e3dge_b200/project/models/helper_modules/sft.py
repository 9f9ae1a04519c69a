from e3dge_b200.local_branch import Fuse_sft_MLP, ResnetBlockFC  # noqa: F401
