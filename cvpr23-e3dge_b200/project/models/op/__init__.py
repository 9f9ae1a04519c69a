from e3dge_b200.op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d  # noqa: F401
