from e3dge_b200.frontend import (Bottleneck, GradualStyleBlock, SEModule, bottleneck_IR,  # noqa: F401
                                 bottleneck_IR_SE, get_blocks)
