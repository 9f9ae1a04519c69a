from e3dge_b200.stylesdf_model import *  # noqa: F401,F403
from e3dge_b200.stylesdf_model import (Blur, Decoder, Downsample, EqualLinear,  # noqa: F401
                                       G_pred_latents, Generator, MappingLinear,
                                       ModulatedConv2d, NoiseInjection, PixelNorm, StyledConv,
                                       ToRGB, Upsample, make_kernel)
