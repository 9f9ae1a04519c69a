from e3dge_b200.stylesdf_model import *  # noqa: F401,F403
from e3dge_b200.stylesdf_model import (Blur, Decoder, Downsample, EqualLinear,  # noqa: F401
                                       G_pred_latents, Generator, MappingLinear,
                                       ModulatedConv2d, NoiseInjection, PixelNorm, StyledConv,
                                       ToRGB, Upsample, make_kernel)
from e3dge_b200.frontend import (AddCoords, CoordConv2d, CoordConvLayer, VolumeRenderDiscConv2d,  # noqa: F401,E402
                                 VolumeRenderDiscriminator, VolumeRenderResBlock)
