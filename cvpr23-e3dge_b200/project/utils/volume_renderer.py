from e3dge_b200.volume_renderer import *  # noqa: F401,F403
from e3dge_b200.volume_renderer import (FiLMSiren, LinearLayer, SirenGenerator,  # noqa: F401
                                        SirenLocalGlobal, UniformBoxWarp, VolumeFeatureRenderer)
