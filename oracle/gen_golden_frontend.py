#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — records tests/golden/frontend.npz from the REAL reference: its
`HybridGradualStyleEncoder_V2(50, 'ir_se', -1, opts)` (project/models/encoders/fpn_encoders.py:266-431), its pose net
`VolumeRenderDiscriminator` (project/models/stylesdf_model.py:1369-1419; FusedLeakyReLU takes its pure-PyTorch CPU
branch) and `generate_camera_params(..., locations=..., return_calibs=True)` (project/utils/camera_utils.py:8-151),
filled with the deterministic synthetic weights of synthetic_inputs.py, eval mode, fp32, CPU.
Run from the repository root:   python oracle/gen_golden_frontend.py"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import ref_harness as H  # noqa: E402
import synthetic_inputs as P  # noqa: E402

SEED = 51


def fill(module, prefix):
    sd = {k: torch.from_numpy(np.ascontiguousarray(P.make_param(SEED, prefix + k, v.shape))).to(v.dtype)
          for k, v in module.state_dict().items()}
    module.load_state_dict(sd, strict=True)


def inputs():
    g = np.random.Generator(np.random.PCG64(SEED))
    return torch.from_numpy(g.uniform(-1, 1, (2, 3, 256, 256)).astype(np.float32))


def main():
    for n in ("munch", "omegaconf", "omegaconf.dictconfig", "IPython", "IPython.display"):
        if n not in sys.modules:
            H._stub_module(n)
    ref = H.load_reference()
    H._shell_package("project.models.helper_modules", os.path.join(H.REFERENCE_ROOT, "project", "models", "helper_modules"))
    H._shell_package("project.models.encoders", os.path.join(H.REFERENCE_ROOT, "project", "models", "encoders"))
    # pytorch3d is stubbed: camera_utils only needs it for create_cameras, not for generate_camera_params
    from project.models.encoders.fpn_encoders import HybridGradualStyleEncoder_V2
    from project.utils.camera_utils import generate_camera_params
    opts = H.Opt(input_nc=3, fpn_pigan_geo_layer_dim=32, fpn_pigan_tex_layer_dim=32, full_pipeline=True,
                 disable_decoder_fpn=False, single_decoder_layer=True)
    enc = HybridGradualStyleEncoder_V2(50, "ir_se", -1, opts).eval()
    fill(enc, "encoder.")
    pose = ref.stylesdf_model.VolumeRenderDiscriminator(H.model_opt(renderer_spatial_output_dim=64)).eval()
    fill(pose, "volume_discriminator.")
    x = inputs()
    with torch.no_grad():
        thumb_lat, dec_lat = enc(x)
        thumb = torch.nn.functional.adaptive_avg_pool2d(x, (64, 64))
        gan, loc = pose(thumb)
        cams = generate_camera_params(64, torch.device("cpu"), 2, locations=loc, return_calibs=True)
    rec = {"enc.keys": np.array(sorted(enc.state_dict().keys())), "pose.keys": np.array(sorted(pose.state_dict().keys())),
           "thumb_latents": thumb_lat.numpy(), "decoder_latents": dec_lat.numpy(), "gan": gan.numpy(),
           "locations": loc.numpy()}
    for k in ("poses", "extrinsics", "focal", "near", "far", "viewpoint", "intrinsics", "calibs"):
        rec["cam." + k] = cams[k].numpy()
    path = os.path.join(ROOT, "tests", "golden", "frontend.npz")
    np.savez_compressed(path, **rec)
    print("thumb latents", tuple(thumb_lat.shape), float(thumb_lat.std()), "decoder latents", tuple(dec_lat.shape),
          float(dec_lat.std()), "locations", loc.tolist())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
