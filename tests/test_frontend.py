"""Front end of an inversion frame against tests/golden/frontend.npz, recorded from the
reference's own HybridGradualStyleEncoder_V2 / VolumeRenderDiscriminator / generate_camera_params
(oracle/gen_golden_frontend.py).  CPU part: parameter names, the encoder forward, the camera construction."""
import numpy as np
import pytest
import torch

import synthetic_inputs as P
from helpers import load_golden, rel_linf

SEED = 51


def fill(module, prefix):
    return P.fill_module(module, prefix, SEED)


def images():
    g = np.random.Generator(np.random.PCG64(SEED))
    return torch.from_numpy(g.uniform(-1, 1, (2, 3, 256, 256)).astype(np.float32))


def test_front_end_state_dict_names_match_the_reference():
    from e3dge_b200 import model_options
    from e3dge_b200.frontend import HybridGradualStyleEncoder_V2, VolumeRenderDiscriminator
    gold, _ = load_golden("frontend")
    enc = HybridGradualStyleEncoder_V2(50, "ir_se", -1)
    pose = VolumeRenderDiscriminator(model_options(renderer_spatial_output_dim=64))
    assert sorted(enc.state_dict().keys()) == list(gold["enc.keys"])
    assert sorted(pose.state_dict().keys()) == list(gold["pose.keys"])


def test_encoder_forward_matches_the_reference_fixture():
    from e3dge_b200.frontend import HybridGradualStyleEncoder_V2
    gold, _ = load_golden("frontend")
    torch.set_num_threads(8)
    enc = fill(HybridGradualStyleEncoder_V2(50, "ir_se", -1).eval(), "encoder.")
    with torch.no_grad():
        thumb, dec = enc(images())
    assert rel_linf(thumb, gold["thumb_latents"]) < 1e-4
    assert rel_linf(dec, gold["decoder_latents"]) < 1e-4


def test_camera_params_match_the_reference_fixture():
    from e3dge_b200.frontend import generate_camera_params
    gold, _ = load_golden("frontend")
    cams = generate_camera_params(64, torch.device("cpu"), 2, locations=torch.from_numpy(gold["locations"]),
                                  return_calibs=True)
    for k in ("poses", "extrinsics", "focal", "near", "far", "viewpoint", "intrinsics", "calibs"):
        assert rel_linf(cams[k], gold["cam." + k]) < 1e-6, k
    poses, focal, near, far, viewpoint = generate_camera_params(
        64, torch.device("cpu"), 2, locations=torch.from_numpy(gold["locations"]))
    assert torch.equal(poses, cams["poses"]) and torch.equal(focal, cams["focal"])
    # sampled cameras: on the unit sphere, looking at the origin
    p, f, n, fr, v = generate_camera_params(64, torch.device("cpu"), 16, generator=torch.Generator().manual_seed(1))
    assert torch.allclose(p[:, :, 3].norm(dim=1), torch.ones(16), atol=1e-6)
    assert torch.allclose(torch.det(p[:, :, :3]), torch.ones(16), atol=1e-5)


def test_camera_sweep_mode():
    """camera_utils.py:36-52: 8 azimuths from -range to +range per identity, one elevation per identity."""
    from e3dge_b200.frontend import generate_camera_params
    cams = generate_camera_params(64, torch.device("cpu"), 3, sweep=True, return_calibs=True,
                                  generator=torch.Generator().manual_seed(2))
    vp = cams["viewpoint"].reshape(3, 8, 2)
    assert cams["poses"].shape == (24, 3, 4) and cams["calibs"].shape == (24, 4, 4) and cams["focal"].shape == (24, 1, 1)
    assert torch.allclose(vp[:, :, 0], torch.linspace(-0.3, 0.3, 8).expand(3, 8), atol=1e-6)
    assert (vp[:, :, 1] == vp[:, :1, 1]).all() and (vp[:, 0, 1].abs() <= 0.15).all()
