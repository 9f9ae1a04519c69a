"""`-m gpu`: the front end on the device (encoder and pose net against the reference fixture) and the whole
inversion frame — encoder -> pose net -> cameras -> renderer -> decoder — eager fp32 against the oracle's generator
on the frame's own predicted latents / cameras, and the bf16 + CUDA-graph form bench.py times."""
import numpy as np
import pytest
import torch

from helpers import decoder_layout, load_golden, rel_linf, synthetic_state_dict
from oracle import stylesdf_oracle as O
from test_frontend import fill, images

pytestmark = pytest.mark.gpu


def _pipeline(amp):
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.frontend import HybridGradualStyleEncoder_V2, InversionPipeline, VolumeRenderDiscriminator
    from e3dge_b200.stylesdf_model import G_pred_latents
    size, res, seed = 256, 64, 52
    sd = synthetic_state_dict(size, res, seed, "sharp")
    G = G_pred_latents(model_options(size=size, renderer_spatial_output_dim=res), rendering_options(),
                       full_pipeline=True).eval()
    G.load_state_dict(sd, strict=True)
    enc = fill(HybridGradualStyleEncoder_V2(50, "ir_se", -1).eval(), "encoder.")
    pose = fill(VolumeRenderDiscriminator(model_options(renderer_spatial_output_dim=64)).eval(), "volume_discriminator.")
    g = np.random.Generator(np.random.PCG64(seed))
    mean = [torch.from_numpy(g.standard_normal((1, 256)).astype(np.float32)) * 0.3,
            torch.from_numpy(g.standard_normal((1, 512)).astype(np.float32)) * 0.3]
    pipe = InversionPipeline(G, enc, pose, mean_latents=mean, amp=amp).cuda().eval()
    return pipe, sd


def test_encoder_and_pose_net_on_the_device_match_the_reference_fixture(monkeypatch):
    gold, _ = load_golden("frontend")
    pipe, _ = _pipeline(amp=False)
    x = images().cuda()
    # cuDNN convolutions default to TF32 (10-bit mantissa); the fixture is the reference's fp32 CPU result
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    with torch.no_grad():
        thumb, dec = pipe.encoder(x)
        gan, loc = pipe.volume_discriminator(torch.nn.functional.adaptive_avg_pool2d(x, (64, 64)))
    assert rel_linf(thumb.cpu(), gold["thumb_latents"]) < 1e-3
    assert rel_linf(dec.cpu(), gold["decoder_latents"]) < 1e-3
    assert rel_linf(loc.cpu(), gold["locations"]) < 1e-3
    assert rel_linf(gan.cpu(), gold["gan"]) < 1e-3


def test_inversion_frame_end_to_end():
    pipe, sd = _pipeline(amp=False)
    x = images().cuda()
    with torch.no_grad():
        out = pipe(x, randomize_noise=False, return_xyz=True, return_sdf=True)
    w, wd = out["pred_latents"]
    cams = out["pred_cam_settings"]
    assert w.shape == (2, 9, 256) and wd.shape == (2, decoder_layout(256, 64), 512)
    assert out["gen_imgs"].shape == (2, 3, 256, 256) and torch.isfinite(out["gen_imgs"]).all()
    with torch.no_grad():
        ref = O.generator_forward(sd, w.cpu(), wd.cpu(), cams["poses"].cpu(), cams["focal"].cpu(), cams["near"].cpu(),
                                  cams["far"].cpu(), res=64, n_samples=24)
    for k in ("features", "gen_thumb_imgs", "sdf", "gen_imgs"):
        assert rel_linf(out[k].cpu(), ref[k]) < 1e-3, k
    # the form bench.py times: bf16 channels-last front end, whole frame in one CUDA graph
    from e3dge_b200.graphed import GraphedCall
    pipe16, _ = _pipeline(amp=True)
    static = x.clone()

    def core():
        with torch.no_grad():
            return pipe16(static, randomize_noise=False)
    call = GraphedCall(core)
    o16 = call()
    torch.cuda.synchronize()
    assert torch.isfinite(o16["gen_imgs"]).all()
    # bf16 front end: the latents move by ~1e-2 relative, the frame stays the same picture
    assert rel_linf(o16["pred_latents"][0], w) < 5e-2
    assert rel_linf(o16["gen_imgs"], out["gen_imgs"]) < 0.25
