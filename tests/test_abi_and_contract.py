"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/e3dge_b200.h declares; the Python modules expose the reference's class names
and state_dict keys."""
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, PKG
from helpers import generator_state_dict_spec


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "e3dge_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(e3_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from e3dge_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.exported_symbols()) == set(declared), "ctypes prototypes drifted from the header"
    assert lib.e3_abi_version() == 4
    assert lib.e3_siren_packed_bytes() % 128 == 0


def test_bad_arguments_return_status_not_crash():
    from e3dge_b200 import _lib
    lib = _lib.load()
    # null tensors / unsupported enum: status < 0 and a message, no CUDA call is made
    rc = lib.e3_fused_bias_act(None, None, None, None, 16, 1, 0, 7, 0, 0.2, 1.0, None)
    assert rc < 0 and b"e3_fused_bias_act" in lib.e3_last_error()
    rc = lib.e3_upfirdn2d(None, None, None, 1, 4, 4, 1, 4, 4, 0, 1, 1, 1, 0, 0, 0, 0, None)
    assert rc < 0
    rc = lib.e3_film_fwd(None, None, 1, 3, None, None)
    assert rc < 0
    with pytest.raises(RuntimeError):
        _lib.check(rc, "e3_film_fwd")


def test_ops_refuse_cpu_tensors():
    from e3dge_b200.op import fused_leaky_relu, upfirdn2d
    with pytest.raises(RuntimeError):
        fused_leaky_relu(torch.zeros(2, 3))
    with pytest.raises(RuntimeError):
        upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))


@pytest.mark.parametrize("size,res,local", [(256, 64, False), (1024, 64, False), (64, 16, True)])
def test_state_dict_contract(size, res, local):
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    G = G_pred_latents(model_options(size=size, renderer_spatial_output_dim=res),
                       rendering_options(enable_local_model=local))
    got = {k: tuple(v.shape) for k, v in G.state_dict().items()}
    want = generator_state_dict_spec(size, res, local)
    assert set(got) == set(want), (sorted(set(got) ^ set(want))[:10])
    for k in want:
        assert got[k] == tuple(want[k]), (k, got[k], want[k])
    assert G.decoder.n_latent == 2 * (size.bit_length() - res.bit_length()) + 2


def test_host_side_latent_logic():
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    G = G_pred_latents(model_options(size=256), rendering_options())
    w = torch.randn(3, 9, 256)
    wbar = torch.randn(1, 9, 256)
    out = G.styles_and_noise_forward([w], truncation=0.7, truncation_latent=[wbar],
                                     input_is_latent=True)
    torch.testing.assert_close(out[0], wbar + 0.7 * (w - wbar))
    lat, noise = G.decoder.styles_and_noise_forward([torch.randn(3, 6, 512)], None,
                                                    input_is_latent=True, randomize_noise=False)
    assert lat.shape == (3, 6, 512) and len(noise) == 5 and noise[4].shape == (1, 1, 256, 256)
    lat, noise = G.decoder.styles_and_noise_forward([torch.randn(3, 512)], None,
                                                    input_is_latent=True, randomize_noise=True)
    assert lat.shape == (3, 6, 512) and noise == [None] * 5


def test_reference_import_paths_resolve_to_this_package():
    code = ("import sys; sys.path.insert(0, %r); "
            "from project.utils.volume_renderer import VolumeFeatureRenderer, SirenGenerator; "
            "from project.models.stylesdf_model import G_pred_latents, Generator, Decoder; "
            "from project.models.op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d; "
            "from project.utils.volume_renderer import SirenLocalGlobal; "
            "from project.models.helper_modules.sft import Fuse_sft_MLP; "
            "from project.models.helper_modules.resnetfc import ResnetBlockFC; "
            "from project.utils.misc_utils import PosEncoding; "
            "from project.models.encoders.fpn_encoders import HybridGradualStyleEncoder_V2; "
            "from project.models.stylesdf_model import VolumeRenderDiscriminator; "
            "from project.utils.camera_utils import generate_camera_params; "
            "import e3dge_b200.stylesdf_model as m; assert G_pred_latents is m.G_pred_latents; "
            "print('ok')" % PKG)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_packed_weight_images_are_invalidated_by_every_route_that_changes_weights():
    """in-place `.data` updates do not bump tensor versions, so the packed kernel images carry an epoch
    that load_state_dict / .to() / train() / accumulate() / invalidate_packed() advance."""
    import e3dge_b200
    from e3dge_b200 import _lib, model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    G = G_pred_latents(model_options(size=64, renderer_spatial_output_dim=16), rendering_options())
    G2 = G_pred_latents(model_options(size=64, renderer_spatial_output_dim=16), rendering_options())
    e0 = _lib.pack_epoch
    G.load_state_dict(G2.state_dict())
    e1 = _lib.pack_epoch
    G.eval()
    e2 = _lib.pack_epoch
    G.float()
    e3 = _lib.pack_epoch
    e3dge_b200.accumulate(G, G2, decay=0.5)
    e4 = _lib.pack_epoch
    e3dge_b200.invalidate_packed()
    assert e0 < e1 < e2 < e3 < e4 < _lib.pack_epoch
    # accumulate() is the reference's EMA (training_utils.py:40-45)
    for p, q in zip(G.parameters(), G2.parameters()):
        assert torch.equal(p, q)  # both started from G2's weights: the average of equals stays equal


def test_ctypes_layer_bookkeeping():
    from e3dge_b200 import _lib
    assert all(isinstance(v, int) for v in _lib.KERNELS_PER_CALL.values())
    assert set(_lib.KERNELS_PER_CALL) <= set(_lib._PROTOTYPES)

    class Fake:  # stands in for tensors of two devices
        def __init__(self, idx):
            self.device = type("D", (), {"index": idx})()
    _lib._tls.dev = None
    _lib._note_device(Fake(1))
    _lib._note_device(Fake(1))
    assert _lib._tls.dev == 1
    import pytest
    with pytest.raises(RuntimeError, match="different devices"):
        _lib._note_device(Fake(0))
    assert _lib._tls.dev is None


def test_render_host_control_flow_without_a_device(monkeypatch):
    """The Python half of VolumeFeatureRenderer.render / forward (dict contract of volume_renderer.py:1694-1729,
    1865-1972) with the kernel call stubbed out: key set, optional entries, and the surface path handing the
    frustum-aligned SDF volume to the host extractor (ImportError here: scikit-image is not installed)."""
    import importlib.util
    from e3dge_b200 import rendering_options
    from e3dge_b200.volume_renderer import VolumeFeatureRenderer
    R = VolumeFeatureRenderer(rendering_options(N_samples=6), out_im_res=8).eval()
    B, n, S = 1, 8, 6

    def fake_raw(self, styles, cam_poses, focal, near, far, z_jitter=None, local_mod=None, flags_over=None,
                 want_taps=False, film=None, train=False):
        z = lambda *s: torch.zeros(*s)
        o = dict(features=z(B, 256, n, n), gen_thumb_imgs=z(B, 3, n, n), xyz=z(B, 3, n, n), mask=z(B, 1, n, n, 1),
                 depth=z(B, n, n, 1, 1), sdf=torch.linspace(-1, 1, S).expand(B, n, n, S).unsqueeze(-1).clone(),
                 hit_prob=z(B, n, n, S, 1), visibility=z(B, n, n, S, 1), dists=z(B, n, n, S), points=z(B, n, n, S, 3),
                 rays_o=z(B, n, n, 3), rays_d=z(B, n, n, 3), viewdirs=z(B, n, n, 3), raw_rgb=z(B, n, n, S, 3))
        o["near"], o["far"] = z(B, n, n, 1), z(B, n, n, 1)
        return o
    monkeypatch.setattr(VolumeFeatureRenderer, "_render_raw", fake_raw)
    cam = torch.eye(4)[:3].unsqueeze(0)
    args = (cam, torch.full((B, 1, 1), 300.), torch.full((B, 1, 1), 0.88), torch.full((B, 1, 1), 1.12))
    with torch.no_grad():
        out = R(*args, styles=torch.zeros(B, 256))
    want = {"rays_o", "rays_d", "dists", "near", "far", "hit_prob", "surface_eikonal_term", "points", "sdf",
            "gen_thumb_imgs", "features", "mask", "xyz", "eikonal_term", "depth", "mesh", "shading_mesh", "debug_mesh",
            "viewdirs"}
    assert set(out) == want and out["mesh"] is None and out["eikonal_term"] is None
    have_skimage = importlib.util.find_spec("skimage") is not None
    with torch.no_grad():
        if have_skimage:
            m = R(*args, styles=torch.zeros(B, 256), return_mesh=True)
            assert m["mesh"] is not None and "shaded_mesh" in m
        else:
            with pytest.raises(ImportError, match="scikit-image"):
                R(*args, styles=torch.zeros(B, 256), return_mesh=True)


def test_local_branch_modules_have_no_cpu_path():
    from e3dge_b200.local_branch import Fuse_sft_MLP, ResnetBlockFC, local_tex_modulation
    with pytest.raises(RuntimeError, match="no CPU path"):
        ResnetBlockFC(301, 512)(torch.zeros(2, 301))
    with pytest.raises(RuntimeError, match="no CPU path"):
        Fuse_sft_MLP(257, 256)(torch.zeros(2, 257), torch.zeros(2, 256))
    with pytest.raises(RuntimeError, match="CUDA"):
        local_tex_modulation(Fuse_sft_MLP(257, 256), ResnetBlockFC(301, 512), torch.zeros(2, 257), torch.zeros(2, 256),
                             torch.zeros(2, 3))
