"""GPU parity of the local branch's pixel-aligned feature query (e3_local_feature_query)
against the fixture recorded from the reference's HGPIFuNetGAN.query and against the oracle at the real
size (256-channel 128x128 map, 64x64x24 sample points per image)."""
import os

import numpy as np
import pytest
import torch

from oracle import local_query_oracle as LQ
from helpers import rel_linf

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "local_query.npz")


def _away_from_border(xy, eps=1e-5):
    return ((xy.abs() - 1).abs() > eps).all(1)


@pytest.mark.parametrize("name", ["neg_z", "pos_z"])
def test_query_vs_reference_fixture(name):
    from e3dge_b200 import local_query
    z = np.load(GOLD)
    t = lambda k: torch.from_numpy(z[f"{name}.{k}"])
    out = local_query.query(t("points").cuda(), t("calibs").cuda(), im_feat=t("feat").cuda())
    assert out["feats"].shape == t("feats").shape and out["feats"].permute(0, 2, 1).is_contiguous()
    assert rel_linf(out["proj_xy"].cpu(), t("proj_xy")) < 1e-5
    assert rel_linf(out["depth"].cpu(), t("depth")) < 1e-6
    ok = _away_from_border(t("proj_xy"))
    assert torch.equal(out["in_img"].cpu()[ok], t("in_img")[ok])
    assert rel_linf(out["feats"].cpu(), t("feats")) < 1e-5
    assert out["interp_feats"] is out["feats"]


def test_query_at_full_size_both_point_layouts_and_projection_only():
    from e3dge_b200 import local_query
    g = torch.Generator().manual_seed(3)
    B, C, H, W, N = 2, 256, 128, 128, 64 * 64 * 24
    feat = torch.randn(B, C, H, W, generator=g)
    pts_bn3 = (torch.rand(B, N, 3, generator=g) - 0.5) * 0.3    # the renderer's `points` layout [B,N,3]
    z = np.load(GOLD)
    calibs = torch.from_numpy(z["neg_z.calibs"])[:B]
    ref = LQ.local_feature_query(pts_bn3.permute(0, 2, 1).contiguous(), calibs, feat)
    d = pts_bn3.cuda()
    a = local_query.query(d.permute(0, 2, 1), calibs.cuda(), im_feat=feat.cuda())            # strided view, in place
    b = local_query.query(d.permute(0, 2, 1).contiguous(), calibs[:, :3].cuda(), im_feat=feat.cuda())  # [B,3,4] calibs
    assert torch.equal(a["feats"], b["feats"]) and torch.equal(a["proj_xy"], b["proj_xy"])
    assert rel_linf(a["feats"].cpu(), ref["feats"]) < 1e-4
    assert rel_linf(a["proj_xy"].cpu(), ref["proj_xy"]) < 1e-5
    ok = _away_from_border(ref["proj_xy"])
    assert torch.equal(a["in_img"].cpu()[ok], ref["in_img"][ok])
    frac = a["in_img"].float().mean().item()
    assert 0.3 < frac < 0.99, frac
    p = local_query.query(d.permute(0, 2, 1), calibs.cuda(), return_projection_only=True)
    assert set(p) == {"proj_xy", "depth", "in_img"} and torch.equal(p["proj_xy"], a["proj_xy"])
    # points outside the image read zeros where all four taps fall off the map
    far = (ref["proj_xy"].abs() > 1.0 + 2.0 / W).any(1)
    assert far.any() and a["feats"].cpu().permute(0, 2, 1)[far].abs().max().item() == 0.0


def test_query_rejects_cpu_tensors_and_empty_batches_are_fine():
    from e3dge_b200 import local_query
    with pytest.raises(RuntimeError):
        local_query.query(torch.zeros(1, 3, 4), torch.eye(4)[None], im_feat=torch.zeros(1, 4, 2, 2))
    out = local_query.query(torch.zeros(1, 3, 0, device="cuda"), torch.eye(4, device="cuda")[None],
                            im_feat=torch.zeros(1, 4, 2, 2, device="cuda"))
    assert out["feats"].shape == (1, 4, 0)


def test_install_patches_a_reference_style_module():
    from e3dge_b200 import local_query

    class Net:
        def query(self, points, calibs, feat_key, **kw):
            return "original"
    net = local_query.install(Net())
    z = np.load(GOLD)
    t = lambda k: torch.from_numpy(z[f"neg_z.{k}"]).cuda()
    with torch.no_grad():
        out = net.query(points=t("points"), calibs=t("calibs"), feat_key="ref_view", return_eikonal=False,
                        return_feat_only=True, im_feat=t("feat"))
        assert rel_linf(out["feats"].cpu(), torch.from_numpy(z["neg_z.feats"])) < 1e-5
        assert net.query(t("points"), t("calibs"), "ref_view") == "original"   # stored-feature path: untouched


def test_query_backward_to_the_feature_map_matches_the_oracle_autograd():
    """Stage-2 training reaches netLocal's hourglass filter through the gather: d loss / d feature map from
    e3_local_feature_query_bwd against autograd through the oracle's bilinear sampling (float64)."""
    from e3dge_b200 import local_query as lq
    from e3dge_b200.frontend import generate_camera_params
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(2, 16, 12, 20, generator=g)
    pts = (torch.rand(2, 3, 300, generator=g) - 0.5) * 0.3
    pts[:, :, -40:] *= 6.0  # a few far outside the frustum: zero padding, no gradient
    calibs = generate_camera_params(64, torch.device("cpu"), 2, return_calibs=True, generator=g)["calibs"]
    cot = torch.randn(2, 16, 300, generator=g)
    f64 = feat.double().requires_grad_(True)
    ref = LQ.local_feature_query(pts.double(), calibs.double(), f64)
    gref, = torch.autograd.grad((ref["feats"] * cot.double()).sum(), [f64])
    fc = feat.cuda().requires_grad_(True)
    out = lq.query(pts.cuda(), calibs.cuda(), im_feat=fc)
    assert out["feats"].requires_grad
    ggot, = torch.autograd.grad((out["feats"] * cot.cuda()).sum(), [fc])
    assert rel_linf(out["feats"].detach().cpu(), ref["feats"].detach()) < 1e-5
    assert rel_linf(ggot.cpu(), gref) < 1e-5
    # no graph when nothing asks for one
    with torch.no_grad():
        assert not lq.query(pts.cuda(), calibs.cuda(), im_feat=fc)["feats"].requires_grad
