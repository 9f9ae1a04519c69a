"""Pins oracle/stylesdf_oracle.py against fixtures recorded from the REAL reference
(oracle/gen_golden.py).  CPU only.  Same ATen kernels, same op order => the bar here is
float32 round-off (1e-5 rel), far below the 1e-3 product tolerance."""
import json

import numpy as np
import pytest
import torch

from oracle import params as P
from oracle import stylesdf_oracle as O
from helpers import load_golden, rel_linf, synthetic_state_dict, decoder_layout

TOL = 2e-5

GEN_CASES = ["small_wplus", "small_sharp_w", "small_s18_rayd_viewdirs",
             "small_stratified_ss2", "full_256", "small_no_sdf"]


def _sub(t, k, stride, first=0):
    if not stride:
        return t
    if k in ("features", "gen_thumb_imgs", "xyz", "gen_imgs", "mask"):
        return t[:, :, first::stride, first::stride]
    return t[:, first::stride, first::stride]


@pytest.mark.parametrize("name", GEN_CASES)
def test_generator_cases(name):
    gold, cfg = load_golden(name)
    torch.set_num_threads(8)
    sd = synthetic_state_dict(cfg["size"], cfg["res"], cfg["seed"], cfg["variant"])
    n_lat = decoder_layout(cfg["size"], cfg["res"])
    inp = P.make_inputs(cfg["seed"], cfg["batch"], n_lat, cfg["res"], wplus=cfg["wplus"])
    ro = cfg["ropt"]
    kw = dict(res=cfg["res"], n_samples=cfg["n_samples"],
              spatial_ss=ro.get("spatial_super_sampling_factor", 1),
              static_viewdirs=ro.get("static_viewdirs", True),
              offset_sampling=not ro.get("no_offset_sampling", False),
              force_background=ro.get("force_background", True),
              with_sdf=not ro.get("no_sdf", False))
    with torch.no_grad():
        if cfg.get("renderer_only"):
            out = O.renderer_forward(sd, inp["cam_poses"], inp["focal"], inp["near"],
                                     inp["far"], inp["w"], **kw)
        else:
            out = O.generator_forward(sd, inp["w"], inp["w_dec"], inp["cam_poses"],
                                      inp["focal"], inp["near"], inp["far"], **kw)
    stride = cfg.get("stride")
    checked = 0
    for k, g in gold.items():
        if k.startswith("sum."):
            t = out[k[4:]].double()
            got = np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])
            np.testing.assert_allclose(got[1:], g[1:], rtol=1e-5, err_msg=k)
            continue
        if k.startswith("off."):  # second sub-sample lattice, offset by cfg["offset"]
            got, k = _sub(out[k[4:]], k[4:], stride, cfg["offset"]), k
        else:
            got = _sub(out[k], k, stride)
        assert tuple(got.shape) == g.shape, (k, got.shape, g.shape)
        err = rel_linf(got, g)
        assert err < TOL, f"{name}:{k} rel-Linf {err:.3e}"
        checked += 1
    assert checked >= 14


def test_localmod_case():
    gold, cfg = load_golden("small_localmod")
    sd = synthetic_state_dict(cfg["size"], cfg["res"], cfg["seed"], cfg["variant"])
    inp = P.make_inputs(cfg["seed"], cfg["batch"], decoder_layout(cfg["size"], cfg["res"]),
                        cfg["res"])
    rng = np.random.Generator(np.random.PCG64(cfg["seed"]))
    shp = (cfg["batch"], cfg["res"], cfg["res"], cfg["n_samples"], 256)
    alpha = torch.from_numpy(rng.standard_normal(shp).astype(np.float32) * 0.3)
    beta = torch.from_numpy(rng.standard_normal(shp).astype(np.float32) * 0.3)
    with torch.no_grad():
        out = O.renderer_forward(sd, inp["cam_poses"], inp["focal"], inp["near"],
                                 inp["far"], inp["w"], res=cfg["res"],
                                 n_samples=cfg["n_samples"], local_mod=(alpha, beta))
    for k in ("features", "gen_thumb_imgs", "sdf", "xyz", "hit_prob"):
        assert rel_linf(out[k], gold[k]) < TOL, k


def test_query_and_no_force_stop_case():
    gold, cfg = load_golden("small_query_nfs")
    sd = synthetic_state_dict(cfg["size"], cfg["res"], cfg["seed"], cfg["variant"])
    inp = P.make_inputs(cfg["seed"], cfg["batch"], 1, cfg["res"])
    pts = torch.from_numpy(gold["points"])
    with torch.no_grad():
        sdf = O.sdf_query(sd, pts, inp["w"])
        rays_o, rays_d, vd = O.get_rays(inp["focal"], inp["cam_poses"], cfg["res"])
        vd = vd / torch.norm(vd, dim=-1, keepdim=True)
        near = inp["near"].unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        far = inp["far"].unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        z = O.sample_z(near, far, cfg["n_samples"])
        p = rays_o.unsqueeze(3) + rays_d.unsqueeze(3) * z.unsqueeze(-1)
        raw = O.run_network(p, vd, inp["w"], sd)
        vi = O.volume_integration(raw, z, rays_d, p, sd["renderer.sigmoid_beta"],
                                  no_force_stop=True)
    assert rel_linf(sdf, gold["sdf_query"]) < TOL
    assert rel_linf(vi["feature_map"].permute(0, 3, 1, 2), gold["nfs_features"]) < TOL
    assert rel_linf(vi["weights"], gold["nfs_hit_prob"]) < TOL
    assert rel_linf(vi["visibility"], gold["nfs_visibility"]) < TOL
    assert rel_linf(vi["dists"], gold["nfs_dists"]) < TOL


def test_ops_case():
    gold, _ = load_golden("ops")
    t = lambda k: torch.from_numpy(gold[k])
    assert rel_linf(O.fused_leaky_relu(t("flr.x"), t("flr.b")), gold["flr.y"]) < 1e-6
    assert rel_linf(O.fused_leaky_relu(t("flr.x"), None, scale=1.0),
                    gold["flr.y_nobias_scale1"]) < 1e-6
    assert rel_linf(O.fused_leaky_relu(t("flr.x2"), t("flr.b2"), scale=1.0),
                    gold["flr.y2"]) < 1e-6
    for n, k, up, dn, pd in json.loads(bytes(gold["ufd.cfg"]).decode()):
        y = O.upfirdn2d(t("ufd.x"), torch.tensor(k, dtype=torch.float32), up, dn, tuple(pd))
        assert tuple(y.shape) == gold["ufd.y." + n].shape, n
        assert rel_linf(y, gold["ufd.y." + n]) < 1e-6, n
    for tag, (cin, cout, ksz, upsample, demod) in {
            "conv3": (16, 24, 3, False, True), "conv3_up": (16, 8, 3, True, True),
            "conv1_nodemod": (16, 3, 1, False, False)}.items():
        sd = {}
        for leaf, shape in (("weight", (1, cout, cin, ksz, ksz)),
                            ("modulation.weight", (cin, 512)), ("modulation.bias", (cin,))):
            name = "decoder.x.conv." + leaf
            sd[name] = torch.from_numpy(P.make_param(52, name, shape)).float()
        y = O.modulated_conv2d(t(f"mc.{tag}.x"), t(f"mc.{tag}.style"), sd, "decoder.x.conv.",
                               demodulate=demod, upsample=upsample)
        assert rel_linf(y, gold[f"mc.{tag}.y"]) < 1e-5, tag
    sd = synthetic_state_dict(64, 16, 53)
    w = O.mapping_network(t("map.z"), sd)
    assert rel_linf(w, gold["map.w"]) < 1e-5
    assert rel_linf(O.decoder_mapping(w, sd), gold["map.w_dec"]) < 1e-5


GRAD_KEYS = ["gen_imgs", "gen_thumb_imgs", "features", "sdf", "xyz", "depth", "hit_prob"]


def cotangent(key, shape):
    """The deterministic upstream gradients of oracle/gen_golden.py::cotangent."""
    n = int(np.prod(shape))
    k = GRAD_KEYS.index(key)
    return torch.from_numpy(np.cos(np.arange(n, dtype=np.float64) * 0.37 + k).astype(np.float32)
                            .reshape(shape))


def test_gradients_and_eikonal_term_vs_reference_autograd():
    """Autograd through the oracle == autograd through the real reference (fixture
    small_grad): d<R_k, out_k>/d w+, d image / d decoder latent, and the eikonal term."""
    gold, cfg = load_golden("small_grad")
    torch.set_num_threads(8)
    sd = synthetic_state_dict(cfg["size"], cfg["res"], cfg["seed"], cfg["variant"])
    n_lat = decoder_layout(cfg["size"], cfg["res"])
    inp = P.make_inputs(cfg["seed"], cfg["batch"], n_lat, cfg["res"], wplus=True)
    w = inp["w"].clone().requires_grad_(True)
    wd = inp["w_dec"].clone().requires_grad_(True)
    out = O.generator_forward(sd, w, wd, inp["cam_poses"], inp["focal"], inp["near"], inp["far"],
                              res=cfg["res"], n_samples=cfg["n_samples"])
    for k in GRAD_KEYS:
        loss = (cotangent(k, tuple(out[k].shape)) * out[k]).sum()
        gw, gd = torch.autograd.grad(loss, [w, wd], retain_graph=True, allow_unused=True)
        assert rel_linf(gw, gold["dw." + k]) < 2e-4, k
        if k == "gen_imgs":
            assert rel_linf(gd, gold["dwdec." + k]) < 2e-4
    pts = out["points"].detach().clone().requires_grad_(True)
    sdf = O.sdf_query(sd, pts.reshape(cfg["batch"], -1, 3), inp["w"])
    eik = torch.autograd.grad(sdf.sum(), pts)[0]
    assert rel_linf(eik, gold["eikonal_term"]) < 2e-4


def test_visibility_queries_vs_reference_fixture():
    """query_hitting_probability_{fixed,adapted}_interval (volume_renderer.py:1326-1621)."""
    gold, cfg = load_golden("small_visibility")
    sd = synthetic_state_dict(cfg["size"], cfg["res"], cfg["seed"], cfg["variant"])
    pts, info = P.visibility_case_inputs(cfg)
    cs, ro = info["cam_settings"], info["global_render_out"]
    args = (sd, pts, cs["poses"], cs["extrinsics"], ro["near"], ro["far"], info["pred_latents"][0])
    with torch.no_grad():
        got = dict(fixed_weights=O.query_hitting_probability(*args, n_samples=cfg["n_samples"]),
                   fixed_visibility=O.query_hitting_probability(*args, n_samples=cfg["n_samples"],
                                                                return_type="visibility"),
                   adapted=O.query_hitting_probability(*args, n_samples=cfg["n_samples"], mode="adapted"))
    for k, v in got.items():
        assert tuple(v.shape) == gold[k].shape
        assert rel_linf(v, gold[k]) < 1e-4, k


def test_local_feature_query_oracle_vs_reference_fixture():
    """oracle/local_query_oracle.py against tests/golden/local_query.npz, recorded from the reference's own
    HGPIFuNetGAN.query / geometry.perspective / geometry.index (oracle/gen_golden_local_query.py): both
    signs of the camera's z axis, points inside and far outside the frustum."""
    import os
    import numpy as np
    from oracle import local_query_oracle as LQ
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "local_query.npz"))
    for name in ("neg_z", "pos_z"):
        t = lambda k: torch.from_numpy(z[f"{name}.{k}"])
        out = LQ.local_feature_query(t("points"), t("calibs"), t("feat"))
        assert torch.equal(out["in_img"], t("in_img"))
        for k in ("proj_xy", "depth", "feats"):
            ref = t(k)
            assert (out[k] - ref).abs().max().item() <= 2e-6 * max(ref.abs().max().item(), 1.0), (name, k)
        assert out["in_img"].float().mean().item() > 0.2 and (~out["in_img"]).float().mean().item() > 0.2


def test_local_mlp_case():
    """oracle/local_mlp_oracle.py against the reference's own Fuse_sft_MLP / PosEncoding / ResnetBlockFC and
    its `--enable_local_model` renderer (tests/golden/local_mlp.npz, oracle/gen_golden_local_mlp.py)."""
    from oracle import local_mlp_oracle as L
    from helpers import local_mlp_state_dict, synthetic_local_feats
    gold, cfg = load_golden("local_mlp")
    sd = local_mlp_state_dict(41)
    f2, f3, pts = (torch.from_numpy(gold["mlp." + k]) for k in ("feat_2d", "feat_3d", "points"))
    with torch.no_grad():
        feats = L.local_feats(f2, f3, pts, sd)
        alpha, beta = L.tex_modulation(feats, sd)
    assert rel_linf(feats, gold["mlp.feats"]) < TOL
    assert rel_linf(alpha, gold["mlp.alpha"]) < TOL
    assert rel_linf(beta, gold["mlp.beta"]) < TOL
    # the whole local renderer pass: global pass -> points -> 301-d features -> modulated render
    seed = cfg["seed"]
    gsd = synthetic_state_dict(cfg["size"], cfg["res"], seed, cfg["variant"], local=True)
    gsd = {k.replace("renderer.network.netGlobal.", "renderer.network."): v for k, v in gsd.items()}
    lsd = local_mlp_state_dict(seed, cfg["variant"])
    inp = P.make_inputs(seed, cfg["batch"], 2, cfg["res"])
    pts = torch.from_numpy(gold["render.points"])
    f2, f3 = synthetic_local_feats(seed, pts.shape[:-1])
    with torch.no_grad():
        mod = L.local_tex_modulation(f2, f3, pts, lsd)
        out = O.renderer_forward(gsd, inp["cam_poses"], inp["focal"], inp["near"], inp["far"], inp["w"],
                                 res=cfg["res"], n_samples=cfg["n_samples"], local_mod=mod)
        glob = O.renderer_forward(gsd, inp["cam_poses"], inp["focal"], inp["near"], inp["far"], inp["w"],
                                  res=cfg["res"], n_samples=cfg["n_samples"])
    assert rel_linf(out["points"], gold["render.points"]) < TOL
    assert rel_linf(glob["features"], gold["render.global_features"]) < TOL
    for k in ("features", "gen_thumb_imgs", "sdf", "hit_prob", "xyz", "depth"):
        assert rel_linf(out[k], gold["render." + k]) < 5e-5, k
