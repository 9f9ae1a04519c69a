"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz from the REAL reference.

Run inside the build container (needs /root/reference):

    python -m oracle.gen_golden            # all cases
    python -m oracle.gen_golden small_wplus

Each case builds the reference's own modules (through oracle/ref_harness.py, nothing is
copied), overwrites their parameters with oracle/params.py's deterministic values, runs
the reference forward on CPU in float32 under torch.no_grad(), and stores the outputs
(full tensors for small cases; strided sub-samples plus float64 checksums for the
full-size case) together with the case's configuration.  The fixtures are what pins
oracle/stylesdf_oracle.py — and, on the GPU box where the reference does not exist,
they are the reference's voice in the `-m gpu` parity tests.
"""
import json
import os
import sys

import numpy as np
import torch

from oracle import params as P
from oracle import ref_harness as rh

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _checksums(t):
    d = t.detach().double()
    return np.array([d.sum().item(), d.abs().sum().item(), (d * d).sum().item()])


def build_generator(ref, size, res, n_samples, seed, variant, full_pipeline=True, **ropt):
    G = ref.stylesdf_model.G_pred_latents(
        rh.model_opt(size=size, renderer_spatial_output_dim=res),
        rh.rendering_opt(N_samples=n_samples, **ropt), full_pipeline=full_pipeline)
    G.eval()
    sd = P.fill_state_dict(G.state_dict(), seed=seed, variant=variant)
    G.load_state_dict(sd, strict=True)
    return G


# ------------------------------------------------------------------------------------
# cases
# ------------------------------------------------------------------------------------
GEN_CASES = {
    # name: (config, sub-sample stride for big tensors or None)
    "small_wplus": dict(size=64, res=16, n_samples=24, batch=2, seed=11,
                        variant="default", wplus=True, ropt={}),
    "small_sharp_w": dict(size=64, res=16, n_samples=24, batch=1, seed=12,
                          variant="sharp", wplus=False, ropt={}),
    "small_s18_rayd_viewdirs": dict(size=64, res=16, n_samples=18, batch=2, seed=13,
                                    variant="sharp", wplus=True,
                                    ropt=dict(static_viewdirs=False)),
    "small_stratified_ss2": dict(size=64, res=8, n_samples=12, batch=1, seed=14,
                                 variant="default", wplus=True,
                                 ropt=dict(no_offset_sampling=True,
                                           spatial_super_sampling_factor=2,
                                           force_background=False), renderer_only=True),
    "full_256": dict(size=256, res=64, n_samples=24, batch=2, seed=21, variant="sharp",
                     wplus=True, ropt={}, stride=8, offset=3),
    # density branch without the SDF activation (volume_renderer.py:862-867): alpha = 1-exp(-softplus(raw)*dist)
    "small_no_sdf": dict(size=64, res=16, n_samples=24, batch=2, seed=15, variant="sharp",
                         wplus=True, ropt=dict(no_sdf=True), renderer_only=True),
}

RENDER_KEYS = ["rays_o", "rays_d", "dists", "near", "far", "hit_prob", "points", "sdf",
               "gen_thumb_imgs", "features", "mask", "xyz", "depth", "viewdirs"]


def run_generator_case(ref, name, cfg):
    torch.manual_seed(0)
    G = build_generator(ref, cfg["size"], cfg["res"], cfg["n_samples"], cfg["seed"],
                        cfg["variant"], **cfg["ropt"])
    ss = cfg["ropt"].get("spatial_super_sampling_factor", 1)
    inp = P.make_inputs(cfg["seed"], cfg["batch"], G.decoder.n_latent, cfg["res"],
                        wplus=cfg["wplus"])
    with torch.no_grad():
        out = G([inp["w"], inp["w_dec"]], inp["cam_poses"], inp["focal"], inp["near"],
                inp["far"], input_is_latent=True, randomize_noise=False,
                return_xyz=True, return_sdf=True,
                renderer_only=cfg.get("renderer_only", False))
    stride = cfg.get("stride")
    arrays = {}
    keys = RENDER_KEYS + ([] if cfg.get("renderer_only") else ["gen_imgs"])

    def sub(t, k, first):
        if k in ("features", "gen_thumb_imgs", "xyz", "gen_imgs", "mask"):  # [B,C,H,W(,1)]
            return t[:, :, first::stride, first::stride]
        return t[:, first::stride, first::stride]  # [B,H,W,...]

    for k in keys:
        t = out[k]
        arrays["sum." + k] = _checksums(t)
        if stride:
            # a second sub-sample on an offset grid: together the two lattices touch every 128-row tile of
            # the renderer (5 rays) and every 8x16 / 4x32 conv tile of the decoder
            if cfg.get("offset"):
                arrays["off." + k] = _np(sub(t, k, cfg["offset"])).astype(np.float32)
            t = sub(t, k, 0)
        arrays[k] = _np(t).astype(np.float32)
    arrays["config"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    return arrays


def run_localmod_case(ref):
    """Texture modulation of the local branch, driven through the reference's own
    SirenGenerator.forward_backbone / forward_geo / forward_tex and
    VolumeFeatureRenderer.volume_integration (volume_renderer.py:196-238, 809-943)."""
    cfg = dict(size=64, res=8, n_samples=24, batch=2, seed=31, variant="sharp", wplus=True)
    torch.manual_seed(0)
    G = build_generator(ref, cfg["size"], cfg["res"], cfg["n_samples"], cfg["seed"],
                        cfg["variant"], local_modulation_layer=True)
    R = G.renderer
    inp = P.make_inputs(cfg["seed"], cfg["batch"], G.decoder.n_latent, cfg["res"])
    rng = np.random.Generator(np.random.PCG64(cfg["seed"]))
    shp = (cfg["batch"], cfg["res"], cfg["res"], cfg["n_samples"], 256)
    alpha = torch.from_numpy(rng.standard_normal(shp).astype(np.float32) * 0.3)
    beta = torch.from_numpy(rng.standard_normal(shp).astype(np.float32) * 0.3)
    with torch.no_grad():
        rays_o, rays_d, viewdirs = R.get_rays(inp["focal"], inp["cam_poses"])
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        near = inp["near"].unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        far = inp["far"].unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        z_vals = near * (1. - R.t_vals) + far * R.t_vals
        pts = rays_o.unsqueeze(3) + rays_d.unsqueeze(3) * z_vals.unsqueeze(-1)
        net = R.network
        x = R.grid_warper(pts)
        h = net.forward_backbone(x.contiguous(), inp["w"])
        sdf = net.forward_geo(h)
        rgb, feat = net.forward_tex(h, viewdirs.unsqueeze(3).expand(pts.shape), inp["w"],
                                    conditions=dict(tex=[alpha, beta]))
        raw = torch.cat([rgb, sdf, feat], -1)
        R.local_batch = None
        vi = R.volume_integration(raw, z_vals, rays_d, pts, return_eikonal=False,
                                  return_surface_eikonal=False)
    rgb_map, feature_map, sdf_out, mask, xyz = vi[:5]
    arrays = dict(features=_np(feature_map.permute(0, 3, 1, 2)),
                  gen_thumb_imgs=_np(rgb_map.permute(0, 3, 1, 2)), sdf=_np(sdf_out),
                  xyz=_np(xyz.permute(0, 3, 1, 2)), hit_prob=_np(vi[11]))
    arrays["config"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    return arrays


def run_query_case(ref):
    """SDF-only point query and the no_force_stop composite used by the visibility
    queries (volume_renderer.py:955-957, 830-836, 1935-1943)."""
    cfg = dict(size=64, res=8, n_samples=24, batch=2, seed=41, variant="sharp", wplus=True,
               n_points=200)
    torch.manual_seed(0)
    G = build_generator(ref, cfg["size"], cfg["res"], cfg["n_samples"], cfg["seed"],
                        cfg["variant"], full_pipeline=False)
    R = G.renderer
    inp = P.make_inputs(cfg["seed"], cfg["batch"], 1, cfg["res"])
    rng = np.random.Generator(np.random.PCG64(cfg["seed"]))
    pts = torch.from_numpy(rng.uniform(-0.12, 0.12, (cfg["batch"], cfg["n_points"], 3))
                           .astype(np.float32))
    with torch.no_grad():
        R.local_batch = None
        p5 = pts.reshape(cfg["batch"], -1, 1, 1, 3)
        sdf = R.run_network(p5, torch.zeros_like(p5), styles=inp["w"])[..., 3:4]
        # no_force_stop composite on the regular ray batch
        rays_o, rays_d, viewdirs = R.get_rays(inp["focal"], inp["cam_poses"])
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        near = inp["near"].unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        far = inp["far"].unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        z_vals = near * (1. - R.t_vals) + far * R.t_vals
        rp = rays_o.unsqueeze(3) + rays_d.unsqueeze(3) * z_vals.unsqueeze(-1)
        raw = R.run_network(rp, viewdirs, styles=inp["w"])
        vi = R.volume_integration(raw, z_vals, rays_d, rp, return_eikonal=False,
                                  return_surface_eikonal=False, no_force_stop=True)
    arrays = dict(points=_np(pts), sdf_query=_np(sdf.reshape(cfg["batch"], -1, 1)),
                  nfs_features=_np(vi[1].permute(0, 3, 1, 2)), nfs_hit_prob=_np(vi[11]),
                  nfs_visibility=_np(vi[10]), nfs_dists=_np(vi[9]))
    arrays["config"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    return arrays


def run_ops_case(ref):
    """fused_leaky_relu / upfirdn2d CPU branches (op/fused_act.py:107-115,
    op/upfirdn2d.py:157-200) and ModulatedConv2d / StyledConv / ToRGB / mapping nets."""
    rng = np.random.Generator(np.random.PCG64(51))
    f32 = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    sm, fa, ud = ref.stylesdf_model, ref.fused_act, ref.upfirdn2d
    arrays = {}
    with torch.no_grad():
        x = f32(2, 6, 9, 11)
        b = f32(6)
        arrays["flr.x"], arrays["flr.b"] = _np(x), _np(b)
        arrays["flr.y"] = _np(fa.fused_leaky_relu(x, b))
        arrays["flr.y_nobias_scale1"] = _np(fa.fused_leaky_relu(x, None, scale=1))
        x2 = f32(3, 7)
        b2 = f32(7)
        arrays["flr.x2"], arrays["flr.b2"] = _np(x2), _np(b2)
        arrays["flr.y2"] = _np(fa.fused_leaky_relu(x2, b2, scale=1))
        k4 = sm.make_kernel([1, 3, 3, 1])
        k3 = sm.make_kernel([1, 2, 1])
        u = f32(2, 3, 10, 12)
        arrays["ufd.x"] = _np(u)
        cfgs = [("up2_pad21", k4 * 4, 2, 1, (2, 1)), ("blur_pad11", k4 * 4, 1, 1, (1, 1)),
                ("blur_pad22", k4, 1, 1, (2, 2)), ("down2_pad11", k4, 1, 2, (1, 1)),
                ("k3_up2_pad10", k3 * 4, 2, 1, (1, 0)), ("crop_padm1", k4, 1, 1, (-1, 2)),
                ("up2_down2", k4, 2, 2, (2, 1))]
        arrays["ufd.cfg"] = np.frombuffer(
            json.dumps([(n, _np(k).tolist(), up, dn, list(pd)) for n, k, up, dn, pd in cfgs])
            .encode(), dtype=np.uint8)
        for n, k, up, dn, pd in cfgs:
            arrays["ufd.y." + n] = _np(ud.upfirdn2d(u, k, up=up, down=dn, pad=pd))

        # modulated convs, driven module by module
        for tag, (cin, cout, ksz, upsample, demod, hw) in {
                "conv3": (16, 24, 3, False, True, 10), "conv3_up": (16, 8, 3, True, True, 6),
                "conv1_nodemod": (16, 3, 1, False, False, 7)}.items():
            m = sm.ModulatedConv2d(cin, cout, ksz, 512, demodulate=demod, upsample=upsample)
            sd = P.fill_state_dict({"decoder.x.conv." + k: v for k, v in m.state_dict().items()},
                                   seed=52)
            m.load_state_dict({k[len("decoder.x.conv."):]: v for k, v in sd.items()})
            xi, st = f32(2, cin, hw, hw), f32(2, 512)
            arrays[f"mc.{tag}.x"], arrays[f"mc.{tag}.style"] = _np(xi), _np(st)
            arrays[f"mc.{tag}.y"] = _np(m(xi, st))
        # mapping networks
        G = build_generator(ref, 64, 16, 24, 53, "default")
        z = f32(4, 256)
        arrays["map.z"] = _np(z)
        w = G.style(z)
        arrays["map.w"] = _np(w)
        arrays["map.w_dec"] = _np(G.decoder.style(w))
    return arrays


GRAD_KEYS = ["gen_imgs", "gen_thumb_imgs", "features", "sdf", "xyz", "depth", "hit_prob"]


def cotangent(key, shape):
    """Deterministic upstream gradient for output `key` (no RNG: reproducible anywhere)."""
    n = int(np.prod(shape))
    k = GRAD_KEYS.index(key)
    return torch.from_numpy(np.cos(np.arange(n, dtype=np.float64) * 0.37 + k).astype(np.float32)
                            .reshape(shape))


def run_grad_case(ref):
    """Autograd of the REAL reference through G_pred_latents.forward: for every differentiable
    output k, d <R_k, out_k> / d w+ (and / d decoder latent for the image), plus the eikonal
    term of return_eikonal=True (volume_renderer.py:796-802, 855-856)."""
    cfg = dict(size=32, res=8, n_samples=12, batch=2, seed=61, variant="sharp", wplus=True, ropt={})
    torch.manual_seed(0)
    G = build_generator(ref, cfg["size"], cfg["res"], cfg["n_samples"], cfg["seed"], cfg["variant"])
    for p in G.parameters():
        p.requires_grad_(False)
    inp = P.make_inputs(cfg["seed"], cfg["batch"], G.decoder.n_latent, cfg["res"], wplus=True)
    w = inp["w"].clone().requires_grad_(True)
    wd = inp["w_dec"].clone().requires_grad_(True)
    out = G([w, wd], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], input_is_latent=True,
            randomize_noise=False, return_xyz=True, return_sdf=True, return_eikonal=True)
    arrays = {"eikonal_term": _np(out["eikonal_term"]).astype(np.float32)}
    for k in GRAD_KEYS:
        loss = (cotangent(k, tuple(out[k].shape)) * out[k]).sum()
        gw, gd = torch.autograd.grad(loss, [w, wd], retain_graph=True, allow_unused=True)
        arrays["dw." + k] = _np(gw).astype(np.float32)
        if gd is not None and k == "gen_imgs":
            arrays["dwdec." + k] = _np(gd).astype(np.float32)
    arrays["config"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    return arrays


def run_visibility_case(ref):
    """query_hitting_probability_{fixed,adapted}_interval of the REAL reference
    (volume_renderer.py:1326-1621)."""
    cfg = P.VIS_CFG
    torch.manual_seed(0)
    G = build_generator(ref, cfg["size"], cfg["res"], cfg["n_samples"], cfg["seed"], cfg["variant"],
                        full_pipeline=False)
    R = G.renderer
    R.local_batch = None
    pts, info = P.visibility_case_inputs(cfg)
    with torch.no_grad():
        arrays = dict(
            fixed_weights=_np(R.query_hitting_probability_fixed_interval(pts, info, "weights")),
            fixed_visibility=_np(R.query_hitting_probability_fixed_interval(pts, info, "visibility")),
            adapted=_np(R.query_hitting_probability_adapted_interval(pts, info)))
    arrays["config"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    return arrays


def main(argv):
    if not rh.reference_available():
        raise SystemExit("reference tree not available: cannot regenerate goldens")
    ref = rh.load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    want = set(argv[1:])
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    jobs = {n: (lambda n=n, c=c: run_generator_case(ref, n, c)) for n, c in GEN_CASES.items()}
    jobs["small_localmod"] = lambda: run_localmod_case(ref)
    jobs["small_query_nfs"] = lambda: run_query_case(ref)
    jobs["ops"] = lambda: run_ops_case(ref)
    jobs["small_grad"] = lambda: run_grad_case(ref)
    jobs["small_visibility"] = lambda: run_visibility_case(ref)
    for name, job in jobs.items():
        if want and name not in want:
            continue
        arrays = job()
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {len(arrays)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv)
