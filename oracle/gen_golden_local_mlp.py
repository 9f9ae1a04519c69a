#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — records tests/golden/local_mlp.npz from the REAL reference code, executed in
place from /root/reference (no file copied; SURVEY.md §8c recipe 2b):

  case "mlp"     the reference's own `Fuse_sft_MLP(257, 256)` (project/models/helper_modules/sft.py:84-109),
                 `PosEncoding(3, N_freqs=7)` (project/utils/misc_utils.py:148-184) and the `ResnetBlockFC(301, 512)`
                 that `HGPIFuNetGANResidualResnetFC.build_modulation_net` creates, chained exactly as
                 `E3DGE_Full_Runner.que_render_given_ref` does (e3dge_full_runner.py:282-297) and split into
                 (alpha, beta) as `SirenLocalGlobal.forward_backbone` does (volume_renderer.py:327-336);
  case "render"  the reference's `VolumeFeatureRenderer` built with `--enable_local_model` (real
                 `SirenLocalGlobal` + real 15.3 M-parameter `HGPIFuNetGANResidualResnetFC`), called with
                 `local_data_batch={'feats': <301-d features of case "mlp" at the renderer's own sample points>}`.

Weights are the deterministic synthetic ones of synthetic_inputs.py.  Run in the build container from the
repository root:   python oracle/gen_golden_local_mlp.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
np.deprecate = lambda f=None, *a, **k: (f if callable(f) else (lambda g: g))  # vendor/pifu/lib/geometry.py:1
import torch  # noqa: E402
from oracle import ref_harness as H  # noqa: E402
import synthetic_inputs as P  # noqa: E402

PIFU = os.path.join(H.REFERENCE_ROOT, "project", "vendor", "pifu")


def load():
    """The reference with its vendored PIFu importable (`lib.*` real except `lib.data`)."""
    sys.path.insert(0, PIFU)
    stub = H._stub_module
    H._stub_module = lambda name: None if (name.startswith("lib") and name != "lib.data") else stub(name)
    for n in ("munch", "omegaconf", "omegaconf.dictconfig", "IPython", "IPython.display"):
        if n not in sys.modules:
            stub(n)
    os.chdir(H.REFERENCE_ROOT)  # volume_renderer.py:15 appends a cwd-relative path
    ref = H.load_reference()
    H._stub_module = stub
    H._shell_package("project.models.helper_modules",
                     os.path.join(H.REFERENCE_ROOT, "project", "models", "helper_modules"))
    from project.models.helper_modules.sft import Fuse_sft_MLP
    from project.utils.misc_utils import PosEncoding
    return ref, Fuse_sft_MLP, PosEncoding


def pifu_opt():
    """vendor/pifu/lib/options.py:162-216 defaults with the overrides of demo_view_synthesis.sh:9-10,45-46,80-84."""
    return H.Opt(num_views=1, enforce_minmax=False, uniform_pts_loss="l1", loadSize=256, z_size=1.12, norm="group",
                 num_stack=4, num_hourglass=2, skip_hourglass=False, hg_input_channel=64, hg_down="ave_pool",
                 hourglass_dim=256, mlp_dim=[257, 1024, 512, 256, 128, 1], init_type="normal", no_residual=False,
                 mlp_dim_color=[513, 1024, 512, 256, 128, 3], use_tanh=False, debug=False)


def local_rendering_opt(**over):
    return H.rendering_opt(enable_local_model=True, local_modulation_layer=True, L_pred_tex_modulations=True,
                           L_pred_geo_modulations=False, netLocal_type="HGPIFuNetGANResidualResnetFC",
                           residual_local_feats_dim=301, tex_predictition_strategy="global_local",
                           geo_predictition_strategy="global", pifu=pifu_opt(), **over)


def fill(module, prefix, seed):
    sd = {k: torch.from_numpy(np.ascontiguousarray(P.make_param(seed, prefix + k, v.shape))).float()
          for k, v in module.state_dict().items()}
    module.load_state_dict(sd, strict=True)


def synthetic_feats(seed, shape_prefix):
    """feature_2dAlign | visibility mask [...,257] and feature_3dprojection [...,256] (e3dge_full_runner.py:229-288)."""
    rng = np.random.Generator(np.random.PCG64([seed, 77]))
    f2 = rng.standard_normal(shape_prefix + (256,)).astype(np.float32)
    vis = (rng.uniform(size=shape_prefix + (1,)) < 0.7).astype(np.float32)
    f3 = rng.standard_normal(shape_prefix + (256,)).astype(np.float32)
    return torch.from_numpy(np.concatenate([f2, vis], -1)), torch.from_numpy(f3)


def main():
    ref, Fuse_sft_MLP, PosEncoding = load()
    torch.manual_seed(0)
    rec = {}
    # ---- case "mlp": ragged row count (2*3*5*7 = 210 rows) ----
    seed = 41
    fuse, pe = Fuse_sft_MLP(256 + 1, 256).eval(), PosEncoding(3, N_freqs=7)
    fill(fuse, "fuse_sft_block.", seed)
    R = ref.volume_renderer.VolumeFeatureRenderer(local_rendering_opt(), style_dim=256, out_im_res=8).eval()
    tex = R.network.netLocal.local_feat_to_tex_modulations_linear
    fill(tex, "renderer.network.netLocal.local_feat_to_tex_modulations_linear.", seed)
    shp = (2, 3, 5, 7)
    f2, f3 = synthetic_feats(seed, shp)
    rng = np.random.Generator(np.random.PCG64([seed, 78]))
    pts = torch.from_numpy(rng.uniform(-0.15, 0.15, shp + (3,)).astype(np.float32))
    with torch.no_grad():
        fused = fuse(f2, f3)                                   # e3dge_full_runner.py:289-290
        feats = torch.cat((fused, pe(pts)), -1)                # :293-294
        mods = tex(feats)                                      # volume_renderer.py:329-330
        alpha, beta = torch.split(mods, 256, dim=-1)           # :332-334
    rec.update({"mlp.feat_2d": f2.numpy(), "mlp.feat_3d": f3.numpy(), "mlp.points": pts.numpy(),
                "mlp.feats": feats.numpy(), "mlp.alpha": alpha.numpy(), "mlp.beta": beta.numpy()})
    print("mlp: feats", tuple(feats.shape), "alpha std", alpha.std().item(), "beta std", beta.std().item())

    # ---- case "render": the whole local renderer pass ----
    cfg = dict(size=64, res=8, n_samples=24, batch=2, seed=42, variant="sharp")
    seed = cfg["seed"]
    R = ref.volume_renderer.VolumeFeatureRenderer(local_rendering_opt(N_samples=cfg["n_samples"]), style_dim=256,
                                                  out_im_res=cfg["res"]).eval()
    sd = R.state_dict()
    new = {}
    for k, v in sd.items():
        if "netLocal" in k and "local_feat_to_tex_modulations_linear" not in k:
            new[k] = v  # the hourglass filter is not on this path (feats are given)
        else:
            new[k] = torch.from_numpy(np.ascontiguousarray(
                P.make_param(seed, "renderer." + k, v.shape, cfg["variant"]))).float()
    R.load_state_dict(new, strict=True)
    fill(fuse, "fuse_sft_block.", seed)
    inp = P.make_inputs(seed, cfg["batch"], 2, cfg["res"])
    with torch.no_grad():
        g = R(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=inp["w"])  # global pass -> points
        pts = g["points"]
        f2, f3 = synthetic_feats(seed, tuple(pts.shape[:-1]))
        feats = torch.cat((fuse(f2, f3), pe(pts)), -1)
        out = R(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=inp["w"],
                local_data_batch={"feats": feats})
    assert torch.equal(out["sdf"], g["sdf"])  # texture modulation leaves the geometry alone
    for k in ("features", "gen_thumb_imgs", "sdf", "hit_prob", "xyz", "depth", "points"):
        rec["render." + k] = out[k].numpy()
    rec["render.global_features"] = g["features"].numpy()
    rec["config"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    print("render: features", tuple(out["features"].shape), "local vs global features rel diff",
          ((out["features"] - g["features"]).abs().max() / g["features"].abs().max()).item())
    path = os.path.join(ROOT, "tests", "golden", "local_mlp.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
