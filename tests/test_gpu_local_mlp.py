"""`-m gpu` parity of the local branch's per-sample MLP tail: SFT fusion MLP +
positional encoding + texture-modulation ResnetBlockFC on the tensor cores (e3_local_mlp_fwd), and the
`SirenLocalGlobal` renderer around it, against fixtures recorded from the reference's own modules
(tests/golden/local_mlp.npz, oracle/gen_golden_local_mlp.py) and against the oracle.  Tolerance 1e-3 rel-Linf."""
import numpy as np
import pytest
import torch

from helpers import (load_golden, local_mlp_state_dict, rel_linf, synthetic_local_feats, synthetic_state_dict)
from oracle import local_mlp_oracle as L
from oracle import params as P
from oracle import stylesdf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3
TEX = "renderer.network.netLocal.local_feat_to_tex_modulations_linear."


def _modules(seed, variant="default"):
    from e3dge_b200.local_branch import Fuse_sft_MLP, LocalBranch
    sd = local_mlp_state_dict(seed, variant)
    fuse = Fuse_sft_MLP(257, 256)
    fuse.load_state_dict({k[len("fuse_sft_block."):]: v for k, v in sd.items() if k.startswith("fuse_sft_block.")})
    net = LocalBranch(None)
    net.local_feat_to_tex_modulations_linear.load_state_dict({k[len(TEX):]: v for k, v in sd.items() if k.startswith(TEX)})
    return fuse.cuda().eval(), net.cuda().eval(), sd


def test_local_mlp_vs_reference_fixture():
    from e3dge_b200.local_branch import local_tex_modulation, tex_modulation
    gold, _ = load_golden("local_mlp")
    fuse, net, _ = _modules(41)
    tex = net.local_feat_to_tex_modulations_linear
    f2, f3, pts = (torch.from_numpy(gold["mlp." + k]).cuda() for k in ("feat_2d", "feat_3d", "points"))
    with torch.no_grad():
        alpha, beta, feats = local_tex_modulation(fuse, tex, f2, f3, pts, return_feats=True)
        a2, b2 = tex_modulation(tex, torch.from_numpy(gold["mlp.feats"]).cuda())
        m = tex(torch.from_numpy(gold["mlp.feats"]).cuda())  # the module's forward routes to the same kernel
    assert alpha.shape == gold["mlp.alpha"].shape and feats.shape == gold["mlp.feats"].shape
    assert rel_linf(feats.cpu(), gold["mlp.feats"]) < TOL
    assert rel_linf(alpha.cpu(), gold["mlp.alpha"]) < TOL
    assert rel_linf(beta.cpu(), gold["mlp.beta"]) < TOL
    assert rel_linf(a2.cpu(), gold["mlp.alpha"]) < TOL and rel_linf(b2.cpu(), gold["mlp.beta"]) < TOL
    assert torch.equal(m, torch.cat([a2, b2], -1))


@pytest.mark.parametrize("rows", [1, 127, 129, 1000, 5 * 128 * 3 + 17])
def test_local_mlp_ragged_rows_vs_oracle(rows, monkeypatch):
    from e3dge_b200 import local_branch as lb
    fuse, net, sd = _modules(43)
    tex = net.local_feat_to_tex_modulations_linear
    f2, f3 = synthetic_local_feats(rows, (rows,))
    pts = torch.from_numpy(np.random.Generator(np.random.PCG64(rows)).uniform(-0.15, 0.15, (rows, 3)).astype(np.float32))
    with torch.no_grad():
        ra, rb = L.local_tex_modulation(f2, f3, pts, sd)
        alpha, beta = lb.local_tex_modulation(fuse, tex, f2.cuda(), f3.cuda(), pts.cuda())
        # the same rows walked in chunks of 256 through a small workspace: not a bit changes
        monkeypatch.setattr(lb, "CHUNK_ROWS", 256)
        a_c, b_c = lb.local_tex_modulation(fuse, tex, f2.cuda(), f3.cuda(), pts.cuda())
    assert rel_linf(alpha.cpu(), ra) < TOL and rel_linf(beta.cpu(), rb) < TOL
    assert torch.equal(alpha, a_c) and torch.equal(beta, b_c)


def test_local_mlp_modules_forward_match_the_fused_chain():
    """The modules' own (autograd-capable) forwards and the fused kernel agree; in grad mode with trainable
    parameters the fused entry refuses instead of silently dropping gradients."""
    from e3dge_b200.local_branch import PosEncoding, local_tex_modulation
    fuse, net, _ = _modules(44)
    tex = net.local_feat_to_tex_modulations_linear
    f2, f3 = (t.cuda() for t in synthetic_local_feats(44, (3, 50)))
    pts = (torch.rand(3, 50, 3, device="cuda") - 0.5) * 0.3
    pe = PosEncoding(3, N_freqs=7)
    feats = torch.cat((fuse(f2, f3), pe(pts)), -1)      # grad mode, trainable modules: PyTorch route
    assert feats.requires_grad
    mods = tex(feats)
    with torch.no_grad():
        alpha, beta = local_tex_modulation(fuse, tex, f2, f3, pts)
    assert rel_linf(alpha, mods[..., :256].detach()) < TOL and rel_linf(beta, mods[..., 256:].detach()) < TOL
    with pytest.raises(RuntimeError, match="fused inference path"):
        local_tex_modulation(fuse, tex, f2, f3, pts)


def _local_generator(cfg):
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    seed = cfg["seed"]
    sd = synthetic_state_dict(cfg["size"], cfg["res"], seed, cfg["variant"], local=True)
    lsd = local_mlp_state_dict(seed, cfg["variant"])
    sd.update({k: v for k, v in lsd.items() if k.startswith("renderer.")})
    G = G_pred_latents(model_options(size=cfg["size"], renderer_spatial_output_dim=cfg["res"]),
                       rendering_options(N_samples=cfg["n_samples"], enable_local_model=True,
                                         local_modulation_layer=True, L_pred_tex_modulations=True,
                                         residual_local_feats_dim=301), full_pipeline=True).eval()
    G.load_state_dict(sd, strict=True)  # netGlobal.* and netLocal.* names of the local-branch checkpoints
    return G.cuda(), sd, lsd


def test_local_renderer_vs_reference_fixture():
    from e3dge_b200.local_branch import local_tex_modulation
    gold, cfg = load_golden("local_mlp")
    G, sd, lsd = _local_generator(cfg)
    fuse, _, _ = _modules(cfg["seed"], cfg["variant"])
    R = G.renderer
    tex = R.network.netLocal.local_feat_to_tex_modulations_linear
    inp = {k: v.cuda() for k, v in P.make_inputs(cfg["seed"], cfg["batch"], 2, cfg["res"]).items()}
    with torch.no_grad():
        g = R(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=inp["w"])
        pts = g["points"]
        f2, f3 = (t.cuda() for t in synthetic_local_feats(cfg["seed"], tuple(pts.shape[:-1])))
        alpha, beta, feats = local_tex_modulation(fuse, tex, f2, f3, pts, return_feats=True)
        out = R(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=inp["w"],
                local_data_batch={"feats": feats})
        out2 = R(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=inp["w"],
                 local_data_batch={"tex_modulation": (alpha, beta)})
    assert rel_linf(g["features"].cpu(), gold["render.global_features"]) < TOL
    for k in ("features", "gen_thumb_imgs", "sdf", "hit_prob", "xyz", "depth", "points"):
        assert rel_linf(out[k].cpu(), gold["render." + k]) < TOL, k
    assert torch.equal(out["sdf"], g["sdf"])                    # texture modulation leaves the geometry alone
    assert rel_linf(out2["features"], out["features"]) < 1e-5   # feats -> (alpha, beta) twice: same stages 5-6


def test_siren_local_global_stage_methods_compose_to_the_fused_forward():
    """forward_backbone -> retrieve_feats_for_rendering -> forward_rendering (volume_renderer.py:527-558) against
    the one-kernel `forward`, and both against the oracle's network."""
    gold, cfg = load_golden("local_mlp")
    G, sd, lsd = _local_generator(cfg)
    net = G.renderer.network
    B, n = 2, 333
    g = np.random.Generator(np.random.PCG64(5))
    x = torch.from_numpy(g.uniform(-0.9, 0.9, (B, n, 1, 1, 3)).astype(np.float32))
    v = torch.nn.functional.normalize(torch.from_numpy(g.standard_normal((B, n, 1, 1, 3)).astype(np.float32)), dim=-1)
    feats = torch.from_numpy(g.standard_normal((B, n, 1, 1, 301)).astype(np.float32)) * 0.5
    w = P.make_inputs(cfg["seed"], B, 2, cfg["res"])["w"]
    net_inputs = torch.cat([x, v], -1).cuda()
    with torch.no_grad():
        raw = net(net_inputs, w.cuda(), local_data_batch={"feats": feats.cuda()})
        fo = net.forward_backbone(net_inputs[..., :3], w.cuda(), {"feats": feats.cuda()})
        fr = net.retrieve_feats_for_rendering(fo, sample_mode=False)
        staged = net.forward_rendering(fr, net_inputs[..., 3:], w.cuda())
        glob = net(net_inputs, w.cuda())
        # oracle: the same network on normalised points (dist_radius such that the box warp is the identity)
        gsd = {k.replace("renderer.network.netGlobal.", "renderer.network."): t for k, t in sd.items()}
        mod = L.tex_modulation(feats, lsd)
        ref = O.run_network(x, v, w, gsd, dist_radius=1.0, local_mod=mod)
        ref_glob = O.run_network(x, v, w, gsd, dist_radius=1.0)
    assert raw.shape == (B, n, 1, 1, 260)
    for sl, name in ((slice(0, 3), "rgb"), (slice(3, 4), "sdf"), (slice(4, 260), "features")):
        assert rel_linf(raw[..., sl].cpu(), ref[..., sl]) < TOL, name
        assert rel_linf(staged[..., sl].cpu(), ref[..., sl]) < TOL, name
        assert rel_linf(glob[..., sl].cpu(), ref_glob[..., sl]) < TOL, name
    assert torch.equal(raw[..., 3], glob[..., 3])


def test_local_branch_full_size_image_vs_oracle():
    """One 64 x 64 x 24 image (98 304 samples) through the fused tail, every value against the oracle."""
    from e3dge_b200.local_branch import local_tex_modulation
    fuse, net, sd = _modules(45)
    tex = net.local_feat_to_tex_modulations_linear
    shp = (1, 64, 64, 24)
    f2, f3 = synthetic_local_feats(45, shp)
    pts = torch.from_numpy(np.random.Generator(np.random.PCG64(45)).uniform(-0.15, 0.15, shp + (3,)).astype(np.float32))
    torch.set_num_threads(max(1, min(16, torch.get_num_threads())))
    with torch.no_grad():
        alpha, beta = local_tex_modulation(fuse, tex, f2.cuda(), f3.cuda(), pts.cuda())
        ra, rb = L.local_tex_modulation(f2, f3, pts, sd)
    assert rel_linf(alpha.cpu(), ra) < TOL and rel_linf(beta.cpu(), rb) < TOL
