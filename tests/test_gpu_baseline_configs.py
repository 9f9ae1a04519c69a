"""`-m gpu` parity at the shapes of BASELINE.json configs[2] and configs[4] (configs[0]/[1] are
covered by test_gpu_parity.py, configs[3] — the training step — by test_gpu_backward.py):

  configs[2]  novel-view sweep (demo_view_synthesis shapes): full 64x64x24 renders along the
              `create_trajectory` azimuth sweep (trainer.py:2356-2369), global-only AND with the local
              branch's texture modulation, 4 (identity, view) pairs per GPU;
  configs[4]  512^2 high-res render with 48 samples per ray, and the 128^3 SDF grid of the surface
              extractor (train_setup.py:118-119).

At these sizes the CPU oracle cannot run the whole tensor in seconds, so each test checks the oracle
on a sub-sample (one image of the batch / a strided set of rays / a random subset of grid points)
plus size-independent properties of the full result.  Tolerance 1e-3 rel-Linf."""
import math

import numpy as np
import pytest
import torch

from helpers import decoder_layout, rel_linf, synthetic_state_dict
from oracle import params as P
from oracle import stylesdf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _build(size, res, seed, n_samples, full_pipeline=True, **ropt):
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    sd = synthetic_state_dict(size, 64, seed, "sharp")
    G = G_pred_latents(model_options(size=size, renderer_spatial_output_dim=res),
                       rendering_options(N_samples=n_samples, **ropt), full_pipeline=full_pipeline).eval()
    G.load_state_dict(sd, strict=False)
    return G.cuda(), sd


def _trajectory_poses(n, res, fov_deg=6.0):
    """create_trajectory: azim = 1.5 * 0.3 * cos(pi t), t = linspace(0, 1, n), elev 0 (trainer.py:2356-2369)."""
    t = np.linspace(0.0, 1.0, n)
    azim = 1.5 * 0.3 * np.cos(np.pi * t)
    cam_dir = np.stack([np.sin(azim), np.zeros(n), np.cos(azim)], 1)
    up = np.tile(np.array([[0.0, 1.0, 0.0]]), (n, 1))
    z_axis = cam_dir / np.linalg.norm(cam_dir, axis=1, keepdims=True)
    x_axis = np.cross(up, z_axis)
    x_axis /= np.linalg.norm(x_axis, axis=1, keepdims=True)
    y_axis = np.cross(z_axis, x_axis)
    poses = np.concatenate([np.stack([x_axis, y_axis, z_axis], 2), cam_dir[:, :, None]], 2)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()
    return dict(cam_poses=f32(poses), focal=f32(np.full((n, 1, 1), 0.5 * res / math.tan(math.radians(fov_deg)))),
                near=f32(np.full((n, 1, 1), 0.88)), far=f32(np.full((n, 1, 1), 1.12)))


@pytest.mark.parametrize("local", [False, True])
def test_config2_novel_view_sweep(local):
    size, res, S, B, seed = 256, 64, 24, 4, 301
    G, sd = _build(size, res, seed, S)
    lat = P.make_inputs(seed, 1, decoder_layout(size, res), res)
    cams = _trajectory_poses(9, res)
    pick = [0, 3, 5, 8]  # 4 of the 9 sweep frames of one identity on this GPU
    inp = {k: v[pick] for k, v in cams.items()}
    inp["w"] = lat["w"].expand(B, -1, -1).contiguous()
    inp["w_dec"] = lat["w_dec"].expand(B, -1, -1).contiguous()
    mod = None
    if local:
        rng = np.random.Generator(np.random.PCG64(seed))
        shp = (B, res, res, S, 256)
        mod = tuple(torch.from_numpy(rng.standard_normal(shp).astype(np.float32) * 0.3) for _ in range(2))
    d = {k: v.cuda() for k, v in inp.items()}
    kw = dict(local_tex_modulation=tuple(t.cuda() for t in mod)) if local else {}
    with torch.no_grad():
        out = G([d["w"], d["w_dec"]], d["cam_poses"], d["focal"], d["near"], d["far"], input_is_latent=True,
                randomize_noise=False, return_xyz=True, return_sdf=True, **kw)
    # properties of the whole batch
    assert (out["hit_prob"].sum(3) - 1).abs().max().item() < 1e-5
    assert torch.isfinite(out["gen_imgs"]).all() and out["gen_imgs"].shape == (B, 3, size, size)
    # same identity from different views: the sdf at a fixed world point does not depend on the camera
    pts = torch.from_numpy(np.random.Generator(np.random.PCG64(7)).uniform(-0.1, 0.1, (B, 500, 3)).astype(np.float32))
    pts[:] = pts[0]
    q = G.renderer.sdf_query(pts.cuda(), d["w"])
    assert torch.equal(q[0], q[3])
    # one frame of the sweep against the oracle, renderer and decoder
    i = 2
    one = {k: v[i:i + 1] for k, v in inp.items()}
    with torch.no_grad():
        ref = O.generator_forward(sd, one["w"], one["w_dec"], one["cam_poses"], one["focal"], one["near"],
                                  one["far"], res=res, n_samples=S,
                                  local_mod=tuple(t[i:i + 1] for t in mod) if local else None)
    for k in ("features", "gen_thumb_imgs", "sdf", "hit_prob", "xyz", "depth", "gen_imgs"):
        assert rel_linf(out[k][i:i + 1].cpu(), ref[k]) < TOL, k


def test_config4_highres_512_render_48_samples():
    res, S, seed = 512, 48, 302
    G, sd = _build(1024, res, seed, S, full_pipeline=False)
    inp = P.make_inputs(seed, 1, 1, res)
    d = {k: v.cuda() for k, v in inp.items()}
    with torch.no_grad():
        out = G.renderer(d["cam_poses"], d["focal"], d["near"], d["far"], styles=d["w"])
    assert out["features"].shape == (1, 256, res, res) and out["sdf"].shape == (1, res, res, S, 1)
    assert (out["hit_prob"].sum(3) - 1).abs().max().item() < 1e-5
    assert out["depth"].min().item() >= 0.88 - 1e-5 and out["depth"].max().item() <= 1.12 + 1e-5
    assert torch.isfinite(out["features"]).all()
    # a strided 8x8 set of rays through the oracle's network + composite, on the kernel's own samples
    st = 64
    pts = out["points"][:, ::st, ::st].cpu()
    vd = out["viewdirs"][:, ::st, ::st].cpu()
    rays_d = out["rays_d"][:, ::st, ::st].cpu()
    ones = torch.ones_like(rays_d[..., :1])
    z = O.sample_z(inp["near"].unsqueeze(-1) * ones, inp["far"].unsqueeze(-1) * ones, S)
    with torch.no_grad():
        raw = O.run_network(pts, vd, inp["w"], sd)
        vi = O.volume_integration(raw, z, rays_d, pts, sd["renderer.sigmoid_beta"])
    assert rel_linf(out["sdf"][:, ::st, ::st].cpu(), vi["sdf"]) < TOL
    assert rel_linf(out["hit_prob"][:, ::st, ::st].cpu(), vi["weights"]) < TOL
    assert rel_linf(out["features"][:, :, ::st, ::st].cpu(), vi["feature_map"].permute(0, 3, 1, 2)) < TOL
    assert rel_linf(out["gen_thumb_imgs"][:, :, ::st, ::st].cpu(), vi["rgb_map"].permute(0, 3, 1, 2)) < TOL


def test_config4_sdf_grid_128_cubed():
    seed, n = 303, 128
    G, sd = _build(256, 64, seed, 24, full_pipeline=False)
    R = G.renderer
    w = P.make_inputs(seed, 1, 1, 64)["w"]
    lin = torch.linspace(-0.12, 0.12, n)
    grid = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(1, -1, 3)  # 2 097 152 points
    gc = grid.cuda()
    with torch.no_grad():
        sdf = R.sdf_query(gc, w.cuda())
        half = grid.shape[1] // 2
        a = R.sdf_query(gc[:, :half].contiguous(), w.cuda())
        b = R.sdf_query(gc[:, half:].contiguous(), w.cuda())
    assert sdf.shape == (1, n ** 3, 1) and torch.isfinite(sdf).all()
    assert torch.equal(sdf, torch.cat([a, b], 1))  # chunking the grid does not change a bit
    idx = torch.from_numpy(np.random.Generator(np.random.PCG64(seed)).choice(n ** 3, 4096, replace=False))
    with torch.no_grad():
        ref = O.sdf_query(sd, grid[:, idx], w)
    assert rel_linf(sdf[:, idx.cuda()].cpu(), ref) < TOL


def test_config1_the_benched_batch_all_frames_full_tensors_through_the_graph():
    """bench.py's own generator, its own seeded B=8 batch, through the CUDA-graph replay it times
    (e3dge_b200.graphed.GraphedCall), every frame and every element against the oracle."""
    import bench
    from e3dge_b200.graphed import GraphedCall
    dev = torch.device("cuda", torch.cuda.current_device())
    G, sd = bench.build_generator(dev)
    inp = bench.make_inputs(0)
    static = {k: v.to(dev) for k, v in inp.items()}
    assert static["w"].shape[0] == bench.BATCH == 8

    def core():
        with torch.no_grad():
            return G([static["w"], static["w_dec"]], static["cam_poses"], static["focal"], static["near"],
                     static["far"], input_is_latent=True, randomize_noise=False, return_xyz=True, return_sdf=True)
    call = GraphedCall(core)
    out = call()
    torch.cuda.synchronize()
    got = {k: out[k].cpu() for k in ("features", "gen_thumb_imgs", "sdf", "hit_prob", "xyz", "depth", "gen_imgs")}
    torch.set_num_threads(max(1, min(16, torch.get_num_threads())))
    with torch.no_grad():
        ref = O.generator_forward(sd, inp["w"], inp["w_dec"], inp["cam_poses"], inp["focal"], inp["near"],
                                  inp["far"], res=bench.RES, n_samples=bench.N_SAMPLES)
    for k, t in got.items():
        assert t.shape == ref[k].shape, k
        for b in range(bench.BATCH):  # per frame: a frame with small values cannot hide behind a bright one
            err = rel_linf(t[b], ref[k][b])
            assert err < TOL, f"frame {b} {k}: rel-Linf {err:.3e}"
    # the replay is the eager pass, bit for bit
    eager = core()
    for k in got:
        assert torch.equal(eager[k], out[k]), k


def test_size_1024_generator_full_frame():
    """The shipped scripts run --size 1024 (demo_view_synthesis.sh:35-36): 64^2 features -> 4 up-sampling stages
    down to 32 channels at 1024^2 (stylesdf_model.py:614-624).  One full frame, every pixel, against the oracle."""
    size, res, S, seed = 1024, 64, 24, 304
    G, sd = _build(size, res, seed, S)
    sd = synthetic_state_dict(size, res, seed, "sharp")
    G.load_state_dict(sd, strict=True)
    inp = P.make_inputs(seed, 1, decoder_layout(size, res), res)
    d = {k: v.cuda() for k, v in inp.items()}
    with torch.no_grad():
        out = G([d["w"], d["w_dec"]], d["cam_poses"], d["focal"], d["near"], d["far"], input_is_latent=True,
                randomize_noise=False, return_xyz=True, return_sdf=True)
        ref = O.generator_forward(sd, inp["w"], inp["w_dec"], inp["cam_poses"], inp["focal"], inp["near"],
                                  inp["far"], res=res, n_samples=S)
    assert out["gen_imgs"].shape == (1, 3, size, size)
    for k in ("features", "gen_thumb_imgs", "sdf", "hit_prob", "gen_imgs"):
        err = rel_linf(out[k].cpu(), ref[k])
        assert err < TOL, f"{k}: rel-Linf {err:.3e}"


def test_perturbed_samples_match_the_oracle_on_the_same_depths():
    """perturb > 0 (training, volume_renderer.py:1213-1228): the jittered depths are drawn on the device and
    handed to the kernel explicitly; the oracle's network + composite on those very depths must agree."""
    size, res, S, seed = 64, 16, 24, 305
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    sd = synthetic_state_dict(size, res, seed, "sharp")
    G = G_pred_latents(model_options(size=size, renderer_spatial_output_dim=res),
                       rendering_options(N_samples=S, perturb=1.0), full_pipeline=False).eval()
    G.load_state_dict(sd, strict=False)
    G = G.cuda()
    R = G.renderer
    R.test, R.perturb = False, 1.0
    inp = P.make_inputs(seed, 2, 1, res)
    d = {k: v.cuda() for k, v in inp.items()}
    torch.manual_seed(seed)
    zj = R._make_z_jitter(d["near"], d["far"], 2, d["w"].device)
    with torch.no_grad():
        o = R._render_raw(d["w"], d["cam_poses"], d["focal"], d["near"], d["far"], z_jitter=zj)
    z = zj.cpu()
    # stratified offsets stay inside their own bins, one offset per ray with offset sampling
    near, far = inp["near"].reshape(-1, 1, 1, 1), inp["far"].reshape(-1, 1, 1, 1)
    step = (far - near) / S
    assert (z[..., 1:] > z[..., :-1]).all() and (z >= near).all() and (z <= far).all()
    assert ((z[..., 1:] - z[..., :-1]) - step).abs().max().item() < 1e-6
    rays_o, rays_d, viewdirs = O.get_rays(inp["focal"], inp["cam_poses"], res)
    viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
    pts = rays_o.unsqueeze(3) + rays_d.unsqueeze(3) * z.unsqueeze(-1)
    with torch.no_grad():
        raw = O.run_network(pts, viewdirs, inp["w"], sd)
        vi = O.volume_integration(raw, z, rays_d, pts, sd["renderer.sigmoid_beta"])
    assert rel_linf(o["points"].cpu(), pts) < 1e-5
    assert rel_linf(o["sdf"].cpu(), vi["sdf"]) < TOL
    assert rel_linf(o["hit_prob"].cpu(), vi["weights"]) < TOL
    assert rel_linf(o["dists"].cpu(), vi["dists"]) < TOL
    assert rel_linf(o["features"].cpu(), vi["feature_map"].permute(0, 3, 1, 2)) < TOL
    assert rel_linf(o["gen_thumb_imgs"].cpu(), vi["rgb_map"].permute(0, 3, 1, 2)) < TOL
    assert rel_linf(o["xyz"].cpu(), vi["xyz"].permute(0, 3, 1, 2)) < TOL
