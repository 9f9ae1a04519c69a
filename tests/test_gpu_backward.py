"""`-m gpu` parity of the backward kernels (e3_render_bwd, e3_siren_points_bwd, e3_film_bwd and
the decoder backward), through the Python surface that binds the C ABI, against

  (a) gradients recorded from autograd through the REAL reference (tests/golden/small_grad.npz,
      written by oracle/gen_golden.py::run_grad_case), and
  (b) autograd through the float64 oracle on the same seeded inputs.

Tolerance: 1e-3 rel-Linf of each gradient tensor (BASELINE.json north_star), fp32."""
import numpy as np
import pytest
import torch

from helpers import decoder_layout, load_golden, rel_linf, synthetic_state_dict
from oracle import params as P
from oracle import stylesdf_oracle as O
from test_oracle_golden import GRAD_KEYS, cotangent

pytestmark = pytest.mark.gpu
TOL = 1e-3
RENDER_KEYS = [k for k in GRAD_KEYS if k != "gen_imgs"]


def _build(size, res, seed, variant, n_samples=24, full_pipeline=True, **ropt):
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    sd = synthetic_state_dict(size, res, seed, variant)
    G = G_pred_latents(model_options(size=size, renderer_spatial_output_dim=res),
                       rendering_options(N_samples=n_samples, **ropt),
                       full_pipeline=full_pipeline).eval()
    G.load_state_dict(sd, strict=full_pipeline)
    for p in G.parameters():
        p.requires_grad_(False)
    return G.cuda(), sd


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def test_renderer_gradients_vs_reference_fixture():
    gold, cfg = load_golden("small_grad")
    G, sd = _build(cfg["size"], cfg["res"], cfg["seed"], cfg["variant"], cfg["n_samples"])
    inp = _cuda(P.make_inputs(cfg["seed"], cfg["batch"], decoder_layout(cfg["size"], cfg["res"]),
                              cfg["res"], wplus=True))
    w = inp["w"].clone().requires_grad_(True)
    out = G.renderer(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=w,
                     return_eikonal=True)
    worst = {}
    for k in RENDER_KEYS:
        loss = (cotangent(k, tuple(out[k].shape)).cuda() * out[k]).sum()
        gw, = torch.autograd.grad(loss, [w], retain_graph=True)
        worst[k] = rel_linf(gw.cpu(), gold["dw." + k])
    worst["eikonal_term"] = rel_linf(out["eikonal_term"].cpu(), gold["eikonal_term"])
    bad = {k: v for k, v in worst.items() if v >= TOL}
    assert not bad, f"{bad} (all: {worst})"


@pytest.mark.parametrize("case", ["wplus_s24", "w_s7_nfb", "wplus_s18_rayd_local"])
def test_renderer_gradients_vs_oracle_autograd(case):
    res, seed, B = 16, 71, 2
    S = {"wplus_s24": 24, "w_s7_nfb": 7, "wplus_s18_rayd_local": 18}[case]
    wplus = case != "w_s7_nfb"
    ropt = {}
    if case == "w_s7_nfb":
        ropt = dict(force_background=False)
    if case == "wplus_s18_rayd_local":
        ropt = dict(static_viewdirs=False)
    G, sd = _build(64, res, seed, "sharp", S, full_pipeline=False, **ropt)
    inp = P.make_inputs(seed, B, 1, res, wplus=wplus)
    rng = np.random.Generator(np.random.PCG64(seed))
    local = None
    if case == "wplus_s18_rayd_local":
        shp = (B, res, res, S, 256)
        local = [torch.from_numpy(rng.standard_normal(shp).astype(np.float32) * 0.3) for _ in range(2)]

    # float64 oracle autograd
    sd64 = O.cast_state_dict(sd, torch.float64)
    w64 = inp["w"].double().requires_grad_(True)
    l64 = [t.double().requires_grad_(True) for t in local] if local else None
    ref = O.renderer_forward(sd64, inp["cam_poses"].double(), inp["focal"].double(), inp["near"].double(),
                             inp["far"].double(), w64, res=res, n_samples=S,
                             static_viewdirs=ropt.get("static_viewdirs", True),
                             force_background=ropt.get("force_background", True), local_mod=l64)
    d = _cuda(inp)
    w = d["w"].clone().requires_grad_(True)
    lc = [t.cuda().requires_grad_(True) for t in local] if local else None
    out = G.renderer(d["cam_poses"], d["focal"], d["near"], d["far"], styles=w,
                     local_tex_modulation=tuple(lc) if lc else None)
    worst = {}
    for k in RENDER_KEYS:
        ct = cotangent(k, tuple(out[k].shape))
        ins64 = [w64] + (l64 or [])
        g64 = torch.autograd.grad((ct.double() * ref[k]).sum(), ins64, retain_graph=True, allow_unused=True)
        ins = [w] + (lc or [])
        g = torch.autograd.grad((ct.cuda() * out[k]).sum(), ins, retain_graph=True, allow_unused=True)
        for name, a, b in zip(["w", "alpha", "beta"], g, g64):
            if b is None or b.abs().max() == 0:
                assert a is None or a.abs().max().item() == 0, (k, name)
                continue
            worst[f"{k}.{name}"] = rel_linf(a.cpu(), b)
    bad = {k: v for k, v in worst.items() if v >= TOL}
    assert not bad, f"{case}: {bad} (all: {worst})"


def test_point_query_gradients_and_eikonal_vs_oracle_autograd():
    seed, B, N = 73, 2, 333  # ragged: 333 = 2*128 + 77
    G, sd = _build(64, 16, seed, "sharp", full_pipeline=False)
    R = G.renderer
    inp = P.make_inputs(seed, B, 1, 16, wplus=True)
    rng = np.random.Generator(np.random.PCG64(seed))
    pts = torch.from_numpy(rng.uniform(-0.12, 0.12, (B, N, 3)).astype(np.float32))
    vd = torch.from_numpy(rng.standard_normal((B, N, 3)).astype(np.float32))
    vd = vd / vd.norm(dim=-1, keepdim=True)
    sd64 = O.cast_state_dict(sd, torch.float64)
    w64 = inp["w"].double().requires_grad_(True)
    p64 = pts.double().requires_grad_(True)
    raw64 = O.run_network(p64.reshape(B, N, 1, 1, 3), vd.double().reshape(B, N, 1, 1, 3), w64, sd64)
    w = inp["w"].cuda().requires_grad_(True)
    p = pts.cuda().requires_grad_(True)
    raw = R.run_network(p.reshape(B, N, 1, 1, 3), vd.cuda().reshape(B, N, 1, 1, 3), styles=w)
    assert rel_linf(raw.detach().cpu(), raw64.detach()) < TOL
    ct = torch.from_numpy(np.cos(np.arange(raw64.numel()) * 0.37).reshape(raw64.shape))
    # rgb | sdf | features cotangents separately and together
    for name, sl in {"rgb": slice(0, 3), "sdf": slice(3, 4), "feat": slice(4, 260), "all": slice(0, 260)}.items():
        g64 = torch.autograd.grad((ct[..., sl] * raw64[..., sl]).sum(), [w64, p64], retain_graph=True)
        g = torch.autograd.grad((ct[..., sl].float().cuda() * raw[..., sl]).sum(), [w, p], retain_graph=True)
        assert rel_linf(g[0].cpu(), g64[0]) < TOL, name
        assert rel_linf(g[1].cpu(), g64[1]) < TOL, name
    # sdf-only graph (7 GEMMs) + the eikonal seed
    s64 = O.sdf_query(sd64, p64, w64)
    g64 = torch.autograd.grad(s64.sum(), [w64, p64])
    s = R.sdf_query(p, w)
    g = torch.autograd.grad(s.sum(), [w, p])
    assert rel_linf(g[0].cpu(), g64[0]) < TOL and rel_linf(g[1].cpu(), g64[1]) < TOL
    sdf_v, grad_v = R.sdf_and_gradient(pts.cuda(), inp["w"].cuda())
    assert rel_linf(sdf_v.cpu(), s64.detach()) < TOL and rel_linf(grad_v.cpu(), g64[1]) < TOL
    # w-space styles share one latent across the nine layers
    w1 = inp["w"][:, 0].double().requires_grad_(True)
    g64 = torch.autograd.grad(O.sdf_query(sd64, pts.double(), w1).sum(), [w1])
    w1c = inp["w"][:, 0].cuda().requires_grad_(True)
    g = torch.autograd.grad(R.sdf_query(pts.cuda(), w1c).sum(), [w1c])
    assert rel_linf(g[0].cpu(), g64[0]) < TOL


def test_renderer_backward_is_deterministic_and_batch_independent_at_full_size():
    """Size-independent properties at the BASELINE shape (64x64 rays x 24 samples): bit-exact
    repeatability (no atomics) and independence of each image's gradient from its batch."""
    G, sd = _build(256, 64, 81, "sharp", full_pipeline=False)
    inp = _cuda(P.make_inputs(81, 3, 1, 64, wplus=True))
    ct = torch.cos(torch.arange(3 * 256 * 64 * 64, device="cuda") * 0.37).reshape(3, 256, 64, 64)

    def grad(sel):
        w = inp["w"][sel].clone().requires_grad_(True)
        out = G.renderer(inp["cam_poses"][sel], inp["focal"][sel], inp["near"][sel], inp["far"][sel], styles=w)
        loss = (ct[sel] * out["features"]).sum() + out["gen_thumb_imgs"].sum() + (out["depth"] ** 2).sum()
        return torch.autograd.grad(loss, [w])[0]

    g_all = grad(slice(0, 3))
    assert torch.isfinite(g_all).all() and g_all.abs().max() > 0
    assert torch.equal(g_all, grad(slice(0, 3)))
    g1 = grad(slice(1, 2))
    assert rel_linf(g1, g_all[1:2]) < 1e-5
