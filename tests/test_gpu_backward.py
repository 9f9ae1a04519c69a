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


# --------------------------------------------------------------------------------------
# decoder backward
# --------------------------------------------------------------------------------------
def _decoder_acts(G, features, latent):
    """The StyledConv outputs of Decoder.forward (NCHW), layer by layer through the same module
    calls the decoder makes."""
    from e3dge_b200.stylesdf_model import _nchw, _nhwc
    dec = G.decoder
    noise = [getattr(dec.noises, f"noise_{i}") for i in range(dec.num_layers)]
    acts = []
    with torch.no_grad():
        out = dec.conv1.forward_nhwc(_nhwc(features), latent[:, 0], noise[0])
        acts.append(_nchw(out))
        i = 1
        for c1, c2, n1, n2 in zip(dec.convs[::2], dec.convs[1::2], noise[1::2], noise[2::2]):
            out = c1.forward_nhwc(out, latent[:, i], n1)
            acts.append(_nchw(out))
            out = c2.forward_nhwc(out, latent[:, i + 1], n2)
            acts.append(_nchw(out))
            i += 2
    return acts


def _check_gates(acts_cuda, acts_ref):
    """Signs may differ only where the reference activation sits on the kink (|y| tiny)."""
    for a, r in zip(acts_cuda, acts_ref):
        mism = (a.cpu() > 0) != (r > 0)
        if mism.any():
            assert r[mism].abs().max().item() < 1e-4 * r.abs().max().item()


def test_generator_gradients_vs_reference_fixture():
    """Image-space cotangent through decoder AND renderer: d<R, gen_imgs>/d w+ and /d decoder
    latent.  The leaky-ReLU kinks make this gradient discontinuous: elements whose pre-activation
    is zero within float32 round-off take either slope depending on the summation order, and the
    fixture (reference fp32 on CPU) differs from the float64 oracle by 4e-3 / 6e-3 for exactly
    that reason (measured, scratch study recorded in DESIGN.md).  So: (1) against the fixture a
    bound above that ambiguity; (2) the tight 1e-3 bound against the float64 oracle evaluated
    with the branch decisions of the CUDA forward, after checking that those decisions differ
    from the oracle's own only on the kink."""
    gold, cfg = load_golden("small_grad")
    G, sd = _build(cfg["size"], cfg["res"], cfg["seed"], cfg["variant"], cfg["n_samples"])
    inp0 = P.make_inputs(cfg["seed"], cfg["batch"], decoder_layout(cfg["size"], cfg["res"]), cfg["res"],
                         wplus=True)
    inp = _cuda(inp0)
    w = inp["w"].clone().requires_grad_(True)
    wd = inp["w_dec"].clone().requires_grad_(True)
    out = G([w, wd], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], input_is_latent=True,
            randomize_noise=False, return_xyz=True, return_sdf=True)
    ct = cotangent("gen_imgs", tuple(out["gen_imgs"].shape))
    gw, gd = torch.autograd.grad((ct.cuda() * out["gen_imgs"]).sum(), [w, wd])
    e_w, e_d = rel_linf(gw.cpu(), gold["dw.gen_imgs"]), rel_linf(gd.cpu(), gold["dwdec.gen_imgs"])
    assert e_w < 5e-2 and e_d < 5e-2, (e_w, e_d)

    acts = _decoder_acts(G, out["features"].detach(), inp["w_dec"])
    sd64 = O.cast_state_dict(sd, torch.float64)
    w64 = inp0["w"].double().requires_grad_(True)
    wd64 = inp0["w_dec"].double().requires_grad_(True)
    r64 = O.renderer_forward(sd64, inp0["cam_poses"].double(), inp0["focal"].double(), inp0["near"].double(),
                             inp0["far"].double(), w64, res=cfg["res"], n_samples=cfg["n_samples"])
    ref_acts = []
    img64 = O.decoder_forward(sd64, r64["features"], wd64, gates=[a.cpu() > 0 for a in acts], acts=ref_acts)
    _check_gates(acts, [t.detach() for t in ref_acts])
    assert rel_linf(out["gen_imgs"].detach().cpu(), img64.detach()) < TOL
    g64 = torch.autograd.grad((ct.double() * img64).sum(), [w64, wd64])
    e_w, e_d = rel_linf(gw.cpu(), g64[0]), rel_linf(gd.cpu(), g64[1])
    assert e_w < TOL and e_d < TOL, (e_w, e_d)


def _module_sd(mod, prefix):
    return {prefix + k: v.detach().cpu().double() for k, v in mod.state_dict().items()}


@pytest.mark.parametrize("cin,cout,hw,batch,up,backend", [
    (64, 128, 8, 3, False, "auto"), (128, 128, 16, 2, True, "auto"), (256, 128, 8, 5, True, "auto"),
    (16, 32, 10, 2, False, "auto"), (64, 128, 8, 2, False, "fp32"),
    (32, 16, 6, 2, True, "auto"), (128, 128, 8, 2, True, "fp32"),      # up-conv backward on the CUDA cores
    (64, 32, 32, 1, True, "auto")])                                    # the 64 -> 32 tail of a size-1024 decoder
def test_styled_conv_backward_vs_oracle_autograd(cin, cout, hw, batch, up, backend):
    from e3dge_b200.stylesdf_model import StyledConv
    torch.manual_seed(5)
    m = StyledConv(cin, cout, 3, 512, upsample=up).cuda()
    with torch.no_grad():
        m.noise.weight.fill_(0.3)
        m.activate.bias.normal_(0, 0.2)
    for p in m.parameters():
        p.requires_grad_(False)
    m.conv.backend = backend
    x = torch.randn(batch, cin, hw, hw)
    st = torch.randn(batch, 512)
    oh = 2 * hw if up else hw
    noise = torch.randn(batch, 1, oh, oh)
    ct = torch.cos(torch.arange(batch * cout * oh * oh) * 0.37).reshape(batch, cout, oh, oh)
    sd = _module_sd(m, "k.")
    x64, s64 = x.double().requires_grad_(True), st.double().requires_grad_(True)
    xc, sc = x.cuda().requires_grad_(True), st.cuda().requires_grad_(True)
    y = m(xc, sc, noise=noise.cuda())
    # the oracle takes the leaky-ReLU branch decisions of the CUDA forward (see fused_leaky_relu)
    ref = O.styled_conv(x64, s64, noise.double(), sd, "k.", upsample=up, gate=y.detach().cpu() > 0)
    _check_gates([y.detach()], [O.styled_conv(x64, s64, noise.double(), sd, "k.", upsample=up).detach()])
    g64 = torch.autograd.grad((ct.double() * ref).sum(), [x64, s64])
    assert rel_linf(y.detach().cpu(), ref.detach()) < TOL
    g = torch.autograd.grad((ct.cuda() * y).sum(), [xc, sc])
    e = (rel_linf(g[0].cpu(), g64[0]), rel_linf(g[1].cpu(), g64[1]))
    assert max(e) < TOL, e


def test_torgb_and_bare_modconv_backward_vs_oracle_autograd():
    from e3dge_b200.stylesdf_model import ModulatedConv2d, ToRGB
    torch.manual_seed(6)
    B, cin, hw = 3, 64, 16
    m = ToRGB(cin, 512, upsample=True).cuda()
    with torch.no_grad():
        m.bias.normal_(0, 0.1)
    for p in m.parameters():
        p.requires_grad_(False)
    x, st, skip = torch.randn(B, cin, hw, hw), torch.randn(B, 512), torch.randn(B, 3, hw // 2, hw // 2)
    ct = torch.cos(torch.arange(B * 3 * hw * hw) * 0.37).reshape(B, 3, hw, hw)
    sd = _module_sd(m, "k.")
    ins64 = [t.double().requires_grad_(True) for t in (x, st, skip)]
    ref = O.to_rgb(ins64[0], ins64[1], ins64[2], sd, "k.", upsample=True)
    g64 = torch.autograd.grad((ct.double() * ref).sum(), ins64)
    ins = [t.cuda().requires_grad_(True) for t in (x, st, skip)]
    y = m(*ins)
    g = torch.autograd.grad((ct.cuda() * y).sum(), ins)
    e = [rel_linf(a.cpu(), b) for a, b in zip(g, g64)]
    assert max(e) < TOL, e
    # bare modulated conv (no noise / activation), demodulated
    mc = ModulatedConv2d(64, 128, 3, 512).cuda()
    for p in mc.parameters():
        p.requires_grad_(False)
    sd = _module_sd(mc, "k.")
    x, st = torch.randn(2, 64, 8, 8), torch.randn(2, 512)
    ct = torch.cos(torch.arange(2 * 128 * 64) * 0.37).reshape(2, 128, 8, 8)
    ins64 = [t.double().requires_grad_(True) for t in (x, st)]
    g64 = torch.autograd.grad((ct.double() * O.modulated_conv2d(ins64[0], ins64[1], sd, "k.")).sum(), ins64)
    ins = [t.cuda().requires_grad_(True) for t in (x, st)]
    g = torch.autograd.grad((ct.cuda() * mc(*ins)).sum(), ins)
    e = [rel_linf(a.cpu(), b) for a, b in zip(g, g64)]
    assert max(e) < TOL, e


def test_full_size_generator_backward_runs_and_is_deterministic():
    """BASELINE shape (size 256, 64x64x24): one full backward, finite, non-zero, bit-repeatable."""
    G, sd = _build(256, 64, 91, "sharp")
    inp = _cuda(P.make_inputs(91, 2, decoder_layout(256, 64), 64, wplus=True))

    def grads():
        w = inp["w"].clone().requires_grad_(True)
        wd = inp["w_dec"].clone().requires_grad_(True)
        out = G([w, wd], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], input_is_latent=True,
                randomize_noise=False)
        loss = (out["gen_imgs"] ** 2).mean() + (out["gen_thumb_imgs"] ** 2).mean()
        return torch.autograd.grad(loss, [w, wd])

    a, b = grads(), grads()
    for t, u in zip(a, b):
        assert torch.isfinite(t).all() and t.abs().max() > 0
        assert torch.equal(t, u)


def test_local_branch_feats_hook_trains_the_callers_netlocal():
    """`local_data_batch={'feats': ...}` with a caller-supplied netLocal (SirenLocalGlobal.forward_backbone,
    volume_renderer.py:323-336): same result as handing (alpha, beta) explicitly, and the image-space loss
    reaches netLocal's parameters through d_local_alpha / d_local_beta (trainer.py:1611 trains them)."""
    res, S, B, seed = 8, 12, 2, 95
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    G = G_pred_latents(model_options(size=64, renderer_spatial_output_dim=res),
                       rendering_options(N_samples=S, enable_local_model=True, L_pred_tex_modulations=True),
                       full_pipeline=False).eval().cuda()
    for p in G.parameters():
        p.requires_grad_(False)

    class NetLocal(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.local_feat_to_tex_modulations_linear = torch.nn.Linear(301, 512)

    torch.manual_seed(seed)
    net = NetLocal().cuda()
    G.renderer.network.netLocal = net
    inp = _cuda(P.make_inputs(seed, B, 1, res, wplus=True))
    feats = torch.randn(B, res, res, S, 301, device="cuda") * 0.2
    out = G.renderer(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=inp["w"],
                     local_data_batch={"feats": feats})
    with torch.no_grad():
        mods = net.local_feat_to_tex_modulations_linear(feats)
        ref = G.renderer(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=inp["w"],
                         local_tex_modulation=tuple(torch.split(mods, 256, dim=-1)))
    assert torch.equal(out["features"].detach(), ref["features"])
    loss = (out["features"] ** 2).mean() + out["gen_thumb_imgs"].mean()
    gW, gb = torch.autograd.grad(loss, [net.local_feat_to_tex_modulations_linear.weight,
                                        net.local_feat_to_tex_modulations_linear.bias])
    assert torch.isfinite(gW).all() and gW.abs().max() > 0 and gb.abs().max() > 0
    # sdf does not depend on the texture modulation (volume_renderer.py:206-220)
    assert torch.equal(out["sdf"].detach(), G.renderer(inp["cam_poses"], inp["focal"], inp["near"], inp["far"],
                                                       styles=inp["w"])["sdf"])


def test_eikonal_term_is_differentiable_to_the_latents_second_order():
    """`get_eikonal_term` = autograd.grad(sdf, pts, create_graph=True) (volume_renderer.py:796-802).
    Stage-1 shapes (N_samples = 18, eikonal_lambda > 0, stage1.sh:46-50): the eikonal loss of trainer.py:618-624
    on the renderer's `eikonal_term`, differentiated with respect to the w+ latents, against the float64 oracle's
    double backward; plus the mixed loss (image term + eikonal term) through both backward paths at once."""
    res, S, B, seed = 8, 18, 2, 97
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    sd = synthetic_state_dict(64, res, seed, "sharp")
    G = G_pred_latents(model_options(size=64, renderer_spatial_output_dim=res), rendering_options(N_samples=S),
                       full_pipeline=False).eval()
    G.load_state_dict(sd, strict=False)
    G = G.cuda()
    for p in G.parameters():
        p.requires_grad_(False)
    inp = P.make_inputs(seed, B, 1, res, wplus=True)
    d = _cuda(inp)
    w = d["w"].clone().requires_grad_(True)
    out = G.renderer(d["cam_poses"], d["focal"], d["near"], d["far"], styles=w, return_eikonal=True)
    eik = out["eikonal_term"]
    assert eik.shape == (B, res, res, S, 3) and eik.requires_grad
    loss_e = ((eik.norm(dim=-1) - 1) ** 2).mean()
    loss_i = (out["features"] ** 2).mean()
    g_e, = torch.autograd.grad(loss_e, [w], retain_graph=True)
    g_mix, = torch.autograd.grad(loss_i + 0.1 * loss_e, [w])
    # float64 oracle
    sd64 = O.cast_state_dict(sd, torch.float64)
    w64 = inp["w"].double().requires_grad_(True)
    r64 = O.renderer_forward(sd64, inp["cam_poses"].double(), inp["focal"].double(), inp["near"].double(),
                             inp["far"].double(), w64, res=res, n_samples=S)
    e64 = O.eikonal_term(sd64, r64["points"].detach().reshape(B, -1, 3), w64).reshape(B, res, res, S, 3)
    l64 = ((e64.norm(dim=-1) - 1) ** 2).mean()
    ge64, = torch.autograd.grad(l64, [w64], retain_graph=True)
    gm64, = torch.autograd.grad((r64["features"] ** 2).mean() + 0.1 * l64, [w64])
    assert rel_linf(eik.detach().cpu(), e64.detach()) < TOL
    assert ge64.abs().max() > 0
    assert rel_linf(g_e.cpu(), ge64) < TOL, rel_linf(g_e.cpu(), ge64)
    assert rel_linf(g_mix.cpu(), gm64) < TOL, rel_linf(g_mix.cpu(), gm64)
    # no graph is built when nothing asks for one
    with torch.no_grad():
        plain = G.renderer(d["cam_poses"], d["focal"], d["near"], d["far"], styles=d["w"], return_eikonal=True)
    assert not plain["eikonal_term"].requires_grad
    assert torch.equal(plain["eikonal_term"], eik.detach())


def test_tc_linear_matches_fp32_matmul():
    """e3_tc_linear_fwd (the generic tcgen05 GEMM behind the local MLP tail and the eikonal sweeps): split-bf16
    products against a float64 matmul, ragged row counts, with and without bias."""
    from e3dge_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    for rows, n, k in ((1, 128, 64), (300, 256, 256), (1027, 384, 320)):
        x = torch.randn(rows, k, generator=g)
        w = torch.randn(n, k, generator=g) / k ** 0.5
        b = torch.randn(n, generator=g)
        xc, wc, bc = x.cuda(), w.cuda(), b.cuda()
        packed = torch.empty(lib.e3_tc_linear_packed_bytes(n, k) // 4, device="cuda")
        _lib.check(lib.e3_tc_linear_pack(_lib.ptr(wc), n, k, _lib.ptr(packed), _lib.cur_stream()), "e3_tc_linear_pack")
        nbytes = lib.e3_tc_linear_workspace_bytes(rows, k)
        ws = torch.empty(nbytes // 4 + 1, device="cuda")
        for bias in (None, bc):
            y = torch.empty(rows, n, device="cuda")
            _lib.check(lib.e3_tc_linear_fwd(_lib.ptr(packed), n, k, _lib.ptr(xc), rows, _lib.ptr(bias), _lib.ptr(y),
                                            _lib.ptr(ws), nbytes, _lib.cur_stream()), "e3_tc_linear_fwd")
            ref = x.double() @ w.double().t() + (b.double() if bias is not None else 0)
            assert rel_linf(y.cpu(), ref) < 2e-5, (rows, n, k)
