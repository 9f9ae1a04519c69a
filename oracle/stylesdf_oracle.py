"""TEST INFRASTRUCTURE ONLY — CPU restatement of the E3DGE / StyleSDF generator hot path.

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and the CPU-baseline legs of ``bench.py`` may import it.  The shipped path is the CUDA
library behind ``include/e3dge_b200.h``; it fails loudly when that library is missing.

It restates, as pure functions over a plain ``{reference state_dict key: tensor}``
mapping, what the reference computes on the generator path.
Every function cites the reference lines it follows (paths relative to /root/reference).
It runs in float32 (bit-comparable with the reference's CPU run, same ATen kernels) or
float64 (noise-floor studies).

PARITY PIN: the reference holds no golden vectors.  This restatement is
pinned against outputs of the reference itself, executed in the build container through
``oracle/ref_harness.py``: ``oracle/gen_golden.py`` wrote ``tests/golden/*.npz`` from the
REAL reference modules, and ``tests/test_oracle_golden.py`` checks this file against
them (and, when /root/reference is present, ``tests/test_oracle_vs_reference.py`` checks
it live, key by key).
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# --------------------------------------------------------------------------------------
# renderer: rays, samples, FiLM-SIREN, composite
# --------------------------------------------------------------------------------------
def siren_prefix(sd):
    """`renderer.network.` or, with the local branch, `renderer.network.netGlobal.`
    (train_setup.py:245-260)."""
    if any(k.startswith("renderer.network.netGlobal.") for k in sd):
        return "renderer.network.netGlobal."
    return "renderer.network."


def linear_layer(x, w, b, std_init=1.0, bias_init=0.0):
    """LinearLayer.forward — volume_renderer.py:76-80."""
    return std_init * F.linear(x, w, b) + bias_init


def film_siren(x, style, sd, key):
    """FiLMSiren.forward — volume_renderer.py:116-132 (gamma/beta :107-114).

    x [B,H,W,S,K], style [B,256] -> [B,H,W,S,256]
    """
    b, feat = style.shape
    out = F.linear(x, sd[key + "weight"], sd[key + "bias"])
    gamma = linear_layer(style, sd[key + "gamma.weight"], sd[key + "gamma.bias"],
                         std_init=15.0, bias_init=30.0).reshape(b, 1, 1, 1, -1)
    beta = linear_layer(style, sd[key + "beta.weight"], sd[key + "beta.bias"],
                        std_init=0.25, bias_init=0.0).reshape(b, 1, 1, 1, -1)
    return torch.sin(gamma * out + beta)


def siren_generator(net_inputs, styles, sd, prefix=None, local_mod=None, depth=8,
                    return_taps=None):
    """SirenGenerator.forward — volume_renderer.py:240-264 (+ :168-238).

    net_inputs [B,H,W,S,6] (normalised points | view dirs); styles [B,9,256] (w+) or
    [B,256] (w).  local_mod = (alpha, beta_loc) each [B,H,W,S,256] is the texture
    modulation of the local branch, applied after the sdf head and before the view
    layer (volume_renderer.py:217-220).  Returns raw [B,H,W,S,260] = [rgb|sdf|feat]
    (and the detached taps of `return_taps` layers, :179-180).
    """
    prefix = prefix or siren_prefix(sd)
    pts, views = net_inputs[..., :3], net_inputs[..., 3:6]
    h = pts
    taps = []
    for i in range(depth):
        st = styles[:, i] if styles.ndim == 3 else styles
        h = film_siren(h, st, sd, f"{prefix}pts_linears.{i}.")
        if return_taps and (i + 1) in return_taps:
            taps.append(h.detach())
    sdf = linear_layer(h, sd[prefix + "sigma_linear.weight"],
                       sd[prefix + "sigma_linear.bias"])
    if local_mod is not None:
        alpha, beta_loc = local_mod
        h = (alpha + 1) * h + beta_loc
    st_view = styles[:, -1] if styles.ndim == 3 else styles
    feat = film_siren(torch.cat([h, views], -1), st_view, sd, prefix + "views_linears.")
    rgb = linear_layer(feat, sd[prefix + "rgb_linear.weight"],
                       sd[prefix + "rgb_linear.bias"])
    raw = torch.cat([rgb, sdf, feat], -1)
    if return_taps:
        return raw, taps
    return raw


def get_rays(focal, c2w, res, spatial_ss=1, static_viewdirs=True):
    """VolumeFeatureRenderer.get_rays — volume_renderer.py:666-674, 769-794.

    Pixel centres linspace(0.5, res-0.5, res*ss); tensors are indexed [B, y, x, .].
    """
    n = res * spatial_ss
    lin = torch.linspace(0.5, res - 0.5, n, dtype=focal.dtype)
    # reference: i, j = meshgrid(lin, lin) (ij indexing) then transposed => i varies
    # along x (last axis), j along y.
    i = lin.view(1, 1, n).expand(1, n, n)
    j = lin.view(1, n, 1).expand(1, n, n)
    dirs = torch.stack([(i - res * .5) / focal, -(j - res * .5) / focal,
                        -torch.ones_like(i).expand(focal.shape[0], n, n)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:, None, None, :3, :3], -1)
    rays_o = c2w[:, None, None, :3, -1].expand(rays_d.shape)
    viewdirs = dirs if static_viewdirs else rays_d
    return rays_o, rays_d, viewdirs


def sample_z(near, far, n_samples, offset_sampling=True):
    """t_vals and z_vals — volume_renderer.py:690-698, 1211 (perturb == 0)."""
    if offset_sampling:
        t = torch.linspace(0., 1. - 1 / n_samples, steps=n_samples, dtype=near.dtype)
    else:
        t = torch.linspace(0., 1., steps=n_samples, dtype=near.dtype)
    t = t.reshape(1, 1, 1, -1)
    return near * (1. - t) + far * t


def volume_integration(raw, z_vals, rays_d, pts, sigmoid_beta, force_background=True,
                       no_force_stop=False, with_sdf=True, return_xyz=True,
                       feature_size=256):
    """VolumeFeatureRenderer.volume_integration — volume_renderer.py:809-943."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    rays_d_norm = torch.norm(rays_d.unsqueeze(3), dim=-1)  # [B,H,W,1]
    if not no_force_stop:
        dists = torch.cat([dists, torch.full_like(rays_d_norm, 1e10)], -1)
    else:
        dists = torch.cat([dists, dists[..., 0:1]], -1)
    dists = dists * rays_d_norm
    rgb, sdf, features = torch.split(raw, [3, 1, feature_size], dim=-1)
    if with_sdf:
        sigma = torch.sigmoid(-sdf / sigmoid_beta) / sigmoid_beta  # :804-807, 853
        alpha = 1 - torch.exp(-sigma * dists.unsqueeze(-1))  # :860
    else:
        alpha = 1 - torch.exp(-F.softplus(sdf) * dists.unsqueeze(-1))  # :866-867
    vis = torch.cumprod(
        torch.cat([torch.ones_like(alpha[..., :1, :]), 1. - alpha + 1e-10], 3), 3)
    vis = vis[..., :-1, :]
    weights = alpha * vis
    if force_background and not no_force_stop:
        weights = weights.clone()
        weights[..., -1, :] = 1 - weights[..., :-1, :].sum(3)  # :884-886
    rgb_map = -1 + 2 * torch.sum(weights * torch.sigmoid(rgb), 3)
    feature_map = torch.sum(weights * features, 3)
    xyz = depth = mask = None
    if return_xyz:
        xyz = torch.sum(weights * pts, 3)
        depth = torch.sum(weights * z_vals.unsqueeze(-1), 3, keepdim=True)
        mask = (depth < 1.08).type_as(weights)
    return dict(rgb_map=rgb_map, feature_map=feature_map, sdf=sdf, mask=mask, xyz=xyz,
                depth=depth, dists=dists, visibility=vis, weights=weights,
                rays_d_norm=rays_d_norm)


def run_network(pts, viewdirs, styles, sd, dist_radius=0.12, local_mod=None,
                return_taps=None):
    """VolumeFeatureRenderer.run_network — volume_renderer.py:1052-1128.

    pts [B,H,W,S,3] world space; viewdirs [B,H,W,3] (broadcast over S) or same shape.
    """
    if viewdirs.shape != pts.shape:
        if viewdirs.ndim != pts.ndim:
            viewdirs = viewdirs.unsqueeze(3)
        viewdirs = viewdirs.expand(pts.shape)
    scale = 2 / (dist_radius * 2)  # UniformBoxWarp — :23-30, 720
    net_inputs = torch.cat([pts * scale, viewdirs], -1)
    return siren_generator(net_inputs, styles, sd, local_mod=local_mod,
                           return_taps=return_taps)


def renderer_forward(sd, cam_poses, focal, near, far, styles, res=64, n_samples=24,
                     spatial_ss=1, static_viewdirs=True, offset_sampling=True,
                     force_background=True, with_sdf=True, dist_radius=0.12,
                     local_mod=None, return_taps=None):
    """VolumeFeatureRenderer.forward -> render -> render_rays —
    volume_renderer.py:1865-1972, 1666-1701, 1183-1298 (inference wiring: perturb 0)."""
    dt = sd["renderer.sigmoid_beta"].dtype
    rays_o, rays_d, viewdirs = get_rays(focal, cam_poses, res, spatial_ss,
                                        static_viewdirs)
    viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)  # :1679
    _near = near.unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
    _far = far.unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
    # `rays = rays.float()` (:1688) quantises the packed ray batch to fp32.
    q = lambda t: t.float().to(dt)
    rays_o, rays_d, _near, _far, viewdirs = map(q, (rays_o, rays_d, _near, _far,
                                                    viewdirs))
    z_vals = sample_z(_near, _far, n_samples, offset_sampling)
    pts = rays_o.unsqueeze(3) + rays_d.unsqueeze(3) * z_vals.unsqueeze(-1)  # :1231
    raw = run_network(pts, viewdirs, styles, sd, dist_radius, local_mod, return_taps)
    taps = None
    if return_taps:
        raw, taps = raw
    vi = volume_integration(raw, z_vals, rays_d, pts, sd["renderer.sigmoid_beta"],
                            force_background=force_background, with_sdf=with_sdf)
    out = dict(rays_o=rays_o, rays_d=rays_d, dists=vi["dists"], near=_near, far=_far,
               hit_prob=vi["weights"], points=pts, sdf=vi["sdf"],
               gen_thumb_imgs=vi["rgb_map"].permute(0, 3, 1, 2).contiguous(),
               features=vi["feature_map"].permute(0, 3, 1, 2).contiguous(),
               mask=vi["mask"].permute(0, 4, 1, 2, 3).contiguous(),
               xyz=vi["xyz"].permute(0, 3, 1, 2).contiguous(), depth=vi["depth"],
               viewdirs=viewdirs, visibility=vi["visibility"])
    if taps is not None:
        out["all_feats"] = taps
    return out


def sdf_query(sd, points, styles, dist_radius=0.12):
    """SDF-only query at arbitrary world-space points with zero view dirs —
    volume_renderer.py:955-957, 1935-1943 (`run_network(samples, 0)[..., 3:4]`).

    points [B,N,3] -> sdf [B,N,1]
    """
    p5 = points.reshape(points.shape[0], -1, 1, 1, 3)
    raw = run_network(p5, torch.zeros_like(p5), styles, sd, dist_radius)
    return raw[..., 3:4].reshape(points.shape[0], -1, 1)


def eikonal_term(sd, points, styles, dist_radius=0.12, create_graph=True):
    """get_eikonal_term — volume_renderer.py:796-802: d sdf / d point by autograd, with a graph of its own
    (`create_graph=True`) so that an eikonal loss back-propagates to the latents.  points [B,N,3]."""
    pts = points.detach().requires_grad_(True)
    with torch.enable_grad():
        sdf = sdf_query(sd, pts, styles, dist_radius)
        return torch.autograd.grad(sdf, pts, grad_outputs=torch.ones_like(sdf), create_graph=create_graph)[0]


def query_hitting_probability(sd, wd_space_pts, poses, extrinsics, near, far, styles, n_samples=24,
                              mode="fixed", return_type="weights", static_viewdirs=True,
                              offset_sampling=True, dist_radius=0.12):
    """VolumeFeatureRenderer.query_hitting_probability_fixed_interval / _adapted_interval —
    volume_renderer.py:1326-1493 / 1495-1621: re-march the reference camera's ray through every
    world-space query point and read the composite (no_force_stop) at the point.

    wd_space_pts [B,H,W,S,3]; poses / extrinsics [B,3,4] (camera-to-world / world-to-camera of the
    reference view); near, far [B,H,W,1] -> [B,H,W,S,1]."""
    B, H, W, S = wd_space_pts.shape[:4]
    dt = wd_space_pts.dtype
    pts = wd_space_pts.reshape(B, H * W, S, 3)
    homo = torch.cat([pts, torch.ones_like(pts[..., :1])], -1).unsqueeze(-1)          # :1360-1362
    w2c = torch.cat([extrinsics, torch.zeros_like(extrinsics[:, :1])], 1)
    w2c[:, -1, -1] = 1
    pts = pts.unsqueeze(-2)
    rays_o = poses[..., 3:4].permute(0, 2, 1).reshape(B, 1, 1, 1, 3)
    ref_pts = w2c.reshape(B, 1, 1, 4, 4) @ homo
    d_ref = ref_pts[..., :3, :] / (-ref_pts[..., 2:3, :])
    d_wd = (poses.reshape(B, 1, 1, 3, 4)[..., :3] @ d_ref).permute(0, 1, 2, 4, 3)       # B HW S 1 3
    nr, fr = near.reshape(B, H * W, 1, 1, 1), far.reshape(B, H * W, 1, 1, 1)
    if mode == "fixed":
        if offset_sampling:
            t = torch.linspace(0., 1. - 1 / n_samples, steps=n_samples, dtype=dt)
        else:
            t = torch.linspace(0., 1., steps=n_samples, dtype=dt)
        z = nr * (1. - t) + fr * t                                                      # B HW 1 1 T
        interval = (z[..., 1:2] - z[..., 0:1]) * d_wd.norm(dim=-1, keepdim=True).permute(0, 1, 2, 4, 3)
        z = z.permute(0, 1, 2, 4, 3)
        ray_pts = rays_o + d_wd * z
        idx = (pts - ray_pts[..., 0:1, :]).norm(dim=-1, keepdim=True) / interval + 1e-5
        lo = idx.floor().long().clamp(0, n_samples - 1)
        hi = idx.ceil().long().clamp(0, n_samples - 1)
        z = z.squeeze(-1)
    else:
        near_pts = rays_o + d_wd * nr
        t = torch.linspace(0., 1., steps=n_samples, dtype=dt).reshape(1, 1, 1, -1, 1)
        ray_pts = near_pts * (1 - t) + pts * t
        z = (ray_pts - rays_o).norm(dim=-1)
    d_ref = d_ref.squeeze(-1)
    viewdirs = F.normalize(d_ref if static_viewdirs else d_wd.squeeze(3), dim=-1)
    raw = run_network(ray_pts, viewdirs, styles, sd, dist_radius)
    vi = volume_integration(raw, z, viewdirs, None, sd["renderer.sigmoid_beta"], no_force_stop=True,
                            return_xyz=False)
    if mode == "fixed":
        info = vi["weights"] if return_type == "weights" else vi["visibility"]
        out = torch.lerp(torch.gather(info, 3, lo), torch.gather(info, 3, hi), idx - lo)
    else:
        out = vi["weights"][..., -1:, :]
    return out.reshape(B, H, W, S, 1)


# --------------------------------------------------------------------------------------
# StyleGAN2 ops
# --------------------------------------------------------------------------------------
def fused_leaky_relu(x, bias=None, negative_slope=0.2, scale=SQRT2, gate=None):
    """fused_bias_act semantics — op/fused_bias_act_kernel.cu:36-47, op/fused_act.py:107-115:
    y = leaky_relu(x + b[c]) * scale, bias along dim 1.

    `gate` (bool tensor, test aid): take the branch decisions x + b > 0 from outside instead of
    from this evaluation.  Gradient parity tests pass the signs of the CUDA forward here, so that
    elements sitting on the kink within float32 round-off (where the derivative is ambiguous)
    do not decide the comparison."""
    if bias is not None:
        x = x + bias.reshape(1, -1, *([1] * (x.ndim - 2)))
    if gate is not None:
        return torch.where(gate, x, x * negative_slope) * scale
    return F.leaky_relu(x, negative_slope) * scale


def fused_leaky_relu_grad(grad_out, out, negative_slope=0.2, scale=SQRT2):
    """grad variant (act=3, grad=1): gated by the sign of the saved OUTPUT —
    op/fused_bias_act_kernel.cu:40-42, op/fused_act.py:29-30."""
    return torch.where(out > 0, grad_out, grad_out * negative_slope) * scale


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """upfirdn2d — op/upfirdn2d.py:157-200 / op/upfirdn2d_kernel.cu:107-207:
    zero-insert upsample, pad (negative pad crops), correlate with the FLIPPED kernel,
    decimate.  x [N,C,H,W], kernel [kh,kw], pad=(p0,p1) used for both axes."""
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    p0, p1 = pad
    y = x.reshape(n * c, 1, h, 1, w, 1)
    y = F.pad(y, [0, up - 1, 0, 0, 0, up - 1])
    y = y.reshape(n * c, 1, h * up, w * up)
    y = F.pad(y, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    y = y[:, :, max(-p0, 0):y.shape[2] - max(-p1, 0),
          max(-p0, 0):y.shape[3] - max(-p1, 0)]
    y = F.conv2d(y, torch.flip(kernel, [0, 1]).reshape(1, 1, kh, kw).to(y.dtype))
    y = y[:, :, ::down, ::down]
    return y.reshape(n, c, y.shape[2], y.shape[3])


def make_kernel(k, dtype=torch.float32):
    """stylesdf_model.py:85-93."""
    k = torch.tensor(k, dtype=dtype)
    k = k[None, :] * k[:, None]
    return k / k.sum()


def equal_linear(x, w, b, lr_mul=1.0, activation=False):
    """EqualLinear.forward — stylesdf_model.py:234-244."""
    scale = (1 / math.sqrt(w.shape[1])) * lr_mul
    if activation:
        return fused_leaky_relu(F.linear(x, w * scale), b * lr_mul)
    return F.linear(x, w * scale, b * lr_mul)


def modulated_conv2d(x, style, sd, key, demodulate=True, upsample=False,
                     blur_kernel=(1, 3, 3, 1)):
    """ModulatedConv2d.forward — stylesdf_model.py:317-362 (+ Blur :148-165, :283-291)."""
    weight = sd[key + "weight"]  # [1,O,I,k,k]
    _, o, i, k, _ = weight.shape
    b, _, h, w = x.shape
    s = equal_linear(style, sd[key + "modulation.weight"], sd[key + "modulation.bias"])
    s = s.reshape(b, 1, i, 1, 1)
    wgt = (1 / math.sqrt(i * k * k)) * weight * s
    if demodulate:
        d = torch.rsqrt(wgt.pow(2).sum([2, 3, 4]) + 1e-8)
        wgt = wgt * d.reshape(b, o, 1, 1, 1)
    if upsample:
        xin = x.reshape(1, b * i, h, w)
        wt = wgt.transpose(1, 2).reshape(b * i, o, k, k)
        out = F.conv_transpose2d(xin, wt, padding=0, stride=2, groups=b)
        out = out.reshape(b, o, out.shape[2], out.shape[3])
        factor = 2
        p = (len(blur_kernel) - factor) - (k - 1)
        pad = ((p + 1) // 2 + factor - 1, p // 2 + 1)
        kern = make_kernel(list(blur_kernel), x.dtype) * (factor ** 2)
        return upfirdn2d(out, kern, pad=pad)
    xin = x.reshape(1, b * i, h, w)
    out = F.conv2d(xin, wgt.reshape(b * o, i, k, k), padding=k // 2, groups=b)
    return out.reshape(b, o, out.shape[2], out.shape[3])


def styled_conv(x, style, noise, sd, key, upsample=False, gate=None):
    """StyledConv.forward — stylesdf_model.py:494-507 (noise :459-466, act op/fused_act.py)."""
    out = modulated_conv2d(x, style, sd, key + "conv.", upsample=upsample)
    out = out + sd[key + "noise.weight"] * noise
    return fused_leaky_relu(out, sd[key + "activate.bias"], gate=gate)


def to_rgb(x, style, skip, sd, key, upsample=True):
    """ToRGB.forward — stylesdf_model.py:531-541 (Upsample :96-119)."""
    out = modulated_conv2d(x, style, sd, key + "conv.", demodulate=False)
    out = out + sd[key + "bias"]
    if skip is not None:
        if upsample:
            kern = make_kernel([1, 3, 3, 1], x.dtype) * 4
            skip = upfirdn2d(skip, kern, up=2, pad=(2, 1))
        out = out + skip
    return out


def decoder_num_layers(sd):
    n = 0
    while f"decoder.noises.noise_{n}" in sd:
        n += 1
    return n


def decoder_forward(sd, features, latent, noises=None, gates=None, acts=None):
    """Decoder.forward with input_is_latent=True, randomize_noise=False —
    stylesdf_model.py:742-797 (latent indexing :764-792; noise buffers :652-656,707-710).

    features [B,256,R,R]; latent [B,n_latent,512] -> image [B,3,size,size]
    gates (test aid): per StyledConv layer the branch decisions of its leaky-ReLU (see
    fused_leaky_relu); acts: a list that receives the StyledConv outputs in order.
    """
    n_layers = decoder_num_layers(sd)
    if noises is None:
        noises = [sd[f"decoder.noises.noise_{i}"] for i in range(n_layers)]
    g = (lambda k: gates[k]) if gates is not None else (lambda k: None)
    keep = acts.append if acts is not None else (lambda t: None)
    out = styled_conv(features, latent[:, 0], noises[0], sd, "decoder.conv1.", gate=g(0))
    keep(out)
    skip = to_rgb(out, latent[:, 1], None, sd, "decoder.to_rgb1.", upsample=False)
    i = 1
    for stage in range((n_layers - 1) // 2):
        out = styled_conv(out, latent[:, i], noises[2 * stage + 1], sd,
                          f"decoder.convs.{2 * stage}.", upsample=True, gate=g(2 * stage + 1))
        keep(out)
        out = styled_conv(out, latent[:, i + 1], noises[2 * stage + 2], sd,
                          f"decoder.convs.{2 * stage + 1}.", gate=g(2 * stage + 2))
        keep(out)
        skip = to_rgb(out, latent[:, i + 2], skip, sd, f"decoder.to_rgbs.{stage}.")
        i += 2
    return skip


def mapping_network(z, sd):
    """Generator.style (3 x MappingLinear, fused lrelu scale 1) — stylesdf_model.py:70-77,822-830."""
    h = z
    for i in range(3):
        h = fused_leaky_relu(F.linear(h, sd[f"style.{i}.weight"]), sd[f"style.{i}.bias"],
                             scale=1.0)
    return h


def decoder_mapping(w, sd, lr_mul=0.01):
    """Decoder.style = PixelNorm + 5 x EqualLinear(fused lrelu) — stylesdf_model.py:30-37,596-611."""
    h = w * torch.rsqrt(torch.mean(w ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(1, 6):
        h = equal_linear(h, sd[f"decoder.style.{i}.weight"], sd[f"decoder.style.{i}.bias"],
                         lr_mul=lr_mul, activation=True)
    return h


def generator_forward(sd, w, w_dec, cam_poses, focal, near, far, res=64, n_samples=24,
                      **render_kw):
    """G_pred_latents.forward(styles=[w, w_dec], input_is_latent=True,
    randomize_noise=False) — stylesdf_model.py:1034-1172."""
    out = renderer_forward(sd, cam_poses, focal, near, far, w, res=res,
                           n_samples=n_samples, **render_kw)
    out["styles"] = w
    out["gen_imgs"] = decoder_forward(sd, out["features"], w_dec)
    return out


def cast_state_dict(sd, dtype):
    return {k: v.to(dtype) for k, v in sd.items()}
