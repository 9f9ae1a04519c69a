"""`project.utils.volume_renderer` surface on top of the fused sm_100a render kernel.

Classes, constructor arguments, attribute names, state_dict keys and the returned dict
follow the reference (project/utils/volume_renderer.py:23-166, 636-749, 1865-1972) so that
its runners can call this module unchanged; the arithmetic between `forward()` and the
returned dict is ONE CUDA kernel (csrc/render_siren.cu) instead of ~150 ATen launches.

Not provided (SURVEY.md §8 out of scope / "next" rows): a GPU surface extractor (`return_mesh`
aligns the SDF volume on the device and hands it to scikit-image on the host, as the reference does:
mesh_utils.py), the 2-D hourglass image filter of the PIFu `netLocal` (an encoder; its per-sample
half — feature query, SFT fusion, positional encoding, texture-modulation MLP — is local_query.py /
local_branch.py).  Second-order gradients: the eikonal terms carry a graph to the latents (eikonal.py);
nothing else is twice differentiable.

Training (encoders against the frozen generator, trainer.py:881-900, generator frozen at :1569): `_FilmFn`, `_RenderFn`
and `_PointsFn` bind e3_film_bwd / e3_render_bwd / e3_siren_points_bwd, so gradients reach the
w / w+ latents, the local texture modulation and explicit query points.  The generator's own
parameters and the cameras receive no gradient.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .eikonal import eikonal_term


class UniformBoxWarp(nn.Module):
    """volume_renderer.py:23-30."""

    def __init__(self, sidelength):
        super().__init__()
        self.scale_factor = 2 / sidelength

    def forward(self, coordinates):
        return coordinates * self.scale_factor


def _kaiming_leaky(out_dim, in_dim, gain_mul=1.0):
    w = torch.randn(out_dim, in_dim)
    return gain_mul * nn.init.kaiming_normal_(w, a=0.2, mode="fan_in", nonlinearity="leaky_relu")


class LinearLayer(nn.Module):
    """std_init * (W x + b) + bias_init — volume_renderer.py:42-80."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, std_init=1, freq_init=False,
                 is_first=False):
        super().__init__()
        if is_first:
            w = torch.empty(out_dim, in_dim).uniform_(-1 / in_dim, 1 / in_dim)
        elif freq_init:
            lim = math.sqrt(6 / in_dim) / 25
            w = torch.empty(out_dim, in_dim).uniform_(-lim, lim)
        else:
            w = _kaiming_leaky(out_dim, in_dim, 0.25)
        self.weight = nn.Parameter(w)
        lim = math.sqrt(1 / in_dim)
        self.bias = nn.Parameter(torch.empty(out_dim).uniform_(-lim, lim))
        self.bias_init = bias_init
        self.std_init = std_init

    def forward(self, input):
        return self.std_init * F.linear(input, self.weight, bias=self.bias) + self.bias_init


class FiLMSiren(nn.Module):
    """sin(gamma(style) * (W x + b) + beta(style)) — volume_renderer.py:84-132.

    Holds the parameters under the reference's names; inside the renderer these layers are
    executed by the fused kernel, `forward` here serves stand-alone use.
    """

    def __init__(self, in_channel, out_channel, style_dim, is_first=False):
        super().__init__()
        self.in_channel, self.out_channel = in_channel, out_channel
        lim = 1 / 3 if is_first else math.sqrt(6 / in_channel) / 25
        self.weight = nn.Parameter(torch.empty(out_channel, in_channel).uniform_(-lim, lim))
        blim = math.sqrt(1 / in_channel)
        self.bias = nn.Parameter(torch.empty(out_channel).uniform_(-blim, blim))
        self.activation = torch.sin
        self.gamma = LinearLayer(style_dim, out_channel, bias_init=30, std_init=15)
        self.beta = LinearLayer(style_dim, out_channel, bias_init=0, std_init=0.25)

    def forward(self, input, style):
        batch, features = style.shape
        out = F.linear(input, self.weight, bias=self.bias)
        shape = (batch,) + (1,) * (out.ndim - 2) + (features,)
        return torch.sin(self.gamma(style).reshape(shape) * out + self.beta(style).reshape(shape))


class SirenGenerator(nn.Module):
    """8 x FiLMSiren + view layer + rgb / sdf heads — volume_renderer.py:136-264."""

    def __init__(self, opt=None, D=8, W=256, style_dim=256, input_ch=3, input_ch_views=3,
                 output_ch=4, output_features=True, scene_scale=0.12, **kwargs):
        super().__init__()
        if D != 8 or W != 256 or style_dim != 256 or input_ch != 3 or input_ch_views != 3:
            raise NotImplementedError(
                "e3dge_b200: the fused kernel is built for depth=8, width=256, style_dim=256 "
                f"(got D={D}, W={W}, style_dim={style_dim})")
        self.opt, self.D, self.W = opt, D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.style_dim, self.output_features = style_dim, output_features
        self.pts_linears = nn.ModuleList(
            [FiLMSiren(3, W, style_dim=style_dim, is_first=True)] +
            [FiLMSiren(W, W, style_dim=style_dim) for _ in range(D - 1)])
        self.views_linears = FiLMSiren(input_ch_views + W, W, style_dim=style_dim)
        self.rgb_linear = LinearLayer(W, 3, freq_init=True)
        self.sigma_linear = LinearLayer(W, 1, freq_init=True)

    def weight_tensors(self):
        """Parameters in the order of `e3_siren_weights` (include/e3dge_b200.h)."""
        pl, vl = self.pts_linears, self.views_linears
        films = list(pl) + [vl]
        return dict(pts_w=[l.weight for l in pl], pts_b=[l.bias for l in pl],
                    gamma_w=[f.gamma.weight for f in films], gamma_b=[f.gamma.bias for f in films],
                    beta_w=[f.beta.weight for f in films], beta_b=[f.beta.bias for f in films],
                    views_w=vl.weight, views_b=vl.bias, rgb_w=self.rgb_linear.weight,
                    rgb_b=self.rgb_linear.bias, sigma_w=self.sigma_linear.weight,
                    sigma_b=self.sigma_linear.bias)

    # ---- kernel plumbing ----
    def packed_weights(self):
        holder = self.__dict__.setdefault("_e3_packed", _PackedSiren())
        return holder.get(self)

    def film(self, styles):
        """styles [B,256] (w) / [B,9,256] (w+) -> FiLM table [B,9,3,256] (gamma, beta, beta' = gamma*b + beta)."""
        lib = _lib.load()
        styles = _lib.as_f32c(styles)
        if styles.ndim == 2:
            b, spi = styles.shape[0], 1
        elif styles.ndim == 3 and styles.shape[1] == 9:
            b, spi = styles.shape[0], 9
        else:
            raise RuntimeError(f"styles must be [B,256] (w) or [B,9,256] (w+), got {tuple(styles.shape)}")
        film = torch.empty(b, 9, 3, 256, device=styles.device, dtype=torch.float32)
        _lib.check(lib.e3_film_fwd(_lib.ptr(self.packed_weights()), _lib.ptr(styles), b, spi,
                                   _lib.ptr(film), _lib.cur_stream()), "e3_film_fwd")
        return film

    def _points(self, x, views, styles, local_mod=None, want=("sdf", "rgb", "feat"), scale=1.0):
        """One e3_siren_points_fwd_ex launch over network inputs x [B,...,3] (already box-normalised when
        scale == 1): any of sdf [B,...,1], rgb [B,...,3], feat [B,...,256], h8 [B,...,256]."""
        lib = _lib.load()
        shp = tuple(x.shape[:-1])
        B = shp[0]
        pts = _lib.as_f32c(x.detach()).reshape(B, -1, 3)
        N = pts.shape[1]
        dev = pts.device
        with_view = "rgb" in want or "feat" in want
        vd = None
        if views is not None and with_view:
            vd = _lib.as_f32c(views.detach()).expand(*shp, 3).reshape(B, -1, 3).contiguous()
        new = lambda c: torch.empty(B, N, c, device=dev, dtype=torch.float32)
        sdf = torch.empty(B, N, device=dev, dtype=torch.float32)
        rgb = new(3) if with_view else None
        feat = new(256) if with_view else None
        h8 = new(256) if "h8" in want else None
        la = lb = None
        if local_mod is not None:
            la, lb = (_lib.as_f32c(t.detach()).reshape(B, N, 256) for t in local_mod)
        film = self.film(styles.detach())
        _lib.check(lib.e3_siren_points_fwd_ex(
            _lib.ptr(self.packed_weights()), _lib.ptr(film), _lib.ptr(pts), _lib.ptr(vd), B, N,
            float(scale), _lib.ptr(la), _lib.ptr(lb), _lib.ptr(sdf), _lib.ptr(rgb), _lib.ptr(feat), _lib.ptr(h8), 0,
            _lib.cur_stream()), "e3_siren_points_fwd_ex")
        out = {"sdf": sdf.reshape(*shp, 1)}
        if with_view:
            out.update(rgb=rgb.reshape(*shp, 3), feat=feat.reshape(*shp, 256))
        if h8 is not None:
            out["h8"] = h8.reshape(*shp, 256)
        return out

    # ---- the reference's network-level API (inference; volume_renderer.py:168-264) ----
    @torch.no_grad()
    def forward_generator(self, input_pts, styles, conditions=None):
        """Backbone features after the eight FiLM layers, [...,256] (:168-194)."""
        return self._points(input_pts, None, styles, want=("h8",))["h8"]

    forward_backbone = forward_generator  # older name of the same method (:196-204)

    def forward_geo(self, feats):
        """sdf head on given backbone features (:206-208)."""
        return self.sigma_linear(feats)

    def forward_tex(self, mlp_out, input_views, styles, conditions=None):
        """View layer + rgb head on given backbone features, optional local texture modulation (:210-238).
        Stand-alone PyTorch (FiLMSiren.forward); the renderer runs this inside the fused kernel."""
        if isinstance(mlp_out, dict):
            conditions = mlp_out.get("conditions", conditions)
            mlp_out = mlp_out["mlp_out"]
        if conditions and "tex" in conditions:
            alpha, beta = conditions["tex"]
            mlp_out = (alpha + 1) * mlp_out + beta
        w = styles[:, -1] if styles.ndim == 3 else styles
        out_features = self.views_linears(torch.cat([mlp_out, input_views.expand(*mlp_out.shape[:-1], 3)], -1), w)
        return self.rgb_linear(out_features), out_features

    @torch.no_grad()
    def forward(self, x, styles, local_modulation=None):
        """raw = [rgb | sdf | features] at network inputs x = [normalised point | view direction] (:240-264)."""
        o = self._points(x[..., :3], x[..., 3:6], styles, local_mod=local_modulation)
        return torch.cat([o["rgb"], o["sdf"], o["feat"]], -1)


class SirenLocalGlobal(nn.Module):
    """Global FiLM-SIREN + local branch — volume_renderer.py:267-558.  `netGlobal.*` / `netLocal.*` give the
    state_dict names of the local-branch checkpoints (train_setup.py:245-260); both are reached from outside
    (trainer.py:1611, e3dge_full_runner.py:166, 219).  `forward` is the fused route (texture modulation from
    e3_local_mlp_fwd handed to e3_siren_points_fwd_ex); the three stage methods the reference composes it
    from are provided with the same dict contracts."""

    def __init__(self, opt=None, D=8, W=256, style_dim=256, input_ch=3, input_ch_views=3,
                 output_ch=4, output_features=True, scene_scale=0.12, local_options=None, **kw):
        super().__init__()
        from .local_branch import LocalBranch
        self.opt = opt
        self.netGlobal = SirenGenerator(opt, D, W, style_dim, input_ch, input_ch_views, output_ch,
                                        output_features, scene_scale)
        self.netLocal = LocalBranch(opt, local_options)

    def forward_local(self, data_batch):
        """:439-476 — local features of the sample points: given (`feats`) or queried from filtered images."""
        if data_batch.get("feats") is not None:
            return data_batch
        points, images, calibs = (data_batch[k] for k in ("world_space_pts", "gen_imgs", "calibs"))
        shp = points.shape
        self.netLocal.filter(images)
        q = self.netLocal.query(points=points.reshape(shp[0], -1, 3).permute(0, 2, 1), calibs=calibs,
                                feat_key="ref_view", return_feat_only=True)
        out = dict(q)
        out["feats"] = q["feats"].permute(0, 2, 1).reshape(*shp[:-1], -1)
        return out

    def _tex_conditions(self, local_output):
        from .local_branch import ResnetBlockFC, tex_modulation
        tex = self.netLocal.local_feat_to_tex_modulations_linear
        feats = local_output["feats"]
        fused = (isinstance(tex, ResnetBlockFC) and (tex.size_in, tex.size_out) == (301, 512) and feats.is_cuda
                 and not VolumeFeatureRenderer._wants_grad(feats, *tex.parameters()))
        if fused:  # inference: stages 5-6 of e3_local_mlp_fwd
            return list(tex_modulation(tex, feats))
        return list(torch.split(tex(feats), 256, dim=-1))  # training / a caller's own module: autograd path

    def forward_backbone(self, input_pts, styles, local_data_batch):
        """:313-369."""
        conditions = {}
        local_output = None
        if local_data_batch is not None and "sampling" not in local_data_batch:
            local_output = self.forward_local(local_data_batch)
            if getattr(self.opt, "L_pred_tex_modulations", False):
                conditions["tex"] = self._tex_conditions(local_output)
                local_output["tex_modulate_conditions"] = True
        global_feats = self.netGlobal.forward_generator(input_pts, styles, conditions)
        return dict(global_output={"feats": global_feats}, local_output=local_output,
                    local_modulation_conditions=conditions)

    def retrieve_feats_for_rendering(self, forward_out, sample_mode):
        """:371-428 (strategies of the shipped configuration: geometry 'global', texture 'global_local')."""
        global_feats = forward_out["global_output"]["feats"]
        if forward_out["local_output"] is None or sample_mode:
            return dict(feats_to_geo=global_feats, feats_to_tex=global_feats)
        geo = getattr(self.opt, "geo_predictition_strategy", "global")
        tex = getattr(self.opt, "tex_predictition_strategy", "global_local")
        if "global" not in geo or "global" not in tex:
            raise NotImplementedError("local-only prediction strategies are not adopted in the paper (:401, 414)")
        feats_to_geo = global_feats
        conditions = forward_out.get("local_modulation_conditions") or {}
        if "local" in geo:
            alpha, beta = conditions["geo"]
            feats_to_geo = (alpha + 1) * feats_to_geo + beta
        feats_to_tex = dict(mlp_out=global_feats)
        if "tex" in conditions:
            feats_to_tex["conditions"] = conditions
        return dict(feats_to_geo=feats_to_geo, feats_to_tex=feats_to_tex)

    def forward_rendering(self, feats_for_render_dict, input_views, styles):
        """:478-517."""
        rgb, out_features = self.netGlobal.forward_tex(feats_for_render_dict["feats_to_tex"], input_views, styles)
        sdf = self.netGlobal.forward_geo(feats_for_render_dict["feats_to_geo"])
        outputs = torch.cat([rgb, sdf], -1)
        return torch.cat([outputs, out_features], -1) if self.netGlobal.output_features else outputs

    def feats_to_geo_query(self, *args):
        return self.netGlobal.forward_geo(*args)

    def feats_to_tex_query(self, *args, **kwargs):
        return self.netGlobal.forward_tex(*args, **kwargs)

    @torch.no_grad()
    def forward(self, net_inputs, styles, local_data_batch=None, sample_mode=False):
        """:527-558, fused: (alpha, beta) from the local features, then ONE kernel for backbone, sdf head,
        modulation, view layer and rgb head."""
        mod = None
        if (local_data_batch is not None and "sampling" not in local_data_batch and not sample_mode
                and getattr(self.opt, "L_pred_tex_modulations", False)):
            mod = tuple(self._tex_conditions(self.forward_local(local_data_batch)))
        return self.netGlobal(net_inputs, styles, local_modulation=mod)


class _PackedSiren:
    """Device image of the SIREN weights in the kernel's layout, rebuilt when any
    parameter changes (version counters) — frozen generator => packed once."""

    def __init__(self):
        self.buf = None
        self.key = None

    def get(self, net):
        wt = net.weight_tensors()
        flat = []
        for v in wt.values():
            flat.extend(v if isinstance(v, list) else [v])
        key = (_lib.pack_epoch,) + tuple((t.data_ptr(), t._version) for t in flat)
        if self.buf is not None and key == self.key:
            return self.buf
        lib = _lib.load()
        dev = flat[0].device
        if dev.type != "cuda":
            raise RuntimeError("e3dge_b200: renderer weights must live on a CUDA device")
        keep = []

        def p(t):
            t = _lib.as_f32c(t.detach())
            keep.append(t)
            return t.data_ptr()

        sw = _lib.SirenWeights()
        for name, v in wt.items():
            if isinstance(v, list):
                arr = getattr(sw, name)
                for i, t in enumerate(v):
                    arr[i] = p(t)
            else:
                setattr(sw, name, p(v))
        nbytes = lib.e3_siren_packed_bytes()
        buf = torch.empty(nbytes // 4, device=dev, dtype=torch.float32)
        import ctypes
        _lib.check(lib.e3_siren_pack(ctypes.byref(sw), _lib.ptr(buf), _lib.cur_stream()),
                   "e3_siren_pack")
        self.buf, self.key = buf, key
        return buf


def _scratch(batch, dev):
    lib = _lib.load()
    n = lib.e3_render_bwd_scratch_bytes(batch)
    return torch.empty(max(n // 4, 1), device=dev, dtype=torch.float32), n


class _FilmFn(torch.autograd.Function):
    """styles -> FiLM table [B,9,3,256] (e3_film_fwd) with its adjoint (e3_film_bwd)."""

    @staticmethod
    def forward(ctx, renderer, styles):
        ctx.renderer, ctx.shape = renderer, tuple(styles.shape)
        return renderer._film(styles)

    @staticmethod
    def backward(ctx, d_film):
        lib = _lib.load()
        shape = ctx.shape
        b, spi = shape[0], (1 if len(shape) == 2 else shape[1])
        d2 = d_film[:, :, :2].contiguous()
        d_styles = torch.empty(b, spi, 256, device=d_film.device, dtype=torch.float32)
        _lib.check(lib.e3_film_bwd(_lib.ptr(ctx.renderer.packed_weights()), _lib.ptr(d2), b, spi,
                                   _lib.ptr(d_styles), _lib.cur_stream()), "e3_film_bwd")
        return None, d_styles.reshape(shape)


class _RenderFn(torch.autograd.Function):
    """Forward = the fused kernel writing its backward stash; backward = e3_render_bwd.
    Differentiable inputs: the FiLM table and the local texture modulation."""

    DIFF = ("features", "gen_thumb_imgs", "xyz", "depth", "sdf", "hit_prob")

    @staticmethod
    def forward(ctx, renderer, film, local_alpha, local_beta, cam_poses, focal, near, far, z_jitter,
                flags_over, want_taps):
        _lib.warn_if_trainable(renderer, "VolumeFeatureRenderer")
        local_mod = None if local_alpha is None else (local_alpha, local_beta)
        out, call = renderer._render_raw(None, cam_poses, focal, near, far, z_jitter, local_mod,
                                         flags_over, want_taps, film=film, train=True)
        names = list(out.keys())
        ctx.names, ctx.renderer, ctx.call = names, renderer, call
        # outputs go through save_for_backward (a plain attribute would tie them to the graph node
        # in a reference cycle and keep the multi-GB stash alive until the cyclic GC runs)
        ctx.save_for_backward(out["sdf"], out["hit_prob"], out["raw_rgb"])
        ctx.want_local = local_alpha is not None
        ctx.mark_non_differentiable(*[out[k] for k in names if k not in _RenderFn.DIFF])
        renderer._last_names = names
        return tuple(out[k] for k in names)

    @staticmethod
    def backward(ctx, *grads):
        import ctypes
        lib = _lib.load()
        call = ctx.call
        g = {k: v for k, v in zip(ctx.names, grads)}
        gin = [(_lib.as_f32c(g[k]) if g.get(k) is not None else None) for k in _RenderFn.DIFF]
        film = call["film"]
        B, dev = film.shape[0], film.device
        d_film = torch.empty(B, 9, 2, 256, device=dev, dtype=torch.float32)
        d_la = d_lb = None
        if ctx.want_local and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3]):
            d_la = torch.empty_like(call["la"])
            d_lb = torch.empty_like(call["lb"])
        scratch, nbytes = _scratch(B, dev)
        sv_sdf, sv_hit, sv_rgb = ctx.saved_tensors
        saved = _lib.RenderSaved(_lib.ptr(call["stash"]), _lib.ptr(sv_sdf), _lib.ptr(sv_hit), _lib.ptr(sv_rgb))
        gr = _lib.RenderGrads(*[_lib.ptr(t) for t in gin])
        bo = _lib.RenderBwdOutputs(_lib.ptr(d_film), _lib.ptr(d_la), _lib.ptr(d_lb), None)
        _lib.check(lib.e3_render_bwd(_lib.ptr(ctx.renderer.packed_weights()), ctypes.byref(call["prm"]),
                                     ctypes.byref(call["inp"]), ctypes.byref(saved), ctypes.byref(gr),
                                     ctypes.byref(bo), _lib.ptr(scratch), nbytes, _lib.cur_stream()),
                   "e3_render_bwd")
        d_film3 = torch.zeros(B, 9, 3, 256, device=dev, dtype=torch.float32)
        d_film3[:, :, :2] = d_film
        return (None, d_film3, d_la, d_lb) + (None,) * 7


class _PointsFn(torch.autograd.Function):
    """FiLM-SIREN at explicit points with a backward (e3_siren_points_fwd_train / _bwd).
    Differentiable inputs: the FiLM table and the points."""

    @staticmethod
    def forward(ctx, renderer, film, pts, viewdirs, with_view):
        lib = _lib.load()
        B, N = pts.shape[0], pts.shape[1]
        dev = pts.device
        sdf = torch.empty(B, N, device=dev, dtype=torch.float32)
        rgb = torch.empty(B, N, 3, device=dev, dtype=torch.float32) if with_view else None
        feat = torch.empty(B, N, 256, device=dev, dtype=torch.float32) if with_view else None
        stash = torch.empty(max(lib.e3_render_stash_bytes(1, N, B) // 4, 1), device=dev,
                            dtype=torch.float32)
        scale = float(renderer.grid_warper.scale_factor)
        _lib.check(lib.e3_siren_points_fwd_train(
            _lib.ptr(renderer.packed_weights()), _lib.ptr(film), _lib.ptr(pts), _lib.ptr(viewdirs), B, N,
            scale, _lib.ptr(sdf), _lib.ptr(rgb), _lib.ptr(feat), _lib.ptr(stash), _lib.cur_stream()),
            "e3_siren_points_fwd_train")
        ctx.renderer, ctx.film, ctx.stash, ctx.with_view, ctx.scale = renderer, film, stash, with_view, scale
        ctx.dims = (B, N)
        if with_view:
            return sdf, rgb, feat
        return sdf, None, None

    @staticmethod
    def backward(ctx, d_sdf, d_rgb, d_feat):
        lib = _lib.load()
        B, N = ctx.dims
        dev = ctx.film.device
        c = lambda t: _lib.as_f32c(t) if t is not None else None
        d_sdf, d_rgb, d_feat = c(d_sdf), c(d_rgb), c(d_feat)
        d_film = torch.empty(B, 9, 2, 256, device=dev, dtype=torch.float32)
        d_pts = torch.empty(B, N, 3, device=dev, dtype=torch.float32) if ctx.needs_input_grad[2] else None
        scratch, nbytes = _scratch(B, dev)
        _lib.check(lib.e3_siren_points_bwd(
            _lib.ptr(ctx.renderer.packed_weights()), _lib.ptr(ctx.film), B, N, ctx.scale,
            _lib.ptr(ctx.stash), int(ctx.with_view), 0, _lib.ptr(d_sdf), _lib.ptr(d_rgb), _lib.ptr(d_feat),
            _lib.ptr(d_film), _lib.ptr(d_pts), _lib.ptr(scratch), nbytes, _lib.cur_stream()),
            "e3_siren_points_bwd")
        d_film3 = torch.zeros(B, 9, 3, 256, device=dev, dtype=torch.float32)
        d_film3[:, :, :2] = d_film
        return None, d_film3, d_pts, None, None


class VolumeFeatureRenderer(nn.Module):
    """volume_renderer.py:636-749 (construction), :1865-1972 (forward)."""

    def __init__(self, opt, style_dim=256, out_im_res=64, mode="train"):
        super().__init__()
        self.test = mode != "train"
        self.opt = opt
        self.perturb = opt.perturb
        self.offset_sampling = not opt.no_offset_sampling
        self.N_samples = opt.N_samples
        self.raw_noise_std = opt.raw_noise_std
        self.return_xyz = opt.return_xyz
        self.return_sdf = True
        self.static_viewdirs = opt.static_viewdirs
        self.z_normalize = not opt.no_z_normalize
        self.out_im_res = out_im_res
        self.spatial_ss = opt.spatial_super_sampling_factor
        self.force_background = opt.force_background
        self.with_sdf = not opt.no_sdf
        self.add_fg_mask = opt.add_fg_mask
        self.output_features = "no_features_output" not in opt.keys()
        if not self.z_normalize:
            raise NotImplementedError("no_z_normalize crashes in the reference as well "
                                      "(volume_renderer.py:1074-1078)")
        if self.with_sdf:
            self.sigmoid_beta = nn.Parameter(0.1 * torch.ones(1))
        n = out_im_res * self.spatial_ss
        lin = torch.linspace(0.5, out_im_res - 0.5, n)
        self.register_buffer("i", lin.view(1, 1, n).repeat(1, n, 1), persistent=False)
        self.register_buffer("j", lin.view(1, n, 1).repeat(1, 1, n), persistent=False)
        self.register_buffer("_pix", lin.clone(), persistent=False)
        if self.offset_sampling:
            t_vals = torch.linspace(0., 1. - 1 / self.N_samples, steps=self.N_samples)
        else:
            t_vals = torch.linspace(0., 1., steps=self.N_samples)
        self.register_buffer("t_vals", t_vals.reshape(1, 1, 1, -1), persistent=False)
        self.register_buffer("inf", torch.Tensor([1e10]), persistent=False)
        if self.test:
            self.perturb = False
            self.raw_noise_std = 0.
        self.channel_dim, self.samples_dim = -1, 3
        self.input_ch = self.input_ch_views = 3
        self.feature_out_size = opt.width
        self.grid_warper = UniformBoxWarp(opt.camera.dist_radius * 2)
        self.grid_un_warper = UniformBoxWarp(1 / opt.camera.dist_radius * 2)
        self.enable_local_model = bool(opt.enable_local_model)
        net_cls = SirenLocalGlobal if self.enable_local_model else SirenGenerator
        extra = {"local_options": getattr(opt, "pifu", None)} if self.enable_local_model else {}  # :741
        self.network = net_cls(opt=opt, D=opt.depth, W=opt.width, style_dim=style_dim,
                               input_ch=3, output_ch=4, input_ch_views=3,
                               output_features=self.output_features, **extra)
        r = opt.camera.dist_radius
        self.register_buffer("B_MAX", torch.Tensor([r] * 3), persistent=False)
        self.register_buffer("B_MIN", -torch.Tensor([r] * 3), persistent=False)
        self.local_batch = None
        self.sample_mode = False
        # arithmetic of the 256x256 hidden layers: "tensor_cores" (tcgen05 split-bf16, default) or
        # "fp32" (exact-fp32 FFMA kernel) — include/e3dge_b200.h E3_RENDER_FP32_CUDA_CORES
        self.backend = "tensor_cores"
        self._last_names = None

    # ------------------------------------------------------------------ internals
    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): new storages -> new packed image
        _lib.invalidate_packed()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        _lib.invalidate_packed()
        return super()._load_from_state_dict(*args, **kwargs)

    def train(self, mode=True):
        _lib.invalidate_packed()
        return super().train(mode)

    def invalidate_packed(self):
        """Call after updating weights in place through `.data` (EMA `accumulate`, Ranger): such updates
        do not bump the version counters the packed image is keyed on."""
        _lib.invalidate_packed()

    @property
    def siren(self):
        return self.network.netGlobal if self.enable_local_model else self.network

    def packed_weights(self):
        return self.siren.packed_weights()

    def _film(self, styles):
        return self.siren.film(styles)

    def _point_flags(self):
        return _lib.RENDER_FP32_CUDA_CORES if self.backend == "fp32" else 0

    def _flags(self, no_force_stop=False):
        f = 0
        if self.static_viewdirs:
            f |= _lib.RENDER_STATIC_VIEWDIRS
        if self.force_background:
            f |= _lib.RENDER_FORCE_BACKGROUND
        if no_force_stop:
            f |= _lib.RENDER_NO_FORCE_STOP
        if not self.with_sdf:
            f |= _lib.RENDER_NO_SDF
        if self.backend == "fp32":
            f |= _lib.RENDER_FP32_CUDA_CORES
        return f

    def _render_raw(self, styles, cam_poses, focal, near, far, z_jitter=None, local_mod=None,
                    flags_over=None, want_taps=False, film=None, train=False):
        import ctypes
        lib = _lib.load()
        dev = cam_poses.device
        B = cam_poses.shape[0]
        n = self.out_im_res * self.spatial_ss
        S = self.N_samples
        if film is None:
            film = self._film(styles)
        cam = _lib.as_f32c(cam_poses[:, :3, :4])
        def vec(t):  # python float | 0-d | [B,1,1] ... -> contiguous [B]
            t = torch.as_tensor(t, device=dev, dtype=torch.float32).reshape(-1)
            if t.numel() == 1:
                t = t.expand(B)
            if t.numel() != B:
                raise RuntimeError(f"expected one value per image ({B}), got {t.numel()}")
            return t.contiguous()

        focal_v, near_v, far_v = vec(focal), vec(near), vec(far)
        new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        o = dict(features=new(B, 256, n, n), gen_thumb_imgs=new(B, 3, n, n), xyz=new(B, 3, n, n),
                 mask=new(B, 1, n, n, 1), depth=new(B, n, n, 1, 1), sdf=new(B, n, n, S, 1),
                 hit_prob=new(B, n, n, S, 1), visibility=new(B, n, n, S, 1), dists=new(B, n, n, S),
                 points=new(B, n, n, S, 3), rays_o=new(B, n, n, 3), rays_d=new(B, n, n, 3),
                 viewdirs=new(B, n, n, 3), raw_rgb=new(B, n, n, S, 3))
        if want_taps:
            o["all_feats"] = new(4, B, n, n, S, 256)
        prm = _lib.RenderParams(B, n, n, self.out_im_res, S,
                                self._flags() if flags_over is None else flags_over,
                                float(self.grid_warper.scale_factor), 1.08)
        tv = _lib.as_f32c(self.t_vals.reshape(-1))
        pix = _lib.as_f32c(self._pix)
        sb = _lib.as_f32c(self.sigmoid_beta.detach()) if self.with_sdf else None
        la = lb = None
        if local_mod is not None:
            la, lb = (_lib.as_f32c(t) for t in local_mod)
            if tuple(la.shape) != (B, n, n, S, 256) or la.shape != lb.shape:
                raise RuntimeError("local texture modulation must be two [B,H,W,S,256] tensors")
        zj = _lib.as_f32c(z_jitter) if z_jitter is not None else None
        inp = _lib.RenderInputs(*[_lib.ptr(t) for t in (cam, focal_v, near_v, far_v, pix, pix, tv, zj,
                                                        sb, film, la, lb)])
        stash = None
        if train:
            if self.backend != "tensor_cores":
                raise RuntimeError("e3dge_b200: training (backward) needs backend='tensor_cores'")
            stash = torch.empty(max(lib.e3_render_stash_bytes(S, n * n, B) // 4, 1), device=dev,
                                dtype=torch.float32)
        outs = _lib.RenderOutputs(*[_lib.ptr(o[k]) for k in (
            "features", "gen_thumb_imgs", "xyz", "mask", "depth", "sdf", "hit_prob", "visibility",
            "dists", "points", "rays_o", "rays_d", "viewdirs", "raw_rgb")],
            _lib.ptr(o["all_feats"]) if want_taps else None, _lib.ptr(stash))
        _lib.check(lib.e3_render_fwd(_lib.ptr(self.packed_weights()), ctypes.byref(prm),
                                     ctypes.byref(inp), ctypes.byref(outs), _lib.cur_stream()),
                   "e3_render_fwd")
        o["near"] = near_v.reshape(B, 1, 1, 1).expand(B, n, n, 1)
        o["far"] = far_v.reshape(B, 1, 1, 1).expand(B, n, n, 1)
        if train:
            # everything the backward call reads again (keeps the argument tensors alive)
            call = dict(prm=prm, inp=inp, film=film, stash=stash, la=la, lb=lb,
                        keep=(cam, focal_v, near_v, far_v, pix, tv, zj, sb))
            return o, call
        return o

    def _make_z_jitter(self, near, far, B, dev):
        """perturb > 0 (training): the random offsets of volume_renderer.py:1213-1228, drawn
        with torch on the device and handed to the kernel as explicit z values."""
        n = self.out_im_res * self.spatial_ss
        nr = torch.as_tensor(near, device=dev, dtype=torch.float32).reshape(-1, 1, 1, 1)
        fr = torch.as_tensor(far, device=dev, dtype=torch.float32).reshape(-1, 1, 1, 1)
        z = (nr * (1. - self.t_vals) + fr * self.t_vals).expand(B, n, n, self.N_samples)
        if self.offset_sampling:
            upper = torch.cat([z[..., 1:], fr.expand(B, n, n, 1)], -1)
            lower = z
            t_rand = torch.rand(B, n, n, 1, device=dev)
        else:
            mids = .5 * (z[..., 1:] + z[..., :-1])
            upper = torch.cat([mids, z[..., -1:]], -1)
            lower = torch.cat([z[..., :1], mids], -1)
            t_rand = torch.rand(z.shape, device=dev)
        return (lower + (upper - lower) * t_rand).contiguous()

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def get_rays(self, focal, c2w, dirs=None):
        """volume_renderer.py:769-794 (stand-alone use; `forward` computes rays in-kernel)."""
        if dirs is None:
            n = self.out_im_res * self.spatial_ss
            dirs = torch.stack([(self.i - self.out_im_res * .5) / focal,
                                -(self.j - self.out_im_res * .5) / focal,
                                -torch.ones_like(self.i).expand(focal.shape[0], n, n)], -1)
        rays_d = torch.sum(dirs[..., None, :] * c2w[:, None, None, :3, :3], -1)
        rays_o = c2w[:, None, None, :3, -1].expand(rays_d.shape)
        return rays_o, rays_d, (dirs if self.static_viewdirs else rays_d)

    def sdf_activation(self, input):
        return torch.sigmoid(input / self.sigmoid_beta) / self.sigmoid_beta

    def run_network(self, inputs, viewdirs, normalize=True, styles=None, global_only=False,
                    return_sdf_only=False, **kwargs):
        """FiLM-SIREN at explicit world-space samples — volume_renderer.py:1052-1128.
        inputs [B,H,W,S,3]; viewdirs [B,H,W,3] / [B,H,W,S,3]; returns raw [...,260]
        (rgb | sdf | features) or the sdf slice."""
        import ctypes  # noqa: F401
        lib = _lib.load()
        shp = inputs.shape
        B = shp[0]
        pts = _lib.as_f32c(inputs).reshape(B, -1, 3)
        N = pts.shape[1]
        if viewdirs.shape != inputs.shape:
            if viewdirs.ndim != inputs.ndim:
                viewdirs = viewdirs.unsqueeze(self.samples_dim)
            viewdirs = viewdirs.expand(shp)
        vd = _lib.as_f32c(viewdirs).reshape(B, -1, 3)
        if self._wants_grad(styles, inputs):
            film = _FilmFn.apply(self, styles)
            sdf, rgb, feat = _PointsFn.apply(self, film, pts, vd, not return_sdf_only)
            if return_sdf_only:
                return sdf.reshape(*shp[:-1], 1)
            return torch.cat([rgb, sdf.unsqueeze(-1), feat], -1).reshape(*shp[:-1], 260)
        film = self._film(styles)
        sdf = torch.empty(B, N, device=pts.device, dtype=torch.float32)
        rgb = feat = None
        if not return_sdf_only:
            rgb = torch.empty(B, N, 3, device=pts.device, dtype=torch.float32)
            feat = torch.empty(B, N, 256, device=pts.device, dtype=torch.float32)
        _lib.check(lib.e3_siren_points_fwd(_lib.ptr(self.packed_weights()), _lib.ptr(film),
                                           _lib.ptr(pts), _lib.ptr(vd), B, N,
                                           float(self.grid_warper.scale_factor), _lib.ptr(sdf),
                                           _lib.ptr(rgb), _lib.ptr(feat), self._point_flags(),
                                           _lib.cur_stream()),
                   "e3_siren_points_fwd")
        if return_sdf_only:
            return sdf.reshape(*shp[:-1], 1)
        return torch.cat([rgb, sdf.unsqueeze(-1), feat], -1).reshape(*shp[:-1], 260)

    def sdf_query(self, points, styles):
        """sdf [B,N,1] at world-space points [B,N,3] with zero view directions (the
        geometry queries of volume_renderer.py:955-957, 1935-1943), view layer skipped."""
        lib = _lib.load()
        pts = _lib.as_f32c(points)
        B, N = pts.shape[0], pts.shape[1]
        if self._wants_grad(styles, points):
            film = _FilmFn.apply(self, styles)
            return _PointsFn.apply(self, film, pts, None, False)[0].unsqueeze(-1)
        film = self._film(styles)
        sdf = torch.empty(B, N, device=pts.device, dtype=torch.float32)
        _lib.check(lib.e3_siren_points_fwd(_lib.ptr(self.packed_weights()), _lib.ptr(film),
                                           _lib.ptr(pts), None, B, N,
                                           float(self.grid_warper.scale_factor), _lib.ptr(sdf), None,
                                           None, self._point_flags(), _lib.cur_stream()),
                   "e3_siren_points_fwd")
        return sdf.unsqueeze(-1)

    @staticmethod
    def _wants_grad(*tensors):
        return torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors)

    @torch.no_grad()
    def sdf_and_gradient(self, points, styles):
        """(sdf [B,N,1], d sdf / d point [B,N,3]) — the value of get_eikonal_term
        (volume_renderer.py:796-802): one sdf-only forward with stash + e3_siren_points_bwd seeded
        with dL/dsdf = 1.  Returned without a graph (no second-order gradients)."""
        lib = _lib.load()
        pts = _lib.as_f32c(points.detach())
        B, N = pts.shape[0], pts.shape[1]
        dev = pts.device
        film = self._film(styles.detach())
        sdf = torch.empty(B, N, device=dev, dtype=torch.float32)
        stash = torch.empty(max(lib.e3_render_stash_bytes(1, N, B) // 4, 1), device=dev, dtype=torch.float32)
        scale = float(self.grid_warper.scale_factor)
        _lib.check(lib.e3_siren_points_fwd_train(_lib.ptr(self.packed_weights()), _lib.ptr(film), _lib.ptr(pts),
                                                 None, B, N, scale, _lib.ptr(sdf), None, None,
                                                 _lib.ptr(stash), _lib.cur_stream()),
                   "e3_siren_points_fwd_train")
        d_film = torch.empty(B, 9, 2, 256, device=dev, dtype=torch.float32)
        d_pts = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
        scratch, nbytes = _scratch(B, dev)
        _lib.check(lib.e3_siren_points_bwd(_lib.ptr(self.packed_weights()), _lib.ptr(film), B, N, scale,
                                           _lib.ptr(stash), 0, 1, None, None, None, _lib.ptr(d_film),
                                           _lib.ptr(d_pts), _lib.ptr(scratch), nbytes, _lib.cur_stream()),
                   "e3_siren_points_bwd")
        return sdf.unsqueeze(-1), d_pts

    def sample_uniform_grid(self, batch_size, num_sample_inout, device, styles):
        """volume_renderer.py:945-963."""
        length = self.B_MAX - self.B_MIN
        pts = torch.rand(batch_size, num_sample_inout, 3, device=device) * length + self.B_MIN
        sdf = self.sdf_query(pts, styles)
        return pts, sdf, torch.ones_like(sdf)

    def sample_near_surface_grid(self, surface_points, viewdirs, normal_stdv, styles, multiplier=1):
        """volume_renderer.py:965-1003 (surface_points [B,H,W,3])."""
        pert = torch.randn_like(surface_points) * normal_stdv
        pts = (surface_points + pert).unsqueeze(-2)
        valid = (torch.abs(pts).max(dim=-1)[0] < self.opt.camera.dist_radius).int()
        sdf = self.run_network(pts, viewdirs, styles=styles)[..., 3]
        return pts, sdf, valid

    def render(self, focal, c2w, near, far, styles, return_eikonal=False, return_mesh=False,
               mesh_with_shading=True, **kwargs):
        """volume_renderer.py:1666-1701."""
        B, dev = c2w.shape[0], c2w.device
        zj = self._make_z_jitter(near, far, B, dev) if (self.perturb and self.perturb > 0) else None
        local_mod = kwargs.get("local_tex_modulation")
        want_taps = bool(getattr(self.opt, "return_feats", False))
        need_grad = self._wants_grad(styles, *(local_mod or ()))
        if need_grad:
            film = _FilmFn.apply(self, styles)
            la, lb = local_mod if local_mod is not None else (None, None)
            vals = _RenderFn.apply(self, film, la, lb, c2w, focal, near, far, zj, None, want_taps)
            o = dict(zip(self._last_names, vals))
        else:
            o = self._render_raw(styles, c2w, focal, near, far, zj, local_mod, None, want_taps)
        n = self.out_im_res * self.spatial_ss
        eik = surf_eik = None
        if return_eikonal or kwargs.get("return_surface_eikonal", False):
            # d sdf / d sample position (volume_renderer.py:796-802, 855-856); with latents that require grad it
            # carries a graph to them (create_graph=True semantics: eikonal.py)
            eik = eikonal_term(self, o["points"].detach().reshape(B, -1, 3), styles).reshape(B, n, n, -1, 3)
        if kwargs.get("return_surface_eikonal", False):
            if getattr(self.opt, "use_integrated_surface_normal", False):
                surf_eik = torch.sum(o["hit_prob"].detach() * eik, 3).unsqueeze(-2)  # :932-934
            else:  # sdf gradient at the integrated surface point (:921-930)
                xyz_pts = o["xyz"].detach().permute(0, 2, 3, 1).reshape(B, -1, 3)
                surf_eik = eikonal_term(self, xyz_pts, styles).reshape(B, n, n, 1, 3)
        if not return_eikonal:
            eik = None
        out = {
            "rays_o": o["rays_o"], "rays_d": o["rays_d"], "dists": o["dists"], "near": o["near"],
            "far": o["far"], "hit_prob": o["hit_prob"], "surface_eikonal_term": surf_eik,
            "points": o["points"], "sdf": o["sdf"] if self.return_sdf else None,
            "gen_thumb_imgs": o["gen_thumb_imgs"],
            "features": o["features"] if self.output_features else None,
            "mask": o["mask"] if self.return_xyz else None,
            "xyz": o["xyz"] if self.return_xyz else None, "eikonal_term": eik,
            "depth": o["depth"] if self.return_xyz else None, "mesh": None, "shading_mesh": None,
            "debug_mesh": None, "viewdirs": o["viewdirs"],
        }
        if return_mesh:
            # volume_renderer.py:1703-1727: frustum-aligned SDF volume -> marching cubes on the host (skimage,
            # like the reference; mesh_utils.py raises ImportError when it is not installed).  Batch 1 only (:1736).
            from .mesh_utils import align_volume, extract_mesh_with_marching_cubes
            out.pop("shading_mesh"), out.pop("debug_mesh")
            try:
                mesh, verts, faces = extract_mesh_with_marching_cubes(align_volume(o["sdf"].detach()))
                out["mesh"], out["shaded_mesh"] = mesh, mesh
            except ValueError:  # no zero crossing in the volume
                print("Marching cubes extraction failed.")
                print("Please check whether the SDF values are all larger (or all smaller) than 0.")
                out["mesh"], out["shaded_mesh"] = None, None
        if want_taps:
            out["all_feats"] = list(o["all_feats"].unbind(0))
        if kwargs.get("sample_without_grad", False):
            out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
        out["_visibility"] = o["visibility"]
        out["_raw_rgb"] = o["raw_rgb"]
        return out

    def forward(self, cam_poses, focal, near, far, styles=None, return_eikonal=False,
                geometry_sample=None, return_surface_eikonal=False, local_data_batch=None,
                sample_mode=False, return_mesh=False, mesh_with_shading=True,
                return_sdf_only=False, **kwargs):
        """volume_renderer.py:1865-1972.  Maps are returned NCHW exactly as the reference
        does after its permutes (:1957-1968); the kernel writes them in that layout."""
        self.sample_mode = sample_mode
        self.local_batch = local_data_batch if self.enable_local_model else None
        if local_data_batch is not None and "tex_modulation" in local_data_batch:
            kwargs.setdefault("local_tex_modulation", local_data_batch["tex_modulation"])
        elif (local_data_batch is not None and local_data_batch.get("feats") is not None and not sample_mode
              and getattr(self.network, "netLocal", None) is not None
              and getattr(self.opt, "L_pred_tex_modulations", True)):
            # SirenLocalGlobal.forward_backbone (volume_renderer.py:323-336): netLocal maps the per-sample local
            # features [B,H,W,S,301] to the texture modulation (alpha | beta) — stages 5-6 of e3_local_mlp_fwd
            # at inference, the module's own autograd-capable forward when it is being trained
            net_local = self.network.netLocal
            if hasattr(self.network, "_tex_conditions") and hasattr(net_local, "local_feat_to_tex_modulations_linear"):
                mods = tuple(self.network._tex_conditions(local_data_batch))
            else:  # a caller-supplied netLocal
                mods = tuple(torch.split(net_local.local_feat_to_tex_modulations_linear(local_data_batch["feats"]),
                                         256, dim=-1))
            kwargs.setdefault("local_tex_modulation", mods)
        out = self.render(focal, c2w=cam_poses, near=near, far=far, styles=styles,
                          return_eikonal=return_eikonal,
                          return_surface_eikonal=return_surface_eikonal, return_mesh=return_mesh,
                          mesh_with_shading=mesh_with_shading, **kwargs)
        out.pop("_visibility", None)
        out.pop("_raw_rgb", None)
        if geometry_sample:
            for k in ["uniform_pts"] + (["xyz"] if geometry_sample.get("xyz") is not None else []):
                if k not in geometry_sample:
                    continue
                samples = geometry_sample[k]
                if samples.ndim == 4:
                    samples = samples.unsqueeze(self.samples_dim)
                shp = samples.shape
                sdf = self.sdf_query(samples.reshape(shp[0], -1, 3), styles)
                out[f"{k}_rec"] = sdf.reshape(*shp[:-1], 1)
                if return_surface_eikonal and k == "xyz":  # :1945-1949
                    out[f"{k}_rec_eikonal_term"] = eikonal_term(
                        self, samples.detach().reshape(shp[0], -1, 3), styles).reshape(*shp[:-1], 3)
        if sample_mode:
            out = self._sample_and_collate(out, styles)
            self.sample_mode = False
        return out

    def _sample_and_collate(self, out, styles):
        """sample_mode tail of render_rays + collate_fn — volume_renderer.py:1296-1324,1976-2043.
        (In sample mode the reference leaves xyz / mask in [B,H,W,.] layout.)"""
        B = out["gen_thumb_imgs"].shape[0]
        dev = out["gen_thumb_imgs"].device
        pts_l, sdf_l, msk_l = [], [], []
        xyz_hw = out["xyz"].permute(0, 2, 3, 1).contiguous()
        if self.opt.sample_near_surface:
            p, s, m = self.sample_near_surface_grid(xyz_hw, out["viewdirs"],
                                                    self.opt.surface_sampling_stdv, styles)
            out.update(points_near_surface=p, points_near_surface_sdf=s,
                       points_near_surface_valid_mask=m)
            pts_l.append(p.reshape(B, -1, 3)), sdf_l.append(s.reshape(B, -1, 1))
            msk_l.append(m.reshape(B, -1, 1).float())
        if self.opt.sample_uniform_grid:
            p, s, m = self.sample_uniform_grid(B, self.opt.uniform_grid_sampling_num, dev, styles)
            out.update(grid_random_pts=p, grid_random_pts_sdf=s, grid_sample_valid_mask=m)
            pts_l.append(p.reshape(B, -1, 3)), sdf_l.append(s.reshape(B, -1, 1))
            msk_l.append(m.reshape(B, -1, 1))
        cat = lambda l, c: torch.cat(l, 1) if l else torch.empty(B, 0, c, device=dev)
        out["uniform_pts"] = cat(pts_l, 3).reshape(B, -1, 1, 1, 3)
        out["uniform_points_sdf"] = cat(sdf_l, 1).reshape(B, -1, 1, 1, 1)
        out["uniform_points_valid_mask"] = cat(msk_l, 1).reshape(B, -1, 1, 1, 1)
        out["xyz"] = xyz_hw
        out["mask"] = out["mask"].permute(0, 2, 3, 4, 1).contiguous()
        return out

    def sdf_sample_pass(self, cam_poses, focal, near, far, styles, return_grad=False,
                        merge_spatial_dim=True):
        """sdf at stratified-jittered samples along the camera rays — volume_renderer.py:1760-1831.
        (The reference's version dereferences an undefined `normalized_pts` (:1811) and cannot run;
        this one returns what its docstring promises: box-normalised points [B,3,N] and sdf [B,1,N].)"""
        B, dev = cam_poses.shape[0], cam_poses.device
        rays_o, rays_d, _ = self.get_rays(focal, cam_poses)
        nr = near.unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        fr = far.unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        z = nr * (1. - self.t_vals) + fr * self.t_vals
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * torch.rand(z.shape, device=dev)
        pts = rays_o.unsqueeze(3) + rays_d.unsqueeze(3) * z.unsqueeze(-1)
        normalized = self.grid_warper(pts)
        if return_grad:  # `sdf_norm` of :1824 = d sdf / d (world-space point)
            sdf, grad = self.sdf_and_gradient(pts.reshape(B, -1, 3), styles)
            extra = {"sdf_grad": grad.permute(0, 2, 1) if merge_spatial_dim else grad.reshape(*z.shape, 3)}
        else:
            sdf, extra = self.sdf_query(pts.reshape(B, -1, 3), styles), {}
        if merge_spatial_dim:
            return {"points": normalized.reshape(B, -1, 3).permute(0, 2, 1), "sdf": sdf.reshape(B, 1, -1),
                    **extra}
        return {"points": normalized, "sdf": sdf.reshape(z.shape), **extra}

    def mlp_init_pass(self, cam_poses, focal, near, far, styles=None):
        """Sphere-init pass: sdf at stratified-jittered samples and its target
        |p| - (far-near)/4 — volume_renderer.py:1833-1863."""
        B, dev = cam_poses.shape[0], cam_poses.device
        rays_o, rays_d, _ = self.get_rays(focal, cam_poses)
        nr = near.unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        fr = far.unsqueeze(-1) * torch.ones_like(rays_d[..., :1])
        z = nr * (1. - self.t_vals) + fr * self.t_vals
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * torch.rand(z.shape, device=dev)
        pts = rays_o.unsqueeze(3) + rays_d.unsqueeze(3) * z.unsqueeze(-1)
        sdf = self.sdf_query(pts.reshape(B, -1, 3), styles).reshape(z.shape)
        return sdf, pts.detach().norm(dim=-1) - ((fr - nr) / 4)

    # ------------------------------------------------------------------ visibility queries (a17)
    def _weights_from_sdf(self, sdf, z_vals, rays_d_norm, no_force_stop=True):
        """The density half of volume_integration (volume_renderer.py:822-886) for sdf [...,S,1],
        z_vals [...,S], rays_d_norm [...,1]: (visibility, weights), both [...,S,1]."""
        dists = z_vals[..., 1:] - z_vals[..., :-1]
        tail = dists[..., 0:1] if no_force_stop else self.inf.expand(dists[..., 0:1].shape)
        dists = torch.cat([dists, tail], -1) * rays_d_norm
        if self.with_sdf:
            alpha = 1 - torch.exp(-self.sdf_activation(-sdf) * dists.unsqueeze(-1))
        else:
            alpha = 1 - torch.exp(-F.softplus(sdf) * dists.unsqueeze(-1))
        vis = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1, :]), 1. - alpha + 1e-10], -2), -2)
        vis = vis[..., :-1, :]
        w = alpha * vis
        if self.force_background and not no_force_stop:
            w = torch.cat([w[..., :-1, :], 1 - w[..., :-1, :].sum(-2, keepdim=True)], -2)
        return vis, w

    def _reference_view_rays(self, wd_space_pts, ref_img_info):
        """Shared head of the two visibility queries (volume_renderer.py:1340-1376, 1512-1545): the ray
        from the reference camera through every query point.  Returns points [B,N,S,1,3], the ray
        origin [B,1,1,1,3], world-space directions [B,N,S,1,3] (unit depth along -z of the reference
        camera) and the points in the reference camera frame [B,N,S,3]."""
        B, H, W, S = wd_space_pts.shape[:4]
        poses = ref_img_info["cam_settings"]["poses"][:, :3, :4].float()
        extr = ref_img_info["cam_settings"]["extrinsics"][:, :3, :4].float()
        pts = wd_space_pts.reshape(B, H * W, S, 3).float()
        ref = torch.einsum("bij,bnsj->bnsi", extr[:, :, :3], pts) + extr[:, None, None, :, 3]
        d_ref = ref / (-ref[..., 2:3])
        d_wd = torch.einsum("bij,bnsj->bnsi", poses[:, :, :3], d_ref)
        rays_o = poses[:, :, 3].reshape(B, 1, 1, 1, 3)
        return pts.unsqueeze(-2), rays_o, d_wd.unsqueeze(-2), d_ref, ref

    def _march_reference_rays(self, ray_pts, z_vals, viewdirs, styles):
        """sdf along the query rays -> (visibility, weights) [B,N,S,T,1] with the no_force_stop composite
        (volume_renderer.py:1428-1470).  Only the density is needed, so the sdf-only kernel runs (the
        reference evaluates the full network and discards rgb / features)."""
        B, N, S, T = ray_pts.shape[:4]
        vis = torch.empty(B, N, S, T, 1, device=ray_pts.device)
        w = torch.empty_like(vis)
        step = 64 ** 2
        for lo in range(0, N, step):
            chunk = ray_pts[:, lo:lo + step]
            sdf = self.sdf_query(chunk.reshape(B, -1, 3), styles).reshape(*chunk.shape[:4], 1)
            ones = torch.ones_like(z_vals[:, lo:lo + step, :, :1])  # normalised view dirs: |d| = 1
            vis[:, lo:lo + step], w[:, lo:lo + step] = self._weights_from_sdf(sdf, z_vals[:, lo:lo + step], ones)
        return vis, w

    @torch.no_grad()
    def query_hitting_probability_fixed_interval(self, wd_space_pts, ref_img_info, return_type="weights"):
        """Hit probability / visibility of world-space points [B,H,W,S,3] seen from a reference view:
        re-march the reference camera's ray through every point with the renderer's own depth samples
        and interpolate at the point's depth — volume_renderer.py:1326-1493."""
        assert return_type in ("weights", "visibility")
        assert wd_space_pts.ndim == 5
        B, H, W, S = wd_space_pts.shape[:4]
        styles = ref_img_info["pred_latents"][0]
        out = ref_img_info["global_render_out"]
        near = out["near"].reshape(B, H * W, 1, 1, 1).float()
        far = out["far"].reshape(B, H * W, 1, 1, 1).float()
        pts, rays_o, d_wd, d_ref, _ = self._reference_view_rays(wd_space_pts, ref_img_info)
        t = self.t_vals.reshape(1, 1, 1, 1, -1)
        z = near * (1. - t) + far * t                                     # [B,N,1,1,T]
        interval = (z[..., 1:2] - z[..., 0:1]) * d_wd.norm(dim=-1, keepdim=True)  # [B,N,S,1,1]
        z = z.permute(0, 1, 2, 4, 3)                                      # [B,N,1,T,1]
        ray_pts = rays_o + d_wd * z                                       # [B,N,S,T,3]
        idx = (pts - ray_pts[..., 0:1, :]).norm(dim=-1, keepdim=True) / interval + 1e-5  # [B,N,S,1,1]
        T = self.t_vals.shape[-1]
        lo_i = idx.floor().long().clamp(0, T - 1)
        hi_i = idx.ceil().long().clamp(0, T - 1)
        viewdirs = F.normalize(d_ref if self.static_viewdirs else d_wd.squeeze(-2), dim=-1)
        vis, w = self._march_reference_rays(ray_pts, z.squeeze(-1).expand(B, H * W, S, T), viewdirs, styles)
        info = w if return_type == "weights" else vis
        lo_v = torch.gather(info, 3, lo_i)
        hi_v = torch.gather(info, 3, hi_i)
        return torch.lerp(lo_v, hi_v, idx - lo_i).reshape(B, H, W, S, 1)

    @torch.no_grad()
    def query_hitting_probability_adapted_interval(self, wd_space_pts, ref_img_info):
        """Same question, marching N_samples steps from the reference near plane exactly up to each
        point and returning the last sample's weight — volume_renderer.py:1495-1621."""
        assert wd_space_pts.ndim == 5
        B, H, W, S = wd_space_pts.shape[:4]
        styles = ref_img_info["pred_latents"][0]
        near = ref_img_info["global_render_out"]["near"].reshape(B, H * W, 1, 1, 1).float()
        pts, rays_o, d_wd, d_ref, _ = self._reference_view_rays(wd_space_pts, ref_img_info)
        near_pts = rays_o + d_wd * near                                   # [B,N,S,1,3]
        t = torch.linspace(0., 1., steps=self.N_samples, device=pts.device).reshape(1, 1, 1, -1, 1)
        ray_pts = near_pts * (1 - t) + pts * t                            # [B,N,S,T,3]
        z = (ray_pts - rays_o).norm(dim=-1)                               # [B,N,S,T]
        viewdirs = F.normalize(d_ref if self.static_viewdirs else d_wd.squeeze(-2), dim=-1)
        _, w = self._march_reference_rays(ray_pts, z, viewdirs, styles)
        return w[..., -1:, :].reshape(B, H, W, S, 1)

    def volume_integration(self, raw, z_vals, rays_d, pts, return_eikonal=False,
                           return_surface_eikonal=False, return_mesh=True, c2w=None,
                           no_force_stop=False, **kwargs):
        """Stand-alone composite of a caller-provided `raw` (volume_renderer.py:809-943); only
        callers outside the fused kernel use it, so it stays device-side PyTorch host code.  The hot path composites inside the kernel."""
        if isinstance(raw, dict):
            raw = raw["raw"]
        if return_eikonal or return_surface_eikonal:
            raise NotImplementedError("eikonal terms need the renderer backward")
        dists = z_vals[..., 1:] - z_vals[..., :-1]
        rd_norm = torch.norm(rays_d.unsqueeze(3), dim=-1)
        tail = dists[..., 0:1] if no_force_stop else self.inf.expand(rd_norm.shape)
        dists = torch.cat([dists, tail], -1) * rd_norm
        rgb, sdf, feats = torch.split(raw, [3, 1, self.feature_out_size], dim=-1)
        if self.with_sdf:
            alpha = 1 - torch.exp(-self.sdf_activation(-sdf) * dists.unsqueeze(-1))
        else:
            alpha = 1 - torch.exp(-F.softplus(sdf) * dists.unsqueeze(-1))
        vis = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1, :]), 1. - alpha + 1e-10], 3), 3)
        vis = vis[..., :-1, :]
        w = alpha * vis
        if self.force_background and not no_force_stop:
            w = torch.cat([w[..., :-1, :], 1 - w[..., :-1, :].sum(3, keepdim=True)], 3)
        rgb_map = -1 + 2 * torch.sum(w * torch.sigmoid(rgb), 3)
        feat_map = torch.sum(w * feats, 3)
        xyz = depth = mask = None
        if self.return_xyz and pts is not None:
            xyz = torch.sum(w * pts, 3)
            depth = torch.sum(w * z_vals.unsqueeze(-1), 3, keepdim=True)
            mask = (depth < 1.08).type_as(w)
        return rgb_map, feat_map, sdf, mask, xyz, None, None, rd_norm, depth, dists, vis, w
