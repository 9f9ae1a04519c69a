"""`project.models.stylesdf_model` surface: Generator / G_pred_latents / Decoder and their
building blocks, on top of the sm_100a kernels.

Class names, constructor arguments, forward signatures, attribute names and state_dict keys
follow the reference (project/models/stylesdf_model.py:30-1172) so StyleSDF `g_ema`
checkpoints load unchanged and the runners can call `generator(...)` as they do today
(trainer.py:881, 1399).  Discriminators and the legacy encoders of that file (lines
1193-1765) are not part of the generator path and are not provided.

Inside `Decoder.forward` activations stay channels-last fp32 between layers; the NCHW
contract of the reference holds at the module boundaries.
"""
import ctypes
import math
import random

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib
from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d
from .volume_renderer import VolumeFeatureRenderer


class PixelNorm(nn.Module):
    """stylesdf_model.py:30-37."""

    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


class MappingLinear(nn.Module):
    """stylesdf_model.py:40-82 (z -> w mapping of the renderer; fused lrelu with scale 1)."""

    def __init__(self, in_dim, out_dim, bias=True, activation=None, is_last=False):
        super().__init__()
        std = 0.25 if is_last else 1
        w = torch.empty(out_dim, in_dim)
        nn.init.kaiming_normal_(w, a=0.2, mode="fan_in", nonlinearity="leaky_relu")
        self.weight = nn.Parameter(std * w)
        lim = math.sqrt(1 / in_dim)
        self.bias = nn.Parameter(torch.empty(out_dim).uniform_(-lim, lim)) if bias else None
        self.activation = activation

    def forward(self, input):
        if self.activation is not None:
            return fused_leaky_relu(F.linear(input, self.weight), self.bias, scale=1)
        return F.linear(input, self.weight, bias=self.bias)


def make_kernel(k):
    """stylesdf_model.py:85-93."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


class Upsample(nn.Module):
    """stylesdf_model.py:96-119."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel) * (factor ** 2)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    """stylesdf_model.py:122-145."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    """stylesdf_model.py:148-165."""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualLinear(nn.Module):
    """stylesdf_model.py:210-249."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if self.activation:
            out = F.linear(input, self.weight * self.scale)
            return fused_leaky_relu(out, self.bias * self.lr_mul)
        return F.linear(input, self.weight * self.scale, bias=self.bias * self.lr_mul)


class _PackedConv:
    """Kernel-layout image of one conv weight + sum_k W^2, rebuilt when the weight changes."""

    def __init__(self):
        self.key = None
        self.wp = self.wsq = None

    def get(self, weight, layout):
        """layout: 0 plain / 1 upsampling forward, 2 / 3 their backward images (E3_CONV_PACK_*)."""
        layout = int(layout)
        key = (weight.data_ptr(), weight._version, layout, _lib.pack_epoch)
        if key == self.key:
            return self.wp, self.wsq
        lib = _lib.load()
        w = _lib.as_f32c(weight.detach())
        _, cout, cin, k, _ = w.shape
        wsq = torch.empty(cout, cin, device=w.device, dtype=torch.float32)
        _lib.check(lib.e3_modconv_weight_sq(_lib.ptr(w), cout, cin, k, _lib.ptr(wsq),
                                            _lib.cur_stream()), "e3_modconv_weight_sq")
        wp = None
        if k == 3:
            wp = torch.empty(lib.e3_conv_packed_bytes(cout, cin) // 4, device=w.device,
                             dtype=torch.float32)
            _lib.check(lib.e3_conv_pack_weight(_lib.ptr(w), cout, cin, layout,
                                               _lib.ptr(wp), _lib.cur_stream()),
                       "e3_conv_pack_weight")
        self.key, self.wp, self.wsq = key, wp, wsq
        return wp, wsq


CONV_BACKENDS = {"auto": _lib.CONV_AUTO, "fp32": _lib.CONV_FP32_CUDA_CORES,
                 "tensor_cores": _lib.CONV_TENSOR_CORES}


def _to_nhwc(x):
    lib = _lib.load()
    x = _lib.as_f32c(x)
    b, c, h, w = x.shape
    y = torch.empty(b, h, w, c, device=x.device, dtype=torch.float32)
    _lib.check(lib.e3_nchw_to_nhwc(_lib.ptr(x), _lib.ptr(y), b, c, h, w, _lib.cur_stream()),
               "e3_nchw_to_nhwc")
    return y


def _to_nchw(x):
    lib = _lib.load()
    b, h, w, c = x.shape
    y = torch.empty(b, c, h, w, device=x.device, dtype=torch.float32)
    _lib.check(lib.e3_nhwc_to_nchw(_lib.ptr(x), _lib.ptr(y), b, c, h, w, _lib.cur_stream()),
               "e3_nhwc_to_nchw")
    return y


def _wants_grad(*tensors):
    return torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors)


def _latent_grad(conv, ds, dd, s, d):
    """(ds, dd) -> dlatent [B,512] (e3_modconv_styles_bwd)."""
    lib = _lib.load()
    b = ds.shape[0]
    _, wsq = conv._packed.get(conv.weight, int(conv.upsample))
    dlat = torch.empty(b, 512, device=ds.device, dtype=torch.float32)
    _lib.check(lib.e3_modconv_styles_bwd(_lib.ptr(ds), _lib.ptr(dd), _lib.ptr(s), _lib.ptr(d),
                                         _lib.ptr(wsq), _lib.ptr(_lib.as_f32c(conv.modulation.weight.detach())),
                                         b, conv.in_channel, conv.out_channel, conv.kernel_size,
                                         _lib.ptr(dlat), _lib.cur_stream()), "e3_modconv_styles_bwd")
    return dlat


class _StyledConvFn(torch.autograd.Function):
    """StyledConv / bare ModulatedConv2d on NHWC tensors with its backward
    (e3_styled_conv3x3_bwd + e3_modconv_styles_bwd): gradients for the input and the latent."""

    @staticmethod
    def forward(ctx, conv, x, latent, noise, noise_w, act_bias):
        _lib.warn_if_trainable(conv, "ModulatedConv2d")
        y, saved = _styled_conv_nhwc(conv, x, latent, noise, noise_w, act_bias, want_saved=True)
        ctx.conv, ctx.saved = conv, saved
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        conv, (s, d, has_d, noise, nstride, noise_w, act_bias) = ctx.conv, ctx.saved
        x, y = ctx.saved_tensors
        b, h, w, cin = x.shape
        cout, up = conv.out_channel, int(conv.upsample)
        dy = _lib.as_f32c(dy)
        wp, _ = conv._packed_bwd.get(conv.weight, 2 + up)
        dx = torch.empty_like(x)
        ds = torch.empty(b, cin, device=x.device, dtype=torch.float32)
        dd = torch.empty(b, cout, device=x.device, dtype=torch.float32) if has_d else None
        nbytes = lib.e3_styled_conv_bwd_scratch_bytes(b, h, w, cin, cout, up)
        scratch = torch.empty(max(nbytes // 4, 1), device=x.device, dtype=torch.float32)
        _lib.check(lib.e3_styled_conv3x3_bwd(
            _lib.ptr(dy), _lib.ptr(y), _lib.ptr(x), _lib.ptr(wp), _lib.ptr(s), _lib.ptr(d), _lib.ptr(noise),
            nstride, _lib.ptr(noise_w), _lib.ptr(act_bias), _lib.ptr(dx), _lib.ptr(ds), _lib.ptr(dd), b, h, w,
            cin, cout, up, _lib.ptr(scratch), nbytes, CONV_BACKENDS[conv.backend], _lib.cur_stream()),
            "e3_styled_conv3x3_bwd")
        dlat = _latent_grad(conv, ds, dd, s, d) if ctx.needs_input_grad[2] else None
        return None, dx, dlat, None, None, None


class _ToRGBFn(torch.autograd.Function):
    """ToRGB on an NHWC input with its backward (e3_torgb_bwd; the skip gradient is the
    e3_upfirdn2d adjoint of the FIR upsampling)."""

    @staticmethod
    def forward(ctx, conv, x, latent, bias, skip, upsample_skip, up_kernel):
        rgb, s = _torgb_nhwc(conv, x, latent, bias, skip, upsample_skip, want_saved=True)
        ctx.conv, ctx.s = conv, s
        ctx.save_for_backward(x)
        ctx.has_skip, ctx.upsample_skip, ctx.up_kernel = skip is not None, bool(upsample_skip), up_kernel
        return rgb

    @staticmethod
    def backward(ctx, drgb):
        lib = _lib.load()
        conv, s = ctx.conv, ctx.s
        x, = ctx.saved_tensors
        b, h, w, cin = x.shape
        drgb = _lib.as_f32c(drgb)
        dx = torch.empty_like(x)
        ds = torch.empty(b, cin, device=x.device, dtype=torch.float32)
        nbytes = lib.e3_torgb_bwd_scratch_bytes(b, cin)
        scratch = torch.empty(max(nbytes // 4, 1), device=x.device, dtype=torch.float32)
        wt = _lib.as_f32c(conv.weight.detach().reshape(3, cin))
        _lib.check(lib.e3_torgb_bwd(_lib.ptr(drgb), _lib.ptr(x), _lib.ptr(wt), _lib.ptr(s), _lib.ptr(dx),
                                    _lib.ptr(ds), b, h, w, cin, _lib.ptr(scratch), nbytes, _lib.cur_stream()),
                   "e3_torgb_bwd")
        dlat = _latent_grad(conv, ds, None, s, None) if ctx.needs_input_grad[2] else None
        dskip = None
        if ctx.has_skip and ctx.needs_input_grad[4]:
            if ctx.upsample_skip:  # adjoint of upfirdn2d(up=2, pad=(2,1)): down=2, pad=(1,1), flipped kernel
                with torch.no_grad():
                    dskip = upfirdn2d(drgb, torch.flip(ctx.up_kernel, [0, 1]), up=1, down=2, pad=(1, 1))
            else:
                dskip = drgb
        return None, dx, dlat, None, dskip, None, None


class ModulatedConv2d(nn.Module):
    """stylesdf_model.py:263-362.  Parameters: weight [1,O,I,k,k], modulation.{weight,bias};
    `blur.kernel` buffer when upsample (kept for state_dict compatibility)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True,
                 upsample=False, downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if downsample:
            raise NotImplementedError("downsampling ModulatedConv2d is only used by the "
                                      "discriminators (out of scope)")
        if kernel_size not in (1, 3):
            raise NotImplementedError("kernel_size must be 1 (ToRGB) or 3")
        self.eps = 1e-8
        self.kernel_size, self.in_channel, self.out_channel = kernel_size, in_channel, out_channel
        self.upsample, self.downsample = upsample, downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1),
                             upsample_factor=factor)
            if list(blur_kernel) != [1, 3, 3, 1]:
                raise NotImplementedError("the fused up-conv is built for blur_kernel [1,3,3,1]")
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        # "auto": tcgen05 split-bf16 when the shape allows, else exact-fp32 CUDA cores;
        # "fp32" / "tensor_cores" force one of them (include/e3dge_b200.h E3_CONV_*)
        self.backend = "auto"
        self._packed = _PackedConv()
        self._packed_bwd = _PackedConv()
        self._prefetched = None  # ((latent ptr, batch), s, d) left by Decoder.prepare for the next call

    def _apply(self, fn, *args, **kwargs):
        _lib.invalidate_packed()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        _lib.invalidate_packed()
        return super()._load_from_state_dict(*args, **kwargs)

    def train(self, mode=True):
        _lib.invalidate_packed()
        return super().train(mode)

    def styles(self, latent):
        """s [B,cin] and (if demodulating) d [B,cout] for latent [B,512] (may be a strided
        view latent[:, i] of [B,n_latent,512])."""
        pre, self._prefetched = self._prefetched, None
        if pre is not None and pre[0] == (latent.data_ptr(), latent.shape[0]):
            return pre[1], pre[2]
        lib = _lib.load()
        if latent.stride(-1) != 1 or latent.dtype != torch.float32 or not latent.is_cuda:
            latent = _lib.as_f32c(latent)
        if latent.shape[-1] != 512 or self.modulation.weight.shape[1] != 512:
            raise NotImplementedError("decoder style_dim must be 512")
        b = latent.shape[0]
        _, wsq = self._packed.get(self.weight, int(self.upsample))
        s = torch.empty(b, self.in_channel, device=latent.device, dtype=torch.float32)
        d = torch.empty(b, self.out_channel, device=latent.device,
                        dtype=torch.float32) if self.demodulate else None
        import ctypes
        _lib.check(lib.e3_modconv_styles(_lib.vptr(latent),
                                         latent.stride(0) if b > 1 else 512,
                                         _lib.ptr(_lib.as_f32c(self.modulation.weight.detach())),
                                         _lib.ptr(_lib.as_f32c(self.modulation.bias.detach())),
                                         _lib.ptr(wsq), b, self.in_channel, self.out_channel,
                                         self.kernel_size, _lib.ptr(s), _lib.ptr(d),
                                         _lib.cur_stream()), "e3_modconv_styles",
                   launches=2 if d is not None else 1)
        return s, d

    def forward(self, input, style):
        """input [B,I,H,W] NCHW -> [B,O,H',W'] NCHW, bare modulated conv (no noise / act)."""
        if self.kernel_size == 1:
            zero_b = torch.zeros(3, device=input.device)
            if self.out_channel != 3 or self.demodulate:
                raise NotImplementedError("1x1 ModulatedConv2d is supported as ToRGB (3 outputs, "
                                          "no demodulation)")
            return _torgb_apply(self, _nhwc(input), style, zero_b, None, False, None)
        return _nchw(_styled_conv_apply(self, _nhwc(input), style, None, None, None))


class _Permute(torch.autograd.Function):
    """NCHW <-> NHWC through the transpose kernels; the adjoint is the opposite transpose."""

    @staticmethod
    def forward(ctx, x, to_nhwc):
        ctx.to_nhwc = to_nhwc
        return _to_nhwc(x) if to_nhwc else _to_nchw(x)

    @staticmethod
    def backward(ctx, g):
        g = _lib.as_f32c(g)
        return (_to_nchw(g) if ctx.to_nhwc else _to_nhwc(g)), None


def _nhwc(x):
    return _Permute.apply(x, True) if _wants_grad(x) else _to_nhwc(x)


def _nchw(x):
    return _Permute.apply(x, False) if _wants_grad(x) else _to_nchw(x)


def _styled_conv_apply(conv, x, latent, noise, noise_w, act_bias):
    if _wants_grad(x, latent):
        return _StyledConvFn.apply(conv, x, latent, noise, noise_w, act_bias)
    return _styled_conv_nhwc(conv, x, latent, noise, noise_w, act_bias)


def _torgb_apply(conv, x, latent, bias, skip, upsample_skip, up_kernel):
    if _wants_grad(x, latent, skip):
        return _ToRGBFn.apply(conv, x, latent, bias, skip, upsample_skip, up_kernel)
    return _torgb_nhwc(conv, x, latent, bias, skip, upsample_skip)


def _styled_conv_nhwc(conv, x, latent, noise, noise_w, act_bias, want_saved=False):
    """x [B,H,W,cin] NHWC -> StyledConv output [B,H',W',cout] NHWC.
    act_bias None = bare modulated conv (no noise / bias / activation)."""
    lib = _lib.load()
    b, h, w, cin = x.shape
    cout, up = conv.out_channel, conv.upsample
    x = _lib.as_f32c(x.detach())
    s, d = conv.styles(latent.detach())
    has_d = d is not None
    if d is None:
        d = torch.ones(b, cout, device=x.device, dtype=torch.float32)
    wp, _ = conv._packed.get(conv.weight, int(up))
    oh, ow = (2 * h, 2 * w) if up else (h, w)
    nstride = 0
    if act_bias is not None:
        noise = _lib.as_f32c(noise)
        if noise.numel() == b * oh * ow and b > 1:
            nstride = oh * ow
        elif noise.numel() != oh * ow:
            raise RuntimeError(f"noise of {tuple(noise.shape)} does not match a {oh}x{ow} layer")
        noise_w, act_bias = _lib.as_f32c(noise_w.detach()), _lib.as_f32c(act_bias.detach())
    y = torch.empty(b, oh, ow, cout, device=x.device, dtype=torch.float32)
    nbytes = lib.e3_styled_conv_scratch_bytes(b, h, w, cin, cout, int(up))
    scratch = torch.empty(max(nbytes // 4, 1), device=x.device, dtype=torch.float32)
    fn = lib.e3_styled_conv3x3_up_fwd if up else lib.e3_styled_conv3x3_fwd
    args = (_lib.ptr(x), _lib.ptr(wp), _lib.ptr(s), _lib.ptr(d), _lib.ptr(noise), nstride,
            _lib.ptr(noise_w), _lib.ptr(act_bias), _lib.ptr(y), b, h, w, cin, cout,
            _lib.ptr(scratch), nbytes, CONV_BACKENDS[conv.backend], _lib.cur_stream())
    _lib.check(fn(*args), "e3_styled_conv3x3_up_fwd" if up else "e3_styled_conv3x3_fwd")
    if want_saved:
        return y, (s, d, has_d, noise, nstride, noise_w, act_bias)
    return y


FUSE_CONV_PAIRS = True  # inference: hand the up-conv's output to the next conv as bf16 hi / lo operands


def _pair_fusable(up, plain, x):
    """Can e3_styled_conv3x3_up_fwd_split / _fwd_presplit run this (upsampling, plain) StyledConv pair?"""
    if not FUSE_CONV_PAIRS:
        return False
    cu, cp = up.conv, plain.conv
    if not (cu.upsample and not cp.upsample and cp.in_channel == cu.out_channel == cp.out_channel
            and cu.kernel_size == cp.kernel_size == 3 and cu.demodulate and cp.demodulate):
        return False
    if cu.backend == "fp32" or cp.backend == "fp32":
        return False
    b, h, w, cin = x.shape
    return bool(_lib.load().e3_styled_conv_pair_fusable(b, h, w, cin, cu.out_channel, CONV_BACKENDS["auto"]))


def _styled_conv_pair_nhwc(up, plain, x, lat_up, lat_plain, noise_up, noise_plain):
    """Inference only: StyledConv(upsample) -> StyledConv of one resolution step with the intermediate
    activation handed over as the second conv's bf16 hi / lo operands (include/e3dge_b200.h)."""
    lib = _lib.load()
    cu, cp = up.conv, plain.conv
    b, h, w, cin = x.shape
    c = cu.out_channel
    oh, ow = 2 * h, 2 * w
    x = _lib.as_f32c(x.detach())
    s1, d1 = cu.styles(lat_up.detach())
    s2, d2 = cp.styles(lat_plain.detach())
    wp1, _ = cu._packed.get(cu.weight, 1)
    wp2, _ = cp._packed.get(cp.weight, 0)

    def noise_of(n):
        n = torch.empty(b, 1, oh, ow, device=x.device).normal_() if n is None else _lib.as_f32c(n)
        if n.numel() == b * oh * ow and b > 1:
            return n, oh * ow
        if n.numel() != oh * ow:
            raise RuntimeError(f"noise of {tuple(n.shape)} does not match a {oh}x{ow} layer")
        return n, 0
    n1, st1 = noise_of(noise_up)
    n2, st2 = noise_of(noise_plain)
    xs = torch.empty(2, b, oh, ow, c, device=x.device, dtype=torch.bfloat16)
    nbytes = lib.e3_styled_conv_scratch_bytes(b, h, w, cin, c, 1)
    scratch = torch.empty(max(nbytes // 4, 1), device=x.device, dtype=torch.float32)
    vp = _lib.vptr
    f32 = lambda t: _lib.ptr(_lib.as_f32c(t.detach()))
    _lib.check(lib.e3_styled_conv3x3_up_fwd_split(
        _lib.ptr(x), _lib.ptr(wp1), _lib.ptr(s1), _lib.ptr(d1), _lib.ptr(n1), st1, f32(up.noise.weight),
        f32(up.activate.bias), _lib.ptr(s2), vp(xs[0]), vp(xs[1]), b, h, w, cin, c, _lib.ptr(scratch), nbytes,
        CONV_BACKENDS[cu.backend], _lib.cur_stream()), "e3_styled_conv3x3_up_fwd_split")
    y = torch.empty(b, oh, ow, c, device=x.device, dtype=torch.float32)
    _lib.check(lib.e3_styled_conv3x3_fwd_presplit(
        vp(xs[0]), vp(xs[1]), _lib.ptr(wp2), _lib.ptr(d2), _lib.ptr(n2), st2, f32(plain.noise.weight),
        f32(plain.activate.bias), _lib.ptr(y), b, oh, ow, c, c, CONV_BACKENDS[cp.backend], _lib.cur_stream()),
        "e3_styled_conv3x3_fwd_presplit")
    return y


def _torgb_nhwc(conv, x, latent, bias, skip, upsample_skip, want_saved=False):
    lib = _lib.load()
    b, h, w, cin = x.shape
    x = _lib.as_f32c(x.detach())
    s, _ = conv.styles(latent.detach())
    rgb = torch.empty(b, 3, h, w, device=x.device, dtype=torch.float32)
    wt = _lib.as_f32c(conv.weight.detach().reshape(3, cin))
    sk = _lib.as_f32c(skip) if skip is not None else None
    _lib.check(lib.e3_torgb_fwd(_lib.ptr(x), _lib.ptr(wt), _lib.ptr(s),
                                _lib.ptr(_lib.as_f32c(bias.detach().reshape(3))), _lib.ptr(sk),
                                int(bool(upsample_skip)), _lib.ptr(rgb), b, h, w, cin,
                                _lib.cur_stream()), "e3_torgb_fwd")
    if want_saved:
        return rgb, s
    return rgb


class NoiseInjection(nn.Module):
    """stylesdf_model.py:365-466 (`project_noise` mesh path: out of scope)."""

    def __init__(self, project=False):
        super().__init__()
        if project:
            raise NotImplementedError("project_noise needs pytorch3d mesh rasterisation "
                                      "(out of scope, SURVEY.md §8a a20)")
        self.project = project
        self.weight = nn.Parameter(torch.zeros(1))

    def make_noise(self, image_nchw_shape, device):
        b, _, h, w = image_nchw_shape
        return torch.empty(b, 1, h, w, device=device).normal_()

    def forward(self, image, noise=None, transform=None, mesh_path=None):
        if noise is None:
            noise = self.make_noise(image.shape, image.device)
        return image + self.weight * noise


class StyledConv(nn.Module):
    """stylesdf_model.py:469-507: conv -> + noise_weight*noise -> fused lrelu(x + bias)*sqrt2.
    (`bias` [1,O,1,1] is a dead parameter of the reference; kept for the state_dict.)"""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False,
                 blur_kernel=[1, 3, 3, 1], project_noise=False):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim,
                                    upsample=upsample, blur_kernel=blur_kernel)
        self.noise = NoiseInjection(project=project_noise)
        self.bias = nn.Parameter(torch.zeros(1, out_channel, 1, 1))
        self.activate = FusedLeakyReLU(out_channel)

    def forward_nhwc(self, x, style, noise=None):
        b, h, w, _ = x.shape
        if noise is None:  # fresh noise per call (stylesdf_model.py:461-462)
            oh, ow = (2 * h, 2 * w) if self.conv.upsample else (h, w)
            noise = torch.empty(b, 1, oh, ow, device=x.device).normal_()
        return _styled_conv_apply(self.conv, x, style, noise, self.noise.weight, self.activate.bias)

    def forward(self, input, style, noise=None, transform=None, mesh_path=None):
        return _nchw(self.forward_nhwc(_nhwc(input), style, noise))


class ToRGB(nn.Module):
    """stylesdf_model.py:510-541."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.upsample = Upsample(blur_kernel) if upsample else upsample
        if upsample and list(blur_kernel) != [1, 3, 3, 1]:
            raise NotImplementedError("fused skip upsampling is built for blur_kernel [1,3,3,1]")
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward_nhwc(self, x, style, skip=None):
        up = bool(self.upsample)
        return _torgb_apply(self.conv, x, style, self.bias, skip, up, self.upsample.kernel if up else None)

    def forward(self, input, style, skip=None):
        return self.forward_nhwc(_nhwc(input), style, skip)


class Decoder(nn.Module):
    """stylesdf_model.py:587-797."""

    def __init__(self, model_opt, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.size = model_opt.size
        self.style_dim = model_opt.style_dim * 2
        layers = [PixelNorm(), EqualLinear(self.style_dim // 2, self.style_dim,
                                           lr_mul=model_opt.lr_mapping, activation="fused_lrelu")]
        for _ in range(4):
            layers.append(EqualLinear(self.style_dim, self.style_dim, lr_mul=model_opt.lr_mapping,
                                      activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        cm = model_opt.channel_multiplier
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm,
                         256: 64 * cm, 512: 32 * cm, 1024: 16 * cm}
        decoder_in_size = model_opt.renderer_spatial_output_dim
        self.log_size = int(math.log(self.size, 2))
        self.log_in_size = int(math.log(decoder_in_size, 2))
        self.conv1 = StyledConv(model_opt.feature_encoder_in_channels,
                                self.channels[decoder_in_size], 3, self.style_dim,
                                blur_kernel=blur_kernel, project_noise=model_opt.project_noise)
        self.to_rgb1 = ToRGB(self.channels[decoder_in_size], self.style_dim, upsample=False)
        self.num_layers = (self.log_size - self.log_in_size) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[decoder_in_size]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 2 * self.log_in_size + 1) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(self.log_in_size + 1, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, self.style_dim, upsample=True,
                                         blur_kernel=blur_kernel,
                                         project_noise=model_opt.project_noise))
            self.convs.append(StyledConv(out_channel, out_channel, 3, self.style_dim,
                                         blur_kernel=blur_kernel,
                                         project_noise=model_opt.project_noise))
            self.to_rgbs.append(ToRGB(out_channel, self.style_dim))
            in_channel = out_channel
        self.n_latent = (self.log_size - self.log_in_size) * 2 + 2

    def mean_latent(self, renderer_latent):
        return self.style(renderer_latent).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def styles_and_noise_forward(self, styles, noise, inject_index=None, truncation=1,
                                 truncation_latent=None, input_is_latent=False,
                                 randomize_noise=True):
        """stylesdf_model.py:692-740."""
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        if truncation < 1:
            styles = [truncation_latent[1] + truncation * (s - truncation_latent[1]) for s in styles]
        if len(styles) < 2:
            inject_index = self.n_latent
            latent = styles[0]
            if latent.ndim < 3:
                latent = latent.unsqueeze(1).repeat(1, inject_index, 1)
        else:
            if inject_index is None:
                inject_index = random.randint(1, self.n_latent - 1)
            latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)
        return latent, noise

    def _layer_latents(self):
        """(ModulatedConv2d, latent index) in execution order (stylesdf_model.py:764-795)."""
        out = [(self.conv1.conv, 0), (self.to_rgb1.conv, 1)]
        i = 1
        for conv1, conv2, to_rgb in zip(self.convs[::2], self.convs[1::2], self.to_rgbs):
            out += [(conv1.conv, i), (conv2.conv, i + 1), (to_rgb.conv, i + 2)]
            i += 2
        return out

    def prepare(self, styles, batch, noise=None, inject_index=None, truncation=1, truncation_latent=None,
                input_is_latent=False, randomize_noise=True):
        """Everything of a pass that does not depend on the feature map: the latent (mapping network,
        truncation, style mixing), every layer's modulation s / demodulation d, and the fresh noise maps —
        about 25 tiny latency-bound launches.  G_pred_latents issues them on a side stream while the render
        kernel runs; `forward(prepared=...)` consumes the result."""
        assert isinstance(styles, list), "wrap latent code with list"
        latent, noise = self.styles_and_noise_forward(styles, noise, inject_index, truncation,
                                                      truncation_latent, input_is_latent, randomize_noise)
        latent = _lib.as_f32c(latent)
        keep = [latent]
        for conv, idx in self._layer_latents():
            lat = latent[:, idx]
            conv._prefetched = None
            s, d = conv.styles(lat)
            conv._prefetched = ((lat.data_ptr(), lat.shape[0]), s, d)
            keep += [s] if d is None else [s, d]
        noise = list(noise)
        for k in range(self.num_layers):
            if noise[k] is None:  # fresh noise per call (stylesdf_model.py:461-462)
                res = 2 ** ((k + 2 * self.log_in_size + 1) // 2)
                noise[k] = torch.empty(batch, 1, res, res, device=latent.device).normal_()
                keep.append(noise[k])
        return latent, noise, keep

    def forward(self, features, styles, rgbd_in=None, transform=None, return_latents=False,
                inject_index=None, truncation=1, truncation_latent=None, input_is_latent=False,
                noise=None, randomize_noise=True, mesh_path=None, conditions=None, prepared=None):
        """features [B,256,R,R] NCHW -> (image [B,3,size,size], latent|None)
        — stylesdf_model.py:742-797.  `conditions` is accepted and ignored exactly as in the
        reference (its HFGI hook is dead code)."""
        assert isinstance(styles, list), "wrap latent code with list"
        if prepared is not None:
            latent, noise = prepared[0], prepared[1]
        else:
            latent, noise = self.styles_and_noise_forward(styles, noise, inject_index, truncation,
                                                          truncation_latent, input_is_latent,
                                                          randomize_noise)
            latent = _lib.as_f32c(latent)
        x = _nhwc(features)
        out = self.conv1.forward_nhwc(x, latent[:, 0], noise[0])
        skip = self.to_rgb1.forward_nhwc(out, latent[:, 1], rgbd_in)
        i = 1
        for conv1, conv2, noise1, noise2, to_rgb in zip(self.convs[::2], self.convs[1::2],
                                                        noise[1::2], noise[2::2], self.to_rgbs):
            if not _wants_grad(out, latent) and _pair_fusable(conv1, conv2, out):
                out = _styled_conv_pair_nhwc(conv1, conv2, out, latent[:, i], latent[:, i + 1], noise1, noise2)
            else:
                out = conv1.forward_nhwc(out, latent[:, i], noise1)
                out = conv2.forward_nhwc(out, latent[:, i + 1], noise2)
            skip = to_rgb.forward_nhwc(out, latent[:, i + 2], skip)
            i += 2
        return skip, (latent if return_latents else None)


class Generator(nn.Module):
    """stylesdf_model.py:800-1020."""

    def __init__(self, model_opt, renderer_opt, blur_kernel=[1, 3, 3, 1], ema=False,
                 full_pipeline=True):
        super().__init__()
        self.size = model_opt.size
        self.style_dim = model_opt.style_dim
        self.num_layers = 1
        self.train_renderer = not model_opt.freeze_renderer
        self.full_pipeline = full_pipeline
        model_opt.feature_encoder_in_channels = renderer_opt.width
        self.is_train = not (ema or model_opt.is_test)
        self.style = nn.Sequential(*[MappingLinear(self.style_dim, self.style_dim,
                                                   activation="fused_lrelu") for _ in range(3)])
        self.renderer = VolumeFeatureRenderer(renderer_opt, style_dim=self.style_dim,
                                              out_im_res=model_opt.renderer_spatial_output_dim)
        self.renderer_n_latent = renderer_opt.depth + 1
        if self.full_pipeline:
            self.decoder = Decoder(model_opt)
            self.stylegan_n_latent = 10

    def mean_latent(self, n_latent, device):
        """stylesdf_model.py:854-864."""
        latent_in = torch.randn(n_latent, self.style_dim, device=device)
        renderer_latent = self.style(latent_in)
        renderer_latent_mean = renderer_latent.mean(0, keepdim=True)
        decoder_latent_mean = None
        if self.full_pipeline:
            decoder_latent_mean = self.decoder.mean_latent(renderer_latent)
            self.decoder_latent_mean = decoder_latent_mean.to(device)
        return [renderer_latent_mean, decoder_latent_mean]

    def get_latent(self, input):
        return self.style(input)

    def styles_and_noise_forward(self, styles, inject_index=None, truncation=1,
                                 truncation_latent=None, input_is_latent=False):
        """stylesdf_model.py:869-903."""
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if truncation < 1:
            assert isinstance(truncation_latent, list)
            styles = [truncation_latent[0] + truncation * (s - truncation_latent[0]) for s in styles]
        return styles

    def data_sample_forward(self, styles, cam_poses, focals, near=0.88, far=1.12, **kwargs):
        """stylesdf_model.py:905-921."""
        latent = self.styles_and_noise_forward(styles)
        return self.renderer.sdf_sample_pass(cam_poses, focals, near, far, styles=latent[0], **kwargs)

    def init_forward(self, styles, cam_poses, focals, near=0.88, far=1.12):
        latent = self.styles_and_noise_forward(styles)
        return self.renderer.mlp_init_pass(cam_poses, focals, near, far, styles=latent[0])

    def forward(self, styles, cam_poses, focals, near=0.88, far=1.12, return_latents=False,
                inject_index=None, truncation=1, truncation_latent=None, input_is_latent=False,
                noise=None, randomize_noise=True, return_sdf=False, return_xyz=False,
                return_eikonal=False, project_noise=False, return_mesh=False,
                mesh_with_shading=True, mesh_path=None, pred_decoder_latents=None,
                sample_mode=False, diable_decoder_inference=False):
        """Tuple-returning forward of the base Generator — stylesdf_model.py:933-1020."""
        with torch.set_grad_enabled(self.is_train and self.train_renderer):
            latent = self.styles_and_noise_forward(styles, inject_index, truncation,
                                                   truncation_latent, input_is_latent)
            sample_batch = self.renderer(cam_poses, focals, near, far, styles=latent[0],
                                         return_eikonal=return_eikonal, sample_mode=sample_mode,
                                         return_mesh=return_mesh,
                                         mesh_with_shading=mesh_with_shading)
            if sample_mode:
                return sample_batch
            thumb_rgb, features, sdf, mask, xyz, eikonal_term = [
                sample_batch[k] for k in ["gen_thumb_imgs", "features", "sdf", "mask", "xyz",
                                          "eikonal_term"]]
        rgb, decoder_latent = None, None
        if self.full_pipeline and not diable_decoder_inference:
            decoder_latent = latent if pred_decoder_latents is None else pred_decoder_latents
            rgb, decoder_latent = self.decoder(
                features, decoder_latent, transform=cam_poses if project_noise else None,
                return_latents=return_latents, inject_index=inject_index, truncation=truncation,
                truncation_latent=truncation_latent, noise=noise, input_is_latent=input_is_latent,
                randomize_noise=randomize_noise, mesh_path=mesh_path)
        if return_latents:
            return rgb, decoder_latent
        out = (rgb, thumb_rgb)
        if return_xyz:
            out += (xyz,)
        if return_sdf:
            out += (sdf,)
        if return_eikonal:
            out += (eikonal_term,)
        if return_xyz:
            out += (mask,)
        return out


class G_pred_latents(Generator):
    """Dict-returning generator the E3DGE runners call — stylesdf_model.py:1023-1172."""

    def _side_stream(self, device):
        streams = self.__dict__.setdefault("_side_streams", {})
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key not in streams:
            streams[key] = torch.cuda.Stream(device=device)
        return streams[key]

    def forward(self, styles, cam_poses, focals, near=0.88, far=1.12, return_latents=False,
                inject_index=None, truncation=1, truncation_latent=None, input_is_latent=False,
                noise=None, randomize_noise=True, return_sdf=False, return_xyz=False,
                return_eikonal=False, project_noise=False, return_mesh=False,
                mesh_with_shading=True, mesh_path=None, conditions=None, sample_mode=False,
                geometry_sample=None, sample_with_decoder=False, sample_with_renderer=False,
                return_surface_eikonal=False, renderer_only=False, inference_mode=False,
                sample_without_grad=False, **kwargs):
        if self.full_pipeline:
            assert type(styles) in [list, tuple], "reformat latent to list/tuple"
            if not input_is_latent:
                encoder_latent, decoder_latent = styles[0], None
            else:
                encoder_latent, decoder_latent = styles
        else:
            decoder_latent = None
            encoder_latent = styles[0]
        renderer_latent = self.styles_and_noise_forward([encoder_latent], inject_index, truncation,
                                                        truncation_latent, input_is_latent)
        # Inference: the decoder's latent-only work (styles, demodulation, noise maps: ~25 tiny launches)
        # goes to a side stream and runs under the render kernel instead of after it.
        runs_decoder = (self.full_pipeline or sample_with_decoder) and not sample_with_renderer and not renderer_only
        prepared = None
        if (runs_decoder and not torch.is_grad_enabled() and renderer_latent[0].is_cuda
                and not (sample_mode or project_noise)):
            dl = renderer_latent if decoder_latent is None else (
                decoder_latent if isinstance(decoder_latent, list) else [decoder_latent])
            main = torch.cuda.current_stream()
            side = self._side_stream(renderer_latent[0].device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                prepared = self.decoder.prepare(dl, renderer_latent[0].shape[0], noise, inject_index, truncation,
                                                truncation_latent, input_is_latent, randomize_noise)
            for t in prepared[2]:
                t.record_stream(main)
        try:
            render_out = self.renderer(cam_poses, focals, near, far, styles=renderer_latent[0],
                                       return_eikonal=return_eikonal, return_mesh=return_mesh,
                                       mesh_with_shading=mesh_with_shading, sample_mode=sample_mode,
                                       geometry_sample=geometry_sample,
                                       return_surface_eikonal=return_surface_eikonal,
                                       sample_without_grad=sample_without_grad, **kwargs)
            render_out["styles"] = renderer_latent[0]
            if renderer_only:
                return render_out
            if (self.full_pipeline or sample_with_decoder) and not sample_with_renderer:
                if decoder_latent is None:
                    decoder_latent = renderer_latent
                elif not isinstance(decoder_latent, list):
                    decoder_latent = [decoder_latent]
                if prepared is not None:
                    torch.cuda.current_stream().wait_stream(self._side_stream(render_out["features"].device))
                gen_imgs, decoder_latent = self.decoder(
                    render_out["features"], decoder_latent,
                    transform=cam_poses if project_noise else None, return_latents=return_latents,
                    inject_index=inject_index, truncation=truncation,
                    truncation_latent=truncation_latent, noise=noise, input_is_latent=input_is_latent,
                    randomize_noise=randomize_noise, mesh_path=mesh_path, conditions=conditions,
                    prepared=prepared)
                render_out["gen_imgs"] = gen_imgs
                render_out["decoder_latent"] = decoder_latent
        finally:
            if prepared is not None:  # never leave a half-consumed prefetch behind (exceptions, early returns)
                for conv, _ in self.decoder._layer_latents():
                    conv._prefetched = None
        return render_out
