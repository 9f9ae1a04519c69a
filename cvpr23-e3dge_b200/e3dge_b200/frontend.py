"""Front end of an inversion frame: image -> latents (pSp-style IR-SE50 + FPN encoder) and
image -> camera (CoordConv pose net), then the generator hot path.

    AERunner.image2image                      project/trainers/trainer.py:773-840
      image2latents -> encoder + mean latent  :950-969, 989-1010
      image2camsettings -> pose net -> camera :935-948, 971-987, project/utils/camera_utils.py:8-151
      latent2image -> G_pred_latents          :843-900

These two networks are ordinary conv nets (cuDNN through PyTorch — library code, not hand-written kernels): the
module trees and parameter names follow the reference (`HybridGradualStyleEncoder_V2`,
project/models/encoders/fpn_encoders.py:266-431 with project/models/helper_modules/helpers.py:95-224, 472-497;
`VolumeRenderDiscriminator`, project/models/stylesdf_model.py:1193-1419) so their checkpoints load, and
`InversionPipeline` runs them channels-last under bf16 autocast inside the same CUDA graph as the generator, so
that bench.py can report the whole encoder -> render -> decode frame next to the generator-only figure."""
import math
from collections import namedtuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .op import FusedLeakyReLU
from .stylesdf_model import EqualLinear


# ---------------------------------------------------------------------------------------- cameras
def generate_camera_params(resolution, device, batch=1, locations=None, sweep=False, uniform=False, azim_range=0.3,
                           elev_range=0.15, fov_ang=6, dist_radius=0.12, return_calibs=False, azim_mean=0.,
                           elev_mean=0., generator=None):
    """Cameras on the unit sphere looking at the origin — camera_utils.py:8-151.
    locations [B,2] = (azimuth, elevation) in radians; else `sweep` = 8 evenly spaced azimuths per identity at one
    random elevation each (8*batch cameras); else sampled (normal, or uniform in +-range)."""
    if locations is not None:
        azim, elev = locations[:, 0:1], locations[:, 1:2]
        batch = azim.shape[0]
    elif sweep:
        azim = (-azim_range + (2 * azim_range / 7) * torch.arange(8, device=device)).reshape(-1, 1).repeat(batch, 1)
        elev = (-elev_range + 2 * elev_range *
                torch.rand(batch, 1, device=device, generator=generator).repeat(1, 8).reshape(-1, 1))
        batch = batch * 8
    elif uniform:
        azim = -azim_range + 2 * azim_range * torch.rand(batch, 1, device=device, generator=generator)
        elev = -elev_range + 2 * elev_range * torch.rand(batch, 1, device=device, generator=generator)
    else:
        azim = azim_range * torch.randn(batch, 1, device=device, generator=generator)
        elev = elev_range * torch.randn(batch, 1, device=device, generator=generator)
    dist = torch.ones(batch, 1, device=device)
    near, far = (dist - dist_radius).unsqueeze(-1), (dist + dist_radius).unsqueeze(-1)
    fov = fov_ang * torch.ones(batch, 1, device=device) * math.pi / 180
    focal = 0.5 * resolution / torch.tan(fov).unsqueeze(-1)
    azim, elev = azim_mean + azim, elev_mean + elev
    viewpoint = torch.cat([azim, elev], 1)
    cam_dir = torch.stack([torch.cos(elev) * torch.sin(azim), torch.sin(elev), torch.cos(elev) * torch.cos(azim)],
                          1).reshape(-1, 3)
    cam_loc = dist * cam_dir
    up = torch.zeros(batch, 3, device=device)  # (constants are built by fill kernels, not host copies: the
    up[:, 1] = 1.                               # whole frame records into a CUDA graph)
    z_axis = F.normalize(cam_dir, eps=1e-5)  # -z points into the screen
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    degenerate = torch.isclose(x_axis, torch.zeros((), device=device), atol=5e-3).all(dim=1, keepdim=True)
    x_axis = torch.where(degenerate, F.normalize(torch.cross(y_axis, z_axis, dim=1), eps=1e-5), x_axis)
    w2c_R = torch.stack([x_axis, y_axis, z_axis], 1)
    T = cam_loc[:, :, None]
    poses = torch.cat([w2c_R.transpose(1, 2), T], -1)  # c2w [B,3,4]
    if not return_calibs:
        return poses, focal, near, far, viewpoint
    extrinsics = torch.cat([w2c_R, -w2c_R @ T], -1)  # w2c [B,3,4]
    intrinsics = torch.zeros(batch, 3, 3, device=device)  # uv space: maps camera points to [-1,1]
    intrinsics[:, 0, 0] = intrinsics[:, 1, 1] = focal[0].squeeze() / (resolution / 2)
    intrinsics[:, 2, 2] = 1.
    calibs = intrinsics @ extrinsics
    bottom = torch.zeros(batch, 1, 4, device=device)
    bottom[:, :, 3] = 1.
    return dict(poses=poses, extrinsics=extrinsics, focal=focal, near=near, far=far, viewpoint=viewpoint,
                intrinsics=intrinsics, calibs=torch.cat([calibs, bottom], -2), locations=locations,
                azim_range=azim_range, elev_range=elev_range)


# ---------------------------------------------------------------------------------------- encoder
Bottleneck = namedtuple("Block", ["in_channel", "depth", "stride"])


def get_blocks(num_layers):
    """IR-ResNet stage plan — helpers.py:99-130."""
    units = {50: (3, 4, 14, 3), 100: (3, 13, 30, 3), 152: (3, 8, 36, 3)}
    if num_layers not in units:
        raise ValueError(f"Invalid number of layers: {num_layers}. Must be one of [50, 100, 152]")
    stage = lambda cin, depth, n: [Bottleneck(cin, depth, 2)] + [Bottleneck(depth, depth, 1)] * (n - 1)
    return [stage(cin, depth, n) for (cin, depth), n in zip(((64, 64), (64, 128), (128, 256), (256, 512)),
                                                            units[num_layers])]


class SEModule(nn.Module):
    """helpers.py:133-158."""

    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        return x * self.sigmoid(self.fc2(self.relu(self.fc1(self.avg_pool(x)))))


class bottleneck_IR_SE(nn.Module):
    """helpers.py:204-223."""

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, (1, 1), stride, bias=False),
                                                nn.BatchNorm2d(depth))
        self.res_layer = nn.Sequential(
            nn.BatchNorm2d(in_channel), nn.Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False), nn.PReLU(depth),
            nn.Conv2d(depth, depth, (3, 3), stride, 1, bias=False), nn.BatchNorm2d(depth), SEModule(depth, 16))

    def forward(self, x):
        return self.res_layer(x) + self.shortcut_layer(x)


class bottleneck_IR(nn.Module):
    """helpers.py:161-201 (plain BatchNorm / Conv2d form)."""

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, (1, 1), stride, bias=False),
                                                nn.BatchNorm2d(depth))
        self.res_layer = nn.Sequential(
            nn.BatchNorm2d(in_channel), nn.Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False), nn.PReLU(depth),
            nn.Conv2d(depth, depth, (3, 3), stride, 1, bias=False), nn.BatchNorm2d(depth))

    def forward(self, x):
        return self.res_layer(x) + self.shortcut_layer(x)


class GradualStyleBlock(nn.Module):
    """Feature map [B,in_c,s,s] -> one latent [B,out_c]: log2(s) stride-2 convs, then an EqualLinear —
    helpers.py:472-497."""

    def __init__(self, in_c, out_c, spatial):
        super().__init__()
        self.out_c, self.spatial = out_c, spatial
        mods, cin = [], in_c
        for _ in range(int(math.log2(spatial))):
            mods += [nn.Conv2d(cin, out_c, kernel_size=3, stride=2, padding=1), nn.LeakyReLU()]
            cin = out_c
        self.convs = nn.Sequential(*mods)
        self.linear = EqualLinear(out_c, out_c, lr_mul=1)

    def forward(self, x):
        return self.linear(self.convs(x).reshape(-1, self.out_c))


def encoder_options(**over):
    """The `opt.training` fields the encoder reads (project/utils/options.py), shipped-script values."""
    from .options import Opt
    o = Opt(input_nc=3, fpn_pigan_geo_layer_dim=32, fpn_pigan_tex_layer_dim=32, full_pipeline=True,
            disable_decoder_fpn=False, single_decoder_layer=True)
    o.update(over)
    return o


class HybridGradualStyleEncoder_V2(nn.Module):
    """IR-SE50 trunk with taps after units 2 / 6 / 20 / 23 (128^2 .. 16^2), top-down FPN, nine 256-d renderer
    latents and one 512-d decoder latent repeated ten times — fpn_encoders.py:266-431.  Returns latent OFFSETS;
    the runner adds the generator's mean latents (trainer.py:989-1010)."""

    def __init__(self, num_layers=50, mode="ir_se", n_styles=-1, opts=None):
        super().__init__()
        opts = encoder_options() if opts is None else opts
        assert num_layers in (50, 100, 152) and mode in ("ir", "ir_se")
        unit = bottleneck_IR if mode == "ir" else bottleneck_IR_SE
        self.opts = opts
        self.input_layer = nn.Sequential(nn.Conv2d(opts.input_nc, 64, (3, 3), 1, 1, bias=False), nn.BatchNorm2d(64),
                                         nn.PReLU(64))
        self.full_pipeline = opts.full_pipeline
        self.body = nn.Sequential(*[unit(b.in_channel, b.depth, b.stride) for blk in get_blocks(num_layers) for b in blk])
        self.pigan_geo_layer, self.pigan_tex_layer = 6, 9
        self.styles_pigan = nn.ModuleList(
            [GradualStyleBlock(512, 256, opts.fpn_pigan_geo_layer_dim if i < self.pigan_geo_layer
                               else opts.fpn_pigan_tex_layer_dim) for i in range(9)])
        self.enable_decoder = bool(self.full_pipeline and not opts.disable_decoder_fpn)
        if self.enable_decoder:
            self.stylegan_style_count = 10
            if opts.single_decoder_layer:
                self.styles_stylegan = nn.ModuleList([GradualStyleBlock(512, 512, 128)])
            else:
                self.stylegan_coarse_ind, self.stylegan_middle_ind = 0, 3
                self.styles_stylegan = nn.ModuleList(
                    [GradualStyleBlock(512, 512, 128 if i < self.stylegan_middle_ind else 256)
                     for i in range(self.stylegan_style_count)])
        self.latlayer64 = nn.Conv2d(64, 512, kernel_size=1, stride=1, padding=0)
        self.latlayer128 = nn.Conv2d(128, 512, kernel_size=1, stride=1, padding=0)
        self.latlayer256 = nn.Conv2d(256, 512, kernel_size=1, stride=1, padding=0)

    @staticmethod
    def _upsample_add(x, y):
        return F.interpolate(x, size=y.shape[-2:], mode="bilinear", align_corners=True) + y

    def forward(self, x, return_featmap=False):
        if x.shape[-1] != 256:
            x = F.adaptive_avg_pool2d(x, (256, 256))
        x = self.input_layer(x)
        taps = {}
        for i, layer in enumerate(self.body):
            x = layer(x)
            if i in (2, 6, 20, 23):
                taps[i] = x
        c128, c64, c32, c16 = taps[2], taps[6], taps[20], taps[23]
        p32 = self._upsample_add(c16, self.latlayer256(c32))
        p64 = self._upsample_add(p32, self.latlayer128(c64))
        tex_src = p64 if self.opts.fpn_pigan_tex_layer_dim == 64 else p32
        latents = [self.styles_pigan[j](p32 if j < self.pigan_geo_layer else tex_src) for j in range(self.pigan_tex_layer)]
        thumb_out = torch.stack(latents, dim=1)
        stylegan_out = None
        if self.enable_decoder:
            p128 = self._upsample_add(p64, self.latlayer64(c128))
            stylegan_out = self.styles_stylegan[0](p128).unsqueeze(1).repeat(1, self.stylegan_style_count, 1)
            if return_featmap:
                return {"pred_latents": [thumb_out, stylegan_out], "feat_maps": p64, "p32": p32}
        return [thumb_out, stylegan_out]


# ---------------------------------------------------------------------------------------- pose net
class VolumeRenderDiscConv2d(nn.Module):
    """stylesdf_model.py:1193-1236."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, activate=False):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias and not activate)
        self.activate = activate
        if activate:
            self.activation = FusedLeakyReLU(out_channels, bias=bias, scale=1)
            lim = math.sqrt(1 / (in_channels * kernel_size * kernel_size))
            nn.init.uniform_(self.activation.bias, a=-lim, b=lim)

    def forward(self, input):
        out = self.conv(input)
        return self.activation(out) if self.activate else out


class AddCoords(nn.Module):
    """Appends (y, x) coordinate channels in [-1, 1] — stylesdf_model.py:1239-1271."""

    def forward(self, t):
        b, _, h, w = t.shape
        xx = torch.linspace(-1, 1, w, device=t.device, dtype=t.dtype).view(1, 1, 1, w).expand(b, 1, h, w)
        yy = torch.linspace(-1, 1, h, device=t.device, dtype=t.dtype).view(1, 1, h, 1).expand(b, 1, h, w)
        return torch.cat([t, yy, xx], dim=1)


class CoordConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.addcoords = AddCoords()
        self.conv = nn.Conv2d(in_channels + 2, out_channels, kernel_size, stride=stride, padding=padding, bias=bias)

    def forward(self, t):
        return self.conv(self.addcoords(t))


class CoordConvLayer(nn.Module):
    """stylesdf_model.py:1303-1336."""

    def __init__(self, in_channel, out_channel, kernel_size, bias=True, activate=True):
        super().__init__()
        self.activate = activate
        self.padding = kernel_size // 2 if kernel_size > 2 else 0
        self.conv = CoordConv2d(in_channel, out_channel, kernel_size, padding=self.padding, stride=1,
                                bias=bias and not activate)
        if activate:
            self.activation = FusedLeakyReLU(out_channel, bias=bias, scale=1)
            lim = math.sqrt(1 / (in_channel * kernel_size * kernel_size))
            nn.init.uniform_(self.activation.bias, a=-lim, b=lim)

    def forward(self, input):
        out = self.conv(input)
        return self.activation(out) if self.activate else out


class VolumeRenderResBlock(nn.Module):
    """stylesdf_model.py:1339-1366."""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.conv1 = CoordConvLayer(in_channel, out_channel, 3)
        self.conv2 = CoordConvLayer(out_channel, out_channel, 3)
        self.pooling = nn.AvgPool2d(2)
        self.downsample = nn.AvgPool2d(2)
        self.skip = VolumeRenderDiscConv2d(in_channel, out_channel, 1) if out_channel != in_channel else None

    def forward(self, input):
        out = self.pooling(self.conv2(self.conv1(input)))
        skip = self.downsample(input)
        if self.skip is not None:
            skip = self.skip(skip)
        return (out + skip) / math.sqrt(2)


class VolumeRenderDiscriminator(nn.Module):
    """The pose net of the inversion pipeline: 64^2 thumbnail -> (gan logit, (azimuth, elevation)) —
    stylesdf_model.py:1369-1419."""

    def __init__(self, opt):
        super().__init__()
        init_size = opt.renderer_spatial_output_dim
        self.viewpoint_loss = True
        channels = {2: 400, 4: 400, 8: 400, 16: 400, 32: 256, 64: 128, 128: 64}
        convs = [VolumeRenderDiscConv2d(3, channels[init_size], 1, activate=True)]
        in_channel = channels[init_size]
        for i in range(int(math.log(init_size, 2)) - 1, 0, -1):
            convs.append(VolumeRenderResBlock(in_channel, channels[2 ** i]))
            in_channel = channels[2 ** i]
        self.convs = nn.Sequential(*convs)
        self.final_conv = VolumeRenderDiscConv2d(in_channel, 3, 2)
        self.in_channel = in_channel

    def forward(self, input):
        out = self.final_conv(self.convs(input))
        return out[:, 0:1].reshape(-1, 1), out[:, 1:].reshape(-1, 2)


# ---------------------------------------------------------------------------------------- the frame
class InversionPipeline(nn.Module):
    """One inversion frame, `AERunner.image2image` (trainer.py:773-840) without the visualisation montage:
    images -> pool to 256^2 / 64^2 -> encoder (+ mean latents) and pose net -> cameras -> generator."""

    def __init__(self, generator, encoder=None, pose_net=None, camera_opt=None, renderer_output_size=64,
                 mean_latents=None, amp=True):
        super().__init__()
        from .options import Opt, model_options
        self.generator = generator
        self.encoder = encoder if encoder is not None else HybridGradualStyleEncoder_V2(50, "ir_se", -1)
        self.volume_discriminator = pose_net if pose_net is not None else VolumeRenderDiscriminator(
            model_options(renderer_spatial_output_dim=renderer_output_size))
        self.camera = camera_opt or Opt(dist_radius=0.12, fov=6, azim=0.3, elev=0.15, uniform=False)
        self.renderer_output_size = renderer_output_size
        self.amp = amp
        n_dec = generator.decoder.n_latent if generator.full_pipeline else 0
        ml = mean_latents or [torch.zeros(1, 256), torch.zeros(1, 512)]
        self.register_buffer("mean_renderer_latent", ml[0].reshape(1, 1, 256).clone())
        self.register_buffer("mean_decoder_latent", ml[1].reshape(1, 1, 512).clone())
        self.n_dec = n_dec

    def image2latents(self, images):
        """trainer.py:950-969 + _add_offset2latent (:989-1010): offsets + mean latents; the decoder latent is cut
        or repeated to the decoder's n_latent."""
        x = F.adaptive_avg_pool2d(images, (256, 256)) if images.shape[-1] != 256 else images
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp and x.is_cuda):
            thumb, dec = self.encoder(x.contiguous(memory_format=torch.channels_last))
        w_plus = thumb.float() + self.mean_renderer_latent
        w_dec = None
        if dec is not None and self.n_dec:
            dec = dec.float()
            if dec.shape[1] < self.n_dec:
                dec = torch.cat([dec, dec[:, -1:].expand(-1, self.n_dec - dec.shape[1], -1)], 1)
            w_dec = (dec[:, :self.n_dec] + self.mean_decoder_latent).contiguous()
        return [w_plus.contiguous(), w_dec]

    def image2camsettings(self, thumb):
        """trainer.py:935-948, 971-987."""
        with torch.no_grad():
            _, locations = self.volume_discriminator(thumb)
        return generate_camera_params(self.renderer_output_size, thumb.device, thumb.shape[0],
                                      locations=locations.float(), uniform=self.camera.uniform,
                                      azim_range=self.camera.azim, elev_range=self.camera.elev, fov_ang=self.camera.fov,
                                      dist_radius=self.camera.dist_radius, return_calibs=True)

    def forward(self, images, randomize_noise=True, **gen_kwargs):
        thumb = F.adaptive_avg_pool2d(images, (64, 64))
        latents = self.image2latents(images)
        cams = self.image2camsettings(thumb)
        styles = latents if latents[1] is not None else [latents[0]]
        out = self.generator(styles, cams["poses"], cams["focal"], cams["near"], cams["far"],
                             input_is_latent=latents[1] is not None, randomize_noise=randomize_noise, **gen_kwargs)
        out.update(pred_latents=latents, pred_cam_settings=cams, input_thumb_imgs=thumb)
        return out
