"""The 2-D image filter of the local branch: `HGPIFuNetGANResidualResnetFC.filter` =
residual / depth stems + a stacked-hourglass network producing the 256-channel feature map that
`netLocal.query` samples (vendor/pifu/lib/model/HGPIFuGANNetResidualInputResnetFC.py:28-75,
vendor/pifu/lib/model/HGFilters.py:6-188, vendor/pifu/lib/net_util.py:399-453,
project/models/helper_modules/helpers.py:250-371, 432-455).

Like the encoder of frontend.py this is an ordinary conv net (PyTorch / cuDNN: library code, an image encoder
outside the hand-written hot path); it exists here so that `netLocal` is complete — module tree and parameter
names are the reference's, E3DGE `--enable_local_model` checkpoints load with `strict=True`, and the whole local
branch (filter -> query -> MLP tail -> modulated render) runs without the caller attaching anything.  Built only
when the rendering options carry the PIFu option group (`opt.pifu`, vendor/pifu/lib/options.py)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _norm(kind, channels):
    if kind == "batch":
        return nn.BatchNorm2d(channels)
    if kind == "group":
        return nn.GroupNorm(32, channels)
    raise NotImplementedError(f"norm {kind!r}")


def _conv3x3(cin, cout):
    """net_util.py:235-242 (zero padding, no bias)."""
    return nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1, bias=False)


class ConvBlock(nn.Module):
    """Pre-activation block whose three 3x3 convs (out/2, out/4, out/4 channels) are concatenated — net_util.py:399-453."""

    def __init__(self, in_planes, out_planes, norm="batch"):
        super().__init__()
        h, q = int(out_planes / 2), int(out_planes / 4)
        self.conv1, self.conv2, self.conv3 = _conv3x3(in_planes, h), _conv3x3(h, q), _conv3x3(q, q)
        self.bn1, self.bn2, self.bn3, self.bn4 = _norm(norm, in_planes), _norm(norm, h), _norm(norm, q), _norm(norm, in_planes)
        self.downsample = None
        if in_planes != out_planes:
            self.downsample = nn.Sequential(self.bn4, nn.ReLU(True),
                                            nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=1, bias=False))

    def forward(self, x):
        o1 = self.conv1(F.relu(self.bn1(x), True))
        o2 = self.conv2(F.relu(self.bn2(o1), True))
        o3 = self.conv3(F.relu(self.bn3(o2), True))
        res = x if self.downsample is None else self.downsample(x)
        return torch.cat((o1, o2, o3), 1) + res


class HourGlass(nn.Module):
    """Recursive hourglass of depth `depth` — HGFilters.py:6-69 (bicubic x2 upsampling, align_corners=True)."""

    def __init__(self, num_modules, depth, num_features, norm="batch"):
        super().__init__()
        self.num_modules, self.depth, self.features, self.norm = num_modules, depth, num_features, norm
        self._generate_network(depth)

    def _generate_network(self, level):
        f, n = self.features, self.norm
        self.add_module(f"b1_{level}", ConvBlock(f, f, norm=n))
        self.add_module(f"b2_{level}", ConvBlock(f, f, norm=n))
        if level > 1:
            self._generate_network(level - 1)
        else:
            self.add_module(f"b2_plus_{level}", ConvBlock(f, f, norm=n))
        self.add_module(f"b3_{level}", ConvBlock(f, f, norm=n))

    def _forward(self, level, inp):
        up1 = self._modules[f"b1_{level}"](inp)
        low = self._modules[f"b2_{level}"](F.avg_pool2d(inp, 2, stride=2))
        low = self._forward(level - 1, low) if level > 1 else self._modules[f"b2_plus_{level}"](low)
        low = self._modules[f"b3_{level}"](low)
        return up1 + F.interpolate(low, scale_factor=2, mode="bicubic", align_corners=True)

    def forward(self, x):
        return self._forward(self.depth, x)


class HGFilter(nn.Module):
    """Stem (7x7 stride-2 conv, ConvBlocks, one 2x down-sampling) + `num_stack` hourglasses with intermediate
    heads — HGFilters.py:72-188.  Returns (outputs per stack, stem features, features before conv3)."""

    def __init__(self, opt):
        super().__init__()
        self.num_modules, self.opt = opt.num_stack, opt
        self.conv1 = nn.Conv2d(opt.hg_input_channel, 64, kernel_size=7, stride=2, padding=3)
        self.bn1 = _norm(opt.norm, 64)
        if opt.hg_down == "conv64":
            self.conv2 = ConvBlock(64, 64, opt.norm)
            self.down_conv2 = nn.Conv2d(64, 128, kernel_size=3, stride=2, padding=1)
        elif opt.hg_down == "conv128":
            self.conv2 = ConvBlock(64, 128, opt.norm)
            self.down_conv2 = nn.Conv2d(128, 128, kernel_size=3, stride=2, padding=1)
        elif opt.hg_down == "ave_pool":
            self.conv2 = ConvBlock(64, 128, opt.norm)
        else:
            raise NameError("Unknown Fan Filter setting!")
        self.conv3 = ConvBlock(128, 128, opt.norm)
        self.conv4 = ConvBlock(128, 256, opt.norm)
        for i in range(self.num_modules):
            self.add_module(f"m{i}", HourGlass(1, opt.num_hourglass, 256, opt.norm))
            self.add_module(f"top_m_{i}", ConvBlock(256, 256, opt.norm))
            self.add_module(f"conv_last{i}", nn.Conv2d(256, 256, kernel_size=1, stride=1, padding=0))
            self.add_module(f"bn_end{i}", _norm(opt.norm, 256))
            self.add_module(f"l{i}", nn.Conv2d(256, opt.hourglass_dim, kernel_size=1, stride=1, padding=0))
            if i < self.num_modules - 1:
                self.add_module(f"bl{i}", nn.Conv2d(256, 256, kernel_size=1, stride=1, padding=0))
                self.add_module(f"al{i}", nn.Conv2d(opt.hourglass_dim, 256, kernel_size=1, stride=1, padding=0))

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)), True)
        tmpx = x
        x = self.conv2(x)
        x = F.avg_pool2d(x, 2, stride=2) if self.opt.hg_down == "ave_pool" else self.down_conv2(x)
        normx = x
        previous = self.conv4(self.conv3(x))
        outputs = []
        for i in range(self.num_modules):
            ll = self._modules[f"top_m_{i}"](self._modules[f"m{i}"](previous))
            ll = F.relu(self._modules[f"bn_end{i}"](self._modules[f"conv_last{i}"](ll)), True)
            out = self._modules[f"l{i}"](ll)
            outputs.append(out)
            if i < self.num_modules - 1:
                previous = previous + self._modules[f"bl{i}"](ll) + self._modules[f"al{i}"](out)
        return outputs, tmpx.detach(), normx


# ---- the residual / depth stems (project/models/helper_modules/helpers.py) ----
def conv3x3(in_planes, out_planes, stride=1, groups=1, dilation=1):
    """helpers.py:250-260 (reflect padding, no bias)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=dilation, groups=groups, bias=False,
                     dilation=dilation, padding_mode="reflect")


def conv1x1(in_planes, out_planes, stride=1):
    """helpers.py:263-270."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False, padding_mode="reflect")


class ResidualBlock(nn.Module):
    """helpers.py:318-371 (the normalised form the local branch uses)."""

    def __init__(self, dim_in, dim_out, dim_inter=None, use_norm=True, norm_layer=nn.BatchNorm2d, bias=False):
        super().__init__()
        dim_inter = dim_out if dim_inter is None else dim_inter
        if use_norm:
            self.conv = nn.Sequential(
                norm_layer(dim_in), nn.ReLU(True),
                nn.Conv2d(dim_in, dim_inter, 3, 1, 1, bias=bias, padding_mode="reflect"),
                norm_layer(dim_inter), nn.ReLU(True),
                nn.Conv2d(dim_inter, dim_out, 3, 1, 1, bias=bias, padding_mode="reflect"))
        else:
            self.conv = nn.Sequential(nn.ReLU(True), nn.Conv2d(dim_in, dim_inter, 3, 1, 1), nn.ReLU(True),
                                      nn.Conv2d(dim_inter, dim_out, 3, 1, 1))
        self.short_cut = nn.Conv2d(dim_in, dim_out, 1, 1) if dim_in != dim_out else None

    def forward(self, feats):
        out = self.conv(feats)
        return out + (self.short_cut(feats) if self.short_cut is not None else feats)


class conv(nn.Module):
    """conv + instance / batch norm + ELU — helpers.py:432-455."""

    def __init__(self, num_in_layers, num_out_layers, kernel_size, stride, norm="in"):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv = nn.Conv2d(num_in_layers, num_out_layers, kernel_size=kernel_size, stride=stride,
                              padding=(kernel_size - 1) // 2, padding_mode="reflect")
        self.bn = (nn.InstanceNorm2d(num_out_layers, track_running_stats=False, affine=True) if norm == "in"
                   else nn.BatchNorm2d(num_out_layers, affine=True, track_running_stats=True))

    def forward(self, x):
        return F.elu(self.bn(self.conv(x)), inplace=True)


def build_filter_modules(owner, local_options):
    """Registers on `owner` (netLocal) the four sub-networks of HGPIFuNetGANResidualResnetFC.__init__ (:28-45) under
    the reference's attribute names: image_filter, downsample_channel_conv, depth_conv, residual_conv."""
    owner.image_filter = HGFilter(local_options)
    owner.downsample_channel_conv = conv(512, 64, 3, 1, norm="in")
    depth_dim = 32
    inorm = lambda dim: nn.InstanceNorm2d(dim, track_running_stats=False, affine=True)
    owner.depth_conv = nn.Sequential(conv3x3(1, depth_dim), ResidualBlock(depth_dim, depth_dim, norm_layer=inorm),
                                     conv1x1(depth_dim, depth_dim))
    owner.residual_conv = nn.Sequential(conv3x3(3, depth_dim), ResidualBlock(depth_dim, depth_dim, norm_layer=inorm),
                                        conv1x1(depth_dim, depth_dim))


def run_filter(owner, residual_images, depth_feat=None, ref_feats=None):
    """`filter` of HGPIFuNetGANResidualResnetFC (:47-75) up to the hourglass: 3 -> 32 residual stem (+ 1 -> 32 depth
    stem) -> stacked hourglass; returns the list of per-stack feature maps."""
    feats = owner.residual_conv(residual_images)
    if ref_feats is not None:
        raise DeprecationWarning("deprecated in editing version model.")  # as the reference (:66)
    if depth_feat is not None:
        feats = torch.cat((feats, owner.depth_conv(depth_feat)), 1)
    outputs, tmpx, normx = owner.image_filter(feats)
    return outputs, tmpx, normx
