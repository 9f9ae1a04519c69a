"""The per-sample half of the E3DGE local branch (`--enable_local_model`, what every shipped script runs):

    feats_2d, feats_3d  = netLocal.query(...)                     e3dge_full_runner.py:219-230, 271-281   (local_query.py)
    fused               = Fuse_sft_MLP(257, 256)(feats_2d, feats_3d)       :289-290, sft.py:84-109
    feats               = cat(fused, PosEncoding(3, 7)(points))            :293-294, misc_utils.py:148-184
    alpha, beta         = split(netLocal.local_feat_to_tex_modulations_linear(feats), 256)
                                                                          volume_renderer.py:327-336
    h8'                 = (alpha + 1) * h8 + beta  inside the renderer     :217-220

Module classes, constructor arguments and state_dict names follow the reference so that E3DGE checkpoints load
(`fuse_sft_block.*` of the runner's network dict, `renderer.network.netLocal.local_feat_to_tex_modulations_linear.*`).
On a CUDA device and outside autograd the whole chain from the queried features to (alpha, beta) is
`e3_local_mlp_fwd` (csrc/local_mlp.cu: six tcgen05 GEMM stages, operands handed from epilogue to epilogue as
bf16 hi / lo halves).  The modules' own `forward` methods are plain PyTorch (autograd-capable: the reference
trains these layers in stage 2; the backward of the fused chain is not built)."""
import ctypes
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import local_query as _lq


def _require_cuda(t):
    """The local branch has no CPU path either: its modules run the fused kernel (inference) or device-side
    PyTorch (training, autograd) on CUDA tensors only."""
    if not t.is_cuda:
        raise RuntimeError("e3dge_b200: CUDA tensor required (this framework has no CPU path)")


class ResnetBlockFC(nn.Module):
    """x_s + fc_1(relu(fc_0(relu(x)))) — project/models/helper_modules/resnetfc.py:10-62."""

    def __init__(self, size_in, size_out=None, size_h=None, beta=0.0):
        super().__init__()
        size_out = size_in if size_out is None else size_out
        size_h = min(size_in, size_out) if size_h is None else size_h
        if beta > 0:
            raise NotImplementedError("Softplus ResnetBlockFC is not used on the E3DGE path")
        self.size_in, self.size_h, self.size_out = size_in, size_h, size_out
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        nn.init.constant_(self.fc_0.bias, 0.0)
        nn.init.kaiming_normal_(self.fc_0.weight, a=0, mode="fan_in")
        nn.init.constant_(self.fc_1.bias, 0.0)
        nn.init.zeros_(self.fc_1.weight)
        self.activation = nn.ReLU()
        if size_in == size_out:
            self.shortcut = None
        else:
            self.shortcut = nn.Linear(size_in, size_out, bias=False)
            nn.init.kaiming_normal_(self.shortcut.weight, a=0, mode="fan_in")

    def forward(self, x):
        _require_cuda(x)
        if (self.size_in, self.size_out) == (301, 512) and _fused_ok(x, self):
            a, b = tex_modulation(self, x)
            return torch.cat([a, b], -1)
        net = self.fc_0(self.activation(x))
        dx = self.fc_1(self.activation(net))
        return (self.shortcut(x) if self.shortcut is not None else x) + dx


class Fuse_sft_MLP(nn.Module):
    """dec + w * (dec * scale(e) + shift(e)), e = ResnetBlockFC(cat[enc, dec]) — sft.py:84-109."""

    def __init__(self, in_ch=256 + 1, out_ch=256):
        super().__init__()
        self.encode_enc = ResnetBlockFC(in_ch + out_ch, out_ch)
        self.scale = nn.Sequential(nn.Linear(out_ch, out_ch), nn.LeakyReLU(0.2, True), nn.Linear(out_ch, out_ch))
        self.shift = nn.Sequential(nn.Linear(out_ch, out_ch), nn.LeakyReLU(0.2, True), nn.Linear(out_ch, out_ch))

    def forward(self, enc_feat, dec_feat, w=1):
        _require_cuda(dec_feat)
        enc_feat = self.encode_enc(torch.cat([enc_feat, dec_feat], dim=-1))
        scale = self.scale(enc_feat)
        shift = self.shift(enc_feat)
        return dec_feat + w * (dec_feat * scale + shift)


class PosEncoding(nn.Module):
    """(x, sin(2^k x), cos(2^k x))_k — project/utils/misc_utils.py:148-184."""

    def __init__(self, in_channels, N_freqs, logscale=True):
        super().__init__()
        self.N_freqs, self.in_channels = N_freqs, in_channels
        self.funcs = [torch.sin, torch.cos]
        self.out_channels = in_channels * (len(self.funcs) * N_freqs + 1)
        self.freq_bands = (2 ** torch.linspace(0, N_freqs - 1, N_freqs) if logscale
                           else torch.linspace(1, 2 ** (N_freqs - 1), N_freqs))

    def forward(self, x):
        out = [x]
        for freq in self.freq_bands:
            for func in self.funcs:
                out += [func(freq * x)]
        return torch.cat(out, -1)


class LocalBranch(nn.Module):
    """`netLocal` of SirenLocalGlobal: the part of the reference's HGPIFuNetGANResidualResnetFC
    (vendor/pifu/lib/model/HGPIFuGANNetResidualInputResnetFC.py:19-97) that runs per sample —
    `query` (pixel-aligned feature gather, csrc/local_query.cu) and `local_feat_to_tex_modulations_linear`
    (zero-initialised ResnetBlockFC(301, 512), :84-97) — plus, when the PIFu option group is given, the 2-D
    `filter` (residual / depth stems + stacked hourglass, local_filter.py: PyTorch / cuDNN library code under the
    reference's parameter names).  Without it `attach_filter(module)` plugs the caller's filter in."""

    def __init__(self, opt=None, local_options=None):
        super().__init__()
        self.opt = opt
        # the 2-D image filter (hourglass + stems) under the reference's names, when the PIFu option group is given
        # (`opt.pifu`, volume_renderer.py:741): library-code conv nets, see local_filter.py
        self.local_options = local_options
        if local_options is not None:
            from .local_filter import build_filter_modules
            build_filter_modules(self, local_options)
        dim = int(getattr(opt, "residual_local_feats_dim", 301)) if opt is not None else 301
        if opt is None or getattr(opt, "L_pred_tex_modulations", True):
            m = ResnetBlockFC(dim, 256 * 2)
            for p in (m.fc_0.bias, m.fc_0.weight, m.fc_1.bias, m.shortcut.weight):
                nn.init.zeros_(p)
            self.local_feat_to_tex_modulations_linear = m
        if opt is not None and getattr(opt, "L_pred_geo_modulations", False):
            raise NotImplementedError("geometry modulation of the local branch (volume_renderer.py:338-345) is "
                                      "not built; the shipped scripts predict texture modulation only")
        self.image_filter_module = None
        self.im_feat_dict = {}

    def attach_filter(self, module):
        """module(residual_images, depth_feat=None, ...) -> feature map [B,C,H,W] (the reference's hourglass stack)."""
        object.__setattr__(self, "image_filter_module", module)
        return self

    def filter(self, residual_images, depth_feat=None, ref_feats=None, feat_key="ref_view", return_feat=False,
               *args, **kwargs):
        if self.image_filter_module is not None:  # a caller-attached filter takes precedence
            feat = self.image_filter_module(residual_images, depth_feat=depth_feat, ref_feats=ref_feats, **kwargs)
            feats = list(feat) if isinstance(feat, (list, tuple)) else [feat]
        elif self.local_options is not None:
            from .local_filter import run_filter
            outputs, self.tmpx, self.normx = run_filter(self, residual_images, depth_feat, ref_feats)
            feats = [outputs[-1]]  # only the last stack is kept (HGPIFuNet.py:91-97)
        else:
            raise NotImplementedError("netLocal was built without its image filter (no `pifu` option group in the "
                                      "rendering options); pass one, or attach the caller's filter with "
                                      "LocalBranch.attach_filter(module)")
        self.im_feat_dict[feat_key] = feats
        return feats if return_feat else None

    def query(self, points, calibs, feat_key=None, return_eikonal=False, transforms=None, labels=None,
              return_feat_only=False, im_feat=None, return_projection_only=False):
        """HGPIFuGANNet.py:85-155 on the paths the E3DGE runner uses (`im_feat=` given, stored `feat_key`
        features with `return_feat_only`, or `return_projection_only`)."""
        if transforms is not None or return_eikonal:
            raise NotImplementedError("query: image-space transforms / eikonal terms are not on the E3DGE path")
        if im_feat is None and not return_projection_only:
            if feat_key not in self.im_feat_dict:
                raise RuntimeError(f"query: no filtered features stored under {feat_key!r}")
            im_feat = self.im_feat_dict[feat_key][-1]
        return _lq.query(points, calibs, im_feat=im_feat, return_projection_only=return_projection_only)


# ---------------------------------------------------------------------------------------------- fused chain
class _PackedLocal:
    """Operand image of (fusion MLP, texture MLP) for e3_local_mlp_fwd, rebuilt when a weight changes."""

    def __init__(self):
        self.key, self.buf = None, None

    def get(self, fuse, tex):
        tensors = _fuse_tensors(fuse) + _tex_tensors(tex)
        key = (_lib.pack_epoch,) + tuple((None if t is None else (t.data_ptr(), t._version)) for t in tensors)
        if self.buf is not None and key == self.key:
            return self.buf
        lib = _lib.load()
        dev = tex.fc_0.weight.device
        if dev.type != "cuda":
            raise RuntimeError("e3dge_b200: local-branch weights must live on a CUDA device")
        keep = [None if t is None else _lib.as_f32c(t.detach()) for t in tensors]
        w = _lib.LocalMlpWeights(*[_lib.ptr(t) for t in keep])
        buf = torch.empty(lib.e3_local_mlp_packed_bytes() // 4 + 1, device=dev, dtype=torch.float32)
        _lib.check(lib.e3_local_mlp_pack(ctypes.byref(w), _lib.ptr(buf), _lib.cur_stream()), "e3_local_mlp_pack")
        self.key, self.buf = key, buf
        return buf


def _fuse_tensors(fuse):
    if fuse is None:
        return [None] * 13
    e = fuse.encode_enc
    if (e.fc_0.weight.shape, e.fc_1.weight.shape) != ((256, 513), (256, 256)):
        raise NotImplementedError("the fused local chain is built for Fuse_sft_MLP(257, 256)")
    return [e.fc_0.weight, e.fc_0.bias, e.fc_1.weight, e.fc_1.bias, e.shortcut.weight,
            fuse.scale[0].weight, fuse.scale[0].bias, fuse.scale[2].weight, fuse.scale[2].bias,
            fuse.shift[0].weight, fuse.shift[0].bias, fuse.shift[2].weight, fuse.shift[2].bias]


def _tex_tensors(tex):
    if (tex.fc_0.weight.shape, tex.fc_1.weight.shape) != ((301, 301), (512, 301)):
        raise NotImplementedError("the fused local chain is built for ResnetBlockFC(301, 512)")
    return [tex.fc_0.weight, tex.fc_0.bias, tex.fc_1.weight, tex.fc_1.bias, tex.shortcut.weight]


def _packed_for(fuse, tex):
    holder = tex.__dict__.setdefault("_e3_packed", {})
    slot = holder.setdefault(id(fuse) if fuse is not None else None, _PackedLocal())
    return slot.get(fuse, tex)


def _fused_ok(x, *modules):
    if not (torch.is_tensor(x) and x.is_cuda):
        return False
    if not torch.is_grad_enabled():
        return True
    return not (x.requires_grad or any(p.requires_grad for m in modules if m is not None for p in m.parameters()))


# rows of one e3_local_mlp_fwd pass: its workspace (12.8 KB per row) is sized for this many
CHUNK_ROWS = 128 * 1024


def _run(packed, f2, f3, pts, feats_in, rows, want_feats, dev):
    lib = _lib.load()
    alpha = torch.empty(rows, 256, device=dev, dtype=torch.float32)
    beta = torch.empty(rows, 256, device=dev, dtype=torch.float32)
    feats = torch.empty(rows, 301, device=dev, dtype=torch.float32) if want_feats else None
    nbytes = lib.e3_local_mlp_workspace_bytes(min(max(rows, 128), CHUNK_ROWS))
    ws = torch.empty(nbytes // 4 + 1, device=dev, dtype=torch.float32)
    _lib.check(lib.e3_local_mlp_fwd(_lib.ptr(packed), _lib.ptr(f2), _lib.ptr(f3), _lib.ptr(pts), _lib.ptr(feats_in),
                                    rows, _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(feats), _lib.ptr(ws), nbytes,
                                    _lib.cur_stream()), "e3_local_mlp_fwd",
               launches=(7 if feats_in is None else 3) * max(1, -(-rows // CHUNK_ROWS)))
    return alpha, beta, feats


def local_tex_modulation(fuse, tex, feat_2d, feat_3d, points, return_feats=False):
    """The whole tail in one call: feat_2d [...,257] (2-D-aligned features | visibility mask), feat_3d [...,256],
    points [...,3] (world space) -> (alpha, beta) [...,256] (+ the 301-d `feats` the reference materialises).
    CUDA, inference (no autograd): e3_local_mlp_fwd."""
    if not _fused_ok(feat_2d, fuse, tex) or not _fused_ok(feat_3d) or not _fused_ok(points):
        raise RuntimeError("e3dge_b200: local_tex_modulation is the fused inference path (CUDA tensors, no autograd); "
                           "for training call the modules (Fuse_sft_MLP / PosEncoding / ResnetBlockFC) directly")
    lead = tuple(feat_3d.shape[:-1])
    rows = int(math.prod(lead))
    if feat_2d.shape[-1] != 257 or feat_3d.shape[-1] != 256 or points.shape[-1] != 3:
        raise RuntimeError("local_tex_modulation: expected [...,257], [...,256], [...,3]")
    f2 = _lib.as_f32c(feat_2d.detach()).reshape(rows, 257)
    f3 = _lib.as_f32c(feat_3d.detach()).reshape(rows, 256)
    pts = _lib.as_f32c(points.detach()).reshape(rows, 3)
    alpha, beta, feats = _run(_packed_for(fuse, tex), f2, f3, pts, None, rows, return_feats, f3.device)
    out = (alpha.reshape(*lead, 256), beta.reshape(*lead, 256))
    return out + (feats.reshape(*lead, 301),) if return_feats else out


def tex_modulation(tex, feats):
    """The reference's contract: `local_data_batch['feats']` [...,301] -> (alpha, beta) through
    netLocal.local_feat_to_tex_modulations_linear (volume_renderer.py:327-336), stages 5-6 of the fused chain."""
    lead = tuple(feats.shape[:-1])
    rows = int(math.prod(lead))
    f = _lib.as_f32c(feats.detach()).reshape(rows, 301)
    alpha, beta, _ = _run(_packed_for(None, tex), None, None, None, f, rows, False, f.device)
    return alpha.reshape(*lead, 256), beta.reshape(*lead, 256)
