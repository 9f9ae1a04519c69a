"""Pixel-aligned feature query of the E3DGE local branch — the `query` method of the reference's
`HGPIFuNetGAN` (project/vendor/pifu/lib/model/HGPIFuGANNet.py:85-150) on the path the E3DGE runner uses
(`im_feat=` given or `return_projection_only=True`; e3dge_full_runner.py:219-226, 244-250, 271-278), over
the sm_100a kernel `e3_local_feature_query`.

The hourglass filter that produces `im_feat`, the SFT fusion and the modulation MLP that consume the
queried features stay the caller's PyTorch modules (out of scope)."""
import ctypes

import torch

from . import _lib


def _nhwc(feat):
    """[B,C,H,W] -> channels-last copy through the transpose kernel."""
    lib = _lib.load()
    feat = _lib.as_f32c(feat.detach())
    b, c, h, w = feat.shape
    out = torch.empty(b, h, w, c, device=feat.device, dtype=torch.float32)
    _lib.check(lib.e3_nchw_to_nhwc(_lib.ptr(feat), _lib.ptr(out), b, c, h, w, _lib.cur_stream()), "e3_nchw_to_nhwc")
    return out


class _GatherFn(torch.autograd.Function):
    """Channels-last feature map -> per-point features [B,N,C] with the adjoint for the map
    (e3_local_feature_query / _bwd).  Points and calibration are not differentiated."""

    @staticmethod
    def forward(ctx, fmap, pts, calibs, xy, z, inside):
        lib = _lib.load()
        b, h, w, c = fmap.shape
        n = pts.shape[2]
        feats = torch.empty(b, n, c, device=fmap.device)
        vp = _lib.vptr
        _lib.check(lib.e3_local_feature_query(vp(fmap), vp(pts), pts.stride(0), pts.stride(1), pts.stride(2),
                                              _lib.ptr(calibs), calibs.shape[-2] * 4, b, n, h, w, c, vp(feats),
                                              _lib.ptr(xy), _lib.ptr(z), vp(inside), _lib.cur_stream()),
                   "e3_local_feature_query")
        ctx.save_for_backward(pts, calibs)
        ctx.shape = (b, h, w, c)
        return feats

    @staticmethod
    def backward(ctx, d_feats):
        lib = _lib.load()
        pts, calibs = ctx.saved_tensors
        b, h, w, c = ctx.shape
        d_feats = _lib.as_f32c(d_feats)
        d_map = torch.empty(b, h, w, c, device=d_feats.device)
        vp = _lib.vptr
        _lib.check(lib.e3_local_feature_query_bwd(_lib.ptr(d_feats), vp(pts), pts.stride(0), pts.stride(1),
                                                  pts.stride(2), _lib.ptr(calibs), calibs.shape[-2] * 4, b,
                                                  pts.shape[2], h, w, c, _lib.ptr(d_map), _lib.cur_stream()),
                   "e3_local_feature_query_bwd")
        return d_map, None, None, None, None, None


def query(points, calibs, im_feat=None, im_feat_nhwc=None, return_projection_only=False):
    """points [B,3,N] (any strides: a `.permute(0, 2, 1)` view of the renderer's [B,N,3] points is read in
    place), calibs [B,4,4] or [B,3,4], im_feat [B,C,H,W] (or `im_feat_nhwc` [B,H,W,C], to reuse one
    transposed map for several queries).  Returns the reference's dict: `proj_xy` [B,2,N], `depth` [B,1,N],
    `in_img` [B,N] bool and — unless `return_projection_only` — `feats` / `interp_feats` [B,C,N].
    `feats` is a view of a [B,N,C] buffer, so the `.permute(0, 2, 1)` the runner applies next
    (e3dge_full_runner.py:229-230) is free.  When the feature map requires grad (stage-2 training of the local
    branch) the gather carries its adjoint (e3_local_feature_query_bwd); points / calibs get no gradient."""
    lib = _lib.load()
    if not points.is_cuda:
        raise RuntimeError("e3dge_b200: CUDA tensor required (this framework has no CPU path)")
    if points.ndim != 3 or points.shape[1] != 3:
        raise RuntimeError(f"points must be [B,3,N], got {tuple(points.shape)}")
    if points.dtype != torch.float32:
        points = points.float()
    b, _, n = points.shape
    calibs = _lib.as_f32c(calibs.detach())
    if calibs.shape[0] != b or calibs.shape[-1] != 4 or calibs.shape[-2] not in (3, 4):
        raise RuntimeError(f"calibs must be [B,3,4] or [B,4,4], got {tuple(calibs.shape)}")
    dev = points.device
    xy = torch.empty(b, 2, n, device=dev)
    z = torch.empty(b, 1, n, device=dev)
    inside = torch.empty(b, n, device=dev, dtype=torch.uint8)
    feats = fmap = None
    h = w = c = 4
    pts = points.detach()
    vp = _lib.vptr
    if not return_projection_only:
        if im_feat_nhwc is None:
            if im_feat is None:
                raise RuntimeError("query: im_feat (or im_feat_nhwc) is required unless return_projection_only")
            if torch.is_grad_enabled() and im_feat.requires_grad:  # training: the map's producer gets a gradient
                from .stylesdf_model import _Permute
                im_feat_nhwc = _Permute.apply(im_feat, True)
            else:
                im_feat_nhwc = _nhwc(im_feat)
        if torch.is_grad_enabled() and im_feat_nhwc.requires_grad:
            fmap = im_feat_nhwc if (im_feat_nhwc.dtype == torch.float32 and im_feat_nhwc.is_contiguous()) \
                else im_feat_nhwc.float().contiguous()
            if fmap.shape[0] != b:
                raise RuntimeError("query: one feature map per image of the batch is required")
            feats = _GatherFn.apply(fmap, pts, calibs, xy, z, inside)
        else:
            fmap = _lib.as_f32c(im_feat_nhwc.detach())
            if fmap.shape[0] != b:
                raise RuntimeError("query: one feature map per image of the batch is required")
            _, h, w, c = fmap.shape
            feats = torch.empty(b, n, c, device=dev)
            _lib.check(lib.e3_local_feature_query(vp(fmap), vp(pts), pts.stride(0), pts.stride(1), pts.stride(2),
                                                  _lib.ptr(calibs), calibs.shape[-2] * 4, b, n, h, w, c, vp(feats),
                                                  _lib.ptr(xy), _lib.ptr(z), vp(inside), _lib.cur_stream()),
                       "e3_local_feature_query")
    else:
        _lib.check(lib.e3_local_feature_query(None, vp(pts), pts.stride(0), pts.stride(1), pts.stride(2),
                                              _lib.ptr(calibs), calibs.shape[-2] * 4, b, n, h, w, c, None, _lib.ptr(xy),
                                              _lib.ptr(z), vp(inside), _lib.cur_stream()), "e3_local_feature_query")
    out = {"proj_xy": xy, "depth": z, "in_img": inside.bool()}
    if feats is not None:
        f = feats.permute(0, 2, 1)
        out.update(interp_feats=f, feats=f)
    return out


def install(net_local):
    """Routes `net_local.query(...)` (a reference HGPIFuNetGAN instance) through the CUDA kernel on the two
    paths the E3DGE inference runner uses — `im_feat=` given, or `return_projection_only=True` — and leaves
    every other call (training with eikonal terms, stored `im_feat_dict` features, image-space
    `transforms`) to the module's own method."""
    original = net_local.query

    def patched(points, calibs, feat_key=None, return_eikonal=False, transforms=None, labels=None,
                return_feat_only=False, im_feat=None, return_projection_only=False):
        fast = (transforms is None and not return_eikonal and points.is_cuda
                and not (torch.is_grad_enabled() and points.requires_grad)
                and (return_projection_only or im_feat is not None))
        if not fast:
            return original(points, calibs, feat_key, return_eikonal=return_eikonal, transforms=transforms,
                            labels=labels, return_feat_only=return_feat_only, im_feat=im_feat,
                            return_projection_only=return_projection_only)
        return query(points, calibs, im_feat=im_feat, return_projection_only=return_projection_only)

    net_local.query = patched
    return net_local
