"""CPU restatement of the local branch's pixel-aligned feature query — TEST
INFRASTRUCTURE ONLY: nothing outside tests/, __graft_entry__.smoke() and bench.py's CPU arm may import it.

Follows `HGPIFuNetGAN.query(points, calibs, feat_key, return_feat_only=True, im_feat=...)` as the E3DGE
runner calls it (project/trainers/E3DGE/e3dge_full_runner.py:219-226, 271-278):
  * projection  = `perspective` (netLocal is built with projection_mode='projection',
    project/utils/volume_renderer.py:293-300; vendor/pifu/lib/geometry.py:108-135):
        homo = calib[:, :3, :3] @ p + calib[:, :3, 3];  z = -homo_z if homo[0, 2, 0] < 0 else homo_z
        xy = homo_xy / z
  * y is flipped for grid_sample's top-left origin, `in_img` = |x| <= 1 and |y| <= 1
    (vendor/pifu/lib/model/HGPIFuGANNet.py:113-124)
  * features = bilinear grid_sample, zero padding, align_corners=False
    (vendor/pifu/lib/geometry.py:64-80 `index`; project/models/op/grid_sample_gradfix.py:29-36)
Pinned by tests/golden/local_query.npz, recorded from the real reference functions
(oracle/gen_golden_local_query.py).
"""
import torch


def perspective(points, calibs):
    """points [B,3,N], calibs [B,4,4] or [B,3,4] -> xyz [B,3,N] (x, y in the image plane, depth z)."""
    rot, trans = calibs[:, :3, :3], calibs[:, :3, 3:4]
    homo = trans + rot @ points                                   # geometry.py:118-122 (baddbmm)
    z = homo[:, 2:3] * (-1 if homo[0, -1, 0] < 0 else 1)            # geometry.py:124-127 ("look at -z")
    return torch.cat([homo[:, :2] / z, z], 1)                     # geometry.py:130-135


def bilinear_zero_pad(feat, xy):
    """feat [B,C,H,W], xy [B,2,N] in [-1,1] (x = width axis) -> [B,C,N]; align_corners=False, zeros."""
    B, C, H, W = feat.shape
    ix = ((xy[:, 0] + 1) * W - 1) / 2                              # grid_sampler_unnormalize, align_corners=False
    iy = ((xy[:, 1] + 1) * H - 1) / 2
    x0, y0 = torch.floor(ix), torch.floor(iy)
    out = feat.new_zeros(B, C, xy.shape[-1])
    flat = feat.reshape(B, C, H * W)
    for dy in (0, 1):
        for dx in (0, 1):
            xx, yy = x0 + dx, y0 + dy
            wgt = (1 - (ix - xx).abs()) * (1 - (iy - yy).abs())    # (x1-ix)(y1-iy) etc.
            ok = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
            idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long()
            tap = torch.gather(flat, 2, idx[:, None, :].expand(B, C, -1))
            out = out + tap * (wgt * ok)[:, None, :]
    return out


def local_feature_query(points, calibs, im_feat):
    """The dict HGPIFuNetGAN.query returns on the `im_feat is not None` path (HGPIFuGANNet.py:131-150)."""
    xyz = perspective(points, calibs).clone()
    xyz[:, 1] = -xyz[:, 1]
    xy, z = xyz[:, :2], xyz[:, 2:3]
    in_img = (xy[:, 0] >= -1.0) & (xy[:, 0] <= 1.0) & (xy[:, 1] >= -1.0) & (xy[:, 1] <= 1.0)
    feats = bilinear_zero_pad(im_feat, xy)
    return {"proj_xy": xy, "depth": z, "in_img": in_img, "interp_feats": feats, "feats": feats}
