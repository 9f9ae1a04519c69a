#!/usr/bin/env python
"""Records tests/golden/local_query.npz from the REAL reference code: `HGPIFuNetGAN.query` (unbound, on a
stand-in `self` that carries the reference's own `perspective` / `index` functions — the method only uses
self.projection, self.index, self.normalizer and self.opt on this path), executed in place from
/root/reference (no file copied).  Run in the build container:  python oracle/gen_golden_local_query.py"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
np.deprecate = lambda f=None, *a, **k: (f if callable(f) else (lambda g: g))  # vendor/pifu/lib/geometry.py:1
import torch  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

H.load_reference()
PIFU = os.path.join(H.REFERENCE_ROOT, "project", "vendor", "pifu", "lib")
H._shell_package("pifu_lib", PIFU)
H._shell_package("pifu_lib.model", os.path.join(PIFU, "model"))
import pifu_lib.geometry as G  # noqa: E402
import pifu_lib.model.HGPIFuGANNet as M  # noqa: E402


def make_case(seed, B, C, Hh, Ww, N, look_neg_z):
    """Cameras on a sphere of radius 1 looking at the origin (camera_utils.py:85-151 builds calibs the same
    way: uv-space intrinsics @ w2c extrinsics), points in the [-0.15, 0.15]^3 volume plus a few far outside
    the frustum so that in_img and the zero padding are exercised."""
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, C, Hh, Ww, generator=g)
    pts = (torch.rand(B, 3, N, generator=g) - 0.5) * 0.3
    pts[:, :, -N // 8:] *= 6.0
    az, el = (torch.rand(B, generator=g) - 0.5) * 0.9, (torch.rand(B, generator=g) - 0.5) * 0.45
    loc = torch.stack([torch.sin(az) * torch.cos(el), torch.sin(el), torch.cos(az) * torch.cos(el)], 1)
    zax = loc / loc.norm(dim=1, keepdim=True)
    up = torch.tensor([0.0, 1.0, 0.0]).expand(B, 3)
    xax = torch.cross(up, zax, dim=1)
    xax = xax / xax.norm(dim=1, keepdim=True)
    yax = torch.cross(zax, xax, dim=1)
    w2c = torch.stack([xax, yax, zax], 1)
    ext = torch.cat([w2c, -w2c @ loc[:, :, None]], -1)
    f = 0.5 / np.tan(np.deg2rad(6.0))  # focal / (res / 2), fov 12 degrees (camera_utils.py:29-35)
    K = torch.zeros(B, 3, 3)
    K[:, 0, 0] = K[:, 1, 1] = 2 * f
    K[:, 2, 2] = 1.0
    if not look_neg_z:
        ext = ext * torch.tensor([1.0, 1.0, -1.0])[None, :, None]
    calibs = torch.cat([K @ ext, torch.tensor([0.0, 0, 0, 1]).expand(B, 1, 4)], 1)
    return feat, pts, calibs


def run(feat, pts, calibs):
    fake = types.SimpleNamespace(projection=G.perspective, index=G.index, normalizer=lambda z, calibs=None: z,
                                 opt=types.SimpleNamespace(skip_hourglass=False))
    with torch.no_grad():
        out = M.HGPIFuNetGAN.query(fake, points=pts.clone(), calibs=calibs, feat_key="ref_view",
                                   return_eikonal=False, return_feat_only=True, im_feat=feat)
    return out


def main():
    rec = {}
    for name, args in {"neg_z": (11, 2, 16, 12, 20, 240, True), "pos_z": (12, 3, 8, 16, 16, 128, False)}.items():
        feat, pts, calibs = make_case(*args)
        out = run(feat, pts, calibs)
        rec.update({f"{name}.feat": feat.numpy(), f"{name}.points": pts.numpy(), f"{name}.calibs": calibs.numpy(),
                    f"{name}.proj_xy": out["proj_xy"].numpy(), f"{name}.depth": out["depth"].numpy(),
                    f"{name}.in_img": out["in_img"].numpy(), f"{name}.feats": out["feats"].numpy()})
        print(name, "in_img fraction", out["in_img"].float().mean().item(), "feats", tuple(out["feats"].shape))
    path = os.path.join(ROOT, "tests", "golden", "local_query.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
