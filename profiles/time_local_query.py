#!/usr/bin/env python
"""Timing of e3_local_feature_query at the E3DGE shapes (B = 8 images, 64x64x24 = 98 304 sample points each,
256-channel 128x128 feature map): algorithmic bytes = the [B,N,C] fp32 output written once (805 MB) + the
maps read once (134 MB); the four taps per point are L2 / L1 hits.  Run under gpurun."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from e3dge_b200 import local_query

B, C, H, W, N = 8, 256, 128, 128, 64 * 64 * 24
g = torch.Generator().manual_seed(0)
fmap = torch.randn(B, H, W, C, generator=g).cuda()
pts = ((torch.rand(B, N, 3, generator=g) - 0.5) * 0.2).cuda()
calibs = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "local_query.npz"))["neg_z.calibs"])[:1].repeat(B, 1, 1).cuda()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
ts = []
for i in range(13):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = local_query.query(pts.permute(0, 2, 1), calibs, im_feat_nhwc=fmap)
    b.record()
    torch.cuda.synchronize()
    if i >= 3:
        ts.append(a.elapsed_time(b))
ms = statistics.median(ts)
byts = B * N * C * 4 + B * H * W * C * 4 + B * N * (12 + 13)
print(f"e3_local_feature_query B={B} N={N} C={C} map {H}x{W}: {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s algorithmic "
      f"({byts / 1e6:.0f} MB; in-image fraction {out['in_img'].float().mean().item():.2f})")
