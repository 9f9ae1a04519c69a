"""HBM roofline of the two stand-alone ops of `project.models.op` (bytes = in + out) at the
decoder's largest activation, through the public op API.  Run under gpurun:  python profiles/time_ops.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from e3dge_b200.op import fused_leaky_relu, upfirdn2d  # noqa: E402

peak = 6532.5
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(pk):
    peak = json.load(open(pk))["hbm_gbs"]
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
k = torch.tensor([1., 3., 3., 1.], device=dev)
k = k[None] * k[:, None]
k = k / k.sum()


def timed(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


x = torch.randn(8, 128, 256, 256, device=dev)
xs = torch.randn(8, 128, 128, 128, device=dev)
xt = torch.randn(8, 128, 257, 257, device=dev)
bias = torch.randn(128, device=dev)
with torch.no_grad():
    cases = [
        ("fused_leaky_relu [8,128,256,256] (+bias)", lambda: fused_leaky_relu(x, bias), 2 * x.numel() * 4),
        ("upfirdn2d Blur: up 1, pad (1,1) on [8,128,257,257] -> 256^2", lambda: upfirdn2d(xt, k * 4, pad=(1, 1)),
         (xt.numel() + x.numel()) * 4),
        ("upfirdn2d Upsample: up 2, pad (2,1) on [8,128,128,128] -> 256^2", lambda: upfirdn2d(xs, k * 4, up=2, pad=(2, 1)),
         (xs.numel() + x.numel()) * 4),
        ("upfirdn2d down 2, pad (1,1) on [8,128,256,256] -> 128^2", lambda: upfirdn2d(x, k, down=2, pad=(1, 1)),
         (x.numel() + xs.numel()) * 4),
    ]
    for name, fn, nbytes in cases:
        ms = timed(fn)
        gbs = nbytes / ms / 1e6
        print(f"{name:70s} {ms:7.3f} ms  {gbs:7.0f} GB/s  = {gbs / peak:5.1%} of {peak:.0f} GB/s (bytes = in + out = {nbytes / 1e6:.0f} MB)")
