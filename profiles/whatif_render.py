#!/usr/bin/env python
"""What-if timings of the tensor-core render kernel (B=8, 64x64x24): which side bounds the
k-block-pipelined layer loop?  Uses the kernel's timing-only debug flag bits (results are
numerically wrong with them set).  Run under gpurun; prints one line per variant."""
import os, sys, statistics, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200"), os.path.join(ROOT, "tests")]
import torch
import bench
from e3dge_b200 import _lib

dev = torch.device("cuda", 0)
G, sd = bench.build_generator(dev)
inp = {k: v.to(dev) for k, v in bench.make_inputs(0).items()}
R = G.renderer
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
base = R._flags()
for name, extra in (("full", 0), ("skip_lo_mma", 1 << 30), ("no_sin", 1 << 31), ("skip_lo+no_sin", (1 << 30) | (1 << 31))):
    with torch.no_grad():
        film = R._film(inp["w"])
        ts = []
        for i in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            R._render_raw(inp["w"], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], film=film,
                          flags_over=base | extra)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    print(f"{name:16s} {statistics.median(ts[2:]):.3f} ms")
