#!/usr/bin/env python
"""Timings of the tensor-core render kernel (B=8, 64x64x24) under what-if variants:
  * cluster size of the shared (TMA-multicast) weight stream: E3DGE_RENDER_CLUSTER=1|2|4
    (one process per value: the choice is cached at first launch);
  * timing-only debug flag bits (skip the lo MMA passes / skip the sin / no weight TMA: the MMAs read
    whatever the ring holds and never wait for a tile) — numerically WRONG results.
Run under gpurun:  for c in 1 2 4; do E3DGE_RENDER_CLUSTER=$c python profiles/whatif_render.py; done"""
import os, sys, statistics
os.environ.setdefault("E3DGE_RENDER_EPI", "0")  # the debug flags exist in the scalar-epilogue variant only
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200"), os.path.join(ROOT, "tests")]
import torch
import bench

dev = torch.device("cuda", 0)
G, sd = bench.build_generator(dev)
inp = {k: v.to(dev) for k, v in bench.make_inputs(0).items()}
R = G.renderer
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
base = R._flags()
cl = os.environ.get("E3DGE_RENDER_CLUSTER", "default(2)")
for name, extra in (("full", 0), ("skip_lo_mma", 1 << 30), ("no_sin", 1 << 31), ("skip_lo+no_sin", (1 << 30) | (1 << 31)),
                    ("no_weight_stream", 1 << 29), ("no_wstream+no_sin", (1 << 29) | (1 << 31)),
                    ("no_wstream+skip_lo+no_sin", (1 << 29) | (1 << 30) | (1 << 31))):
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
    print(f"cluster={cl:10s} {name:16s} {statistics.median(ts[2:]):.3f} ms")
