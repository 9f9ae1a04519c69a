#!/usr/bin/env python
"""Eager launches vs CUDA-graph replay of the inference step of bench.py (B=8, size 256, 64x64x24, fresh noise):
how much of a step is launch gaps.  Run under gpurun:  python profiles/time_graph.py"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200"), os.path.join(ROOT, "tests")]
import torch
import bench

dev = torch.device("cuda", 0)
G, sd = bench.build_generator(dev)
buf = {k: v.to(dev) for k, v in bench.make_inputs(0).items()}
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def step():
    with torch.no_grad():
        return G([buf["w"], buf["w_dec"]], buf["cam_poses"], buf["focal"], buf["near"], buf["far"],
                 input_is_latent=True, randomize_noise=True, return_xyz=True, return_sdf=True)["gen_imgs"]


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


eager = timed(step)
s = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        out = step()
torch.cuda.synchronize()
graph = timed(g.replay)
print(f"inference step, eager launches: {eager:.3f} ms   CUDA-graph replay: {graph:.3f} ms   "
      f"difference {eager - graph:.3f} ms ({100 * (eager - graph) / eager:.1f} %)")
