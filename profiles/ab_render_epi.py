"""A/B of the render kernel's MMA-issue variants at the bench batch (one process per variant: the switch is read
once).  Run under gpurun:  python profiles/ab_render_epi.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, statistics
ROOT = %r
for p in (ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, bench
dev = torch.device("cuda")
G, sd = bench.build_generator(dev)
inp = {k: v.to(dev) for k, v in bench.make_inputs(0).items()}
R = G.renderer
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
with torch.no_grad():
    film = R._film(inp["w"])
    for _ in range(3):
        o = R._render_raw(inp["w"], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], film=film)
    ts = []
    for _ in range(15):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        R._render_raw(inp["w"], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], film=film)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
print("EPI", os.environ.get("E3DGE_RENDER_EPI"), "median %%.4f ms  min %%.4f ms" %% (statistics.median(ts), min(ts)),
      "checksum", float(o["features"].double().abs().sum()))
""" % ROOT
for rep in range(2):
    for epi in ("7", "15"):
        env = dict(os.environ, E3DGE_RENDER_EPI=epi)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or r.stderr[-800:])
