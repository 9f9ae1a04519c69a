#!/usr/bin/env python
"""clock64() timeline of one tile of the tensor-core render kernel (CTA 0, second tile).
The trace hook is compiled only into a measurement build (round 2: it used to be a live environment lookup in the
production launch path): run under gpurun as
    E3_TRACE=1 python cvpr23-e3dge_b200/build.py --force && python profiles/trace_render.py
and rebuild without E3_TRACE afterwards (a fresh gpurun box starts from the repository's own library)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200"), os.path.join(ROOT, "tests")]
import torch
trace = torch.zeros(512, dtype=torch.int64, device="cuda")
os.environ["E3DGE_RENDER_TRACE_PTR"] = hex(trace.data_ptr())
import bench
dev = torch.device("cuda", 0)
G, sd = bench.build_generator(dev)
inp = {k: v.to(dev) for k, v in bench.make_inputs(0).items()}
R = G.renderer
with torch.no_grad():
    for _ in range(3):
        R._render_raw(inp["w"], inp["cam_poses"], inp["focal"], inp["near"], inp["far"])
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[0]
rel = lambda v: (v - t0) if v else None
print("cluster", os.environ.get("E3DGE_RENDER_CLUSTER", "default"), " (cycles relative to the start of layer 0 of the tile)")
print("layer | d_ready seen | block0..3 published || MMA: a_ready[kb] seen (kb0..3) | last MMA of kb issued (kb0..3)")
for l in range(8):
    comp = [rel(t[l * 8 + i]) for i in range(5)]
    mw = [rel(t[128 + l * 8 + kb]) for kb in range(4)]
    mi = [rel(t[256 + l * 16 + kb * 4 + 3]) for kb in range(4)]
    print(f"L{l}  {comp[0]}  {comp[1:]}  ||  G{l}: {mw}  {mi}")
print("view-layer d_ready seen:", rel(t[64]), " view epilogue done (this warp):", rel(t[66]), " all warps:", rel(t[67]),
      " tile end:", rel(t[65]), " next tile's layer 0 starts:", rel(t[68]))
# inside layer 3 (warp 2 lane 0): per 64-channel block, cycles since the layer's d_ready was seen
base = t[3 * 8]
if t[384]:
    print("layer 3 block j: tcgen05.ld landed | first 8 ch computed | stored | second 8 computed | stored | published")
    for j in range(4):
        print(f"  j{j}  ", [t[384 + j * 8 + k] - base for k in range(6)])
