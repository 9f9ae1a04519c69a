"""Times the local branch's per-sample tail at the bench batch (8 x 64 x 64 x 24 = 786 432 samples):
e3_local_mlp_fwd (prep + six tcgen05 GEMM stages), the texture-modulation MLP alone, and the renderer with /
without the (alpha, beta) input.  Run under gpurun:  python profiles/time_local_branch.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from e3dge_b200 import local_branch as lb  # noqa: E402
from helpers import local_mlp_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
G, sd = bench.build_generator(dev)
inp = {k: v[:B].to(dev) for k, v in bench.make_inputs(0).items()}
lsd = local_mlp_state_dict(7)
TEX = "renderer.network.netLocal.local_feat_to_tex_modulations_linear."
fuse = lb.Fuse_sft_MLP(257, 256)
fuse.load_state_dict({k[len("fuse_sft_block."):]: v for k, v in lsd.items() if k.startswith("fuse_sft_block.")})
net = lb.LocalBranch(None)
net.local_feat_to_tex_modulations_linear.load_state_dict({k[len(TEX):]: v for k, v in lsd.items() if k.startswith(TEX)})
fuse, net = fuse.to(dev).eval(), net.to(dev).eval()
tex = net.local_feat_to_tex_modulations_linear
shp = (B, 64, 64, 24)
rows = B * 64 * 64 * 24
f2 = torch.randn(*shp, 257, device=dev)
f3 = torch.randn(*shp, 256, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timed(fn, n=5):
    for _ in range(2):
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


with torch.no_grad():
    R = G.renderer
    g = R(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=inp["w"])
    pts = g["points"]
    ms_full = timed(lambda: lb.local_tex_modulation(fuse, tex, f2, f3, pts))
    alpha, beta, feats = lb.local_tex_modulation(fuse, tex, f2, f3, pts, return_feats=True)
    ms_tex = timed(lambda: lb.tex_modulation(tex, feats))
    film = R._film(inp["w"])
    ms_r0 = timed(lambda: R._render_raw(inp["w"], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], film=film))
    ms_r1 = timed(lambda: R._render_raw(inp["w"], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], film=film,
                                        local_mod=(alpha, beta)))
MAC = 989161
MAC_PAD = 576 * 256 + 832 * 256 + 256 * 512 + 512 * 512 + 320 * 384 + 640 * 512
print(f"batch {B}: {rows} samples")
print(f"e3_local_mlp_fwd (whole tail)    {ms_full:8.3f} ms   {rows * MAC * 2 / ms_full / 1e9:7.1f} TFLOP/s algorithmic, "
      f"{rows * MAC_PAD * 6 / ms_full / 1e9:7.1f} TFLOP/s executed (3 bf16 products per padded MAC)")
print(f"texture-modulation MLP alone     {ms_tex:8.3f} ms")
print(f"render kernel, global only       {ms_r0:8.3f} ms")
print(f"render kernel + (alpha, beta)    {ms_r1:8.3f} ms   (+{rows * 2048 / 1e9:.2f} GB read)")
print(f"algorithmic HBM bytes of the tail: in {rows * (257 + 256 + 3) * 4 / 1e9:.2f} GB + out {rows * 2048 / 1e9:.2f} GB")
