"""A/B timing of e3_render_bwd variants inside one process (same box, same clocks):
flag bit 28 of e3_render_params.flags disables the stash prefetch of the backward kernel."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200"), os.path.join(ROOT, "tests")]
import torch
import bench
from e3dge_b200.volume_renderer import _FilmFn, _RenderFn

dev = torch.device("cuda", 0)
G, sd = bench.build_generator(dev)
for p in G.parameters():
    p.requires_grad_(False)
inp = {k: v.to(dev) for k, v in bench.make_inputs(0).items()}
R = G.renderer
base = R._flags()
ct = torch.randn(8, 256, 64, 64, device=dev)


def run(extra, n=12):
    ts = []
    for i in range(n):
        w = inp["w"].clone().requires_grad_(True)
        film = _FilmFn.apply(R, w)
        vals = _RenderFn.apply(R, film, None, None, inp["cam_poses"], inp["focal"], inp["near"], inp["far"], None,
                               base | extra, False)
        o = dict(zip(R._last_names, vals))
        loss = (ct * o["features"]).sum() + o["gen_thumb_imgs"].sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        g, = torch.autograd.grad(loss, [w])
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts[3:]), g


for rep in range(2):
    t1, g1 = run(0)
    t0, g0 = run(1 << 28)
    print(f"backward (autograd.grad: loss grads + e3_render_bwd + film bwd): prefetch {t1:.3f} ms   "
          f"no prefetch {t0:.3f} ms   same result: {torch.equal(g0, g1)}")
