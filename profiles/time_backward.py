"""Times the training step of the generator path on one GPU at the BASELINE shape
(size 256, 64x64 rays x 24 samples, batch 8): forward with stash, backward, per-part CUDA events.

    python profiles/time_backward.py [batch]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cvpr23-e3dge_b200"))
sys.path.insert(0, ROOT)

from e3dge_b200 import model_options, rendering_options  # noqa: E402
from e3dge_b200.stylesdf_model import G_pred_latents  # noqa: E402
import synthetic_inputs as P  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.manual_seed(0)
    G = G_pred_latents(model_options(size=256, renderer_spatial_output_dim=64), rendering_options()).eval().cuda()
    for p in G.parameters():
        p.requires_grad_(False)
    inp = {k: v.cuda() for k, v in P.make_inputs(3, B, G.decoder.n_latent, 64, wplus=True).items()}
    args = (inp["cam_poses"], inp["focal"], inp["near"], inp["far"])

    def fwd_infer():
        with torch.no_grad():
            G([inp["w"], inp["w_dec"]], *args, input_is_latent=True, randomize_noise=False)

    state = {}

    def fwd_train():
        w = inp["w"].clone().requires_grad_(True)
        wd = inp["w_dec"].clone().requires_grad_(True)
        out = G([w, wd], *args, input_is_latent=True, randomize_noise=False)
        state.update(w=w, wd=wd, out=out)

    def fwd_bwd():
        fwd_train()
        loss = (state["out"]["gen_imgs"] ** 2).mean() + (state["out"]["gen_thumb_imgs"] ** 2).mean()
        torch.autograd.grad(loss, [state["w"], state["wd"]])

    def render_train():
        w = inp["w"].clone().requires_grad_(True)
        out = G.renderer(*args, styles=w)
        state.update(rw=w, rout=out)

    def render_fwd_bwd():
        render_train()
        loss = (state["rout"]["features"] ** 2).mean() + (state["rout"]["gen_thumb_imgs"] ** 2).mean()
        torch.autograd.grad(loss, [state["rw"]])

    if len(sys.argv) > 2 and sys.argv[2] == "train_only":  # profiler runs: a few training steps only
        for _ in range(3):
            fwd_bwd()
        torch.cuda.synchronize()
        return
    t = {"generator fwd (inference)": timed(fwd_infer), "generator fwd (training, stash)": timed(fwd_train),
         "generator fwd+bwd": timed(fwd_bwd), "renderer fwd (training, stash)": timed(render_train),
         "renderer fwd+bwd": timed(render_fwd_bwd)}
    for k, v in t.items():
        print(f"{k:36s} {v:8.3f} ms   ({B / v * 1e3:8.1f} frames/s)")


if __name__ == "__main__":
    main()
