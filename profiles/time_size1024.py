"""Generator pass at --size 1024 (what the shipped E3DGE scripts run, demo_view_synthesis.sh:35-36): 64^2 x 24 render,
then four up-sampling stages down to 32 channels at 1024^2 (stylesdf_model.py:614-624).
Run under gpurun:  python profiles/time_size1024.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr23-e3dge_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import synthetic_inputs as P  # noqa: E402
from helpers import decoder_layout, synthetic_state_dict  # noqa: E402
from e3dge_b200 import model_options, rendering_options  # noqa: E402
from e3dge_b200.graphed import GraphedCall  # noqa: E402
from e3dge_b200.stylesdf_model import G_pred_latents  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
size, res, seed = 1024, 64, 2024
dev = torch.device("cuda")
sd = synthetic_state_dict(size, res, seed, "sharp")
G = G_pred_latents(model_options(size=size, renderer_spatial_output_dim=res), rendering_options(), full_pipeline=True).eval()
G.load_state_dict(sd, strict=True)
G = G.to(dev)
inp = {k: v.to(dev) for k, v in P.make_inputs(seed, B, decoder_layout(size, res), res).items()}
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def core():
    with torch.no_grad():
        return G([inp["w"], inp["w_dec"]], inp["cam_poses"], inp["focal"], inp["near"], inp["far"],
                 input_is_latent=True, randomize_noise=True)


def med(fn, n=7):
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


ms_eager = med(core)
if os.environ.get("E3DGE_BENCH_EAGER"):
    print(f"size 1024, batch {B}: eager {ms_eager:.3f} ms per step = {B / ms_eager * 1e3:.1f} frames/s")
else:
    call = GraphedCall(core)
    ms = med(call)
    print(f"size 1024, batch {B}: graph replay {ms:.3f} ms per step = {B / ms * 1e3:.1f} frames/s (eager {ms_eager:.3f} ms); "
          f"decoder MACs 62.8 G/image -> {62.8e9 * 2 * B / ms / 1e9:.1f} TFLOP/s algorithmic")
