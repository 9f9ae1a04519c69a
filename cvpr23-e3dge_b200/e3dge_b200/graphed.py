"""CUDA-graph replay of a generator pass.

Every launch of this package goes to the caller's stream, allocates only through torch's allocator and never
synchronises with the host (SURVEY.md §8b "Stream"), so a whole `G_pred_latents.forward` — the side stream of
the decoder's latent-only work and the fresh noise draws included — records into one CUDA graph.  An eager
pass costs the host ~35 launches plus the Python around them (about as long as the GPU needs for a batch of
8: any hiccup of the host starves the GPU); a replay costs one launch.

    static = {k: v.clone() for k, v in inputs.items()}          # the graph reads these buffers
    call = GraphedCall(lambda: G([static["w"], static["w_dec"]], static["cam_poses"], ...))
    static["w"].copy_(new_w, non_blocking=True)                   # e.g. straight from pinned host memory
    out = call()                                                 # dict of tensors owned by the graph:
                                                                 # valid until the next call
"""
import torch

from . import _lib


class GraphedCall:
    """Records `fn()` (no arguments: it reads caller-owned static tensors) after `warmup` eager calls and
    replays it on the current stream.  `launches` = kernels of this package in one replay."""

    def __init__(self, fn, warmup=2):
        if not torch.cuda.is_available():
            raise RuntimeError("e3dge_b200: CUDA device required (this framework has no CPU path)")
        self.graph = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):   # lazy one-time work (weight packing, attributes) happens here
                fn()
            side.synchronize()
            n0 = _lib.launch_count
            with torch.cuda.graph(self.graph, stream=side):
                self.result = fn()
            self.launches = _lib.launch_count - n0
        cur.wait_stream(side)
        self.pack_epoch = _lib.pack_epoch

    def __call__(self):
        if _lib.pack_epoch != self.pack_epoch:
            raise RuntimeError("e3dge_b200: weights were re-packed (load_state_dict / .to() / train() / "
                               "invalidate_packed()) after this graph was recorded; the graph holds the old "
                               "packed images — record a new GraphedCall")
        self.graph.replay()
        return self.result
