#!/usr/bin/env python
"""bench.py — inversion frames/sec of the StyleSDF generator hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): FFHQ StyleSDF 256^2 generator pass of the inversion
loop — G_pred_latents.forward(styles=[w+, w_dec], input_is_latent=True) = 64x64 rays x 24
samples through the 9-layer FiLM-SIREN + composite, then the modulated-conv decoder to
256^2 — batch 8 per GPU, synthetic random latents / cameras, random-init weights
(no checkpoints offline).  One "step" = one such batch.  Weak scaling: every rank runs its
own batch of 8 and the per-image records are all-gathered once per step.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the
same through the public module API with pinned-host inputs copied in and the image copied
out every step.  `--impl reference` times the CPU implementation (the oracle port of the
reference's PyTorch code; the Python reference itself cannot travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "cvpr23-e3dge_b200")
for _p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

SIZE, RES, N_SAMPLES, BATCH, SEED = 256, 64, 24, 8, 2024
# arithmetic of the timed path, not a precision claim: fp32 tensors in and out; every 256x256 / conv
# contraction as three bf16 tensor-core products of hi/lo operand halves (hi*hi + lo*hi + hi*lo, ~16-17
# mantissa bits) accumulated in fp32; layer 0, FiLM, sin, heads, composite, blur in fp32 on the CUDA cores
DTYPE = "f32 I/O; split-bf16x3 tcgen05 contractions, f32 accumulate"
METRIC = "inversion frames/sec @256^2 StyleSDF (generator pass: 64x64 rays x 24 samples + decoder)"
WORKLOAD = ("FFHQ StyleSDF 256^2 inversion, 64 rays x 24 samples, batch=8 per GPU "
            "(BASELINE.json configs[1])")
# algorithmic figures per image (SURVEY.md §8d / BASELINE.md §4)
RENDER_FLOP_PER_IMAGE = 98304 * 526848 * 2
RENDER_BYTES_PER_IMAGE = 6864896 + 9276  # full dict contract out + styles/camera in


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0 = index, [], None, 0.0

    def mark(self):
        """Samples read before this call (nvidia-smi start-up, idle GPU) are not part of the summary."""
        self.t0 = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for ts, r in self.rows:
            if ts < self.t0:
                continue
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_generator(device):
    from helpers import synthetic_state_dict
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    sd = synthetic_state_dict(SIZE, RES, SEED, "sharp")
    G = G_pred_latents(model_options(size=SIZE, renderer_spatial_output_dim=RES),
                       rendering_options(N_samples=N_SAMPLES), full_pipeline=True).eval()
    G.load_state_dict(sd, strict=True)
    return G.to(device), sd


def set_backends(G, which):
    """"auto" = the tcgen05 split-bf16 kernels wherever the shape allows (default); "fp32" = the exact-fp32
    CUDA-core kernels for renderer and decoder."""
    from e3dge_b200.stylesdf_model import ModulatedConv2d
    G.renderer.backend = "fp32" if which == "fp32" else "tensor_cores"
    for m in G.modules():
        if isinstance(m, ModulatedConv2d):
            m.backend = which


def make_inputs(rank):
    import synthetic_inputs as P  # deterministic synthetic latents / cameras (input generator)
    from helpers import decoder_layout
    return P.make_inputs(SEED + rank, BATCH, decoder_layout(SIZE, RES), RES)


def run_ours(args):
    global BATCH
    import torch.distributed as dist
    from e3dge_b200 import _lib, parallel as par
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.scaling == "strong":  # fixed total work: the global batch is split into contiguous shards
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        BATCH = args.global_batch // world
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun "
                         f"--nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    G, sd = build_generator(dev)
    host = {k: v.pin_memory() for k, v in make_inputs(rank).items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    # e2e: the step's inputs (latents, cameras: 172 KB) travel as ONE pinned buffer and one copy per step
    # instead of six small ones; the device side slices views out of it
    offs, total = {}, 0
    for k, v in host.items():
        offs[k] = (total, v.numel(), tuple(v.shape))
        total += (v.numel() + 3) // 4 * 4  # keep every view 16-byte aligned
    host_packed = torch.empty(total).pin_memory()
    for k, v in host.items():
        o, n, _ = offs[k]
        host_packed[o:o + n].copy_(v.reshape(-1))
    n_lat = resident["w_dec"].shape[1]
    target = torch.zeros(BATCH, 3, SIZE, SIZE, device=dev)
    rec_local = torch.empty(BATCH, par.record_length(n_lat), device=dev)
    rec_all = torch.empty(world * BATCH, par.record_length(n_lat), device=dev)
    img_host = torch.empty(BATCH, 3, SIZE, SIZE).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # > 126 MB L2

    def step(inp):
        with torch.no_grad():
            out = G([inp["w"], inp["w_dec"]], inp["cam_poses"], inp["focal"], inp["near"],
                    inp["far"], input_is_latent=True, randomize_noise=True, return_xyz=True,
                    return_sdf=True)
            par.pack_records(inp["w"], inp["w_dec"], out["gen_imgs"], target, out=rec_local)
            if world > 1:
                par.gather_records(rec_local, out=rec_all, equal_shards=True)
        return out

    # The measured step is a CUDA-graph replay of G_pred_latents.forward + the record kernel
    # (e3dge_b200.graphed.GraphedCall: same kernels, one launch; the eager pass keeps the host busy for about
    # as long as the GPU and is timed below as `eager`).  The graph reads `static`, views of one device
    # buffer that the e2e step fills with ONE copy from pinned host memory.
    from e3dge_b200.graphed import GraphedCall
    packed_dev = host_packed.to(dev)
    static = {k: packed_dev[o:o + n].view(shape) for k, (o, n, shape) in offs.items()}

    def core():
        with torch.no_grad():
            out = G([static["w"], static["w_dec"]], static["cam_poses"], static["focal"], static["near"],
                    static["far"], input_is_latent=True, randomize_noise=True, return_xyz=True,
                    return_sdf=True)
            par.pack_records(static["w"], static["w_dec"], out["gen_imgs"], target, out=rec_local)
        return out
    gather_in_graph = False

    def core_with_gather():
        out = core()
        par.gather_records(rec_local, out=rec_all, equal_shards=True)
        return out
    try:
        if os.environ.get("E3DGE_BENCH_EAGER"):  # profiling aid: every kernel launched from Python, in order
            raise RuntimeError("E3DGE_BENCH_EAGER")
        gcall = None
        if world > 1 and os.environ.get("E3DGE_BENCH_GATHER_IN_GRAPH"):
            # opt-in experiment: the step's one collective recorded into the same graph.  Works (profiles/
            # r02_scale.txt) but NCCL then waits at communicator teardown for the graph that captured it, so the
            # processes hang at exit while the graph object is alive — off unless asked for.
            try:
                gcall = GraphedCall(core_with_gather)
                gather_in_graph = True
            except Exception:
                torch.cuda.synchronize()
                gcall = None
        if gcall is None:
            gcall = GraphedCall(core)
        launch_mode = ("CUDA-graph replay of G_pred_latents.forward + record kernel"
                       + (" + the NCCL all-gather of the records" if gather_in_graph else "")
                       + " (e3dge_b200.graphed.GraphedCall); `eager` = the same step launched from Python")
    except Exception as exc:  # capture refused (e.g. a profiler that forbids it): measure the eager step
        torch.cuda.synchronize()

        class _Eager:
            launches = 0

            def __call__(self):
                return core()
        gcall = _Eager()
        why = "E3DGE_BENCH_EAGER set" if str(exc) == "E3DGE_BENCH_EAGER" else f"graph capture failed: {type(exc).__name__}"
        launch_mode = f"eager launches from Python ({why})"

    def step_graph():
        out = gcall()
        if world > 1 and not gather_in_graph:
            par.gather_records(rec_local, out=rec_all, equal_shards=True)
        return out

    def step_e2e():
        packed_dev.copy_(host_packed, non_blocking=True)
        out = step_graph()
        img_host.copy_(out["gen_imgs"], non_blocking=True)
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        launches0 = _lib.launch_count
        for a, b in ev:
            flush.zero_()  # L2 flush between timed steps, outside the event pair
            a.record()
            fn()
            b.record()
        sync_all()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), _lib.launch_count - launches0

    with ClockSampler(local) as clk:
        time.sleep(1.0)  # nvidia-smi's start-up holds driver locks: keep it out of the timed region
        for _ in range(5):  # a few untimed passes: GPU and host leave their idle states before the W warm-ups
            step_graph()
        sync_all()
        clk.mark()
        ms_total, _ = timed(step_graph, args.steps, args.warmup)
        ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup))
        ms_eager, launches_eager = timed(lambda: step(resident), args.steps, args.warmup)
        # the same step on the exact-fp32 back ends (FFMA renderer, FFMA implicit-GEMM convs): every
        # contraction in plain fp32 like the reference, whole step and end to end
        exact = None
        if not args.no_exact_fp32 and world == 1:  # (secondary arm: single-GPU runs only)
            set_backends(G, "fp32")
            try:
                gcall32 = GraphedCall(core)

                def step32():
                    out = gcall32()
                    if world > 1:
                        par.gather_records(rec_local, out=rec_all, equal_shards=True)
                    return out

                def step32_e2e():
                    packed_dev.copy_(host_packed, non_blocking=True)
                    out = step32()
                    img_host.copy_(out["gen_imgs"], non_blocking=True)
                    return out
                n32 = max(3, min(args.steps, 10))
                ms32, _ = timed(step32, n32, 3)
                ms32_e2e, _ = timed(step32_e2e, n32, 1)
                f32 = n32 * BATCH * world
                exact = {"what": "same step, renderer E3_RENDER_FP32_CUDA_CORES + decoder E3_CONV_FP32_CUDA_CORES "
                                 "(plain fp32 FFMA contractions), CUDA-graph replay",
                         "dtype": "f32", "steps": n32, "ms_per_step": ms32 / n32, "value": f32 / (ms32 / 1e3),
                         "e2e": f32 / (ms32_e2e / 1e3), "unit": "frames/s"}
            except Exception as exc:  # never lose the headline line to the secondary arm
                exact = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
            finally:
                set_backends(G, "auto")
    clocks = clk.summary()
    launches = args.steps * gcall.launches if gcall.launches else launches_eager
    frames = args.steps * BATCH * world
    value = frames / (ms_total / 1e3)
    e2e_value = frames / (ms_e2e / 1e3)

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic (random latents/cameras, random-init weights)",
            "config": {"workload": WORKLOAD if args.scaling == "weak" else WORKLOAD.replace(
                           "batch=8 per GPU", f"global batch {BATCH * world} split over {world} GPU(s)"),
                       "size": SIZE, "render_res": RES,
                       "n_samples": N_SAMPLES, "batch_per_gpu": BATCH, "global_batch": BATCH * world,
                       "parallelism": f"image-parallel dp{world}, 1 all-gather of latents/metrics per step",
                       "l2": "256 MiB memset between timed steps, outside the per-step CUDA-event pairs",
                       "randomize_noise": True,
                       "launch": launch_mode},
            "e2e": {"value": e2e_value, "unit": "frames/s",
                    "h2d_bytes_per_step": host_packed.numel() * 4,
                    "d2h_bytes_per_step": img_host.numel() * 4},
            "gpu_launches": launches, "clocks": clocks,
            "eager": {"ms_per_step": ms_eager / args.steps, "value": frames / (ms_eager / 1e3),
                      "gpu_launches": launches_eager},
            "exact_fp32": exact}

    if rank == 0:
        line.update(kernel_roofline(G, resident, dev, flush, args))
        line["train_step"] = train_step_timing(G, resident, flush, args)
        if world == 1 and not args.no_size1024:
            try:
                line["size_1024"] = size1024_timing(dev, flush, args)
            except Exception as exc:
                line["size_1024"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        if world == 1 and not args.no_full_frame:
            try:
                line["full_frame"] = full_frame_timing(G, dev, flush, args, world, sync_all)
            except Exception as exc:
                line["full_frame"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        if world == 1 and not args.no_local_branch:
            try:
                line["local_branch"] = local_branch_timing(G, sd, resident, dev, flush, args)
            except Exception as exc:  # secondary figure: never lose the headline line to it
                line["local_branch"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sd, make_inputs(0), steps=2)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if gather_in_graph:  # NCCL waits at teardown for graphs that captured its kernels: drop ours first
            import gc
            gcall = None
            gc.collect()
            torch.cuda.synchronize()
        dist.destroy_process_group()


def kernel_roofline(G, inp, dev, flush, args):
    """Dominant kernel = the fused render kernel (tensor-core variant); timed alone through the
    C ABI with CUDA events on the launching stream (after warm-up, L2 flushed between launches).

    `achieved` uses ALGORITHMIC work only (103.58 GFLOP and 6.86 MB per image);
    the kernel *executes* 3x the hidden-layer MACs (split-bf16 products) to stay fp32-faithful,
    which is reported separately as `executed_tflops`."""
    from e3dge_b200 import _lib
    lib = _lib.load()
    peaks, how = _peaks()
    R = G.renderer
    n = max(3, min(args.steps, 10))

    def time_render(backend):
        R.backend = backend
        with torch.no_grad():
            film = R._film(inp["w"])  # e3_film_fwd stays outside the bracket
            for _ in range(3):
                R._render_raw(inp["w"], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], film=film)
            times = []
            for _ in range(n):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                R._render_raw(inp["w"], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], film=film)
                b.record()
                torch.cuda.synchronize()
                times.append(a.elapsed_time(b))
        return statistics.median(times)

    # torch.empty of the outputs is host-side only (caching allocator): the bracket holds 1 kernel
    ms = time_render("tensor_cores")
    ms_fp32 = time_render("fp32")
    R.backend = "tensor_cores"
    gbs = RENDER_BYTES_PER_IMAGE * BATCH / (ms / 1e3) / 1e9
    tflops = RENDER_FLOP_PER_IMAGE * BATCH / (ms / 1e3) / 1e12
    hidden_flop = 98304 * 8 * 256 * 256 * 2  # the eight 256x256 layers, per image
    executed = (RENDER_FLOP_PER_IMAGE + 2 * hidden_flop) * BATCH / (ms / 1e3) / 1e12
    # measured FP32 FFMA peak of this GPU (register-only probe): the roof of the fp32 variant
    sink = torch.empty(lib.e3_ffma_peak_probe_sink_floats(), device=dev)
    iters = 20000
    for _ in range(2):
        _lib.check(lib.e3_ffma_peak_probe(iters, _lib.ptr(sink), _lib.cur_stream()), "e3_ffma_peak_probe")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(lib.e3_ffma_peak_probe(iters, _lib.ptr(sink), _lib.cur_stream()), "e3_ffma_peak_probe")
    b.record()
    torch.cuda.synchronize()
    ffma_peak = sink.numel() * iters * 16 * 2 / (a.elapsed_time(b) / 1e3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "render_kernel_dram_bytes.json")
    if os.path.isfile(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    tf32_equiv = RENDER_FLOP_PER_IMAGE * BATCH / (ms_fp32 / 1e3) / 1e12
    return {"roofline": {
        "kernel": "siren_render_tc_kernel<0, 2, false, 7> (CTA pairs, tcgen05 cta_group::2)", "bound": "tensor", "achieved": tflops,
        "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tflops / peaks["bf16_tflops"],
        "traffic": traffic, "peak_source": how + " (cuBLAS bf16 burst)", "kernel_ms": ms,
        "executed_tflops": executed,
        "note": "algorithmic FLOPs (fp32 semantics) over measured bf16 peak; the kernel executes 3 bf16 "
                "products per hidden-layer MAC (hi*hi + lo*hi + hi*lo) to hold the 1e-3 fp32 parity bar, so "
                "frac <= ~0.36 by construction; the fused renderer moves ~15 kFLOP per HBM byte, so its HBM "
                "fraction (`hbm`) is tiny by construction (SURVEY.md 8d)",
        "hbm": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"]},
        "fp32_variant": {"kernel": "siren_render_kernel<0> (E3_RENDER_FP32_CUDA_CORES)", "kernel_ms": ms_fp32,
                         "achieved": tf32_equiv, "peak": ffma_peak, "unit": "TFLOP/s",
                         "frac": tf32_equiv / ffma_peak,
                         "peak_source": "measured here (register-only FFMA probe)"}}}


def local_branch_timing(G, sd, inp, dev, flush, args):
    """Secondary figure: one novel-view frame batch WITH the local branch (BASELINE.json configs[2] shapes, what
    every shipped E3DGE script runs, e3dge_full_runner.py:185-317) minus the 2-D image encoders: global pass
    (points, depth) -> two pixel-aligned feature queries -> SFT fusion + positional encoding + texture-modulation
    MLP on the tensor cores -> renderer with the (alpha, beta) modulation -> decoder."""
    from helpers import local_mlp_state_dict
    from e3dge_b200 import local_branch as lb, local_query as lq, model_options, rendering_options
    from e3dge_b200.stylesdf_model import G_pred_latents
    peaks, how = _peaks()
    lsd = local_mlp_state_dict(SEED)
    tex_key = "renderer.network.netLocal.local_feat_to_tex_modulations_linear."
    GL = G_pred_latents(model_options(size=SIZE, renderer_spatial_output_dim=RES),
                        rendering_options(N_samples=N_SAMPLES, enable_local_model=True, local_modulation_layer=True,
                                          L_pred_tex_modulations=True, residual_local_feats_dim=301),
                        full_pipeline=True).eval()
    gsd = {k.replace("renderer.network.", "renderer.network.netGlobal."): v for k, v in sd.items()}
    gsd.update({k: v for k, v in lsd.items() if k.startswith("renderer.")})
    GL.load_state_dict(gsd, strict=True)
    GL = GL.to(dev)
    fuse = lb.Fuse_sft_MLP(257, 256)
    fuse.load_state_dict({k[len("fuse_sft_block."):]: v for k, v in lsd.items() if k.startswith("fuse_sft_block.")})
    fuse = fuse.to(dev).eval()
    tex = GL.renderer.network.netLocal.local_feat_to_tex_modulations_linear
    B = inp["w"].shape[0]
    gen = torch.Generator(device=dev).manual_seed(SEED)
    fmap_ref = torch.randn(B, 128, 128, 256, device=dev, generator=gen)  # channels-last filtered feature maps
    fmap_que = torch.randn(B, 128, 128, 256, device=dev, generator=gen)
    # calibration = uv-space intrinsics @ world-to-camera extrinsics (camera_utils.py:85-151)
    c2w = torch.cat([inp["cam_poses"], torch.tensor([0., 0, 0, 1], device=dev).expand(B, 1, 4)], 1)
    K = torch.diag(torch.tensor([2 * 0.5 / 0.10510423526567646] * 2 + [1., 1.], device=dev))
    calibs = (K @ torch.linalg.inv(c2w)).contiguous()
    rows = B * RES * RES * N_SAMPLES
    n = max(3, min(args.steps, 10))

    def frame():
        with torch.no_grad():
            g = GL.renderer(inp["cam_poses"], inp["focal"], inp["near"], inp["far"], styles=inp["w"])
            pts = g["points"]
            p3 = pts.reshape(B, -1, 3).permute(0, 2, 1)
            q3 = lq.query(p3, calibs, im_feat_nhwc=fmap_ref)
            q2 = lq.query(p3, calibs, im_feat_nhwc=fmap_que)
            f3 = q3["feats"].permute(0, 2, 1).reshape(B, RES, RES, N_SAMPLES, 256)
            f2 = torch.cat([q2["feats"].permute(0, 2, 1).reshape(B, RES, RES, N_SAMPLES, 256),
                            q3["in_img"].reshape(B, RES, RES, N_SAMPLES, 1).float()], -1)
            mod = lb.local_tex_modulation(fuse, tex, f2, f3, pts)
            return GL([inp["w"], inp["w_dec"]], inp["cam_poses"], inp["focal"], inp["near"], inp["far"],
                      input_is_latent=True, randomize_noise=True, local_data_batch={"tex_modulation": mod}), f2, f3, pts

    def med(fn):
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
        return statistics.median(ts)

    _, f2, f3, pts = frame()
    in_img = float(f2[..., 256].mean().item())  # fraction of samples that project inside the reference image
    ms_frame = med(frame)
    with torch.no_grad():
        ms_tail = med(lambda: lb.local_tex_modulation(fuse, tex, f2, f3, pts))
    mac, mac_padded = 989161, 576 * 256 + 832 * 256 + 256 * 512 + 512 * 512 + 320 * 384 + 640 * 512
    tflops = rows * mac * 2 / (ms_tail / 1e3) / 1e12
    return {"what": "novel-view frame with the local branch (global pass + 2 feature queries + SFT/PE/texture MLP "
                    "tail + modulated render + decoder), batch %d, eager launches" % B,
            "ms_per_step": ms_frame, "frames_per_s": B / (ms_frame / 1e3), "in_image_fraction": in_img,
            "mlp_tail": {"kernel": "tc_linear_kernel x6 + prep (e3_local_mlp_fwd)", "ms": ms_tail, "samples": rows,
                         "bound": "tensor", "achieved": tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": tflops / peaks["bf16_tflops"],
                         "executed_tflops": rows * mac_padded * 6 / (ms_tail / 1e3) / 1e12,
                         "note": "989 161 algorithmic MACs per sample; 3 bf16 products per MAC and zero padding "
                                 "(K 513->576, 301->320, N 301->384, block-diagonal scale/shift stage) execute "
                                 "3.65x that; operands cross HBM as bf16 hi/lo between the six stages",
                         "peak_source": how + " (cuBLAS bf16 burst)"}}


def full_frame_timing(G, dev, flush, args, world, sync_all):
    """Secondary figure: the WHOLE inversion frame of BASELINE.json's metric — image -> IR-SE50/FPN encoder (+ mean
    latents) and CoordConv pose net -> cameras -> renderer -> decoder (AERunner.image2image, trainer.py:773-840).
    The two front-end conv nets are cuDNN through PyTorch (bf16 autocast, channels-last), recorded into the same
    CUDA graph as the generator; `e2e` copies the batch of images in from pinned host memory and the result out."""
    import numpy as np
    import synthetic_inputs as P
    from e3dge_b200 import model_options
    from e3dge_b200.frontend import HybridGradualStyleEncoder_V2, InversionPipeline, VolumeRenderDiscriminator
    from e3dge_b200.graphed import GraphedCall
    enc = P.fill_module(HybridGradualStyleEncoder_V2(50, "ir_se", -1).eval(), "encoder.", SEED)
    pose = P.fill_module(VolumeRenderDiscriminator(model_options(renderer_spatial_output_dim=RES)).eval(),
                         "volume_discriminator.", SEED)
    pipe = InversionPipeline(G, enc, pose, amp=True).to(dev).eval()
    g = np.random.Generator(np.random.PCG64(SEED))
    img_in = torch.from_numpy(g.uniform(-1, 1, (BATCH, 3, SIZE, SIZE)).astype(np.float32)).pin_memory()
    static = img_in.to(dev)
    img_out = torch.empty(BATCH, 3, SIZE, SIZE).pin_memory()

    def core():
        with torch.no_grad():
            return pipe(static, randomize_noise=True, return_xyz=True, return_sdf=True)

    def front_only():
        with torch.no_grad():
            thumb = torch.nn.functional.adaptive_avg_pool2d(static, (64, 64))
            return pipe.image2latents(static), pipe.image2camsettings(thumb)
    gcall, gfront = GraphedCall(core), GraphedCall(front_only)

    def e2e():
        static.copy_(img_in, non_blocking=True)
        out = gcall()
        img_out.copy_(out["gen_imgs"], non_blocking=True)

    def med(fn, n):
        for _ in range(3):
            fn()
        sync_all()
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
    n = max(3, min(args.steps, 10))
    ms, ms_e2e, ms_front = med(gcall, n), med(e2e, n), med(gfront, n)
    return {"what": "whole inversion frame: encoder (IR-SE50 + FPN, bf16 channels-last cuDNN) + pose net + cameras + "
                    "renderer + decoder in one CUDA graph, batch %d per GPU" % BATCH,
            "ms_per_step": ms, "value": BATCH * world / (ms / 1e3), "e2e": BATCH * world / (ms_e2e / 1e3),
            "unit": "frames/s", "front_end_ms": ms_front, "h2d_bytes_per_step": img_in.numel() * 4,
            "d2h_bytes_per_step": img_out.numel() * 4}


def size1024_timing(dev, flush, args):
    """Secondary figure: the same generator pass at --size 1024, the output size every shipped E3DGE script runs
    (demo_view_synthesis.sh:35-36): four up-sampling stages, the last two 64 and 32 channels wide at 512^2 / 1024^2
    (stylesdf_model.py:614-624; narrow tensor-core tiles and the x-pair view of the 32 -> 32 conv)."""
    import synthetic_inputs as P
    from helpers import decoder_layout, synthetic_state_dict
    from e3dge_b200 import model_options, rendering_options
    from e3dge_b200.graphed import GraphedCall
    from e3dge_b200.stylesdf_model import G_pred_latents
    size, batch = 1024, 4
    sd = synthetic_state_dict(size, RES, SEED, "sharp")
    G = G_pred_latents(model_options(size=size, renderer_spatial_output_dim=RES), rendering_options(N_samples=N_SAMPLES),
                       full_pipeline=True).eval()
    G.load_state_dict(sd, strict=True)
    G = G.to(dev)
    inp = {k: v.to(dev) for k, v in P.make_inputs(SEED, batch, decoder_layout(size, RES), RES).items()}

    def core():
        with torch.no_grad():
            return G([inp["w"], inp["w_dec"]], inp["cam_poses"], inp["focal"], inp["near"], inp["far"],
                     input_is_latent=True, randomize_noise=True)
    call = GraphedCall(core)
    n = max(3, min(args.steps, 10))
    for _ in range(3):
        call()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        call()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    return {"what": "generator pass at size 1024 (64^2 x 24 render + 4 up-sampling stages), batch %d, CUDA-graph replay" % batch,
            "ms_per_step": ms, "frames_per_s": batch / (ms / 1e3),
            "decoder_tflops_algorithmic": 62.8e9 * 2 * batch / (ms / 1e3) / 1e12}


def train_step_timing(G, inp, flush, args):
    """Secondary figure (not the headline metric): the same batch through the generator with the
    backward kernels — forward writing the stash, then dL/d(w+), dL/d(decoder latent) of an image
    loss (frozen generator, as the E3DGE encoder training uses it: BASELINE.json configs[3])."""
    for p in G.parameters():
        p.requires_grad_(False)
    n = max(3, min(args.steps, 10))

    def one():
        w = inp["w"].clone().requires_grad_(True)
        wd = inp["w_dec"].clone().requires_grad_(True)
        out = G([w, wd], inp["cam_poses"], inp["focal"], inp["near"], inp["far"], input_is_latent=True,
                randomize_noise=True)
        loss = (out["gen_imgs"] ** 2).mean() + (out["gen_thumb_imgs"] ** 2).mean()
        torch.autograd.grad(loss, [w, wd])

    for _ in range(3):
        one()
    times = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        one()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = statistics.median(times)
    return {"what": "generator forward (with stash) + backward to the latents, batch 8, 1 GPU",
            "ms_per_step": ms, "frames_per_s": BATCH / (ms / 1e3)}


def best_cpu_threads(sd, inp):
    """torch-CPU is fastest well below the host's hardware-thread count (profiles/
    r01_cpu_port_thread_sweep.txt: 16 threads 1.22 s/frame, 128 threads 13.8 s/frame on the GPU box),
    so the CPU arm gets the thread count that is best for IT, found on one frame."""
    from oracle import stylesdf_oracle as O
    total = os.cpu_count() or 1
    sl = {k: v[:1] for k, v in inp.items()}
    best, best_t = total, float("inf")
    for nt in sorted({min(total, c) for c in (8, 16, 32, 64)}):
        torch.set_num_threads(nt)
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            with torch.no_grad():
                O.generator_forward(sd, sl["w"], sl["w_dec"], sl["cam_poses"], sl["focal"], sl["near"],
                                    sl["far"], res=RES, n_samples=N_SAMPLES)
            ts.append(time.perf_counter() - t0)
        if ts[-1] < best_t:
            best, best_t = nt, ts[-1]
    return best


def cpu_baseline(sd, inp, steps=2, frames_per_step=BATCH):
    """The oracle port of the reference's PyTorch code on the host cores (bounded sample)."""
    from oracle import stylesdf_oracle as O
    sl = {k: v[:frames_per_step] for k, v in inp.items()}

    def one():
        with torch.no_grad():
            return O.generator_forward(sd, sl["w"], sl["w_dec"], sl["cam_poses"], sl["focal"],
                                       sl["near"], sl["far"], res=RES, n_samples=N_SAMPLES)
    cores = best_cpu_threads(sd, inp)
    torch.set_num_threads(cores)
    one()  # warm-up
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
    return {"value": frames_per_step / statistics.median(ts), "unit": "frames/s", "cores": cores,
            "kind": "port",
            "sample": f"{steps} timed + 1 warm-up generator passes of {frames_per_step} frame(s) of the "
                      f"same workload, fp32, torch CPU {torch.__version__}, {cores} threads (best of "
                      f"8/16/32/64 on this host; {os.cpu_count()} hardware threads present)"}


def run_reference(args):
    """Reference arm: the CPU implementation of the path on the box's host cores (the oracle port of the
    reference's PyTorch code — the Python reference itself cannot travel to the GPU box), on the SAME workload as
    the product arm: one step = the whole batch of 8 frames through renderer + decoder."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from helpers import synthetic_state_dict
    sd = synthetic_state_dict(SIZE, RES, SEED, "sharp")
    inp = make_inputs(0)
    from oracle import stylesdf_oracle as O
    cores = best_cpu_threads(sd, inp)
    torch.set_num_threads(cores)

    def one():
        with torch.no_grad():
            O.generator_forward(sd, inp["w"], inp["w_dec"], inp["cam_poses"], inp["focal"], inp["near"],
                                inp["far"], res=RES, n_samples=N_SAMPLES)
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    value = args.steps * BATCH / dt
    sample = (f"the whole batch of {BATCH} frames per step (same workload as the product arm), fp32, torch CPU "
              f"{torch.__version__}, {cores} threads (best of 8/16/32/64 on this host; {os.cpu_count()} hardware "
              f"threads present)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (torch CPU)", "data": "synthetic (random latents/cameras, random-init weights)",
        "config": {"workload": WORKLOAD, "size": SIZE, "render_res": RES, "n_samples": N_SAMPLES,
                   "batch_per_gpu": BATCH, "global_batch": BATCH},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: 8 frames per GPU (default, the driver's scaling run); strong: --global-batch frames "
                         "in total, split over the ranks (BASELINE.json configs[2]: 32 over 8 GPUs)")
    ap.add_argument("--global-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exact-fp32", action="store_true", help="skip the exact-fp32 back-end timing")
    ap.add_argument("--no-local-branch", action="store_true", help="skip the local-branch frame timing")
    ap.add_argument("--no-size1024", action="store_true", help="skip the size-1024 generator timing")
    ap.add_argument("--no-full-frame", action="store_true", help="skip the encoder -> render -> decode frame timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
