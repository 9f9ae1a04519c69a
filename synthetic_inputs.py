"""Deterministic synthetic parameters and inputs (no checkpoints or datasets exist offline).

Shared by bench.py, the tests and oracle/gen_golden.py; it holds no arithmetic of the path — only seeded
random weights, latents and cameras — and is deliberately NOT under oracle/.

The reference ships no golden vectors and its pretrained weights are not available
offline, so every parity case uses *synthetic* weights.  To avoid
committing tens of MB of state_dict, the weights are a pure function of
``(seed, parameter name, shape)`` through numpy's PCG64, which is bit-reproducible
across machines.  ``oracle/gen_golden.py`` fills the REAL reference modules with these
values and records their outputs; the tests regenerate the same values for the oracle
and for the CUDA path.

Ranges follow the reference's initialisers (volume_renderer.py:53-71,91-114;
stylesdf_model.py:54-64,221-224,305-308) so activations live in the regime the
reference runs in (FiLM gamma ~ 12..49, |sin arg| up to ~1e2).  Parameters the
reference initialises to exactly 0 (noise strengths, biases) get small non-zero values
so that every term of the arithmetic is exercised.
"""
import math
import zlib

import numpy as np
import torch


def _rng(seed, name):
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def _uniform(rng, shape, a):
    return rng.uniform(-a, a, size=shape)


def _normal(rng, shape, std):
    return rng.standard_normal(size=shape) * std


def make_param(seed, name, shape, variant="default"):
    """One parameter tensor (float64 numpy) for reference key `name`."""
    rng = _rng(seed, name)
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    kaiming = math.sqrt(2.0 / (1.0 + 0.2 ** 2))

    if name.endswith("sigmoid_beta"):
        return np.full(shape, 0.1 if variant != "sharp" else 0.02)

    # ---- renderer: FiLM-SIREN ------------------------------------------------
    if ".pts_linears." in name or ".views_linears." in name:
        if ".gamma." in name or ".beta." in name:
            fan_in = 256 if leaf == "bias" else shape[-1]
            if leaf == "weight":
                return _normal(rng, shape, 0.25 * kaiming / math.sqrt(fan_in))
            return _uniform(rng, shape, math.sqrt(1.0 / fan_in))
        if leaf == "weight":
            fan_in = shape[-1]
            if fan_in == 3:
                return _uniform(rng, shape, 1.0 / 3.0)
            return _uniform(rng, shape, math.sqrt(6.0 / fan_in) / 25.0)
        # FiLMSiren.bias ~ U(+-sqrt(1/in)); `in` is not recoverable from the bias
        # shape, and 3 vs 256 only matters for layer 0.
        fan_in = 3 if ".pts_linears.0." in name else 256
        return _uniform(rng, shape, math.sqrt(1.0 / fan_in))
    if ".rgb_linear." in name or ".sigma_linear." in name:
        fan_in = 256
        if leaf == "weight":
            w = _uniform(rng, shape, math.sqrt(6.0 / fan_in) / 25.0)
            if variant == "sharp" and ".sigma_linear." in name:
                w = w * 4.0  # sdf crosses zero with larger swings
            return w
        b = _uniform(rng, shape, math.sqrt(1.0 / fan_in))
        if variant == "sharp" and ".sigma_linear." in name:
            b = b * 0.1
        return b

    # ---- local branch: SFT fusion MLP and texture-modulation MLP (sft.py:84-109, resnetfc.py:10-62) ----
    # (the reference zero-initialises the residual / modulation layers; non-zero values exercise every term)
    if name.startswith("fuse_sft_block.") or ".local_feat_to_tex_modulations_linear." in name:
        if leaf == "weight":
            gain = 0.1 if (".local_feat_to_tex_modulations_linear." in name and
                           (".fc_1." in name or ".shortcut." in name)) else 1.0
            return _normal(rng, shape, gain * math.sqrt(2.0 / shape[-1]))
        return _normal(rng, shape, 0.1)

    # ---- front end: IR-SE50 / FPN encoder and the CoordConv pose net (conv nets with BatchNorm / PReLU) ----
    if name.startswith("encoder.") or name.startswith("volume_discriminator."):
        if leaf == "num_batches_tracked":
            return np.zeros(shape)
        if leaf == "running_mean":
            return _normal(rng, shape, 0.05)
        if leaf == "running_var":
            return 1.0 + np.abs(_normal(rng, shape, 0.1))
        if len(shape) >= 2:  # conv / linear weights: variance-preserving through the ~50-layer trunk
            fan_in = int(np.prod(shape[1:]))
            return _normal(rng, shape, math.sqrt(1.0 / fan_in))
        if leaf == "weight":  # BatchNorm scale / PReLU slope (1-d)
            return (0.25 + _normal(rng, shape, 0.02)) if ".res_layer.2." in name or name.endswith("input_layer.2.weight") \
                else 1.0 + _normal(rng, shape, 0.05)
        return _normal(rng, shape, 0.05)

    # ---- z -> w mapping (3 x MappingLinear) ------------------------------------
    if name.startswith("style."):
        if leaf == "weight":
            return _normal(rng, shape, kaiming / math.sqrt(shape[-1]))
        return _uniform(rng, shape, math.sqrt(1.0 / 256))

    # ---- decoder ------------------------------------------------------------------
    if name.startswith("decoder.style."):
        if leaf == "weight":
            return _normal(rng, shape, 100.0)  # randn / lr_mul, lr_mul = 0.01
        return _normal(rng, shape, 5.0)  # used as bias * lr_mul
    if name.startswith("decoder.noises."):
        return _normal(rng, shape, 1.0)
    if ".modulation." in name:
        if leaf == "weight":
            return _normal(rng, shape, 1.0)
        return 1.0 + _normal(rng, shape, 0.1)
    if name.endswith("conv.weight"):
        return _normal(rng, shape, 1.0)
    if name.endswith("noise.weight"):
        return _normal(rng, shape, 0.1)
    if name.endswith("activate.bias"):
        return _normal(rng, shape, 0.1)
    if leaf == "bias":  # ToRGB.bias (live) and StyledConv.bias (dead parameter)
        return _normal(rng, shape, 0.1)
    raise KeyError(f"no synthetic initialiser for parameter {name!r} {shape}")


def fill_state_dict(state_dict, seed=0, variant="default"):
    """Returns a new {name: float32 tensor} for every learnable key of `state_dict`.

    FIR `kernel` buffers are constants of the architecture and are kept as they are.
    """
    out = {}
    for name, ref in state_dict.items():
        if name.split(".")[-1] == "kernel":
            out[name] = ref.detach().clone()
            continue
        v = make_param(seed, name, ref.shape, variant)
        out[name] = torch.from_numpy(np.ascontiguousarray(v)).to(torch.float32)
    return out


def fill_module(module, prefix, seed=0, variant="default"):
    """Loads make_param's values into every entry of `module.state_dict()` (names prefixed by `prefix`)."""
    sd = {k: torch.from_numpy(np.ascontiguousarray(make_param(seed, prefix + k, v.shape, variant))).to(v.dtype)
          for k, v in module.state_dict().items()}
    module.load_state_dict(sd, strict=True)
    return module


def make_inputs(seed, batch, n_dec_latent, res, fov_deg=6.0, dist_radius=0.12,
                wplus=True, frontal=False):
    """Latents + cameras the way the inversion path feeds the generator.

    Cameras follow camera_utils.py:62-111 (azim ~ N(0,0.3), elev ~ N(0,0.15), dist 1,
    look-at origin, up = +y); focal = 0.5*res/tan(fov) (camera_utils.py:74).
    """
    rng = _rng(seed, "inputs")
    if wplus:
        w = _normal(rng, (batch, 9, 256), 0.35)
    else:
        w = _normal(rng, (batch, 256), 0.35)
    w_dec = _normal(rng, (batch, n_dec_latent, 512), 1.0)
    azim = np.zeros((batch, 1)) if frontal else _normal(rng, (batch, 1), 0.3)
    elev = np.zeros((batch, 1)) if frontal else _normal(rng, (batch, 1), 0.15)
    x = np.cos(elev) * np.sin(azim)
    y = np.sin(elev)
    z = np.cos(elev) * np.cos(azim)
    cam_dir = np.concatenate([x, y, z], 1)
    up = np.tile(np.array([[0.0, 1.0, 0.0]]), (batch, 1))

    def _norm(v):
        return v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-5)

    z_axis = _norm(cam_dir)
    x_axis = _norm(np.cross(up, z_axis))
    y_axis = _norm(np.cross(z_axis, x_axis))
    c2w_R = np.stack([x_axis, y_axis, z_axis], axis=2)  # columns
    poses = np.concatenate([c2w_R, cam_dir[:, :, None]], axis=2)
    focal = np.full((batch, 1, 1), 0.5 * res / math.tan(math.radians(fov_deg)))
    near = np.full((batch, 1, 1), 1.0 - dist_radius)
    far = np.full((batch, 1, 1), 1.0 + dist_radius)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(torch.float32)
    return dict(w=f32(w), w_dec=f32(w_dec), cam_poses=f32(poses), focal=f32(focal),
                near=f32(near), far=f32(far))


VIS_CFG = dict(size=64, res=4, n_samples=12, n_query=5, batch=2, seed=71, variant="sharp")


def visibility_case_inputs(cfg):
    """Query points near the surface band + a reference view, shared by gen_golden and the tests."""
    rng = np.random.Generator(np.random.PCG64(cfg["seed"]))
    B, H, S = cfg["batch"], cfg["res"], cfg["n_query"]
    pts = torch.from_numpy(rng.uniform(-0.09, 0.09, (B, H, H, S, 3)).astype(np.float32))
    inp = make_inputs(cfg["seed"], B, 1, H)
    poses = inp["cam_poses"]
    R, t = poses[:, :, :3], poses[:, :, 3:]
    extr = torch.cat([R.transpose(1, 2), -R.transpose(1, 2) @ t], 2)
    near = inp["near"].reshape(B, 1, 1, 1).expand(B, H, H, 1).contiguous()
    far = inp["far"].reshape(B, 1, 1, 1).expand(B, H, H, 1).contiguous()
    info = dict(global_render_out=dict(near=near, far=far), cam_settings=dict(poses=poses, extrinsics=extr),
                pred_latents=[inp["w"]])
    return pts, info
