"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np
import torch

import synthetic_inputs as P

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    arrays = {k: z[k] for k in z.files}
    cfg = None
    if "config" in arrays:
        cfg = json.loads(bytes(arrays.pop("config")).decode())
    return arrays, cfg


def rel_linf(a, b):
    """max |a-b| / max |b|  — the 'rel L-inf' of SURVEY.md §8c(ii)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def decoder_layout(size, res):
    """(n_latent, channels per resolution) — stylesdf_model.py:614-624, 682."""
    import math
    n_up = int(math.log2(size)) - int(math.log2(res))
    return 2 * n_up + 2


def generator_state_dict_spec(size, res, local=False):
    """{key: shape} of the generator state_dict for (size, res) — the contract of
    SURVEY.md §8b(1); built without the reference so it works on the GPU box."""
    import math
    ch = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64,
          1024: 32}
    spec = {}
    for i in range(3):
        spec[f"style.{i}.weight"] = (256, 256)
        spec[f"style.{i}.bias"] = (256,)
    spec["renderer.sigmoid_beta"] = (1,)
    net = "renderer.network." + ("netGlobal." if local else "")

    def film(prefix, cin):
        spec[prefix + "weight"] = (256, cin)
        spec[prefix + "bias"] = (256,)
        for gb in ("gamma", "beta"):
            spec[f"{prefix}{gb}.weight"] = (256, 256)
            spec[f"{prefix}{gb}.bias"] = (256,)

    for i in range(8):
        film(f"{net}pts_linears.{i}.", 3 if i == 0 else 256)
    film(net + "views_linears.", 259)
    spec[net + "rgb_linear.weight"] = (3, 256)
    spec[net + "rgb_linear.bias"] = (3,)
    spec[net + "sigma_linear.weight"] = (1, 256)
    spec[net + "sigma_linear.bias"] = (1,)
    spec["decoder.style.1.weight"] = (512, 256)
    spec["decoder.style.1.bias"] = (512,)
    for i in range(2, 6):
        spec[f"decoder.style.{i}.weight"] = (512, 512)
        spec[f"decoder.style.{i}.bias"] = (512,)

    def styled(prefix, cin, cout, up):
        spec[prefix + "bias"] = (1, cout, 1, 1)
        spec[prefix + "conv.weight"] = (1, cout, cin, 3, 3)
        if up:
            spec[prefix + "conv.blur.kernel"] = (4, 4)
        spec[prefix + "conv.modulation.weight"] = (cin, 512)
        spec[prefix + "conv.modulation.bias"] = (cin,)
        spec[prefix + "noise.weight"] = (1,)
        spec[prefix + "activate.bias"] = (cout,)

    def torgb(prefix, cin, up):
        spec[prefix + "bias"] = (1, 3, 1, 1)
        if up:
            spec[prefix + "upsample.kernel"] = (4, 4)
        spec[prefix + "conv.weight"] = (1, 3, cin, 1, 1)
        spec[prefix + "conv.modulation.weight"] = (cin, 512)
        spec[prefix + "conv.modulation.bias"] = (cin,)

    lr, ls = int(math.log2(res)), int(math.log2(size))
    styled("decoder.conv1.", 256, ch[res], False)
    torgb("decoder.to_rgb1.", ch[res], False)
    cin = ch[res]
    k = 0
    for i in range(lr + 1, ls + 1):
        cout = ch[2 ** i]
        styled(f"decoder.convs.{k}.", cin, cout, True)
        styled(f"decoder.convs.{k + 1}.", cout, cout, False)
        torgb(f"decoder.to_rgbs.{k // 2}.", cout, True)
        cin = cout
        k += 2
    for layer in range((ls - lr) * 2 + 1):
        r = (layer + 2 * lr + 1) // 2
        spec[f"decoder.noises.noise_{layer}"] = (1, 1, 2 ** r, 2 ** r)
    return spec


def synthetic_state_dict(size, res, seed, variant="default", local=False):
    """The same deterministic weights oracle/gen_golden.py loaded into the reference."""
    spec = generator_state_dict_spec(size, res, local)
    sd = {}
    for name, shape in spec.items():
        if name.endswith(".kernel"):
            k = torch.tensor([1., 3., 3., 1.])
            k = k[None, :] * k[:, None]
            sd[name] = k / k.sum() * 4
            continue
        sd[name] = torch.from_numpy(
            np.ascontiguousarray(P.make_param(seed, name, shape, variant))).float()
    return sd


LOCAL_MLP_SPEC = {
    "fuse_sft_block.encode_enc.fc_0.weight": (256, 513), "fuse_sft_block.encode_enc.fc_0.bias": (256,),
    "fuse_sft_block.encode_enc.fc_1.weight": (256, 256), "fuse_sft_block.encode_enc.fc_1.bias": (256,),
    "fuse_sft_block.encode_enc.shortcut.weight": (256, 513),
    "fuse_sft_block.scale.0.weight": (256, 256), "fuse_sft_block.scale.0.bias": (256,),
    "fuse_sft_block.scale.2.weight": (256, 256), "fuse_sft_block.scale.2.bias": (256,),
    "fuse_sft_block.shift.0.weight": (256, 256), "fuse_sft_block.shift.0.bias": (256,),
    "fuse_sft_block.shift.2.weight": (256, 256), "fuse_sft_block.shift.2.bias": (256,),
    "renderer.network.netLocal.local_feat_to_tex_modulations_linear.fc_0.weight": (301, 301),
    "renderer.network.netLocal.local_feat_to_tex_modulations_linear.fc_0.bias": (301,),
    "renderer.network.netLocal.local_feat_to_tex_modulations_linear.fc_1.weight": (512, 301),
    "renderer.network.netLocal.local_feat_to_tex_modulations_linear.fc_1.bias": (512,),
    "renderer.network.netLocal.local_feat_to_tex_modulations_linear.shortcut.weight": (512, 301),
}


def local_mlp_state_dict(seed, variant="default"):
    """Synthetic weights of the local branch's per-sample MLPs (Fuse_sft_MLP of the runner, the texture-
    modulation ResnetBlockFC of netLocal) under the names oracle/gen_golden_local_mlp.py filled them by."""
    return {k: torch.from_numpy(np.ascontiguousarray(P.make_param(seed, k, shp, variant))).float()
            for k, shp in LOCAL_MLP_SPEC.items()}


def synthetic_local_feats(seed, shape_prefix):
    """feature_2dAlign | visibility mask [...,257] and feature_3dprojection [...,256] (e3dge_full_runner.py:229-288)."""
    rng = np.random.Generator(np.random.PCG64([seed, 77]))
    f2 = rng.standard_normal(tuple(shape_prefix) + (256,)).astype(np.float32)
    vis = (rng.uniform(size=tuple(shape_prefix) + (1,)) < 0.7).astype(np.float32)
    f3 = rng.standard_normal(tuple(shape_prefix) + (256,)).astype(np.float32)
    return torch.from_numpy(np.concatenate([f2, vis], -1)), torch.from_numpy(f3)
