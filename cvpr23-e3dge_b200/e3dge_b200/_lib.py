"""ctypes binding of the C-ABI library (include/e3dge_b200.h).

The library is built in-tree by ``cvpr23-e3dge_b200/build.py``.  There is no fallback of
any kind: if the shared object is missing, loading raises, and every op raises on
non-CUDA tensors.
"""
import ctypes
import os
import threading
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint32, c_void_p)

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libe3dge_b200.so")

_fp = c_void_p  # every device pointer crosses the ABI as void*


class SirenWeights(Structure):
    _fields_ = [("pts_w", _fp * 8), ("pts_b", _fp * 8), ("gamma_w", _fp * 9), ("gamma_b", _fp * 9),
                ("beta_w", _fp * 9), ("beta_b", _fp * 9), ("views_w", _fp), ("views_b", _fp),
                ("rgb_w", _fp), ("rgb_b", _fp), ("sigma_w", _fp), ("sigma_b", _fp)]


class LocalMlpWeights(Structure):
    _fields_ = [(n, _fp) for n in ("enc_fc0_w", "enc_fc0_b", "enc_fc1_w", "enc_fc1_b", "enc_shortcut_w",
                                   "scale0_w", "scale0_b", "scale2_w", "scale2_b",
                                   "shift0_w", "shift0_b", "shift2_w", "shift2_b",
                                   "tex_fc0_w", "tex_fc0_b", "tex_fc1_w", "tex_fc1_b", "tex_shortcut_w")]


class RenderParams(Structure):
    _fields_ = [("batch", c_int32), ("height", c_int32), ("width", c_int32), ("res", c_int32),
                ("n_samples", c_int32), ("flags", c_uint32), ("pts_scale", c_float),
                ("mask_depth", c_float)]


class RenderInputs(Structure):
    _fields_ = [(n, _fp) for n in ("cam_poses", "focal", "near", "far", "pix_x", "pix_y", "t_vals",
                                   "z_jitter", "sigmoid_beta", "film", "local_alpha", "local_beta")]


class RenderOutputs(Structure):
    _fields_ = [(n, _fp) for n in ("features", "thumb_rgb", "xyz", "mask", "depth", "sdf", "hit_prob",
                                   "visibility", "dists", "points", "rays_o", "rays_d", "viewdirs",
                                   "raw_rgb", "feats_taps", "bwd_stash")]


class RenderSaved(Structure):
    _fields_ = [(n, _fp) for n in ("stash", "sdf", "hit_prob", "raw_rgb")]


class RenderGrads(Structure):
    _fields_ = [(n, _fp) for n in ("d_features", "d_thumb_rgb", "d_xyz", "d_depth", "d_sdf",
                                   "d_hit_prob")]


class RenderBwdOutputs(Structure):
    _fields_ = [(n, _fp) for n in ("d_film", "d_local_alpha", "d_local_beta", "d_points")]


RENDER_STATIC_VIEWDIRS = 1
RENDER_FORCE_BACKGROUND = 2
RENDER_NO_FORCE_STOP = 4
RENDER_NO_SDF = 8
RENDER_FP32_CUDA_CORES = 16
CONV_AUTO, CONV_FP32_CUDA_CORES, CONV_TENSOR_CORES = 0, 1, 2

_PROTOTYPES = {
    "e3_abi_version": (c_int, []),
    "e3_last_error": (c_char_p, []),
    "e3_siren_packed_bytes": (c_size_t, []),
    "e3_siren_pack": (c_int, [POINTER(SirenWeights), _fp, _fp]),
    "e3_film_fwd": (c_int, [_fp, _fp, c_int, c_int, _fp, _fp]),
    "e3_render_fwd": (c_int, [_fp, POINTER(RenderParams), POINTER(RenderInputs),
                              POINTER(RenderOutputs), _fp]),
    "e3_siren_points_fwd": (c_int, [_fp, _fp, _fp, _fp, c_int, c_int, c_float, _fp, _fp, _fp, c_uint32,
                                    _fp]),
    "e3_siren_points_fwd_ex": (c_int, [_fp, _fp, _fp, _fp, c_int, c_int, c_float, _fp, _fp, _fp, _fp, _fp, _fp,
                                       c_uint32, _fp]),
    "e3_siren_points_fwd_train": (c_int, [_fp, _fp, _fp, _fp, c_int, c_int, c_float, _fp, _fp, _fp, _fp,
                                          _fp]),
    "e3_render_stash_bytes": (c_size_t, [c_int, c_int, c_int]),
    "e3_render_bwd_scratch_bytes": (c_size_t, [c_int]),
    "e3_render_bwd": (c_int, [_fp, POINTER(RenderParams), POINTER(RenderInputs), POINTER(RenderSaved),
                              POINTER(RenderGrads), POINTER(RenderBwdOutputs), _fp, c_size_t, _fp]),
    "e3_siren_points_bwd": (c_int, [_fp, _fp, c_int, c_int, c_float, _fp, c_int, c_int, _fp, _fp, _fp,
                                    _fp, _fp, _fp, c_size_t, _fp]),
    "e3_film_bwd": (c_int, [_fp, _fp, c_int, c_int, _fp, _fp]),
    "e3_fused_bias_act": (c_int, [_fp, _fp, _fp, _fp, c_int64, c_int64, c_int64, c_int, c_int,
                                  c_float, c_float, _fp]),
    "e3_upfirdn2d": (c_int, [_fp, _fp, _fp] + [c_int] * 14 + [_fp]),
    "e3_modconv_weight_sq": (c_int, [_fp, c_int, c_int, c_int, _fp, _fp]),
    "e3_modconv_styles": (c_int, [_fp, c_int64, _fp, _fp, _fp, c_int, c_int, c_int, c_int, _fp, _fp,
                                  _fp]),
    "e3_nchw_to_nhwc": (c_int, [_fp, _fp, c_int, c_int, c_int, c_int, _fp]),
    "e3_nhwc_to_nchw": (c_int, [_fp, _fp, c_int, c_int, c_int, c_int, _fp]),
    "e3_conv_packed_bytes": (c_size_t, [c_int, c_int]),
    "e3_conv_pack_weight": (c_int, [_fp, c_int, c_int, c_int, _fp, _fp]),
    "e3_styled_conv3x3_fwd": (c_int, [_fp, _fp, _fp, _fp, _fp, c_int64, _fp, _fp, _fp, c_int, c_int,
                                      c_int, c_int, c_int, _fp, c_size_t, c_uint32, _fp]),
    "e3_styled_conv3x3_up_fwd": (c_int, [_fp, _fp, _fp, _fp, _fp, c_int64, _fp, _fp, _fp, c_int,
                                         c_int, c_int, c_int, c_int, _fp, c_size_t, c_uint32, _fp]),
    "e3_styled_conv_scratch_bytes": (c_size_t, [c_int] * 6),
    "e3_styled_conv_pair_fusable": (c_int, [c_int] * 5 + [c_uint32]),
    "e3_styled_conv3x3_up_fwd_split": (c_int, [_fp, _fp, _fp, _fp, _fp, c_int64, _fp, _fp, _fp, _fp, _fp,
                                               c_int, c_int, c_int, c_int, c_int, _fp, c_size_t, c_uint32,
                                               _fp]),
    "e3_styled_conv3x3_fwd_presplit": (c_int, [_fp, _fp, _fp, _fp, _fp, c_int64, _fp, _fp, _fp, c_int,
                                               c_int, c_int, c_int, c_int, c_uint32, _fp]),
    "e3_styled_conv_bwd_scratch_bytes": (c_size_t, [c_int] * 6),
    "e3_styled_conv3x3_bwd": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, c_int64, _fp, _fp, _fp, _fp, _fp,
                                      c_int, c_int, c_int, c_int, c_int, c_int, _fp, c_size_t, c_uint32,
                                      _fp]),
    "e3_torgb_bwd_scratch_bytes": (c_size_t, [c_int, c_int]),
    "e3_torgb_bwd": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, _fp, c_size_t, _fp]),
    "e3_modconv_styles_bwd": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, _fp, _fp]),
    "e3_torgb_fwd": (c_int, [_fp, _fp, _fp, _fp, _fp, c_int, _fp, c_int, c_int, c_int, c_int, _fp]),
    "e3_local_feature_query": (c_int, [_fp, _fp, c_int64, c_int64, c_int64, _fp, c_int, c_int, c_int, c_int,
                                       c_int, c_int, _fp, _fp, _fp, _fp, _fp]),
    "e3_local_feature_query_bwd": (c_int, [_fp, _fp, c_int64, c_int64, c_int64, _fp, c_int, c_int, c_int, c_int, c_int,
                                           c_int, _fp, _fp]),
    "e3_local_mlp_packed_bytes": (c_size_t, []),
    "e3_local_mlp_pack": (c_int, [POINTER(LocalMlpWeights), _fp, _fp]),
    "e3_local_mlp_workspace_bytes": (c_size_t, [c_int64]),
    "e3_local_mlp_fwd": (c_int, [_fp, _fp, _fp, _fp, _fp, c_int64, _fp, _fp, _fp, _fp, c_size_t, _fp]),
    "e3_tc_linear_packed_bytes": (c_size_t, [c_int, c_int]),
    "e3_tc_linear_pack": (c_int, [_fp, c_int, c_int, _fp, _fp]),
    "e3_tc_linear_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "e3_tc_linear_fwd": (c_int, [_fp, c_int, c_int, _fp, c_int64, _fp, _fp, _fp, c_size_t, _fp]),
    "e3_pack_inversion_record": (c_int, [_fp, _fp, c_int, _fp, _fp, c_int, c_int64, _fp, _fp]),
    "e3_ffma_peak_probe": (c_int, [c_int, _fp, _fp]),
    "e3_ffma_peak_probe_sink_floats": (c_size_t, []),
}

_lib = None
_tls = threading.local()  # device of the tensors handed to ptr() since the last C call


class _Guarded:
    """The ctypes handle behind a device guard.  The C side launches on the calling thread's CURRENT
    device (SURVEY.md §8b "Threading"); tensors living on another device would otherwise be launched on the
    wrong GPU.  `ptr()` notes the device of every tensor argument, `cur_stream()` returns that device's
    current stream, and the call itself runs with that device current."""

    def __init__(self, cdll):
        object.__setattr__(self, "_cdll", cdll)
        object.__setattr__(self, "_fns", {})

    def __getattr__(self, name):
        fns = object.__getattribute__(self, "_fns")
        fn = fns.get(name)
        if fn is None:
            raw = getattr(object.__getattribute__(self, "_cdll"), name)

            def fn(*args, _raw=raw):
                dev = getattr(_tls, "dev", None)
                _tls.dev = None
                if dev is None or dev == torch.cuda.current_device():
                    return _raw(*args)
                with torch.cuda.device(dev):
                    return _raw(*args)
            fns[name] = fn
        return fn


def load():
    """Loads (once) and returns the guarded ctypes handle.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"e3dge_b200: native library not found at {LIB_PATH}. Build it with "
            "`python cvpr23-e3dge_b200/build.py` (or __graft_entry__.build()); there is no "
            "non-CUDA fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = _Guarded(lib)
    return _lib


def exported_symbols():
    return sorted(_PROTOTYPES)


# kernels launched by one successful call of each entry point (for bench.py's gpu_launches)
KERNELS_PER_CALL = {"e3_siren_pack": 1, "e3_film_fwd": 1, "e3_render_fwd": 1,
                    "e3_siren_points_fwd": 1, "e3_siren_points_fwd_ex": 1, "e3_siren_points_fwd_train": 1, "e3_render_bwd": 3,
                    "e3_siren_points_bwd": 3, "e3_film_bwd": 1, "e3_fused_bias_act": 1, "e3_upfirdn2d": 1,
                    "e3_modconv_weight_sq": 1, "e3_modconv_styles": 2, "e3_nchw_to_nhwc": 1,
                    "e3_nhwc_to_nchw": 1, "e3_conv_pack_weight": 1, "e3_styled_conv3x3_fwd": 2,
                    "e3_styled_conv3x3_up_fwd": 3, "e3_styled_conv3x3_up_fwd_split": 3,
                    "e3_styled_conv3x3_fwd_presplit": 1, "e3_torgb_fwd": 1, "e3_styled_conv3x3_bwd": 7,
                    "e3_torgb_bwd": 2, "e3_modconv_styles_bwd": 1,
    "e3_pack_inversion_record": 1, "e3_ffma_peak_probe": 1, "e3_local_feature_query": 1,
                    "e3_local_feature_query_bwd": 1, "e3_local_mlp_pack": 19, "e3_local_mlp_fwd": 7,
                    "e3_tc_linear_pack": 1, "e3_tc_linear_fwd": 2}
launch_count = 0

# Packed weight images (SIREN stream, conv operands, sum_k W^2) are cached per module and keyed on
# (data_ptr, tensor._version, pack_epoch).  In-place updates through `.data` (the reference's EMA
# `accumulate()`, training_utils.py:40-45; Ranger's `p.data.copy_`) do not bump `_version`, so whoever
# updates weights that way calls `invalidate_packed()` afterwards; `load_state_dict`, `.to()/.cuda()` and
# `train()/eval()` switches do it on their own.
pack_epoch = 0


def invalidate_packed():
    """Forces every packed weight image of this process to be rebuilt at its next use."""
    global pack_epoch
    pack_epoch += 1


def check(rc, what="", launches=None):
    global launch_count
    launch_count += KERNELS_PER_CALL.get(what, 0) if launches is None else launches
    if rc != 0:
        msg = load().e3_last_error()
        raise RuntimeError(f"e3dge_b200 {what} failed (status {rc}): "
                           f"{msg.decode() if msg else 'no message'}")


def _note_device(t):
    idx = t.device.index
    dev = getattr(_tls, "dev", None)
    if dev is None:
        _tls.dev = idx
    elif dev != idx:
        _tls.dev = None
        raise RuntimeError(f"e3dge_b200: tensors of one call live on different devices (cuda:{dev}, cuda:{idx})")


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("e3dge_b200: CUDA tensor required (this framework has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"e3dge_b200: float32 tensor required, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError("e3dge_b200: contiguous tensor required")
    _note_device(t)
    return c_void_p(t.data_ptr())


def vptr(t):
    """Device pointer of a contiguous CUDA tensor of any dtype (bf16 operand images, strided latents)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("e3dge_b200: CUDA tensor required (this framework has no CPU path)")
    _note_device(t)
    return c_void_p(t.data_ptr())


def cur_stream():
    """Current stream of the device the call's tensors live on (the calling thread's current device when
    no tensor has been seen yet)."""
    dev = getattr(_tls, "dev", None)
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def as_f32c(t):
    """fp32 contiguous view/copy (the reference makes inputs .contiguous() inside the op)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_warned_frozen = set()


def warn_if_trainable(module, what):
    """The backward kernels return gradients for latents / features / local modulation only (the E3DGE
    path trains encoders against a frozen generator, trainer.py:881-900): a generator parameter that still
    has requires_grad=True would silently get no gradient, so say it once."""
    if what in _warned_frozen:
        return
    names = [n for n, p in module.named_parameters() if p.requires_grad]
    if names:
        import warnings
        _warned_frozen.add(what)
        warnings.warn(f"e3dge_b200: {what} has {len(names)} parameter(s) with requires_grad=True (e.g. "
                      f"{names[0]}); this path provides no gradients for the generator's own weights — "
                      "freeze them (requires_grad_(False)) or expect None grads", RuntimeWarning, stacklevel=3)
