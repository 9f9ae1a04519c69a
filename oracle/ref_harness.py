"""TEST INFRASTRUCTURE ONLY — read-only import harness for the *real* reference.

Imports ``/root/reference/project/{utils/volume_renderer,models/stylesdf_model}.py``
on CPU without executing the reference's package ``__init__`` files (which fail on
Python >= 3.11) and without its heavy third-party deps
(pytorch3d, skimage, trimesh, mmcv ...), none of which carry hot-path arithmetic.

Nothing is copied from the reference: the modules are loaded *in place* by path.
This only works inside the build container (``/root/reference`` does not exist on
the GPU box), so it is used by exactly two things:

* ``oracle/gen_golden.py``  — writes the fixtures under ``tests/golden/``
* ``tests/test_oracle_vs_reference.py`` — pins ``oracle/stylesdf_oracle.py``

Product code never imports this file.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("E3DGE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(
        os.path.join(REFERENCE_ROOT, "project/utils/volume_renderer.py"))


class _Anything:
    """Permissive placeholder: any attribute / call yields another placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):  # keep inspect / copy / pickle sane
            raise AttributeError(name)
        return _Anything()


def _stub_module(name):
    m = types.ModuleType(name)
    m.__path__ = []

    def _getattr(key, _name=name):
        if key.startswith("__"):
            raise AttributeError(key)
        return _Anything

    m.__getattr__ = _getattr
    # `from lib.x import *` needs __all__ or a plain dict walk; give it nothing.
    m.__all__ = []
    sys.modules[name] = m
    return m


def _shell_package(name, path, **attrs):
    """A package whose sub-modules load from `path` but whose __init__ never runs."""
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_LOADED = {}


def load_reference():
    """Returns a namespace with the reference's hot-path classes (CPU only)."""
    if _LOADED:
        return types.SimpleNamespace(**_LOADED)
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch  # noqa: F401  (must be imported before the stubs go in)
    import torch.utils.cpp_extension as cpp_ext

    for name in ("pytorch3d", "pytorch3d.renderer", "pytorch3d.structures",
                 "pytorch3d.transforms", "skimage", "skimage.measure",
                 "trimesh", "ipdb", "lib", "lib.mesh_util", "lib.sample_util",
                 "lib.train_util", "lib.data", "lib.model"):
        if name not in sys.modules:
            _stub_module(name)

    proj = os.path.join(REFERENCE_ROOT, "project")
    if "project" in sys.modules and not getattr(
            sys.modules["project"], "_e3dge_ref_shell", False):
        raise RuntimeError(
            "a different `project` package is already imported in this process; "
            "run the reference harness in its own process")
    _shell_package("project", proj, _e3dge_ref_shell=True)
    _shell_package("project.models", os.path.join(proj, "models"))
    _shell_package("project.utils", os.path.join(proj, "utils"),
                   align_volume=_Anything(), add_textures=_Anything(),
                   create_cameras=_Anything(),
                   create_depth_mesh_renderer=_Anything())

    # The op wrappers JIT-build CUDA at import; on CPU tensors they take their
    # pure-PyTorch branches, so the build is replaced by a no-op.
    real_load = cpp_ext.load
    cpp_ext.load = lambda *a, **k: _Anything()
    try:
        from project.utils import volume_renderer as vr
        from project.models import stylesdf_model as sm
        import project.models.op  # noqa: F401
        fused_act = sys.modules["project.models.op.fused_act"]
        # (the package re-exports a *function* named upfirdn2d over the sub-module)
        upfirdn2d_mod = sys.modules["project.models.op.upfirdn2d"]
    finally:
        cpp_ext.load = real_load

    _LOADED.update(volume_renderer=vr, stylesdf_model=sm, fused_act=fused_act,
                   upfirdn2d=upfirdn2d_mod)
    return types.SimpleNamespace(**_LOADED)


class Opt(dict):
    """dict with attribute access; AttributeError on a miss (reference reads both ways)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def rendering_opt(**over):
    """`opt.rendering` as the reference's entry points wire it
    (base_setup.py:53-56, options.py:652-933 defaults)."""
    o = Opt(
        N_samples=24, depth=8, width=256, perturb=0., no_offset_sampling=False,
        raw_noise_std=0., return_xyz=True, return_sdf=True,
        static_viewdirs=True, no_z_normalize=False,
        spatial_super_sampling_factor=1, force_background=True, no_sdf=False,
        add_fg_mask=False, enable_local_model=False, return_feats=False,
        return_feats_layers=[1, 3, 5, 7], local_modulation_layer=False,
        local_modulation_layer_in_backbone=False,
        use_integrated_surface_normal=False, sample_near_surface=False,
        sample_uniform_grid=False, uniform_grid_sampling_num=2048,
        surface_sampling_stdv=0.01,
        camera=Opt(dist_radius=0.12, fov=6, azim=0.3, elev=0.15, uniform=False),
    )
    o.update(over)
    return o


def model_opt(**over):
    o = Opt(size=256, style_dim=256, channel_multiplier=2, lr_mapping=0.01,
            renderer_spatial_output_dim=64, project_noise=False,
            freeze_renderer=False, is_test=True)
    o.update(over)
    return o
