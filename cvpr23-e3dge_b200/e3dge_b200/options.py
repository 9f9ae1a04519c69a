"""Option groups the generator path reads (`opt.rendering`, `opt.model`, `opt.camera`).

The reference builds these with configargparse + Munch (project/utils/options.py:546-933,
1499-1534); any attribute-access mapping works.  Defaults below are the reference's, with
the wiring both of its entry points force (base_setup.py:53-56): perturb 0, static view
directions, forced background.
"""


class Opt(dict):
    """dict with attribute access (AttributeError on a miss), like a Munch."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def rendering_options(**over):
    o = Opt(N_samples=24, depth=8, width=256, perturb=0., no_offset_sampling=False,
            raw_noise_std=0., return_xyz=True, return_sdf=True, static_viewdirs=True,
            no_z_normalize=False, spatial_super_sampling_factor=1, force_background=True,
            no_sdf=False, add_fg_mask=False, enable_local_model=False, return_feats=False,
            return_feats_layers=[1, 3, 5, 7], local_modulation_layer=False,
            local_modulation_layer_in_backbone=False, use_integrated_surface_normal=False,
            L_pred_tex_modulations=False, L_pred_geo_modulations=False,
            sample_near_surface=False, sample_uniform_grid=False, uniform_grid_sampling_num=2048,
            surface_sampling_stdv=0.01,
            camera=Opt(dist_radius=0.12, fov=6, azim=0.3, elev=0.15, uniform=False))
    o.update(over)
    return o


def model_options(**over):
    o = Opt(size=256, style_dim=256, channel_multiplier=2, lr_mapping=0.01,
            renderer_spatial_output_dim=64, project_noise=False, freeze_renderer=False,
            is_test=True)
    o.update(over)
    return o
