/*
 * e3dge_b200 — C ABI of the B200-native StyleSDF generator hot path for E3DGE.
 *
 * Drop-in boundary.  The reference has no C ABI of its own: its two
 * pybind shims (project/models/op/fused_bias_act.cpp:11-20, upfirdn2d.cpp:12-23) and the
 * Python surfaces of project/utils/volume_renderer.py / project/models/stylesdf_model.py
 * are what a maintainer binds to.  Every entry point below names the reference interface
 * it replaces.
 *
 * Conventions (derived from the reference's shims, SURVEY.md §8b "Conventions"):
 *   - all pointers are DEVICE pointers to caller-owned, contiguous buffers unless the
 *     name ends in `_host`; the library never allocates or frees caller-visible memory;
 *     scratch is passed in by the caller (sizes from the *_bytes() queries);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), no host sync;
 *   - return value: 0 ok; E3_ERR_* (<0) bad argument; >0 a cudaError_t passthrough;
 *     e3_last_error() gives a thread-local human-readable message;
 *   - re-entrant, no global mutable state; optional outputs may be NULL (= not wanted);
 *   - fp32 in / fp32 out (the reference never runs AMP, SURVEY.md §8b "Dtypes").
 */
#ifndef E3DGE_B200_H_
#define E3DGE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E3_OK 0
#define E3_ERR_BAD_ARG (-1)
#define E3_ERR_UNSUPPORTED (-2)
#define E3_ERR_SCRATCH (-3)

#define E3_SIREN_WIDTH 256 /* rendering.width, options.py (FiLM-SIREN hidden size) */
#define E3_SIREN_DEPTH 8   /* rendering.depth */
#define E3_STYLE_DIM 256   /* model.style_dim */

/* 4 = adds the local branch's per-sample MLP tail (e3_local_mlp_*); nothing of ABI 3 changed.
 * 3 = adds e3_styled_conv_pair_fusable / e3_styled_conv3x3_up_fwd_split / e3_styled_conv3x3_fwd_presplit and
 * e3_local_feature_query; the bf16 part of the up-conv weight image (e3_conv_pack_weight layout 1) is ordered
 * by output parity phase.  Images packed by an ABI-2 library must be re-packed. */
int e3_abi_version(void);
const char* e3_last_error(void);

/* ------------------------------------------------------------------------------------
 * FiLM-SIREN weights (SirenGenerator, volume_renderer.py:136-166) in reference layout.
 * state_dict names: renderer.network[.netGlobal].{pts_linears.i,views_linears}.{weight,
 * bias,gamma.weight,gamma.bias,beta.weight,beta.bias}, rgb_linear.*, sigma_linear.*
 * ---------------------------------------------------------------------------------- */
typedef struct e3_siren_weights {
  const float* pts_w[8];   /* [256,3] for i=0, [256,256] otherwise (out,in) row-major */
  const float* pts_b[8];   /* [256] */
  const float* gamma_w[9]; /* [256,256]; index 8 = views_linears */
  const float* gamma_b[9]; /* [256] */
  const float* beta_w[9];  /* [256,256] */
  const float* beta_b[9];  /* [256] */
  const float* views_w;    /* [256,259] */
  const float* views_b;    /* [256] */
  const float* rgb_w;      /* [3,256] */
  const float* rgb_b;      /* [3] */
  const float* sigma_w;    /* [1,256] */
  const float* sigma_b;    /* [1] */
} e3_siren_weights;

/* Bytes of the library-private packed weight image (TMA-friendly k-major slabs). */
size_t e3_siren_packed_bytes(void);
/* Repack once per weight update (frozen generator => once).  `packed` must be 128-byte
 * aligned and e3_siren_packed_bytes() large. */
int e3_siren_pack(const e3_siren_weights* w, void* packed, void* stream);

/* FiLM frequencies/phases for a batch:  gamma = 15*(G w + g) + 30, beta = 0.25*(H w + h)
 * (FiLMSiren.gamma/.beta, volume_renderer.py:107-114,119-120).
 * styles [batch, styles_per_image, 256]; styles_per_image is 9 (w+: layer i uses style i,
 * the view layer uses the last one, volume_renderer.py:176-178,226-227) or 1 (w).
 * film [batch, 9, 3, 256]: per layer the rows gamma, beta, and beta' = gamma*bias + beta
 * (the layer bias folded in, read by the tensor-core renderer). */
int e3_film_fwd(const void* packed, const float* styles, int batch, int styles_per_image,
                float* film, void* stream);

/* flags of e3_render_params */
#define E3_RENDER_STATIC_VIEWDIRS 1u  /* rendering.static_viewdirs (volume_renderer.py:789-792) */
#define E3_RENDER_FORCE_BACKGROUND 2u /* rendering.force_background (:884-886) */
#define E3_RENDER_NO_FORCE_STOP 4u    /* volume_integration(no_force_stop=True) (:830-836) */
#define E3_RENDER_NO_SDF 8u           /* rendering.no_sdf: softplus density (:862-867) */
/* arithmetic of the eight 256x256 hidden-layer contractions (not a reference option):
 * default = tcgen05 tensor cores with split-bf16 operands (h_hi*W_hi + h_lo*W_hi + h_hi*W_lo,
 * fp32 accumulation in TMEM; ~1e-4 of the fp32 reference end to end); this flag selects the
 * exact-fp32 FFMA kernel instead (~3e-5, the fp32 noise floor), at ~1/6 of the speed. */
#define E3_RENDER_FP32_CUDA_CORES 16u

typedef struct e3_render_params {
  int32_t batch;
  int32_t height;    /* rays per column  = out_im_res * spatial_ss */
  int32_t width;     /* rays per row */
  int32_t res;       /* out_im_res: principal point = res/2 (volume_renderer.py:773-774) */
  int32_t n_samples; /* rendering.N_samples, <= 128 (<= 96 with E3_RENDER_FP32_CUDA_CORES) */
  uint32_t flags;
  float pts_scale;   /* 1/dist_radius, UniformBoxWarp (:23-30,720) */
  float mask_depth;  /* 1.08 (:910) */
} e3_render_params;

/* Inputs of one render call (VolumeFeatureRenderer.forward, volume_renderer.py:1865). */
typedef struct e3_render_inputs {
  const float* cam_poses;    /* [B,3,4] camera-to-world */
  const float* focal;        /* [B] */
  const float* near;         /* [B] */
  const float* far;          /* [B] */
  const float* pix_x;        /* [width]  pixel centres, the `i` buffer (:666-674) */
  const float* pix_y;        /* [height] pixel centres, the `j` buffer */
  const float* t_vals;       /* [n_samples] the `t_vals` buffer (:690-698) */
  const float* z_jitter;     /* NULL, or [B,H,W,S] replacement z_vals (perturb > 0, :1213-1228) */
  const float* sigmoid_beta; /* [1] learned (:662-663) */
  const float* film;         /* [B,9,3,256] from e3_film_fwd */
  const float* local_alpha;  /* NULL or [B,H,W,S,256]: (alpha+1)*h + beta before the view */
  const float* local_beta;   /*                        layer (:217-220) */
} e3_render_inputs;

/* Outputs: the dict contract of render_rays / forward (volume_renderer.py:1270-1287,
 * 1957-1968).  Any pointer may be NULL. */
typedef struct e3_render_outputs {
  float* features;   /* [B,256,H,W] */
  float* thumb_rgb;  /* [B,3,H,W]   'gen_thumb_imgs' */
  float* xyz;        /* [B,3,H,W] */
  float* mask;       /* [B,1,H,W,1] */
  float* depth;      /* [B,H,W,1,1] */
  float* sdf;        /* [B,H,W,S,1] */
  float* hit_prob;   /* [B,H,W,S,1] 'hit_prob' = weights */
  float* visibility; /* [B,H,W,S,1] transmittance T_s */
  float* dists;      /* [B,H,W,S] */
  float* points;     /* [B,H,W,S,3] world-space samples */
  float* rays_o;     /* [B,H,W,3] */
  float* rays_d;     /* [B,H,W,3] */
  float* viewdirs;   /* [B,H,W,3] normalised */
  float* raw_rgb;    /* [B,H,W,S,3] pre-sigmoid radiance (run_network(...)[..., :3]) */
  float* feats_taps; /* NULL or [4,B,H,W,S,256]: hidden states after layers 1,3,5,7
                        (rendering.return_feats, volume_renderer.py:172-193) */
  float* bwd_stash;  /* NULL (inference), or e3_render_stash_bytes() bytes: what e3_render_bwd needs
                        again (the pre-sin phase of every FiLM layer and sample; library-private
                        layout).  Tensor-core renderer only. */
} e3_render_outputs;

/* Fused rays -> samples -> FiLM-SIREN x9 -> SDF->sigma -> alpha composite.
 * Replaces VolumeFeatureRenderer.render/render_rays/run_network/volume_integration
 * (volume_renderer.py:1666-1701, 1183-1298, 1052-1128, 809-943) and
 * SirenGenerator.forward (:240-264) for the inference wiring. */
int e3_render_fwd(const void* packed, const e3_render_params* p, const e3_render_inputs* in,
                  const e3_render_outputs* out, void* stream);

/* FiLM-SIREN evaluated at arbitrary world-space points (run_network on explicit samples:
 * volume_renderer.py:955-957, 996-998, 1935-1943).
 * points [B,N,3]; viewdirs NULL (= zeros, geometry queries) or [B,N,3];
 * sdf [B,N]; raw_rgb NULL or [B,N,3]; feat NULL or [B,N,256].  When raw_rgb and feat are
 * both NULL the view layer is skipped (return_sdf_only, :1125-1126).
 * flags: 0 or E3_RENDER_FP32_CUDA_CORES. */
int e3_siren_points_fwd(const void* packed, const float* film, const float* points,
                        const float* viewdirs, int batch, int n_points, float pts_scale,
                        float* sdf, float* raw_rgb, float* feat, uint32_t flags, void* stream);
/* e3_siren_points_fwd with the two extras SirenLocalGlobal needs (volume_renderer.py:313-369, 515-517):
 * local_alpha / local_beta [B,N,256] (both or neither): the local branch's texture modulation
 * h8' = (alpha + 1) h8 + beta, applied after the sdf head and before the view layer;
 * h8 [B,N,256] (optional output): the backbone features `forward_generator` returns.
 * Tensor-core kernel only (E3_ERR_UNSUPPORTED with E3_RENDER_FP32_CUDA_CORES). */
int e3_siren_points_fwd_ex(const void* packed, const float* film, const float* points,
                           const float* viewdirs, int batch, int n_points, float pts_scale,
                           const float* local_alpha, const float* local_beta, float* sdf,
                           float* raw_rgb, float* feat, float* h8, uint32_t flags, void* stream);

/* Same, additionally writing the backward stash (e3_render_stash_bytes(1, n_points, batch) bytes)
 * for e3_siren_points_bwd.  Tensor-core arithmetic only. */
int e3_siren_points_fwd_train(const void* packed, const float* film, const float* points,
                              const float* viewdirs, int batch, int n_points, float pts_scale,
                              float* sdf, float* raw_rgb, float* feat, float* bwd_stash,
                              void* stream);

/* ------------------------------------------------------------------------------------
 * Backward of the renderer — what autograd gives the reference through
 * VolumeFeatureRenderer.forward when the E3DGE runners train encoders against the frozen
 * generator (trainer.py:881-900, generator frozen at :1569; e3dge_full_runner.py:219-306): gradients with respect to the
 * FiLM table (-> w / w+ through e3_film_bwd), the local texture modulation and the sample
 * positions (volume_renderer.py:796-802, get_eikonal_term).  The generator weights are
 * frozen on this path: no weight gradients.  Camera parameters receive no gradient.
 * ---------------------------------------------------------------------------------- */
size_t e3_render_stash_bytes(int n_samples_per_ray, int rays_per_image, int batch);
size_t e3_render_bwd_scratch_bytes(int batch);

typedef struct e3_render_saved { /* results of the forward call, read again */
  const float* stash;    /* e3_render_outputs.bwd_stash */
  const float* sdf;      /* [B,H,W,S,1] */
  const float* hit_prob; /* [B,H,W,S,1] */
  const float* raw_rgb;  /* [B,H,W,S,3] */
} e3_render_saved;

typedef struct e3_render_grads { /* upstream gradients dL/d(output); any may be NULL (= zero) */
  const float* d_features;  /* [B,256,H,W] */
  const float* d_thumb_rgb; /* [B,3,H,W] */
  const float* d_xyz;       /* [B,3,H,W] */
  const float* d_depth;     /* [B,H,W,1,1] */
  const float* d_sdf;       /* [B,H,W,S,1] */
  const float* d_hit_prob;  /* [B,H,W,S,1] */
} e3_render_grads;

typedef struct e3_render_bwd_outputs {
  float* d_film;        /* [B,9,2,256]: per FiLM layer dL/dgamma, dL/dbeta (required) */
  float* d_local_alpha; /* NULL or [B,H,W,S,256] */
  float* d_local_beta;  /* NULL or [B,H,W,S,256] */
  float* d_points;      /* NULL or [B,H,W,S,3]: dL/d(world-space sample position) */
} e3_render_bwd_outputs;

/* p / in: exactly the arguments of the forward call.  scratch: e3_render_bwd_scratch_bytes(B). */
int e3_render_bwd(const void* packed, const e3_render_params* p, const e3_render_inputs* in,
                  const e3_render_saved* saved, const e3_render_grads* grads,
                  const e3_render_bwd_outputs* out, void* scratch, size_t scratch_bytes,
                  void* stream);

/* Backward of e3_siren_points_fwd_train.  d_sdf [B,N], d_raw_rgb [B,N,3], d_feat [B,N,256] may be
 * NULL; with_view = 0 runs the sdf-only graph (the forward call had raw_rgb == feat == NULL);
 * unit_sdf_seed != 0 uses dL/dsdf = 1 for every point, so that d_points [B,N,3] is the spatial sdf
 * gradient of get_eikonal_term (volume_renderer.py:796-802).  d_film [B,9,2,256] required. */
int e3_siren_points_bwd(const void* packed, const float* film, int batch, int n_points,
                        float pts_scale, const float* stash, int with_view, int unit_sdf_seed,
                        const float* d_sdf, const float* d_raw_rgb, const float* d_feat,
                        float* d_film, float* d_points, void* scratch, size_t scratch_bytes,
                        void* stream);

/* d_film [B,9,2,256] -> d_styles [B,styles_per_image,256] (adjoint of e3_film_fwd). */
int e3_film_bwd(const void* packed, const float* d_film, int batch, int styles_per_image,
                float* d_styles, void* stream);

/* ------------------------------------------------------------------------------------
 * StyleGAN2 ops — same semantics and argument order as the reference's extension ABI.
 * ---------------------------------------------------------------------------------- */

/* fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 * (fused_bias_act.cpp:11-20, fused_bias_act_kernel.cu:19-99).  x/y have `numel` elements;
 * bias (may be NULL) has `size_b` elements and applies along dim 1 with `step_b` =
 * prod(dims[2:]); refer (may be NULL) is the saved output for grad=1.  act: 1 linear,
 * 3 leaky-ReLU; grad: 0,1,2. */
int e3_fused_bias_act(const float* x, const float* bias, const float* refer, float* y,
                      int64_t numel, int64_t step_b, int64_t size_b, int act, int grad,
                      float alpha, float scale, void* stream);

/* upfirdn2d_op.upfirdn2d(input[major,in_h,in_w,minor], kernel[kh,kw], up_x, up_y, down_x,
 * down_y, pad_x0, pad_x1, pad_y0, pad_y1) (upfirdn2d.cpp:12-23, upfirdn2d_kernel.cu).
 * y is [major,out_h,out_w,minor] with out = (in*up + pad0 + pad1 - k)/down + 1. */
int e3_upfirdn2d(const float* x, const float* kernel, float* y, int major, int in_h,
                 int in_w, int minor, int kh, int kw, int up_x, int up_y, int down_x,
                 int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1, void* stream);

/* ------------------------------------------------------------------------------------
 * Modulated-conv decoder (Decoder / StyledConv / ToRGB, stylesdf_model.py:263-362,
 * 469-541, 587-797).
 * ---------------------------------------------------------------------------------- */

/* Per-sample modulation and demodulation factors of one ModulatedConv2d
 * (stylesdf_model.py:319-326):
 *   s[b,i]  = (mod_w[i,:]/sqrt(512)) . latent[b,:] + mod_b[i]
 *   d[b,o]  = rsqrt( sum_{i,k} (W[o,i,k]/sqrt(cin*k*k) * s[b,i])^2 + 1e-8 )   (d may be NULL)
 * latent rows are `latent_stride` floats apart (latent[:, layer] views of [B,n_latent,512]).
 * wsq [cout,cin] = sum_k W[o,i,k]^2, from e3_modconv_weight_sq. */
int e3_modconv_weight_sq(const float* weight, int cout, int cin, int ksize, float* wsq,
                         void* stream);
int e3_modconv_styles(const float* latent, int64_t latent_stride, const float* mod_w,
                      const float* mod_b, const float* wsq, int batch, int cin, int cout,
                      int ksize, float* s, float* d, void* stream);

/* Activations inside the decoder are channels-last (NHWC) fp32.
 * e3_nchw_to_nhwc / e3_nhwc_to_nchw convert at the boundary. */
int e3_nchw_to_nhwc(const float* x, float* y, int batch, int ch, int h, int w, void* stream);
int e3_nhwc_to_nchw(const float* x, float* y, int batch, int ch, int h, int w, void* stream);

/* Library-private packed image of one 3x3 conv weight [cout,cin,3,3] (reference layout,
 * `decoder.*.conv.weight` without its leading 1): the equalised-lr scale 1/sqrt(cin*9)
 * (stylesdf_model.py:301-302) is folded in and the taps are laid out GEMM-major
 * (plain: [tap][cin][cout]; upsample: [cin][tap*cout]), followed by the K-major bf16 hi/lo
 * split the tensor-core path reads through TMA (upsample: taps ordered by output parity phase).  Pack once per weight update. */
size_t e3_conv_packed_bytes(int cout, int cin);
int e3_conv_pack_weight(const float* weight, int cout, int cin, int upsample, void* packed,
                        void* stream);

/* arithmetic selection for the 3x3 modulated convs (flags argument) */
#define E3_CONV_AUTO 0u            /* tensor cores when the shape allows, else CUDA cores */
#define E3_CONV_FP32_CUDA_CORES 1u /* exact-fp32 FFMA implicit GEMM */
#define E3_CONV_TENSOR_CORES 2u    /* tcgen05 split-bf16 (hi*hi + hi*lo + lo*hi, fp32 accumulate in
                                      TMEM); E3_ERR_UNSUPPORTED unless cin % 64 == 0, cout % 128 == 0
                                      and (plain conv only) H, W are powers of two >= 8 */

/* StyledConv forward, plain 3x3 (stylesdf_model.py:356-360, 494-507):
 *   y = lrelu_0.2( d[b,o] * conv3x3(x * s[b,:], W/sqrt(cin*9)) + noise_w*noise[y,x]
 *                  + act_bias[o] ) * sqrt(2)
 * x [B,H,W,cin] NHWC, wpacked from e3_conv_pack_weight(upsample=0), noise [H,W] (the
 * registered noise_k buffer or a caller tensor; per-sample noise: noise_batch_stride=H*W,
 * else 0), y [B,H,W,cout] NHWC.  cin % 16 == 0, cout % 4 == 0.
 * act_bias == NULL selects the bare ModulatedConv2d.forward (y = d * conv(x*s, W), no noise,
 * bias or activation; stylesdf_model.py:317-362); noise / noise_w may then be NULL too. */
int e3_styled_conv3x3_fwd(const float* x, const void* wpacked, const float* s, const float* d,
                          const float* noise, int64_t noise_batch_stride,
                          const float* noise_w, const float* act_bias, float* y, int batch,
                          int h, int w, int cin, int cout, void* scratch,
                          size_t scratch_bytes, uint32_t flags, void* stream);

/* StyledConv forward, upsampling (stylesdf_model.py:331-346, 283-291): conv_transpose2d
 * stride 2 followed by the 4x4 [1,3,3,1] blur (gain 4, pad (1,1)), then noise + bias +
 * leaky-ReLU as above.  x [B,H,W,cin] -> y [B,2H,2W,cout]; noise [2H,2W];
 * wpacked from e3_conv_pack_weight(upsample=1). */
int e3_styled_conv3x3_up_fwd(const float* x, const void* wpacked, const float* s,
                             const float* d, const float* noise,
                             int64_t noise_batch_stride, const float* noise_w,
                             const float* act_bias, float* y, int batch, int h, int w,
                             int cin, int cout, void* scratch, size_t scratch_bytes,
                             uint32_t flags, void* stream);
size_t e3_styled_conv_scratch_bytes(int batch, int h, int w, int cin, int cout, int upsample);

/* Inference fusion of one resolution step of Decoder.forward (stylesdf_model.py:783-790: an upsampling
 * StyledConv followed by a plain one of the same width).  Nothing but the next conv reads the first
 * layer's output, so ..._up_fwd_split writes, instead of y, the next conv's tensor-core operands
 * xs = y * next_s[b,:] as bf16 hi / lo ([B,2H,2W,cout] each; bit-identical to what e3_styled_conv3x3_fwd
 * derives from y), and ..._fwd_presplit is e3_styled_conv3x3_fwd on such operands (cin = the first
 * layer's cout).  Tensor-core path only: e3_styled_conv_pair_fusable says whether the shapes allow it
 * (1) or the caller runs the two layers separately (0).  Scratch as for e3_styled_conv3x3_up_fwd. */
int e3_styled_conv_pair_fusable(int batch, int h, int w, int cin, int cout, uint32_t flags);
int e3_styled_conv3x3_up_fwd_split(const float* x, const void* wpacked, const float* s,
                                   const float* d, const float* noise,
                                   int64_t noise_batch_stride, const float* noise_w,
                                   const float* act_bias, const float* next_s, void* xs_hi,
                                   void* xs_lo, int batch, int h, int w, int cin, int cout,
                                   void* scratch, size_t scratch_bytes, uint32_t flags, void* stream);
int e3_styled_conv3x3_fwd_presplit(const void* xs_hi, const void* xs_lo, const void* wpacked,
                                   const float* d, const float* noise, int64_t noise_batch_stride,
                                   const float* noise_w, const float* act_bias, float* y, int batch,
                                   int h, int w, int cin, int cout, uint32_t flags, void* stream);

/* ToRGB forward (stylesdf_model.py:531-541, Upsample :96-119): 1x1 modulated conv without
 * demodulation + bias + (optionally FIR-upsampled) skip.
 * x [B,H,W,cin] NHWC; weight [3,cin]; skip NULL, or [B,3,H/2,W/2] NCHW when
 * upsample_skip != 0 (upfirdn2d up=2, kernel [1,3,3,1]^2/64*4, pad (2,1)), or [B,3,H,W];
 * rgb [B,3,H,W] NCHW (the image contract of the reference). */
int e3_torgb_fwd(const float* x, const float* weight, const float* s, const float* bias,
                 const float* skip, int upsample_skip, float* rgb, int batch, int h, int w,
                 int cin, void* stream);

/* ------------------------------------------------------------------------------------
 * Decoder backward — what autograd gives the reference through Decoder.forward
 * (stylesdf_model.py:742-797) when image losses are back-propagated into the encoders
 * (trainer.py:881-900, generator frozen at :1569): gradients with respect to the layer input and the layer's latent.
 * The generator weights are frozen on this path: no weight / bias / noise-strength gradients.
 * ---------------------------------------------------------------------------------- */

/* e3_conv_pack_weight layouts: 0 plain forward, 1 upsampling forward, 2 backward of the plain
 * conv, 3 backward of the upsampling conv (pass as the `upsample` argument). */
#define E3_CONV_PACK_FWD 0
#define E3_CONV_PACK_UP_FWD 1
#define E3_CONV_PACK_BWD 2
#define E3_CONV_PACK_UP_BWD 3

/* Backward of e3_styled_conv3x3_fwd (upsample = 0) / e3_styled_conv3x3_up_fwd (upsample = 1).
 * dy, y: [B,H',W',cout] NHWC (H' = 2H when upsampling) — upstream gradient and saved output;
 * x [B,H,W,cin] the saved input; s, d, noise, noise_w, act_bias as in the forward call;
 * wpacked_bwd from e3_conv_pack_weight(layout 2 or 3).
 * Outputs: dx [B,H,W,cin]; ds [B,cin] = dL/ds; dd [B,cout] = dL/dd (NULL when the layer does not
 * demodulate).  Feed ds, dd to e3_modconv_styles_bwd for the latent gradient.
 * Tensor cores when H, W are powers of two >= 8, cout % 64 == 0 and cin % 128 == 0 (flags as in
 * the forward call), else the exact-fp32 CUDA-core GEMM; cin, cout must be multiples of 16. */
size_t e3_styled_conv_bwd_scratch_bytes(int batch, int h, int w, int cin, int cout, int upsample);
int e3_styled_conv3x3_bwd(const float* dy, const float* y, const float* x, const void* wpacked_bwd,
                          const float* s, const float* d, const float* noise,
                          int64_t noise_batch_stride, const float* noise_w, const float* act_bias,
                          float* dx, float* ds, float* dd, int batch, int h, int w, int cin,
                          int cout, int upsample, void* scratch, size_t scratch_bytes,
                          uint32_t flags, void* stream);

/* Backward of e3_torgb_fwd with respect to x and s: drgb [B,3,H,W] NCHW -> dx [B,H,W,cin],
 * ds [B,cin].  (The skip gradient is drgb itself, or its e3_upfirdn2d adjoint: down = 2,
 * pad (1,1).) */
size_t e3_torgb_bwd_scratch_bytes(int batch, int cin);
int e3_torgb_bwd(const float* drgb, const float* x, const float* weight, const float* s, float* dx,
                 float* ds, int batch, int h, int w, int cin, void* scratch, size_t scratch_bytes,
                 void* stream);

/* Adjoint of e3_modconv_styles: dlatent [B,512] from ds [B,cin] and (when demodulating, else
 * NULL) dd [B,cout]; s, d, wsq as produced / consumed by the forward call. */
int e3_modconv_styles_bwd(const float* ds, const float* dd, const float* s, const float* d,
                          const float* wsq, const float* mod_w, int batch, int cin, int cout,
                          int ksize, float* dlatent, void* stream);

/* ------------------------------------------------------------------------------------
 * Image-parallel inversion record: packs, per image, the renderer latent
 * w+ [9*256], the decoder latent [n_latent*512] and K metric scalars into one contiguous
 * fp32 row of the all-gather send buffer, computing the metrics (mean squared error and
 * mean absolute error of `image` vs `target`, both [B,3,H,W]) on the device.
 * record [B, 2304 + n_latent*512 + 2]. */
int e3_pack_inversion_record(const float* w_plus, const float* w_dec, int n_latent,
                             const float* image, const float* target, int batch,
                             int64_t image_numel, float* record, void* stream);

/* Measurement aid (bench.py roofline denominator): runs `iters` rounds of 16 independent
 * FFMA chains per thread on every SM and writes one float per thread to `sink`
 * (sm_count*1024 floats... see e3_ffma_peak_probe_sink_floats).  FLOPs performed =
 * sink_floats * iters * 16 * 2. */
int e3_ffma_peak_probe(int iters, float* sink, void* stream);
size_t e3_ffma_peak_probe_sink_floats(void);

/* ----------------------------------------------------------------------------------
 * Local branch, first row of SURVEY.md section 8(f): pixel-aligned feature query.
 * Replaces HGPIFuNetGAN.query(points, calibs, ..., im_feat=F) of the reference
 * (project/vendor/pifu/lib/model/HGPIFuGANNet.py:85-150; callers e3dge_full_runner.py:219-226,
 * 244-250, 271-278): perspective projection of every point with the [3x4] rows of its image's
 * calibration (vendor/pifu/lib/geometry.py:108-135, including the batch-wide "look at -z" sign taken
 * from point 0 of image 0), y flip, in-image test, and a bilinear zero-padded grid_sample
 * (align_corners=False; geometry.py:64-80) of a C-channel feature map.
 *   feat_nhwc [B,H,W,C] channels-last (e3_nchw_to_nhwc of the reference's [B,C,H,W] map), C % 4 == 0;
 *   points: element (b, k, n) at points[b*pts_batch_stride + k*pts_coord_stride + n*pts_point_stride]
 *           ([B,3,N] as the reference passes them: strides (3N, N, 1); a renderer `points` output
 *           [B,N,3]: (3N, 1, 3));   calibs [B][calib_stride] row-major, first 12 floats of each used;
 *   outputs (each may be NULL): feats [B,N,C] (the layout the caller permutes the reference's [B,C,N]
 *           into), proj_xy [B,2,N], depth [B,1,N], in_img [B,N] (0 / 1).  feats == NULL = the
 *           reference's return_projection_only. */
int e3_local_feature_query(const float* feat_nhwc, const float* points, int64_t pts_batch_stride,
                           int64_t pts_coord_stride, int64_t pts_point_stride, const float* calibs,
                           int calib_stride, int batch, int n_points, int h, int w, int c,
                           float* feats, float* proj_xy, float* depth, unsigned char* in_img,
                           void* stream);

/* Adjoint of e3_local_feature_query with respect to the feature map: d_feat_nhwc [B,H,W,C] (overwritten) =
 * scatter-add of d_feats [B,N,C] over the four bilinear taps of every point (the reference's
 * op/grid_sample_gradfix.py backward; float atomics, so the summation order is not fixed).  Same point /
 * calibration arguments as the forward call. */
int e3_local_feature_query_bwd(const float* d_feats, const float* points, int64_t pts_batch_stride,
                               int64_t pts_coord_stride, int64_t pts_point_stride, const float* calibs,
                               int calib_stride, int batch, int n_points, int h, int w, int c,
                               float* d_feat_nhwc, void* stream);

/* ----------------------------------------------------------------------------------
 * Local branch, rest of SURVEY.md section 8(f) row 1: the per-sample MLP tail between the feature query
 * and the renderer's texture modulation — `Fuse_sft_MLP(257, 256)` of the runner
 * (project/models/helper_modules/sft.py:84-109, built at e3dge_full_runner.py:301-303, called at :289-290),
 * `PosEncoding(3, N_freqs=7)` of the sample position (project/utils/misc_utils.py:148-184,
 * e3dge_full_runner.py:260, 293-294) and netLocal's `local_feat_to_tex_modulations_linear` =
 * `ResnetBlockFC(301, 512)` (project/models/helper_modules/resnetfc.py:10-62,
 * vendor/pifu/lib/model/HGPIFuGANNetResidualInputResnetFC.py:84-86), split into (alpha, beta) as
 * SirenLocalGlobal.forward_backbone does (project/utils/volume_renderer.py:327-336).
 * 989 161 MACs per sample on the tensor cores (tcgen05, split-bf16 operands, fp32 accumulation).
 * Weight pointers are the reference's nn.Linear tensors ([out, in] row-major, fp32). */
typedef struct e3_local_mlp_weights {
  const float *enc_fc0_w, *enc_fc0_b;   /* fuse.encode_enc.fc_0   [256,513], [256] */
  const float *enc_fc1_w, *enc_fc1_b;   /* fuse.encode_enc.fc_1   [256,256], [256] */
  const float* enc_shortcut_w;          /* fuse.encode_enc.shortcut [256,513] (no bias) */
  const float *scale0_w, *scale0_b, *scale2_w, *scale2_b; /* fuse.scale.{0,2} [256,256], [256] */
  const float *shift0_w, *shift0_b, *shift2_w, *shift2_b; /* fuse.shift.{0,2} */
  const float *tex_fc0_w, *tex_fc0_b;   /* local_feat_to_tex_modulations_linear.fc_0 [301,301], [301] */
  const float *tex_fc1_w, *tex_fc1_b;   /* ....fc_1 [512,301], [512] */
  const float* tex_shortcut_w;          /* ....shortcut [512,301] */
} e3_local_mlp_weights;

size_t e3_local_mlp_packed_bytes(void);
/* weights -> packed operand image (bf16 hi / lo planes per GEMM stage, zero padded, + fp32 biases) */
int e3_local_mlp_pack(const e3_local_mlp_weights* w, void* packed, void* stream);
/* workspace for `rows` samples in one pass; a smaller workspace (>= the size for 128 rows) makes
 * e3_local_mlp_fwd walk the rows in chunks */
size_t e3_local_mlp_workspace_bytes(int64_t rows);
/* Either the whole tail — feat_2d [rows,257] (2-D-aligned features | visibility mask), feat_3d [rows,256]
 * (features projected from the reference view), points [rows,3] (world space), feats_in NULL — or, with
 * feats_in [rows,301] given and the three others NULL, the texture-modulation MLP alone (the reference's
 * `local_data_batch['feats']` contract).  Outputs alpha, beta [rows,256]; feats_out [rows,301] optional
 * (whole tail only): the 301-d features the reference materialises. */
int e3_local_mlp_fwd(const void* packed, const float* feat_2d, const float* feat_3d, const float* points,
                     const float* feats_in, int64_t rows, float* alpha, float* beta, float* feats_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Generic fp32-faithful linear layer on the tensor cores (the GEMM kernel of the local MLP tail):
 * y [rows,n] = x [rows,k] * W^T (+ bias [n]), W [n,k] row-major as nn.Linear stores it; split-bf16 operands
 * (hi*hi + hi*lo + lo*hi), fp32 accumulation in TMEM.  n % 128 == 0, k % 64 == 0.  Used by the layer-wise
 * sweeps of the second-order (eikonal) gradient of the SDF network (volume_renderer.py:796-802). */
size_t e3_tc_linear_packed_bytes(int n, int k);
int e3_tc_linear_pack(const float* w, int n, int k, void* packed, void* stream);
size_t e3_tc_linear_workspace_bytes(int64_t rows, int k);
int e3_tc_linear_fwd(const void* packed, int n, int k, const float* x, int64_t rows, const float* bias, float* y,
                     void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* E3DGE_B200_H_ */
