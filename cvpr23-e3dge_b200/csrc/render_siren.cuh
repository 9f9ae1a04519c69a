// Shared by the CUDA-core (render_siren.cu) and tensor-core (render_siren_tc.cu) renderers:
// packed weight image layout and kernel argument block.
#pragma once
#include "common.cuh"

namespace e3 {

constexpr int SW = 256;                        // SIREN width
constexpr int TILE_M = 96;                     // sample rows per tile
constexpr int ACT_LD = 100;                    // padded row stride of h[n][m] (bank-conflict free)
constexpr int KCHUNK = 16;                     // k rows per TMA slab
constexpr int STAGES = 4;
constexpr int CHUNK_FLOATS = KCHUNK * SW;      // 4096 floats = 16 KB
constexpr int CHUNKS_PER_LAYER = SW / KCHUNK;  // 16
constexpr int N_CONSUMER_WARPS = 8;
constexpr int N_CONSUMERS = N_CONSUMER_WARPS * 32;
constexpr int N_THREADS = N_CONSUMERS + 32;

// ---- packed weight image (floats) -----------------------------------------------------
// "p-order": column p of a packed slab holds output channel n(p) so that one lane's 8
// accumulator columns are two LDS.128 and the epilogue stores are conflict free.
constexpr int OFF_W0P = 0;                      // [3][256]   layer 0, p-order
constexpr int OFF_WVD = OFF_W0P + 3 * SW;       // [3][256]   view layer, view-dir inputs, p-order
constexpr int OFF_BIAS = OFF_WVD + 3 * SW;      // [9][256]   natural order (8 trunk + view)
constexpr int OFF_WSIG = OFF_BIAS + 9 * SW;     // [256]
constexpr int OFF_WRGB = OFF_WSIG + SW;         // [3][256]
constexpr int OFF_HEADB = OFF_WRGB + 3 * SW;    // bsig, brgb[3], pad -> 32
constexpr int SMALL_FLOATS = OFF_HEADB + 32;    // 4896
constexpr int OFF_STREAM = SMALL_FLOATS;        // [8][256][256] layers 1..7 + view, p-order
constexpr int OFF_GAMMA_W = OFF_STREAM + 8 * SW * SW;  // [9][256][256] natural (out,in)
constexpr int OFF_GAMMA_B = OFF_GAMMA_W + 9 * SW * SW;
constexpr int OFF_BETA_W = OFF_GAMMA_B + 9 * SW;
constexpr int OFF_BETA_B = OFF_BETA_W + 9 * SW * SW;
constexpr int OFF_W0N = OFF_BETA_B + 9 * SW;    // [3][256]  layer 0, natural channel order
constexpr int OFF_WVDN = OFF_W0N + 3 * SW;      // [3][256]  view-dir inputs, natural order
// tensor-core weight stream: bf16, pre-swizzled SWIZZLE_128B K-major tiles of 128 (n) x 64 (k),
// in consumption order [layer 0..7][k-block 0..3][hi, lo][n-half 0..1]; 16 KB per tile.
constexpr int TC_TILE_BYTES = 128 * 64 * 2;
constexpr int TC_TILES_PER_LAYER = 16;
constexpr int OFF_TC_STREAM = ((OFF_WVDN + 3 * SW + 31) / 32) * 32;  // floats; 8*16*16 KB = 2 MB follow
constexpr int TC_STREAM_FLOATS = 8 * TC_TILES_PER_LAYER * TC_TILE_BYTES / 4;
// backward stream (input gradients dH = G * W): tiles of W^T, 128 (k = forward input) x 64 (n = forward
// output, the contraction index), order [layer 0..7][n-block 0..3][hi, lo][k-half 0..1]; the backward
// kernel consumes the layers from 7 down to 0.
constexpr int OFF_TC_STREAM_BWD = OFF_TC_STREAM + TC_STREAM_FLOATS;
constexpr int PACKED_FLOATS = OFF_TC_STREAM_BWD + TC_STREAM_FLOATS;
// activation stash written by the forward kernel for the backward pass: the pre-sin phases
// arg = gamma*(W h + b) + beta of the 9 FiLM layers, [tile][layer 0..8][channel 256][row 128] fp32
// (rows innermost: a warp's 32 rows are one 128-byte line).
constexpr size_t STASH_FLOATS_PER_TILE = (size_t)9 * SW * 128;
static_assert((OFF_STREAM * 4) % 128 == 0, "weight stream must be 128B aligned");
static_assert((OFF_TC_STREAM * 4) % 128 == 0, "tensor-core weight stream must be 128B aligned");
constexpr int FILM_ROWS = 3;  // per layer: gamma, beta, beta' = gamma*bias + beta (bias folded)

__host__ __device__ __forceinline__ int chan_of_packed_col(int p) {
  // p = 128*q + 4*lane + jj  ->  n = lane + 32*(jj + 4*q)
  const int q = p >> 7, lane = (p & 127) >> 2, jj = p & 3;
  return lane + 32 * (jj + 4 * q);
}

// ---- kernel arguments -------------------------------------------------------------------
struct RenderArgs {
  const float* packed;
  e3_render_params p;
  e3_render_inputs in;
  e3_render_outputs out;
  int rays_per_tile, tiles_per_image, n_tiles;
  // explicit-points mode
  const float* points;
  const float* pviewdirs;
  int n_points;
  float* p_sdf;
  float* p_rgb;
  float* p_feat;
  float* p_h8;    // explicit-points mode, optional: backbone features (layer-8 output before the local modulation) [B,N,256]
  int with_view;  // 0: stop after the sdf head (sdf-only query)
  float* stash;   // NULL, or [n_tiles][9][256][128] pre-sin phases for e3_render_bwd (tensor-core kernel)
  // measurement aid (profiles/trace_render.py): when non-null, CTA 0 records clock64() stamps of the
  // barrier hand-offs of its second tile: [0,128) compute warp, [128,256) MMA waits, [256,384) MMA issue
  unsigned long long* trace;
};


int launch_render_tc(const RenderArgs& a, int mode, cudaStream_t stream);

// ---- backward (render_siren_bwd_tc.cu) ---------------------------------------------------------
// Gradients of the fused renderer with respect to its differentiable inputs: the FiLM table (hence
// the w / w+ latents), the local texture modulation and the sample positions.  The generator's own
// weights are frozen on the E3DGE path (trainer.py:1569 freezes the generator, the encoders train), so no weight
// gradients are produced.
constexpr int BWD_STAT_FLOATS = 9 * 2 * SW;  // per (CTA, image): [layer][sum gpre, sum gpre*arg][256]
struct RenderBwdArgs {
  const float* packed;
  e3_render_params p;
  e3_render_inputs in;   // cameras / film / local_alpha as in the forward call
  const float* stash;    // written by the forward call
  // forward results read again (MODE 0)
  const float* sdf;       // [B,H,W,S]
  const float* hit_prob;  // [B,H,W,S] composite weights
  const float* raw_rgb;   // [B,H,W,S,3]
  // upstream gradients (any may be NULL).  MODE 0: image-space maps; MODE 1: per point.
  const float* d_features;  // [B,256,H,W]
  const float* d_thumb_rgb; // [B,3,H,W]
  const float* d_xyz;       // [B,3,H,W]
  const float* d_depth;     // [B,H,W]
  const float* d_sdf;       // [B,H,W,S] / [B,N]
  const float* d_hit_prob;  // [B,H,W,S]
  const float* d_prgb;      // MODE 1: [B,N,3]
  const float* d_pfeat;     // MODE 1: [B,N,256]
  int n_points;
  // outputs
  float* film_partial;   // [grid][B][9][2][256], zero on entry; owner-CTA accumulation (deterministic)
  float* d_local_alpha;  // NULL or [B,H,W,S,256]
  float* d_local_beta;
  float* d_points;       // NULL or [B,H,W,S,3] / [B,N,3]: d/d(world-space sample position)
  int rays_per_tile, tiles_per_image, n_tiles;
  int with_view;         // 0: sdf-only graph (eikonal / geometry queries): 7 GEMMs
  int unit_sdf_seed;     // 1: d_sdf == 1 for every sample (eikonal term)
};
int render_bwd_grid(int n_tiles);
int launch_render_bwd_tc(const RenderBwdArgs& a, int mode, cudaStream_t stream);

}  // namespace e3
