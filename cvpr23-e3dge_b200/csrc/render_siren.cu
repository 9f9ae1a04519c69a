// Fused StyleSDF volume renderer for sm_100a:
//   rays -> samples -> 8 x FiLM-SIREN -> sdf head -> (local FiLM) -> view layer -> rgb head
//   -> SDF->sigma -> alpha composite -> feature map / thumbnail / depth / xyz
// in ONE persistent kernel.  No [N,256] activation ever reaches HBM (the reference
// materialises ~40 of them).
//
// Replaces (arithmetic spec: SURVEY.md Appendix A.1-A.6):
//   VolumeFeatureRenderer.get_rays / render / render_rays / run_network /
//   volume_integration     project/utils/volume_renderer.py:769-794,1666-1701,1183-1298,
//                          1052-1128,809-943
//   SirenGenerator.forward / FiLMSiren.forward           volume_renderer.py:240-264,116-132
//
// Design (B200):
//   * one persistent CTA per SM, 8 consumer warps + 1 TMA producer warp;
//   * a tile = 96 sample rows = floor(96/S) whole rays, so the composite needs no
//     cross-CTA traffic; hidden state h[256][96] lives in shared memory, k-major;
//   * per layer a 96x256x256 fp32 GEMM on the FFMA pipe, 12x8 register tile per thread
//     (fp32 is required: bf16/tf32 MMA misses the 1e-3 parity bar);
//   * the 2 MB of hidden-layer weights are streamed from L2 in 16 KB k-slabs by
//     cp.async.bulk (TMA, SASS UBLKCP) through a 4-stage full/empty mbarrier ring;
//   * sdf / rgb heads use warp-shuffle reductions; the S-step transmittance scan and the
//     weighted sums run out of shared memory; outputs are written once.
#include <cuda_bf16.h>

#include "render_siren.cuh"

namespace e3 {

__global__ void siren_pack_kernel(e3_siren_weights w, float* __restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= PACKED_FLOATS) return;
  float v = 0.f;
  if (idx < OFF_WVD) {  // W0p[k][p] = W0[n][k], W0 is [256,3]
    const int k = idx / SW, n = chan_of_packed_col(idx % SW);
    v = w.pts_w[0][n * 3 + k];
  } else if (idx < OFF_BIAS) {  // view-dir inputs of the view layer: Wv[n][256+j]
    const int r = idx - OFF_WVD, j = r / SW, n = chan_of_packed_col(r % SW);
    v = w.views_w[n * 259 + 256 + j];
  } else if (idx < OFF_WSIG) {
    const int r = idx - OFF_BIAS, l = r / SW, n = r % SW;
    v = (l < 8) ? w.pts_b[l][n] : w.views_b[n];
  } else if (idx < OFF_WRGB) {
    v = w.sigma_w[idx - OFF_WSIG];
  } else if (idx < OFF_HEADB) {
    v = w.rgb_w[idx - OFF_WRGB];
  } else if (idx < OFF_STREAM) {
    const int r = idx - OFF_HEADB;
    v = (r == 0) ? w.sigma_b[0] : (r < 4 ? w.rgb_b[r - 1] : 0.f);
  } else if (idx < OFF_GAMMA_W) {
    const int r = idx - OFF_STREAM, l = r / (SW * SW), k = (r / SW) % SW,
              n = chan_of_packed_col(r % SW);
    v = (l < 7) ? w.pts_w[l + 1][n * SW + k] : w.views_w[n * 259 + k];
  } else if (idx < OFF_GAMMA_B) {
    const int r = idx - OFF_GAMMA_W;
    v = w.gamma_w[r / (SW * SW)][r % (SW * SW)];
  } else if (idx < OFF_BETA_W) {
    const int r = idx - OFF_GAMMA_B;
    v = w.gamma_b[r / SW][r % SW];
  } else if (idx < OFF_BETA_B) {
    const int r = idx - OFF_BETA_W;
    v = w.beta_w[r / (SW * SW)][r % (SW * SW)];
  } else if (idx < OFF_W0N) {
    const int r = idx - OFF_BETA_B;
    v = w.beta_b[r / SW][r % SW];
  } else if (idx < OFF_WVDN) {
    const int r = idx - OFF_W0N;
    v = w.pts_w[0][(r % SW) * 3 + r / SW];
  } else if (idx < OFF_WVDN + 3 * SW) {
    const int r = idx - OFF_WVDN;
    v = w.views_w[(r % SW) * 259 + 256 + r / SW];
  } else if (idx < OFF_TC_STREAM) {
    v = 0.f;  // alignment padding
  } else if (idx >= OFF_TC_STREAM_BWD) {
    // backward stream: tiles of W^T (rows = forward input k, contraction over forward output n)
    const int e0 = (idx - OFF_TC_STREAM_BWD) * 2;
    uint32_t bits = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = e0 + h;
      const int tile = e / (128 * 64), in_tile = e % (128 * 64);
      const int kh = tile & 1, is_lo = (tile >> 1) & 1, nb = (tile >> 2) & 3, l = tile >> 4;
      const int byte = in_tile * 2;
      const int r = (byte >> 10) * 8 + ((byte >> 7) & 7);
      const int chunk = ((byte >> 4) & 7) ^ (r & 7);
      const int kk = chunk * 8 + ((byte >> 1) & 7);
      const int k = kh * 128 + r, n = nb * 64 + kk;
      const float wv = (l < 7) ? w.pts_w[l + 1][n * SW + k] : w.views_w[n * 259 + k];
      const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
      const __nv_bfloat16 lo = __float2bfloat16_rn(wv - __bfloat162float(hi));
      const uint16_t u = is_lo ? __bfloat16_as_ushort(lo) : __bfloat16_as_ushort(hi);
      bits |= (uint32_t)u << (16 * h);
    }
    v = __uint_as_float(bits);
  } else {
    // two bf16 per float slot of the tensor-core stream
    const int e0 = (idx - OFF_TC_STREAM) * 2;
    uint32_t bits = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = e0 + h;                       // bf16 element index in the stream
      const int tile = e / (128 * 64), in_tile = e % (128 * 64);
      const int nh = tile & 1, is_lo = (tile >> 1) & 1, kb = (tile >> 2) & 3, l = tile >> 4;
      // invert the swizzled placement: byte offset -> (row r, k within block)
      const int byte = in_tile * 2;
      const int r = (byte >> 10) * 8 + ((byte >> 7) & 7);
      const int chunk = ((byte >> 4) & 7) ^ (r & 7);
      const int kk = chunk * 8 + ((byte >> 1) & 7);
      const int n = nh * 128 + r, k = kb * 64 + kk;
      const float wv = (l < 7) ? w.pts_w[l + 1][n * SW + k] : w.views_w[n * 259 + k];
      const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
      const __nv_bfloat16 lo = __float2bfloat16_rn(wv - __bfloat162float(hi));
      const uint16_t u = is_lo ? __bfloat16_as_ushort(lo) : __bfloat16_as_ushort(hi);
      bits |= (uint32_t)u << (16 * h);
    }
    v = __uint_as_float(bits);
  }
  packed[idx] = v;
}

// gamma/beta of all 9 FiLM layers of one image: one warp per output row, lanes over k.  Latency-bound
// (4.7 MB of weights, 9 MFLOP): blockIdx.y splits the 256 rows into 8 chunks and a warp requests its
// four rows' 64 coalesced loads at once.
__global__ void __launch_bounds__(256) film_kernel(const float* __restrict__ packed,
                                                   const float* __restrict__ styles,
                                                   int styles_per_image,
                                                   float* __restrict__ film) {
  const int b = blockIdx.x / 9, l = blockIdx.x % 9;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sidx = (styles_per_image > 1) ? (l < styles_per_image ? l : styles_per_image - 1) : 0;
  const float* st = styles + ((size_t)b * styles_per_image + sidx) * SW;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = st[lane + 32 * i];
  const float* gw = packed + OFF_GAMMA_W + (size_t)l * SW * SW;
  const float* bw = packed + OFF_BETA_W + (size_t)l * SW * SW;
  float* out = film + ((size_t)b * 9 + l) * FILM_ROWS * SW;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = blockIdx.y * 32 + warp * 4 + r;
    float g = 0.f, be = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g = fmaf(__ldg(gw + n * SW + lane + 32 * i), s[i], g);
      be = fmaf(__ldg(bw + n * SW + lane + 32 * i), s[i], be);
    }
    g = warp_sum(g);
    be = warp_sum(be);
    if (lane == 0) {
      // LinearLayer: std_init*(Wx+b)+bias_init  (volume_renderer.py:76-80,107-114)
      const float gam = 15.f * (g + packed[OFF_GAMMA_B + l * SW + n]) + 30.f;
      const float bet = 0.25f * (be + packed[OFF_BETA_B + l * SW + n]);
      out[n] = gam;
      out[SW + n] = bet;
      out[2 * SW + n] = fmaf(gam, packed[OFF_BIAS + l * SW + n], bet);  // layer bias folded
    }
  }
}

struct Smem {
  float act[SW * ACT_LD];             // h[n][m]
  float ring[STAGES * CHUNK_FLOATS];  // weight slabs
  float smallw[SMALL_FLOATS];
  float film[9 * 2 * SW];
  float xin[3 * ACT_LD];   // normalised sample coordinates, k-major
  float vdir[3 * ACT_LD];  // view direction per sample, k-major
  float z[TILE_M], dist[TILE_M], sdf[TILE_M], alpha[TILE_M], wgt[TILE_M], vis[TILE_M];
  float rgb[3 * TILE_M];
  float ray_o[3 * TILE_M], ray_d[3 * TILE_M], ray_v[3 * TILE_M];
  float ray_near[TILE_M], ray_far[TILE_M], ray_dn[TILE_M];
  uint64_t full[STAGES], empty[STAGES];
};
static_assert(sizeof(Smem) <= 227 * 1024, "shared memory budget");
static_assert(offsetof(Smem, ring) % 128 == 0, "ring must be 128B aligned");

__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(N_CONSUMERS) : "memory");
}

// acc[12][8] += A[k][12 rows of this warp] * B[k][8 cols of this lane], k = 0..K-1
template <int K>
__device__ __forceinline__ void mac_rows(float (&acc)[12][8], const float* __restrict__ a_rows,
                                         const float* __restrict__ b_rows, int lane) {
#pragma unroll
  for (int kk = 0; kk < K; ++kk) {
    const float4 a0 = *reinterpret_cast<const float4*>(a_rows + kk * ACT_LD);
    const float4 a1 = *reinterpret_cast<const float4*>(a_rows + kk * ACT_LD + 4);
    const float4 a2 = *reinterpret_cast<const float4*>(a_rows + kk * ACT_LD + 8);
    const float4 b0 = *reinterpret_cast<const float4*>(b_rows + kk * SW + 4 * lane);
    const float4 b1 = *reinterpret_cast<const float4*>(b_rows + kk * SW + 128 + 4 * lane);
    const float a[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// One hidden layer's K=256 contraction, weights arriving through the TMA ring.
__device__ __forceinline__ void gemm_streamed(float (&acc)[12][8], Smem& sm, int warp, int lane,
                                              uint32_t& stage, uint32_t& phase) {
  const float* a_base = sm.act + 12 * warp;
  for (int c = 0; c < CHUNKS_PER_LAYER; ++c) {
    mbar_wait(&sm.full[stage], phase);
    mac_rows<KCHUNK>(acc, a_base + c * KCHUNK * ACT_LD, sm.ring + stage * CHUNK_FLOATS, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[stage]);
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
}

// FiLM + sin epilogue: h[n][m] = sin(gamma[n] * (acc + bias[n]) + beta[n])
__device__ __forceinline__ void film_sin_store(const float (&acc)[12][8], Smem& sm, int layer,
                                               int warp, int lane) {
  const float* bias = sm.smallw + OFF_BIAS + layer * SW;
  const float* gam = sm.film + layer * 2 * SW;
  const float* bet = gam + SW;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane + 32 * j;
    const float bb = bias[n], g = gam[n], be = bet[n];
    float v[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = sin_accurate(fmaf(g, acc[i][j] + bb, be));
    float4* dst = reinterpret_cast<float4*>(sm.act + n * ACT_LD + 12 * warp);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    dst[2] = make_float4(v[8], v[9], v[10], v[11]);
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[12][8]) {
#pragma unroll
  for (int i = 0; i < 12; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// out[m] = sum_n w[c][n] * h[n][m] for the warp's 12 rows, NC heads at once; lane 0 gets sums.
template <int NC>
__device__ __forceinline__ void head_dot(const Smem& sm, const float* __restrict__ w, int warp,
                                         int lane, float (&out)[NC][12]) {
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 12; ++i) out[c][i] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane + 32 * j;
    const float4* src = reinterpret_cast<const float4*>(sm.act + n * ACT_LD + 12 * warp);
    const float4 h0 = src[0], h1 = src[1], h2 = src[2];
    const float h[12] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, h2.x, h2.y, h2.z, h2.w};
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float wc = w[c * SW + n];
#pragma unroll
      for (int i = 0; i < 12; ++i) out[c][i] = fmaf(wc, h[i], out[c][i]);
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 12; ++i) out[c][i] = warp_sum(out[c][i]);
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// MODE 0: camera rays + composite.  MODE 1: explicit points (sdf / raw rgb / features).
template <int MODE>
__global__ void __launch_bounds__(N_THREADS, 1)
siren_render_kernel(const __grid_constant__ RenderArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], N_CONSUMER_WARPS);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < SMALL_FLOATS; i += N_THREADS) sm.smallw[i] = a.packed[i];
  __syncthreads();

  const int n_my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int chunks_per_tile = (a.with_view ? 8 : 7) * CHUNKS_PER_LAYER;

  if (warp == N_CONSUMER_WARPS) {
    // ===== TMA producer: streams layers 1..7 (+ view) for every tile of this CTA =====
    if (lane == 0) {
      const float* stream = a.packed + OFF_STREAM;
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        for (int c = 0; c < chunks_per_tile; ++c) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&sm.full[stage], CHUNK_FLOATS * 4);
          tma_bulk_g2s(sm.ring + stage * CHUNK_FLOATS, stream + (size_t)c * CHUNK_FLOATS,
                       CHUNK_FLOATS * 4, &sm.full[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    return;
  }

  // ===== consumers =====
  const e3_render_params& P = a.p;
  const int S = (MODE == 0) ? P.n_samples : 1;
  const int HW = (MODE == 0) ? P.height * P.width : a.n_points;
  uint32_t stage = 0, phase = 0;
  int cur_b = -1;
  float acc[12][8];

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_image;
    const int t_in = tile - b * a.tiles_per_image;
    const int unit0 = t_in * a.rays_per_tile;  // first ray (MODE 0) / point (MODE 1)
    const int n_units = min(a.rays_per_tile, HW - unit0);
    const int n_valid = n_units * S;
    const size_t samp0 = ((size_t)b * HW + unit0) * S;  // first sample row in [B,HW,S]

    if (b != cur_b) {
      const float* f = a.in.film + (size_t)b * 9 * FILM_ROWS * SW;
      for (int i = tid; i < 9 * 2 * SW; i += N_CONSUMERS)
        sm.film[i] = f[(i / (2 * SW)) * FILM_ROWS * SW + (i % (2 * SW))];
      cur_b = b;
    }

    // ---- phase A: rays (volume_renderer.py:769-794, 1678-1688) ----
    if (MODE == 0) {
      if (tid < n_units) {
        const int ray = unit0 + tid, py = ray / P.width, px = ray - py * P.width;
        const float foc = a.in.focal[b];
        const float half = (float)P.res * 0.5f;
        const float dx = __fdiv_rn(__fsub_rn(a.in.pix_x[px], half), foc);
        const float dy = -__fdiv_rn(__fsub_rn(a.in.pix_y[py], half), foc);
        const float dz = -1.f;
        const float* c2w = a.in.cam_poses + (size_t)b * 12;
        float rd[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)  // torch.sum(dirs * R[r,:], -1): left-to-right fp32
          rd[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w[r * 4 + 0]), __fmul_rn(dy, c2w[r * 4 + 1])),
                            __fmul_rn(dz, c2w[r * 4 + 2]));
        float vx, vy, vz;
        if (P.flags & E3_RENDER_STATIC_VIEWDIRS) {
          vx = dx, vy = dy, vz = dz;
        } else {
          vx = rd[0], vy = rd[1], vz = rd[2];
        }
        const float vn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)),
                                         __fmul_rn(vz, vz)));
        vx = __fdiv_rn(vx, vn), vy = __fdiv_rn(vy, vn), vz = __fdiv_rn(vz, vn);
        const float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])),
                                         __fmul_rn(rd[2], rd[2])));
        const float o[3] = {c2w[3], c2w[7], c2w[11]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          sm.ray_o[r * TILE_M + tid] = o[r];
          sm.ray_d[r * TILE_M + tid] = rd[r];
        }
        sm.ray_v[0 * TILE_M + tid] = vx;
        sm.ray_v[1 * TILE_M + tid] = vy;
        sm.ray_v[2 * TILE_M + tid] = vz;
        sm.ray_near[tid] = a.in.near[b];
        sm.ray_far[tid] = a.in.far[b];
        sm.ray_dn[tid] = dn;
        const size_t ro = ((size_t)b * HW + ray) * 3;
        if (a.out.rays_o) a.out.rays_o[ro] = o[0], a.out.rays_o[ro + 1] = o[1], a.out.rays_o[ro + 2] = o[2];
        if (a.out.rays_d) a.out.rays_d[ro] = rd[0], a.out.rays_d[ro + 1] = rd[1], a.out.rays_d[ro + 2] = rd[2];
        if (a.out.viewdirs) a.out.viewdirs[ro] = vx, a.out.viewdirs[ro + 1] = vy, a.out.viewdirs[ro + 2] = vz;
      }
      consumer_sync();
    }

    // ---- phase B: samples (volume_renderer.py:1211,1231-1233,1074-1079) ----
    if (tid < TILE_M) {
      const int m = tid;
      float x0 = 0.f, x1 = 0.f, x2 = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f;
      if (m < n_valid) {
        if (MODE == 0) {
          const int r = m / S, s = m - r * S;
          const float nr = sm.ray_near[r], fr = sm.ray_far[r];
          float z;
          if (a.in.z_jitter) {
            z = a.in.z_jitter[samp0 + m];
          } else {
            const float t = a.in.t_vals[s];
            z = __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
          }
          float pw[3];
#pragma unroll
          for (int c = 0; c < 3; ++c)
            pw[c] = __fadd_rn(sm.ray_o[c * TILE_M + r], __fmul_rn(sm.ray_d[c * TILE_M + r], z));
          sm.z[m] = z;
          x0 = __fmul_rn(pw[0], P.pts_scale), x1 = __fmul_rn(pw[1], P.pts_scale),
          x2 = __fmul_rn(pw[2], P.pts_scale);
          v0 = sm.ray_v[r], v1 = sm.ray_v[TILE_M + r], v2 = sm.ray_v[2 * TILE_M + r];
          if (a.out.points) {
            float* o = a.out.points + (samp0 + m) * 3;
            o[0] = pw[0], o[1] = pw[1], o[2] = pw[2];
          }
        } else {
          const float* pp = a.points + (samp0 + m) * 3;
          x0 = __fmul_rn(pp[0], P.pts_scale), x1 = __fmul_rn(pp[1], P.pts_scale),
          x2 = __fmul_rn(pp[2], P.pts_scale);
          if (a.pviewdirs) {
            const float* vv = a.pviewdirs + (samp0 + m) * 3;
            v0 = vv[0], v1 = vv[1], v2 = vv[2];
          }
        }
      }
      sm.xin[m] = x0, sm.xin[ACT_LD + m] = x1, sm.xin[2 * ACT_LD + m] = x2;
      sm.vdir[m] = v0, sm.vdir[ACT_LD + m] = v1, sm.vdir[2 * ACT_LD + m] = v2;
    }
    consumer_sync();
    if (MODE == 0 && tid < TILE_M && tid < n_valid) {
      // dists (volume_renderer.py:826-837)
      const int m = tid, r = m / S, s = m - r * S;
      float d;
      if (s + 1 < S) d = __fsub_rn(sm.z[m + 1], sm.z[m]);
      else if (P.flags & E3_RENDER_NO_FORCE_STOP) d = (S > 1) ? __fsub_rn(sm.z[r * S + 1], sm.z[r * S]) : 0.f;
      else d = 1e10f;
      d = __fmul_rn(d, sm.ray_dn[r]);
      sm.dist[m] = d;
      if (a.out.dists) a.out.dists[samp0 + m] = d;
    }

    auto write_tap = [&](int tap) {  // h[n][m] -> feats_taps[tap][b][ray][s][n]
      float* dst = a.out.feats_taps + ((size_t)tap * P.batch * HW * S + samp0) * SW;
      for (int m = 0; m < n_valid; ++m) dst[(size_t)m * SW + tid] = sm.act[tid * ACT_LD + m];
    };

    // ---- layer 0 (K = 3) ----
    zero_acc(acc);
    mac_rows<3>(acc, sm.xin + 12 * warp, sm.smallw + OFF_W0P, lane);
    film_sin_store(acc, sm, 0, warp, lane);
    consumer_sync();
    if (MODE == 0 && a.out.feats_taps) write_tap(0);

    // ---- layers 1..7 ----
    for (int l = 1; l < 8; ++l) {
      zero_acc(acc);
      gemm_streamed(acc, sm, warp, lane, stage, phase);
      consumer_sync();  // every warp is done reading h before it is overwritten
      film_sin_store(acc, sm, l, warp, lane);
      consumer_sync();
      // rendering.return_feats: taps after reference layers (i+1) in {1,3,5,7}, i.e. the
      // outputs of 0-based layers 0,2,4,6 (volume_renderer.py:179-180)
      if (MODE == 0 && a.out.feats_taps && (l & 1) == 0) write_tap(l >> 1);
    }

    // ---- sdf head (volume_renderer.py:206-208) ----
    {
      float s1[1][12];
      head_dot<1>(sm, sm.smallw + OFF_WSIG, warp, lane, s1);
      if (lane == 0) {
        const float bs = sm.smallw[OFF_HEADB];
#pragma unroll
        for (int i = 0; i < 12; ++i) sm.sdf[12 * warp + i] = s1[0][i] + bs;
      }
    }
    consumer_sync();

    if (MODE == 0) {
      // ---- SDF -> sigma -> alpha (volume_renderer.py:804-807,853-867) ----
      if (tid < n_valid) {
        const int m = tid;
        const float sd = sm.sdf[m];
        float al;
        if (P.flags & E3_RENDER_NO_SDF) {
          const float sp = (sd > 20.f) ? sd : log1pf(expf(sd));
          al = 1.f - expf(-sp * sm.dist[m]);
        } else {
          const float beta = a.in.sigmoid_beta[0];
          const float sigma = __fdiv_rn(sigmoidf_acc(__fdiv_rn(-sd, beta)), beta);
          al = 1.f - expf(-sigma * sm.dist[m]);
        }
        sm.alpha[m] = al;
        if (a.out.sdf) a.out.sdf[samp0 + m] = sd;
      }
      consumer_sync();
      // ---- transmittance scan per ray (volume_renderer.py:869-886, 905-910) ----
      if (tid < n_units) {
        const int r = tid;
        float T = 1.f, wsum = 0.f;
        for (int s = 0; s < S; ++s) {
          const int m = r * S + s;
          const float al = sm.alpha[m];
          float w = __fmul_rn(al, T);
          if ((P.flags & E3_RENDER_FORCE_BACKGROUND) && !(P.flags & E3_RENDER_NO_FORCE_STOP) &&
              s == S - 1)
            w = __fsub_rn(1.f, wsum);
          sm.vis[m] = T;
          sm.wgt[m] = w;
          wsum = __fadd_rn(wsum, w);
          T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, al), 1e-10f));
        }
        float depth = 0.f, xs = 0.f, ys = 0.f, zs = 0.f;
        for (int s = 0; s < S; ++s) {
          const int m = r * S + s;
          const float w = sm.wgt[m], z = sm.z[m];
          depth = fmaf(w, z, depth);
          xs = fmaf(w, __fadd_rn(sm.ray_o[r], __fmul_rn(sm.ray_d[r], z)), xs);
          ys = fmaf(w, __fadd_rn(sm.ray_o[TILE_M + r], __fmul_rn(sm.ray_d[TILE_M + r], z)), ys);
          zs = fmaf(w, __fadd_rn(sm.ray_o[2 * TILE_M + r], __fmul_rn(sm.ray_d[2 * TILE_M + r], z)), zs);
        }
        const size_t pix = (size_t)b * HW + unit0 + r;
        if (a.out.depth) a.out.depth[pix] = depth;
        if (a.out.mask) a.out.mask[pix] = (depth < P.mask_depth) ? 1.f : 0.f;
        if (a.out.xyz) {
          float* o = a.out.xyz + (size_t)b * 3 * HW + unit0 + r;
          o[0] = xs, o[HW] = ys, o[2 * (size_t)HW] = zs;
        }
      }
      consumer_sync();
      if (tid < n_valid) {
        if (a.out.hit_prob) a.out.hit_prob[samp0 + tid] = sm.wgt[tid];
        if (a.out.visibility) a.out.visibility[samp0 + tid] = sm.vis[tid];
      }
      // ---- local-branch texture FiLM before the view layer (volume_renderer.py:217-220) ----
      if (a.in.local_alpha) {
        const float* la = a.in.local_alpha + samp0 * SW;
        const float* lb = a.in.local_beta + samp0 * SW;
        for (int m = 0; m < n_valid; ++m) {
          const float h = sm.act[tid * ACT_LD + m];
          sm.act[tid * ACT_LD + m] =
              __fadd_rn(__fmul_rn(__fadd_rn(la[(size_t)m * SW + tid], 1.f), h), lb[(size_t)m * SW + tid]);
        }
        consumer_sync();
      }
    } else {
      if (tid < n_valid) a.p_sdf[samp0 + tid] = sm.sdf[tid];
    }

    if (a.with_view) {
      // ---- view layer: K = 256 (h) + 3 (view dir) (volume_renderer.py:222-233) ----
      zero_acc(acc);
      gemm_streamed(acc, sm, warp, lane, stage, phase);
      mac_rows<3>(acc, sm.vdir + 12 * warp, sm.smallw + OFF_WVD, lane);
      consumer_sync();
      film_sin_store(acc, sm, 8, warp, lane);
      consumer_sync();

      // ---- rgb head (volume_renderer.py:235) ----
      {
        float c3[3][12];
        head_dot<3>(sm, sm.smallw + OFF_WRGB, warp, lane, c3);
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int i = 0; i < 12; ++i)
              sm.rgb[c * TILE_M + 12 * warp + i] = c3[c][i] + sm.smallw[OFF_HEADB + 1 + c];
        }
      }
      consumer_sync();

      if (MODE == 0) {
        // ---- composite (volume_renderer.py:888-894) ----
        if (a.out.raw_rgb && tid < n_valid) {
          float* o = a.out.raw_rgb + (samp0 + tid) * 3;
          o[0] = sm.rgb[tid], o[1] = sm.rgb[TILE_M + tid], o[2] = sm.rgb[2 * TILE_M + tid];
        }
        if (a.out.thumb_rgb && tid < 3 * n_units) {
          const int c = tid / n_units, r = tid - c * n_units;
          float accum = 0.f;
          for (int s = 0; s < S; ++s) {
            const int m = r * S + s;
            accum = fmaf(sm.wgt[m], sigmoidf_acc(sm.rgb[c * TILE_M + m]), accum);
          }
          a.out.thumb_rgb[((size_t)b * 3 + c) * HW + unit0 + r] = -1.f + 2.f * accum;
        }
        if (a.out.features) {
          const int n = tid;  // one output channel per consumer thread
          const float* row = sm.act + n * ACT_LD;
          float* o = a.out.features + ((size_t)b * SW + n) * HW + unit0;
          for (int r = 0; r < n_units; ++r) {
            float accum = 0.f;
            for (int s = 0; s < S; ++s) accum = fmaf(sm.wgt[r * S + s], row[r * S + s], accum);
            o[r] = accum;
          }
        }
      } else {
        if (a.p_rgb && tid < n_valid) {
          float* o = a.p_rgb + (samp0 + tid) * 3;
          o[0] = sm.rgb[tid], o[1] = sm.rgb[TILE_M + tid], o[2] = sm.rgb[2 * TILE_M + tid];
        }
        if (a.p_feat) {
          float* dst = a.p_feat + samp0 * SW;
          for (int m = 0; m < n_valid; ++m) dst[(size_t)m * SW + tid] = sm.act[tid * ACT_LD + m];
        }
      }
    }
    consumer_sync();  // smem is reused by the next tile
  }
}

static int launch_render(const RenderArgs& a, int mode, cudaStream_t stream) {
  static thread_local bool attr_set_dev[E3_MAX_DEVICES][2] = {};  // function attributes are per device
  bool (&attr_set)[2] = attr_set_dev[device_slot()];
  const void* fn = (mode == 0) ? (const void*)siren_render_kernel<0> : (const void*)siren_render_kernel<1>;
  if (!attr_set[mode]) {
    E3_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    attr_set[mode] = true;
  }
  const int grid = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
  if (grid <= 0) return E3_OK;
  if (mode == 0)
    siren_render_kernel<0><<<grid, N_THREADS, sizeof(Smem), stream>>>(a);
  else
    siren_render_kernel<1><<<grid, N_THREADS, sizeof(Smem), stream>>>(a);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

}  // namespace e3

using namespace e3;

extern "C" size_t e3_siren_packed_bytes(void) { return (size_t)PACKED_FLOATS * sizeof(float); }

extern "C" int e3_siren_pack(const e3_siren_weights* w, void* packed, void* stream) {
  E3_REQUIRE(w && packed, E3_ERR_BAD_ARG, "e3_siren_pack: null argument");
  E3_REQUIRE(((uintptr_t)packed & 127) == 0, E3_ERR_BAD_ARG, "e3_siren_pack: packed must be 128B aligned");
  const void* const* ptrs = reinterpret_cast<const void* const*>(w);
  for (size_t i = 0; i < sizeof(e3_siren_weights) / sizeof(void*); ++i)
    E3_REQUIRE(ptrs[i] != nullptr, E3_ERR_BAD_ARG, "e3_siren_pack: weight pointer %zu is null", i);
  const int threads = 256, blocks = (PACKED_FLOATS + threads - 1) / threads;
  siren_pack_kernel<<<blocks, threads, 0, as_stream(stream)>>>(*w, static_cast<float*>(packed));
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" int e3_film_fwd(const void* packed, const float* styles, int batch, int styles_per_image,
                           float* film, void* stream) {
  E3_REQUIRE(batch >= 0 && (styles_per_image == 1 || styles_per_image == 9), E3_ERR_BAD_ARG,
             "e3_film_fwd: styles_per_image must be 1 (w) or 9 (w+), got %d", styles_per_image);
  if (batch == 0) return E3_OK;  // empty batches carry null data pointers
  E3_REQUIRE(packed && styles && film, E3_ERR_BAD_ARG, "e3_film_fwd: null argument");
  film_kernel<<<dim3(batch * 9, SW / 32), 256, 0, as_stream(stream)>>>(static_cast<const float*>(packed), styles,
                                                                      styles_per_image, film);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" int e3_render_fwd(const void* packed, const e3_render_params* p, const e3_render_inputs* in,
                             const e3_render_outputs* out, void* stream) {
  E3_REQUIRE(packed && p && in && out, E3_ERR_BAD_ARG, "e3_render_fwd: null argument");
  E3_REQUIRE(p->batch >= 0 && p->height > 0 && p->width > 0 && p->res > 0, E3_ERR_BAD_ARG,
             "e3_render_fwd: bad geometry B=%d H=%d W=%d res=%d", p->batch, p->height, p->width, p->res);
  if (p->batch == 0) return E3_OK;  // empty batches carry null data pointers
  const bool ffma = (p->flags & E3_RENDER_FP32_CUDA_CORES) != 0;
  const int tile_m = ffma ? TILE_M : 128;
  E3_REQUIRE(p->n_samples >= 1 && p->n_samples <= tile_m, E3_ERR_UNSUPPORTED,
             "e3_render_fwd: n_samples=%d outside [1,%d] (use e3_siren_points_fwd for sdf grids)",
             p->n_samples, tile_m);
  E3_REQUIRE(in->cam_poses && in->focal && in->near && in->far && in->pix_x && in->pix_y && in->film,
             E3_ERR_BAD_ARG, "e3_render_fwd: missing camera / film input");
  E3_REQUIRE(in->t_vals || in->z_jitter, E3_ERR_BAD_ARG, "e3_render_fwd: need t_vals or z_jitter");
  E3_REQUIRE((p->flags & E3_RENDER_NO_SDF) || in->sigmoid_beta, E3_ERR_BAD_ARG,
             "e3_render_fwd: sigmoid_beta missing");
  E3_REQUIRE((in->local_alpha == nullptr) == (in->local_beta == nullptr), E3_ERR_BAD_ARG,
             "e3_render_fwd: local_alpha and local_beta come together");
  if (p->batch == 0) return E3_OK;
  RenderArgs a{};
  a.packed = static_cast<const float*>(packed);
  a.p = *p;
  a.in = *in;
  a.out = *out;
  a.rays_per_tile = tile_m / p->n_samples;
  const int hw = p->height * p->width;
  a.tiles_per_image = (hw + a.rays_per_tile - 1) / a.rays_per_tile;
  a.n_tiles = a.tiles_per_image * p->batch;
  a.with_view = 1;
  E3_REQUIRE(!(ffma && out->bwd_stash), E3_ERR_UNSUPPORTED,
             "e3_render_fwd: bwd_stash (training) needs the tensor-core renderer");
  a.stash = out->bwd_stash;
  return ffma ? launch_render(a, 0, as_stream(stream)) : launch_render_tc(a, 0, as_stream(stream));
}

static int siren_points_fwd_impl(const void* packed, const float* film, const float* points,
                                 const float* viewdirs, int batch, int n_points, float pts_scale,
                                 float* sdf, float* raw_rgb, float* feat, uint32_t flags, float* stash,
                                 void* stream, const float* local_alpha = nullptr, const float* local_beta = nullptr,
                                 float* h8 = nullptr) {
  E3_REQUIRE(batch >= 0 && n_points >= 0, E3_ERR_BAD_ARG, "e3_siren_points_fwd: negative size");
  if (batch == 0 || n_points == 0) return E3_OK;
  E3_REQUIRE(packed && film && points && sdf, E3_ERR_BAD_ARG, "e3_siren_points_fwd: null argument");
  RenderArgs a{};
  a.packed = static_cast<const float*>(packed);
  a.p.batch = batch;
  a.p.pts_scale = pts_scale;
  a.p.n_samples = 1;
  a.in.film = film;
  a.points = points;
  a.pviewdirs = viewdirs;
  a.n_points = n_points;
  a.p_sdf = sdf;
  a.p_rgb = raw_rgb;
  a.p_feat = feat;
  a.p_h8 = h8;
  a.in.local_alpha = local_alpha;
  a.in.local_beta = local_beta;
  const bool ffma = (flags & E3_RENDER_FP32_CUDA_CORES) != 0;
  E3_REQUIRE(!(ffma && (h8 || local_alpha)), E3_ERR_UNSUPPORTED,
             "e3_siren_points_fwd_ex: backbone features / local modulation need the tensor-core kernel");
  E3_REQUIRE((local_alpha == nullptr) == (local_beta == nullptr), E3_ERR_BAD_ARG,
             "e3_siren_points_fwd_ex: local_alpha and local_beta come together");
  const int tile_m = ffma ? TILE_M : 128;
  a.rays_per_tile = tile_m;
  a.tiles_per_image = (n_points + tile_m - 1) / tile_m;
  a.n_tiles = a.tiles_per_image * batch;
  a.with_view = (raw_rgb || feat) ? 1 : 0;
  a.stash = stash;
  return ffma ? launch_render(a, 1, as_stream(stream)) : launch_render_tc(a, 1, as_stream(stream));
}

extern "C" int e3_siren_points_fwd(const void* packed, const float* film, const float* points,
                                   const float* viewdirs, int batch, int n_points, float pts_scale,
                                   float* sdf, float* raw_rgb, float* feat, uint32_t flags,
                                   void* stream) {
  return siren_points_fwd_impl(packed, film, points, viewdirs, batch, n_points, pts_scale, sdf, raw_rgb, feat,
                               flags, nullptr, stream);
}

extern "C" int e3_siren_points_fwd_ex(const void* packed, const float* film, const float* points,
                                      const float* viewdirs, int batch, int n_points, float pts_scale,
                                      const float* local_alpha, const float* local_beta, float* sdf,
                                      float* raw_rgb, float* feat, float* h8, uint32_t flags, void* stream) {
  return siren_points_fwd_impl(packed, film, points, viewdirs, batch, n_points, pts_scale, sdf, raw_rgb, feat,
                               flags, nullptr, stream, local_alpha, local_beta, h8);
}

extern "C" int e3_siren_points_fwd_train(const void* packed, const float* film, const float* points,
                                         const float* viewdirs, int batch, int n_points, float pts_scale,
                                         float* sdf, float* raw_rgb, float* feat, float* bwd_stash,
                                         void* stream) {
  E3_REQUIRE(bwd_stash || batch == 0 || n_points == 0, E3_ERR_BAD_ARG, "e3_siren_points_fwd_train: bwd_stash missing");
  return siren_points_fwd_impl(packed, film, points, viewdirs, batch, n_points, pts_scale, sdf, raw_rgb, feat, 0,
                               bwd_stash, stream);
}
