// Backward of the fused StyleSDF volume renderer on the tensor cores (tcgen05), sm_100a.
//
// What the reference gets from autograd through VolumeFeatureRenderer.forward
// (volume_renderer.py:1865-1972 -> render_rays :1183-1298 -> run_network :1052-1128 ->
// SirenGenerator.forward :240-264 -> volume_integration :809-943) when the E3DGE runners train
// their encoders against the frozen generator (trainer.py:881-900, generator frozen at :1569, e3dge_full_runner.py:219-306):
// gradients with respect to the FiLM frequencies / phases (hence the w / w+ latents), the local
// texture modulation (alpha, beta) of the PIFu branch and the sample positions (eikonal term,
// volume_renderer.py:796-802).  One persistent kernel per call, mirror image of
// render_siren_tc.cu:
//
//   * a tile = the same 128 sample rows = whole rays as the forward tile; the forward kernel
//     stashed the pre-sin phases arg_l = gamma_l*(W_l h + b_l) + beta_l of the nine FiLM layers
//     ([tile][layer][channel][row] fp32, rows innermost so a warp reads 128-byte lines);
//   * head (CUDA cores): upstream image-space gradients -> per-sample dL/df (feature map, rgb
//     thumbnail through the rgb head), dL/dw_s (needs f = sin(arg_8) again) -> the backward
//     transmittance recurrence per ray -> dL/dsdf;
//   * eight GEMMs dH_{l-1} = (dH_l * cos(arg_l) * gamma_l) W_l on the tensor cores, operands split
//     into bf16 hi + lo exactly like the forward pass (three products, fp32 accumulation in TMEM),
//     weights W^T streamed by TMA from a second pre-swizzled image (OFF_TC_STREAM_BWD) through
//     the same 4 x 16 KB mbarrier ring, two TMEM accumulators ping-ponged by GEMM parity;
//   * epilogue warps: tcgen05.ld -> * cos(arg) -> per-channel sums over the tile's rows of
//     g and g*arg (a 16-shuffle transposing butterfly per 16 columns) -> * gamma -> hi/lo split ->
//     swizzled st.shared of the next A operand, published 64 channels at a time;
//   * the per-(image, layer, channel) sums are accumulated by their owner thread into a per-CTA
//     slice of `film_partial` (no atomics: results are bit-reproducible) and folded into
//     dgamma = (S2 - beta*S1)/gamma, dbeta = S1 by film_grad_reduce_kernel.
//
// No weight gradients: the generator is frozen on this path.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "render_siren.cuh"
#include "tcgen05.cuh"

namespace e3 {
namespace {

constexpr int TCM = 128;
constexpr int RING = 4;
constexpr int CWARPS = 16;
constexpr int NCOMP = CWARPS * 32;
constexpr int NTHREADS = 64 + NCOMP;
constexpr int A_KBLOCK_BYTES = TCM * 128;
constexpr int A_BYTES = 4 * A_KBLOCK_BYTES;

struct SmemBwd {
  uint8_t a_hi[A_BYTES];
  uint8_t a_lo[A_BYTES];
  uint8_t ring[RING * TC_TILE_BYTES];
  float gamma[9][SW];
  float wsig[SW];
  float stat[4][2][SW];      // [row quarter][sum g | sum g*arg][channel] of the current layer
  float dw_part[4][TCM];     // partial dL/dw_s of the four column quarters
  float dx_part[4][3][TCM];  // partial dL/dx of the four column quarters
  float alpha[TCM], chain[TCM], dw[TCM], dalpha[TCM], dsdf[TCM];
  uint64_t full[RING], empty[RING];
  uint64_t a_ready[4], d_ready;
  uint32_t tmem_slot;
};
static_assert(sizeof(SmemBwd) + 1024 <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCOMP) : "memory"); }
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ uint32_t a_chunk_off(int m, int c16) {
  return (uint32_t)((m >> 3) * 1024 + (m & 7) * 128 + ((c16 ^ (m & 7)) << 4));
}
__device__ __forceinline__ void store_a8(SmemBwd& sm, int m, int n0, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair_bf16(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  const uint32_t off = (uint32_t)(n0 >> 6) * A_KBLOCK_BYTES + a_chunk_off(m, (n0 & 63) >> 3);
  *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// exact reduction to [-pi, pi] (reduce_2pi, common.cuh), then MUFU sin / cos
__device__ __forceinline__ float cos_reduced(float a) { return __cosf(reduce_2pi(a)); }
__device__ __forceinline__ void sincos_reduced(float a, float& s, float& c) {
  const float r = reduce_2pi(a);
  s = __sinf(r);
  c = __cosf(r);
}

// v[i] = this lane's (row's) value of column i.  Returns, in lane L, the sum over the warp's 32 rows
// of column L >> 1 (both lanes of a pair hold it): a transposing butterfly, 16 shuffles.
__device__ __forceinline__ float colsum16(const float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2];
  const bool u4 = lane & 16, u3 = lane & 8, u2 = lane & 4, u1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = u4 ? v[i] : v[i + 8], keep = u4 ? v[i + 8] : v[i];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = u3 ? a8[i] : a8[i + 4], keep = u3 ? a8[i + 4] : a8[i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = u2 ? a4[i] : a4[i + 2], keep = u2 ? a4[i + 2] : a4[i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = u1 ? a2[0] : a2[1], keep = u1 ? a2[1] : a2[0];
  float a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
siren_render_bwd_tc_kernel(const __grid_constant__ RenderBwdArgs a, const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ uint8_t smem_raw[];
  SmemBwd& sm = *reinterpret_cast<SmemBwd*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < RING; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) mbar_init(&sm.a_ready[j], CWARPS);
    mbar_init(&sm.d_ready, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < SW; i += NTHREADS) sm.wsig[i] = a.packed[OFF_WSIG + i];
  if (warp == 1) tc::tmem_alloc(&sm.tmem_slot, 512);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = sm.tmem_slot;

  const int n_my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_gemm = a.with_view ? 8 : 7;   // GEMM g contracts with stream layer n_gemm-1-g
  const int top_lo = n_gemm - 1;            // GEMM g yields dL/dh of 0-based layer top_lo - g

  if (warp == 0) {
    // ===== TMA producer: W^T tiles, layers from the top down =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        for (int g = 0; g < n_gemm; ++g) {
          const int ls = n_gemm - 1 - g;
          for (int c = 0; c < TC_TILES_PER_LAYER; ++c) {
            mbar_wait(&sm.empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&sm.full[stage], TC_TILE_BYTES);
            tc::tma_load_2d(sm.ring + stage * TC_TILE_BYTES, &wmap, &sm.full[stage], 0,
                            (ls * TC_TILES_PER_LAYER + c) * 128);
            if (++stage == RING) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (same schedule as the forward kernel) =====
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16_f32(128, 256);
      const uint32_t a_hi0 = smem_u32(sm.a_hi), a_lo0 = smem_u32(sm.a_lo);
      uint32_t stage = 0, phase = 0, pa = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        for (int g = 0; g < n_gemm; ++g) {
          const uint32_t dcol = tmem_base + (uint32_t)(g & 1) * 256;
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&sm.a_ready[kb], pa);
            tc::fence_after_thread_sync();
            const uint64_t dAh = tc::make_smem_desc_sw128(a_hi0 + kb * A_KBLOCK_BYTES);
            const uint64_t dAl = tc::make_smem_desc_sw128(a_lo0 + kb * A_KBLOCK_BYTES);
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {  // pr 0: W_hi block (two 128-row halves), pr 1: W_lo block
              mbar_wait(&sm.full[stage], phase);
              mbar_wait(&sm.full[stage + 1], phase);
              tc::fence_after_thread_sync();
              const uint64_t dB = tc::make_smem_desc_sw128(smem_u32(sm.ring + stage * TC_TILE_BYTES));
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t bk = tc::advance_desc_k(dB, ks);
                if (pr == 0) {
                  tc::mma_bf16_ss(dcol, tc::advance_desc_k(dAh, ks), bk, idesc, (kb | ks) != 0);
                  tc::mma_bf16_ss(dcol, tc::advance_desc_k(dAl, ks), bk, idesc, true);
                } else {
                  tc::mma_bf16_ss(dcol, tc::advance_desc_k(dAh, ks), bk, idesc, true);
                }
              }
              tc::mma_commit(&sm.empty[stage]);
              tc::mma_commit(&sm.empty[stage + 1]);
              stage += 2;
              if (stage == RING) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          pa ^= 1;
          tc::mma_commit(&sm.d_ready);
        }
      }
    }
  } else {
    // ===== compute warps =====
    const e3_render_params& P = a.p;
    const int ct = tid - 64;
    const int q = warp & 3;          // TMEM lane quarter
    const int hw = (warp - 2) >> 2;  // 16-column quarter of each 64-channel block
    const int m = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + hw * 16;
    const int S = (MODE == 0) ? P.n_samples : 1;
    const int HW = (MODE == 0) ? P.height * P.width : a.n_points;
    const float* pk = a.packed;
    uint32_t pd = 0;
    int cur_b = -1;
    const bool prefetch = !(P.flags & (1u << 28));  // bit 28: measurement aid, disables the stash prefetch

    auto publish = [&](int j) {
      fence_proxy_async();
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.a_ready[j]);
    };

    for (int slot = 0; slot < n_my_tiles; ++slot) {
      const int tile = blockIdx.x + slot * gridDim.x;
      const int b = tile / a.tiles_per_image;
      const int t_in = tile - b * a.tiles_per_image;
      const int unit0 = t_in * a.rays_per_tile;
      const int n_units = min(a.rays_per_tile, HW - unit0);
      const int n_valid = n_units * S;
      const size_t samp0 = ((size_t)b * HW + unit0) * S;
      const bool valid = m < n_valid;
      const int r = valid ? m / S : 0, s = valid ? m - r * S : 0;
      const int ray = unit0 + r;

      if (b != cur_b) {
        const float* f = a.in.film + (size_t)b * 9 * FILM_ROWS * SW;
        for (int i = ct; i < 9 * SW; i += NCOMP) sm.gamma[i / SW][i % SW] = f[(i / SW) * FILM_ROWS * SW + (i % SW)];
        cur_b = b;
      }

      // ---- per-row geometry, recomputed exactly as the forward kernel does ----
      float z = 0.f, dist = 0.f, pw0 = 0.f, pw1 = 0.f, pw2 = 0.f;
      if (MODE == 0 && valid) {
        const int py = ray / P.width, px = ray - py * P.width;
        const float foc = a.in.focal[b], half = (float)P.res * 0.5f;
        const float dx = __fdiv_rn(__fsub_rn(a.in.pix_x[px], half), foc);
        const float dy = -__fdiv_rn(__fsub_rn(a.in.pix_y[py], half), foc);
        const float dz = -1.f;
        const float* c2w = a.in.cam_poses + (size_t)b * 12;
        float rd[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          rd[c] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w[c * 4 + 0]), __fmul_rn(dy, c2w[c * 4 + 1])),
                            __fmul_rn(dz, c2w[c * 4 + 2]));
        const float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])),
                                         __fmul_rn(rd[2], rd[2])));
        const float nr = a.in.near[b], fr = a.in.far[b];
        auto z_of = [&](int si) -> float {
          if (a.in.z_jitter) return a.in.z_jitter[samp0 + r * S + si];
          const float t = a.in.t_vals[si];
          return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
        };
        z = z_of(s);
        pw0 = __fadd_rn(c2w[3], __fmul_rn(rd[0], z));
        pw1 = __fadd_rn(c2w[7], __fmul_rn(rd[1], z));
        pw2 = __fadd_rn(c2w[11], __fmul_rn(rd[2], z));
        float dd;
        if (s + 1 < S) dd = __fsub_rn(z_of(s + 1), z);
        else if (P.flags & E3_RENDER_NO_FORCE_STOP) dd = (S > 1) ? __fsub_rn(z_of(1), z_of(0)) : 0.f;
        else dd = 1e10f;
        dist = __fmul_rn(dd, dn);
      }

      compute_sync();  // gamma table visible; the previous tile's readers of shared arrays are done

      const float* stash = a.stash + (size_t)tile * STASH_FLOATS_PER_TILE + m;
      float* part = a.film_partial + ((size_t)blockIdx.x * P.batch + b) * BWD_STAT_FLOATS + ct;

      // per 16-column block: channel sums of g and g*arg, then (feed) gamma * g -> next A operand
      auto tail = [&](int lo, int j, const float (&g)[16], const float (&arg)[16], bool feed) {
        float t2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) t2[i] = g[i] * arg[i];
        const float s1 = colsum16(g, lane);
        const float s2 = colsum16(t2, lane);
        sm.stat[q][lane & 1][j * 64 + hw * 16 + (lane >> 1)] = (lane & 1) ? s2 : s1;
        if (feed) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const int n0 = j * 64 + hw * 16 + g8 * 8;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = g[g8 * 8 + i] * sm.gamma[lo][n0 + i];
            store_a8(sm, m, n0, v);
          }
          publish(j);
        }
      };
      // the layer's sums -> this CTA's slice of film_partial (owner thread, plain read-modify-write)
      auto flush = [&](int lo) {
        compute_sync();
        const int st = ct >> 8, n = ct & 255;
        part[lo * 2 * SW] += (sm.stat[0][st][n] + sm.stat[1][st][n]) + (sm.stat[2][st][n] + sm.stat[3][st][n]);
        compute_sync();
      };
      auto load_args = [&](int lo, int nb, float (&arg)[16]) {
#pragma unroll
        for (int i = 0; i < 16; ++i) arg[i] = valid ? __ldg(stash + (size_t)(lo * SW + nb + i) * TCM) : 0.f;
      };

      if (a.with_view) {
        // ---- head: dL/df per sample from the image-space gradients; view layer pre-activation grads ----
        float w_row = (MODE == 0) ? 0.f : 1.f, drgb0 = 0.f, drgb1 = 0.f, drgb2 = 0.f, dw_acc = 0.f;
        const float* dfeat = nullptr;
        if (valid) {
          if (MODE == 0) {
            w_row = a.hit_prob[samp0 + m];
            float dr[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float dT = a.d_thumb_rgb ? a.d_thumb_rgb[((size_t)b * 3 + c) * HW + ray] : 0.f;
              const float sg = sigmoid_acc(a.raw_rgb[(samp0 + m) * 3 + c]);
              dr[c] = w_row * 2.f * dT * sg * (1.f - sg);
              if (hw == 0) dw_acc = fmaf(2.f * dT, sg, dw_acc);
            }
            drgb0 = dr[0], drgb1 = dr[1], drgb2 = dr[2];
            if (a.d_features) dfeat = a.d_features + (size_t)b * SW * HW + ray;
          } else {
            if (a.d_prgb) {
              const float* o = a.d_prgb + (samp0 + m) * 3;
              drgb0 = o[0], drgb1 = o[1], drgb2 = o[2];
            }
            if (a.d_pfeat) dfeat = a.d_pfeat + (samp0 + m) * SW;
          }
        }
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int nb = j * 64 + hw * 16;
          float g[16], arg[16], df[16];
          load_args(8, nb, arg);
          if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) df[i] = dfeat ? __ldg(dfeat + (size_t)(nb + i) * HW) : 0.f;
          } else {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 t = dfeat ? *reinterpret_cast<const float4*>(dfeat + nb + i4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              df[i4 * 4] = t.x, df[i4 * 4 + 1] = t.y, df[i4 * 4 + 2] = t.z, df[i4 * 4 + 3] = t.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = nb + i;
            float gf = w_row * df[i];
            gf = fmaf(__ldg(pk + OFF_WRGB + n), drgb0, gf);
            gf = fmaf(__ldg(pk + OFF_WRGB + SW + n), drgb1, gf);
            gf = fmaf(__ldg(pk + OFF_WRGB + 2 * SW + n), drgb2, gf);
            float sn, cs;
            sincos_reduced(arg[i], sn, cs);
            if (MODE == 0) dw_acc = fmaf(df[i], sn, dw_acc);
            g[i] = gf * cs;
          }
          tail(8, j, g, arg, true);
        }
        if (MODE == 0) sm.dw_part[hw][m] = dw_acc;
        flush(8);
      }

      // ---- dL/dsdf per sample ----
      if (MODE == 0) {
        if (hw == 0) {
          float al = 0.f, chain = 0.f, dw = 0.f;
          if (valid) {
            const float sd = a.sdf[samp0 + m];
            if (P.flags & E3_RENDER_NO_SDF) {
              const float sp = (sd > 20.f) ? sd : log1pf(expf(sd));
              const float e = expf(-sp * dist);
              al = 1.f - e;
              chain = dist * e * sigmoid_acc(sd);
            } else {
              const float beta = a.in.sigmoid_beta[0];
              const float sg = sigmoid_acc(__fdiv_rn(-sd, beta));
              const float sigma = __fdiv_rn(sg, beta);
              const float e = expf(-sigma * dist);
              al = 1.f - e;
              chain = (dist * e) * (-(sg * (1.f - sg)) / (beta * beta));
            }
            dw = (sm.dw_part[0][m] + sm.dw_part[1][m]) + (sm.dw_part[2][m] + sm.dw_part[3][m]);
            const size_t pix = (size_t)b * HW + ray;
            if (a.d_depth) dw = fmaf(a.d_depth[pix], z, dw);
            if (a.d_xyz) {
              const float* o = a.d_xyz + (size_t)b * 3 * HW + ray;
              dw = fmaf(o[0], pw0, dw);
              dw = fmaf(o[HW], pw1, dw);
              dw = fmaf(o[2 * (size_t)HW], pw2, dw);
            }
            if (a.d_hit_prob) dw += a.d_hit_prob[samp0 + m];
          }
          sm.alpha[m] = al;
          sm.chain[m] = chain;
          sm.dw[m] = dw;
        }
        compute_sync();
        if (ct < n_units) {
          // w_s = alpha_s T_s, T_{s+1} = T_s (1 - alpha_s + 1e-10)  (volume_renderer.py:869-886)
          const int rr = ct;
          const bool fb = (P.flags & E3_RENDER_FORCE_BACKGROUND) && !(P.flags & E3_RENDER_NO_FORCE_STOP);
          float T = 1.f;
          for (int si = 0; si < S; ++si) {
            const int mm = rr * S + si;
            sm.dalpha[mm] = T;
            T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, sm.alpha[mm]), 1e-10f));
          }
          const float dlast = fb ? sm.dw[rr * S + S - 1] : 0.f;  // w_{S-1} = 1 - sum_{j<S-1} w_j
          float gT = 0.f;                                       // dL/dT_{s+1}
          for (int si = S - 1; si >= 0; --si) {
            const int mm = rr * S + si;
            float da = 0.f, gs = 0.f;
            if (!(fb && si == S - 1)) {
              const float al = sm.alpha[mm], Ts = sm.dalpha[mm];
              const float dwe = sm.dw[mm] - dlast;
              da = (dwe - gT) * Ts;
              gs = fmaf(dwe, al, gT * __fadd_rn(__fsub_rn(1.f, al), 1e-10f));
            }
            sm.dalpha[mm] = da;
            gT = gs;
          }
        }
        compute_sync();
        if (hw == 0) {
          float ds = 0.f;
          if (valid) {
            ds = sm.dalpha[m] * sm.chain[m];
            if (a.d_sdf) ds += a.d_sdf[samp0 + m];
          }
          sm.dsdf[m] = ds;
        }
        compute_sync();
      } else {
        if (hw == 0) sm.dsdf[m] = valid ? (a.unit_sdf_seed ? 1.f : (a.d_sdf ? a.d_sdf[samp0 + m] : 0.f)) : 0.f;
        compute_sync();
      }

      if (!a.with_view) {
        // sdf-only graph: dL/dh7 = w_sigma * dsdf, straight into layer 7's pre-activation gradient
        const float ds = sm.dsdf[m];
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int nb = j * 64 + hw * 16;
          float g[16], arg[16];
          load_args(7, nb, arg);
#pragma unroll
          for (int i = 0; i < 16; ++i) g[i] = sm.wsig[nb + i] * ds * cos_reduced(arg[i]);
          tail(7, j, g, arg, true);
        }
        flush(7);
      }

      // ---- GEMM epilogues, layers from the top down ----
      for (int gi = 0; gi < n_gemm; ++gi) {
        const int lo = top_lo - gi;
        const bool feed = lo >= 1;
        const bool l7 = (lo == 7);  // only with the view layer: local modulation + sdf head join here
        float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
        const float ds = l7 ? sm.dsdf[m] : 0.f;
        const float* la = nullptr;
        if (MODE == 0 && l7 && a.in.local_alpha && valid) la = a.in.local_alpha + (samp0 + m) * SW;
        // The stash phases come from HBM.  While this warp would only be waiting for the GEMM, request
        // block 0's phases into registers and pull the lines of blocks 1..3 into L2.
        float arg0[16];
        if (prefetch) {
          load_args(lo, hw * 16, arg0);
          if (valid) {
#pragma unroll
            for (int jj = 1; jj < 4; ++jj)
#pragma unroll
              for (int i = 0; i < 16; ++i)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(stash + (size_t)(lo * SW + jj * 64 + hw * 16 + i) * TCM));
          }
        }
        mbar_wait(&sm.d_ready, pd);
        pd ^= 1;
        tc::fence_after_thread_sync();
        const uint32_t dsrc = trow + (uint32_t)(gi & 1) * 256;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int nb = j * 64 + hw * 16;
          float acc[16], arg[16], g[16];
          if (prefetch && j == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) arg[i] = arg0[i];
          } else {
            load_args(lo, nb, arg);
          }
          tc::tmem_ld_32x16(dsrc + j * 64, acc);
          if (l7) {
            if (la) {
              // h7' = (alpha + 1) * h7 + beta_loc  (volume_renderer.py:217-220)
              float* oa = a.d_local_alpha ? a.d_local_alpha + (samp0 + m) * SW + nb : nullptr;
              float* ob = a.d_local_beta ? a.d_local_beta + (samp0 + m) * SW + nb : nullptr;
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 av = *reinterpret_cast<const float4*>(la + nb + i4 * 4);
                const float aa[4] = {av.x, av.y, av.z, av.w};
                float da4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float h7 = __sinf(reduce_2pi(arg[i4 * 4 + i]));
                  da4[i] = acc[i4 * 4 + i] * h7;
                }
                if (oa) *reinterpret_cast<float4*>(oa + i4 * 4) = make_float4(da4[0], da4[1], da4[2], da4[3]);
                if (ob)
                  *reinterpret_cast<float4*>(ob + i4 * 4) =
                      make_float4(acc[i4 * 4], acc[i4 * 4 + 1], acc[i4 * 4 + 2], acc[i4 * 4 + 3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i4 * 4 + i] *= __fadd_rn(aa[i], 1.f);
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(sm.wsig[nb + i], ds, acc[i]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) g[i] = acc[i] * cos_reduced(arg[i]);
          if (lo == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = nb + i;
              const float ga = g[i] * sm.gamma[0][n];
              dx0 = fmaf(ga, __ldg(pk + OFF_W0N + n), dx0);
              dx1 = fmaf(ga, __ldg(pk + OFF_W0N + SW + n), dx1);
              dx2 = fmaf(ga, __ldg(pk + OFF_W0N + 2 * SW + n), dx2);
            }
          }
          tail(lo, j, g, arg, feed);
        }
        if (!feed) {
          tc::fence_before_thread_sync();
          sm.dx_part[hw][0][m] = dx0;
          sm.dx_part[hw][1][m] = dx1;
          sm.dx_part[hw][2][m] = dx2;
        }
        flush(lo);
      }

      if (a.d_points && hw == 0 && valid) {
        float* o = a.d_points + (samp0 + m) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          o[c] = P.pts_scale * ((sm.dx_part[0][c][m] + sm.dx_part[1][c][m]) + (sm.dx_part[2][c][m] + sm.dx_part[3][c][m]));
      }
      compute_sync();  // shared memory is reused by the next tile
    }
  }

  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// film_partial [grid][B][9][2][256] -> d_film [B][9][2][256] = (dgamma, dbeta):
// arg = gamma*pre + beta  =>  dbeta = sum g,  dgamma = sum g*pre = (sum g*arg - beta * sum g) / gamma
__global__ void __launch_bounds__(256) film_grad_reduce_kernel(const float* __restrict__ partial, int grid,
                                                               const float* __restrict__ film, int batch,
                                                               float* __restrict__ d_film) {
  const int bl = blockIdx.x, n = threadIdx.x;  // bl = b*9 + l
  float s1 = 0.f, s2 = 0.f;
  for (int c = 0; c < grid; ++c) {
    const float* p = partial + ((size_t)c * batch * 9 + bl) * 2 * SW;
    s1 += p[n];
    s2 += p[SW + n];
  }
  const float gam = film[(size_t)bl * FILM_ROWS * SW + n], bet = film[(size_t)bl * FILM_ROWS * SW + SW + n];
  d_film[(size_t)bl * 2 * SW + n] = (fabsf(gam) > 1e-20f) ? (s2 - bet * s1) / gam : 0.f;
  d_film[(size_t)bl * 2 * SW + SW + n] = s1;
}

// d_film -> d_styles: gamma = 15*(G w + g) + 30, beta = 0.25*(H w + h)  (volume_renderer.py:107-114)
__global__ void __launch_bounds__(256) film_bwd_kernel(const float* __restrict__ packed,
                                                       const float* __restrict__ d_film, int styles_per_image,
                                                       float* __restrict__ d_styles) {
  __shared__ float dg[SW], db[SW];
  const int b = blockIdx.x / styles_per_image, si = blockIdx.x % styles_per_image, k = threadIdx.x;
  float acc = 0.f;
  for (int l = 0; l < 9; ++l) {
    const int sidx = (styles_per_image > 1) ? (l < styles_per_image ? l : styles_per_image - 1) : 0;
    if (sidx != si) continue;
    __syncthreads();
    dg[k] = 15.f * d_film[((size_t)b * 9 + l) * 2 * SW + k];
    db[k] = 0.25f * d_film[((size_t)b * 9 + l) * 2 * SW + SW + k];
    __syncthreads();
    const float* gw = packed + OFF_GAMMA_W + (size_t)l * SW * SW;
    const float* bw = packed + OFF_BETA_W + (size_t)l * SW * SW;
    for (int n = 0; n < SW; ++n) acc = fmaf(dg[n], gw[n * SW + k], fmaf(db[n], bw[n * SW + k], acc));
  }
  d_styles[((size_t)b * styles_per_image + si) * SW + k] = acc;
}

template <int MODE>
int launch_bwd_variant(const RenderBwdArgs& a, cudaStream_t stream) {
  static thread_local bool attr_set_dev[E3_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_set_dev[device_slot()];
  const int smem_bytes = (int)sizeof(SmemBwd) + 1024;
  auto* fn = siren_render_bwd_tc_kernel<MODE>;
  if (!attr_set) {
    E3_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  CUtensorMap wmap;
  const uint64_t wdims[2] = {64, (uint64_t)8 * TC_TILES_PER_LAYER * 128};
  const uint64_t wstr[1] = {128};
  const uint32_t wbox[2] = {64, 128};
  int rc = make_tensor_map_bf16(&wmap, a.packed + OFF_TC_STREAM_BWD, 2, wdims, wstr, wbox, /*swizzle128=*/false);
  if (rc) return rc;
  fn<<<render_bwd_grid(a.n_tiles), NTHREADS, smem_bytes, stream>>>(a, wmap);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

}  // namespace

int render_bwd_grid(int n_tiles) { return n_tiles < sm_count() ? n_tiles : sm_count(); }

int launch_render_bwd_tc(const RenderBwdArgs& a, int mode, cudaStream_t stream) {
  if (a.n_tiles <= 0) return E3_OK;
  return mode == 0 ? launch_bwd_variant<0>(a, stream) : launch_bwd_variant<1>(a, stream);
}

}  // namespace e3

using namespace e3;

extern "C" size_t e3_render_stash_bytes(int n_samples_per_ray, int rays_per_image, int batch) {
  if (n_samples_per_ray < 1 || n_samples_per_ray > 128 || rays_per_image < 0 || batch < 0) return 0;
  const int rpt = 128 / n_samples_per_ray;
  const size_t tiles = (size_t)((rays_per_image + rpt - 1) / rpt) * batch;
  return tiles * STASH_FLOATS_PER_TILE * sizeof(float);
}

extern "C" size_t e3_render_bwd_scratch_bytes(int batch) {
  return (size_t)sm_count() * (batch > 0 ? batch : 0) * BWD_STAT_FLOATS * sizeof(float);
}

static int finish_film_grads(const RenderBwdArgs& a, int grid, float* d_film, cudaStream_t st) {
  film_grad_reduce_kernel<<<a.p.batch * 9, 256, 0, st>>>(a.film_partial, grid, a.in.film, a.p.batch, d_film);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" int e3_render_bwd(const void* packed, const e3_render_params* p, const e3_render_inputs* in,
                             const e3_render_saved* saved, const e3_render_grads* grads,
                             const e3_render_bwd_outputs* out, void* scratch, size_t scratch_bytes,
                             void* stream) {
  E3_REQUIRE(packed && p && in && saved && grads && out, E3_ERR_BAD_ARG, "e3_render_bwd: null argument");
  E3_REQUIRE(p->batch >= 0 && p->height > 0 && p->width > 0 && p->res > 0, E3_ERR_BAD_ARG,
             "e3_render_bwd: bad geometry B=%d H=%d W=%d res=%d", p->batch, p->height, p->width, p->res);
  if (p->batch == 0) return E3_OK;
  E3_REQUIRE(!(p->flags & E3_RENDER_FP32_CUDA_CORES), E3_ERR_UNSUPPORTED,
             "e3_render_bwd: the backward pass exists for the tensor-core renderer only");
  E3_REQUIRE(p->n_samples >= 1 && p->n_samples <= 128, E3_ERR_UNSUPPORTED, "e3_render_bwd: n_samples=%d outside [1,128]",
             p->n_samples);
  E3_REQUIRE(in->cam_poses && in->focal && in->near && in->far && in->pix_x && in->pix_y && in->film,
             E3_ERR_BAD_ARG, "e3_render_bwd: missing camera / film input");
  E3_REQUIRE(in->t_vals || in->z_jitter, E3_ERR_BAD_ARG, "e3_render_bwd: need t_vals or z_jitter");
  E3_REQUIRE((p->flags & E3_RENDER_NO_SDF) || in->sigmoid_beta, E3_ERR_BAD_ARG, "e3_render_bwd: sigmoid_beta missing");
  E3_REQUIRE(saved->stash && saved->sdf && saved->hit_prob && saved->raw_rgb, E3_ERR_BAD_ARG,
             "e3_render_bwd: the forward call must have produced bwd_stash, sdf, hit_prob and raw_rgb");
  E3_REQUIRE(out->d_film, E3_ERR_BAD_ARG, "e3_render_bwd: d_film is required");
  E3_REQUIRE((in->local_alpha != nullptr) || (!out->d_local_alpha && !out->d_local_beta), E3_ERR_BAD_ARG,
             "e3_render_bwd: local modulation gradients requested without a local modulation input");
  E3_REQUIRE(scratch && scratch_bytes >= e3_render_bwd_scratch_bytes(p->batch), E3_ERR_SCRATCH,
             "e3_render_bwd: scratch too small (%zu < %zu)", scratch_bytes, e3_render_bwd_scratch_bytes(p->batch));
  RenderBwdArgs a{};
  a.packed = static_cast<const float*>(packed);
  a.p = *p;
  a.in = *in;
  a.stash = saved->stash;
  a.sdf = saved->sdf;
  a.hit_prob = saved->hit_prob;
  a.raw_rgb = saved->raw_rgb;
  a.d_features = grads->d_features;
  a.d_thumb_rgb = grads->d_thumb_rgb;
  a.d_xyz = grads->d_xyz;
  a.d_depth = grads->d_depth;
  a.d_sdf = grads->d_sdf;
  a.d_hit_prob = grads->d_hit_prob;
  a.film_partial = static_cast<float*>(scratch);
  a.d_local_alpha = out->d_local_alpha;
  a.d_local_beta = out->d_local_beta;
  a.d_points = out->d_points;
  a.rays_per_tile = 128 / p->n_samples;
  const int hw = p->height * p->width;
  a.tiles_per_image = (hw + a.rays_per_tile - 1) / a.rays_per_tile;
  a.n_tiles = a.tiles_per_image * p->batch;
  a.with_view = 1;
  cudaStream_t st = as_stream(stream);
  const int grid = render_bwd_grid(a.n_tiles);
  E3_CUDA(cudaMemsetAsync(scratch, 0, (size_t)grid * p->batch * BWD_STAT_FLOATS * sizeof(float), st));
  int rc = launch_render_bwd_tc(a, 0, st);
  if (rc) return rc;
  return finish_film_grads(a, grid, out->d_film, st);
}

extern "C" int e3_siren_points_bwd(const void* packed, const float* film, int batch, int n_points,
                                   float pts_scale, const float* stash, int with_view, int unit_sdf_seed,
                                   const float* d_sdf, const float* d_raw_rgb, const float* d_feat,
                                   float* d_film, float* d_points, void* scratch, size_t scratch_bytes,
                                   void* stream) {
  E3_REQUIRE(batch >= 0 && n_points >= 0, E3_ERR_BAD_ARG, "e3_siren_points_bwd: negative size");
  if (batch == 0) return E3_OK;
  E3_REQUIRE(packed && film && d_film, E3_ERR_BAD_ARG, "e3_siren_points_bwd: null argument");
  E3_REQUIRE(scratch && scratch_bytes >= e3_render_bwd_scratch_bytes(batch), E3_ERR_SCRATCH,
             "e3_siren_points_bwd: scratch too small (%zu < %zu)", scratch_bytes, e3_render_bwd_scratch_bytes(batch));
  E3_REQUIRE(with_view || (!d_raw_rgb && !d_feat), E3_ERR_BAD_ARG,
             "e3_siren_points_bwd: rgb / feature gradients need the view layer (with_view = 1)");
  RenderBwdArgs a{};
  a.packed = static_cast<const float*>(packed);
  a.p.batch = batch;
  a.p.pts_scale = pts_scale;
  a.p.n_samples = 1;
  a.in.film = film;
  a.stash = stash;
  a.n_points = n_points;
  a.d_sdf = d_sdf;
  a.unit_sdf_seed = unit_sdf_seed ? 1 : 0;  // eikonal: d(sdf)/d(point)
  a.d_prgb = d_raw_rgb;
  a.d_pfeat = d_feat;
  a.film_partial = static_cast<float*>(scratch);
  a.d_points = d_points;
  a.rays_per_tile = 128;
  a.tiles_per_image = (n_points + 127) / 128;
  a.n_tiles = a.tiles_per_image * batch;
  a.with_view = with_view ? 1 : 0;
  cudaStream_t st = as_stream(stream);
  const int grid = render_bwd_grid(a.n_tiles);
  E3_CUDA(cudaMemsetAsync(scratch, 0, (size_t)(grid > 0 ? grid : 0) * batch * BWD_STAT_FLOATS * sizeof(float), st));
  if (a.n_tiles > 0) {
    E3_REQUIRE(stash, E3_ERR_BAD_ARG, "e3_siren_points_bwd: stash missing");
    int rc = launch_render_bwd_tc(a, 1, st);
    if (rc) return rc;
  }
  return finish_film_grads(a, grid > 0 ? grid : 0, d_film, st);
}

extern "C" int e3_film_bwd(const void* packed, const float* d_film, int batch, int styles_per_image,
                           float* d_styles, void* stream) {
  E3_REQUIRE(batch >= 0 && (styles_per_image == 1 || styles_per_image == 9), E3_ERR_BAD_ARG,
             "e3_film_bwd: styles_per_image must be 1 (w) or 9 (w+), got %d", styles_per_image);
  if (batch == 0) return E3_OK;
  E3_REQUIRE(packed && d_film && d_styles, E3_ERR_BAD_ARG, "e3_film_bwd: null argument");
  film_bwd_kernel<<<batch * styles_per_image, 256, 0, as_stream(stream)>>>(static_cast<const float*>(packed), d_film,
                                                                          styles_per_image, d_styles);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}
