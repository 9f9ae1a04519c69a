// Per-sample MLP tail of the E3DGE local branch on the tensor cores (tcgen05), sm_100a.
//
// Reference arithmetic (per sample of the [B,H,W,S] volume; oracle/local_mlp_oracle.py restates it):
//   x      = [feat_2d (256) | visibility (1) | feat_3d (256)]                           e3dge_full_runner.py:285-290
//   e      = shortcut(x) + fc_1(relu(fc_0(relu(x))))           ResnetBlockFC(513 -> 256), resnetfc.py:53-62
//   scale  = L2s(lrelu_0.2(L1s(e))),  shift = L2t(lrelu_0.2(L1t(e)))                    sft.py:103-106
//   o      = feat_3d + feat_3d * scale + shift                                          sft.py:107-109
//   feats  = [o (256) | PE(point) (45)]      PE = (p, sin(2^k p), cos(2^k p)), k < 7    misc_utils.py:166-184
//   m      = shortcut(feats) + fc_1(relu(fc_0(relu(feats))))   ResnetBlockFC(301 -> 512)
//   alpha, beta = m[:256], m[256:]                                                      volume_renderer.py:327-336
// 989 161 MACs per sample — twice the SIREN — which the reference runs as ~15 ATen launches over
// [N,256]..[N,513] fp32 tensors.
//
// Here: six launches of ONE tcgen05 GEMM kernel (`tc_linear_kernel`, 128 x 128 output tiles, K-blocks of 64,
// operands split into bf16 hi + lo with the three products hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM,
// exactly the arithmetic of the decoder's convs) whose epilogues apply bias / ReLU / leaky ReLU / the SFT
// combination and write the NEXT stage's operands directly in their bf16 hi / lo form, so that no fp32
// activation is written between the stages:
//
//   prep     X  = split(x), RX = split(relu(x))  [rows,576];  PE columns of Y / RY             (CUDA cores)
//   stage 1  RNET = split(relu(RX * fc_0^T + b))                                       K = 576, N = 256
//   stage 2  E    = split(X * shortcut^T + RNET * fc_1^T + b)                           K = 576 + 256, N = 256
//   stage 3  U    = split(lrelu(E * [L1s; L1t]^T + b))                                  K = 256, N = 512
//   stage 4  (scale_c, shift_c) in adjacent accumulator columns (rows of L2s / L2t interleaved, each reading its
//            own half of U through zero blocks); epilogue o = d (1 + scale) + shift -> Y, RY = split(o), split(relu(o))
//                                                                                       K = 512, N = 512
//   stage 5  RNET2 = split(relu(RY * fc_0^T + b))                                       K = 320, N = 384 (301 padded)
//   stage 6  [alpha | beta] = Y * shortcut^T + RNET2 * fc_1^T + b   (fp32 out)          K = 320 + 320, N = 512
//
// Warp roles and pipeline are those of tc_conv_kernel (tc_conv.cu): warp 0 TMA producer (2-D tensor maps,
// 128B swizzle, out-of-range rows zero-filled), warp 1 TMEM owner + single-thread MMA issue, 8 epilogue warps,
// 3-stage 64 KB operand ring, two 128-column accumulators in ping-pong.
#include "tcgen05.cuh"

namespace e3 {

constexpr int LM_BM = 128, LM_BN = 128, LM_BK = 64, LM_STAGES = 3;
constexpr int LM_TILE_BYTES = 128 * 128;           // one operand tile: 128 rows x 64 bf16
constexpr int LM_STAGE_BYTES = 4 * LM_TILE_BYTES;  // A_hi, A_lo, B_hi, B_lo
// Epilogue staging (split / fp32 outputs): each group of 4 epilogue warps (one per TMEM lane quarter) owns a
// 128-row x 128-byte buffer in the TMA 128B-swizzled layout and hands it to a TMA tensor store — full-line
// writes instead of 32 lanes x 16 B to 32 different rows per instruction (which held the epilogue of the short-K
// stages at 3x their MMA time, profiles/r02_tc_linear_kernel.txt).
constexpr int LM_OUT_BYTES = LM_BM * 128;
constexpr int LM_SMEM_BYTES = LM_STAGES * LM_STAGE_BYTES + 2 * LM_OUT_BYTES + 256 + 1024;
constexpr int LM_EPI_WARPS = 8;
constexpr int LM_THREADS = 64 + LM_EPI_WARPS * 32;

// geometry of the chain
constexpr int LM_KX = 576;     // 513 padded
constexpr int LM_C = 256;
constexpr int LM_KU = 512;
constexpr int LM_KY = 320;     // 301 padded
constexpr int LM_NR2 = 384;    // 301 padded to whole n-tiles
constexpr int LM_FEATS = 301, LM_IN2D = 257, LM_PE = 45;

enum { LM_EPI_SPLIT = 0, LM_EPI_RELU = 1, LM_EPI_LRELU = 2, LM_EPI_SFT = 3, LM_EPI_F32 = 4, LM_EPI_F32_LD = 5 };

struct LinArgs {
  int M, N;        // rows of this chunk; padded output columns (multiple of 128)
  int nkb0, nkb1;  // k-blocks of the two A segments
  const float* bias;             // [N]
  __nv_bfloat16 *out_hi, *out_lo;    // [M, ldo]
  __nv_bfloat16 *out2_hi, *out2_lo;  // SFT: split(relu(o))
  int ldo;
  float *f32_a, *f32_b;          // F32: columns [0,256) -> f32_a, [256,512) -> f32_b, both [M,256]
  const float* gate_d;           // SFT: feat_3d [M,256]
  float* feats_out;              // SFT, optional: o as fp32 into [M,301]
};

// N (8, 16 or 32) consecutive values -> bf16 hi / lo, 16-byte stores
template <int N>
__device__ __forceinline__ void store_split(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const float* v) {
#pragma unroll
  for (int g = 0; g < N / 8; ++g) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pair_bf16(v[g * 8 + 2 * i], v[g * 8 + 2 * i + 1], h[i], l[i]);
    *reinterpret_cast<uint4*>(hi + off + g * 8) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off + g * 8) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

template <int EPI>
__global__ void __launch_bounds__(LM_THREADS, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmA0_hi, const __grid_constant__ CUtensorMap tmA0_lo,
                 const __grid_constant__ CUtensorMap tmA1_hi, const __grid_constant__ CUtensorMap tmA1_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ CUtensorMap tmO_a, const __grid_constant__ CUtensorMap tmO_b,
                 const __grid_constant__ LinArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* out_stage = smem + LM_STAGES * LM_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_stage + 2 * LM_OUT_BYTES);
  uint64_t* empty = full + LM_STAGES;
  uint64_t* acc_full = empty + LM_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_n = a.N / LM_BN;
  const int n_tiles = ((a.M + LM_BM - 1) / LM_BM) * n_tiles_n;
  const int nkb = a.nkb0 + a.nkb1;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tensormap(&tmA0_hi);
    tc::prefetch_tensormap(&tmA0_lo);
    tc::prefetch_tensormap(&tmA1_hi);
    tc::prefetch_tensormap(&tmA1_lo);
    tc::prefetch_tensormap(&tmB_hi);
    tc::prefetch_tensormap(&tmB_lo);
    if (EPI != LM_EPI_SFT) {
      tc::prefetch_tensormap(&tmO_a);
      tc::prefetch_tensormap(&tmO_b);
    }
#pragma unroll
    for (int s = 0; s < LM_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], LM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * LM_BN);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles_n) * LM_BM, n0 = (tile % n_tiles_n) * LM_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], LM_STAGE_BYTES);
          uint8_t* st = smem + stage * LM_STAGE_BYTES;
          if (kb < a.nkb0) {
            tc::tma_load_2d(st, &tmA0_hi, &full[stage], kb * LM_BK, m0);
            tc::tma_load_2d(st + LM_TILE_BYTES, &tmA0_lo, &full[stage], kb * LM_BK, m0);
          } else {
            tc::tma_load_2d(st, &tmA1_hi, &full[stage], (kb - a.nkb0) * LM_BK, m0);
            tc::tma_load_2d(st + LM_TILE_BYTES, &tmA1_lo, &full[stage], (kb - a.nkb0) * LM_BK, m0);
          }
          tc::tma_load_2d(st + 2 * LM_TILE_BYTES, &tmB_hi, &full[stage], kb * LM_BK, n0);
          tc::tma_load_2d(st + 3 * LM_TILE_BYTES, &tmB_lo, &full[stage], kb * LM_BK, n0);
          if (++stage == LM_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16_f32(LM_BM, LM_BN);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1, use = it >> 1;
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);  // the epilogue has drained this accumulator
        tc::fence_after_thread_sync();
        const uint32_t dcol = tmem_base + buf * LM_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc::fence_after_thread_sync();
          const uint32_t sb = smem_u32(smem + stage * LM_STAGE_BYTES);
          const uint64_t dA_hi = tc::make_smem_desc_sw128(sb);
          const uint64_t dA_lo = tc::make_smem_desc_sw128(sb + LM_TILE_BYTES);
          const uint64_t dB_hi = tc::make_smem_desc_sw128(sb + 2 * LM_TILE_BYTES);
          const uint64_t dB_lo = tc::make_smem_desc_sw128(sb + 3 * LM_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < LM_BK / 16; ++ks) {
            const uint64_t ah = tc::advance_desc_k(dA_hi, ks), al = tc::advance_desc_k(dA_lo, ks);
            const uint64_t bh = tc::advance_desc_k(dB_hi, ks), bl = tc::advance_desc_k(dB_lo, ks);
            tc::mma_bf16_ss(dcol, ah, bh, idesc, (kb | ks) != 0);
            tc::mma_bf16_ss(dcol, ah, bl, idesc, true);
            tc::mma_bf16_ss(dcol, al, bh, idesc, true);
          }
          tc::mma_commit(&empty[stage]);
          if (++stage == LM_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc::mma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===== epilogue warps: TMEM lanes [32q, 32q+32) belong to warp q = warp % 4 =====
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;  // which two of the tile's four 32-column chunks
    const int m = q * 32 + lane;
    uint8_t* obuf = out_stage + chalf * LM_OUT_BYTES;
    const bool store_leader = q == 0 && lane == 0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / n_tiles_n) * LM_BM, n0 = (tile % n_tiles_n) * LM_BN;
      const uint32_t buf = it & 1, use = it >> 1;
      const int64_t row = (int64_t)m0 + m;
      const bool valid = row < a.M;
      mbar_wait(&acc_full[buf], use & 1);
      tc::fence_after_thread_sync();
      uint32_t hw[2][16], lw[2][16];  // split modes: this thread's 64 outputs as bf16 hi / lo pairs
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int chunk = chalf * 2 + cc;
        float v[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * LM_BN + chunk * 32, v);
        if (cc == 1) {  // this warp's share is read: tell the MMA warp
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        const int nb = n0 + chunk * 32;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(a.bias + nb + j4 * 4);
          v[j4 * 4] += b4.x, v[j4 * 4 + 1] += b4.y, v[j4 * 4 + 2] += b4.z, v[j4 * 4 + 3] += b4.w;
        }
        if (EPI == LM_EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (EPI == LM_EPI_LRELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
        }
        if (EPI == LM_EPI_SPLIT || EPI == LM_EPI_RELU || EPI == LM_EPI_LRELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) split_pair_bf16(v[2 * i], v[2 * i + 1], hw[cc][i], lw[cc][i]);
        } else if (EPI == LM_EPI_SFT) {
          if (!valid) continue;
          const int c0 = nb >> 1;  // 16 (scale, shift) pairs -> channels [c0, c0 + 16)
          float o[16], ro[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 d4 = *reinterpret_cast<const float4*>(a.gate_d + (size_t)row * LM_C + c0 + j4 * 4);
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = j4 * 4 + i;
              // dec + (dec * scale + shift), in the reference's order of operations (sft.py:107-108)
              o[c] = __fadd_rn(dd[i], __fadd_rn(__fmul_rn(dd[i], v[2 * c]), v[2 * c + 1]));
              ro[c] = fmaxf(o[c], 0.f);
            }
          }
          store_split<16>(a.out_hi, a.out_lo, (size_t)row * a.ldo + c0, o);
          store_split<16>(a.out2_hi, a.out2_lo, (size_t)row * a.ldo + c0, ro);
          if (a.feats_out) {
            float* f = a.feats_out + (size_t)row * LM_FEATS + c0;
#pragma unroll
            for (int c = 0; c < 16; ++c) f[c] = o[c];
          }
        } else {
          // fp32 outputs: one 32-column chunk = one 128-byte row of the staging buffer -> TMA store
          // (LM_EPI_F32: columns [0,256) go to alpha through tmO_a, [256,512) to beta through tmO_b)
          tc::named_bar_sync(2 + chalf, 128);  // the group's previous store has finished reading the buffer
          tc::stage_row32(obuf, m, v);
          fence_proxy_async();
          tc::named_bar_sync(2 + chalf, 128);
          if (store_leader) {
            if (EPI == LM_EPI_F32 && nb >= LM_C) tc::tma_store_2d(&tmO_b, obuf, nb - LM_C, m0);
            else tc::tma_store_2d(&tmO_a, obuf, nb, m0);
            tc::tma_store_commit_and_wait_read();
          }
        }
      }
      if (EPI == LM_EPI_SPLIT || EPI == LM_EPI_RELU || EPI == LM_EPI_LRELU) {
        // 64 bf16 columns of this group = one 128-byte staging row: hi plane first, then lo through the same buffer
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          tc::named_bar_sync(2 + chalf, 128);
#pragma unroll
          for (int u = 0; u < 8; ++u) {  // 16-byte unit u = columns [8u, 8u+8) of the 64
            const uint32_t* w = plane ? lw[u >> 2] : hw[u >> 2];
            const int p = (u & 3) * 4;
            *reinterpret_cast<uint4*>(obuf + m * 128 + ((u ^ (m & 7)) << 4)) = make_uint4(w[p], w[p + 1], w[p + 2], w[p + 3]);
          }
          fence_proxy_async();
          tc::named_bar_sync(2 + chalf, 128);
          if (store_leader) {
            tc::tma_store_2d(plane ? &tmO_b : &tmO_a, obuf, n0 + chalf * 64, m0);
            tc::tma_store_commit_and_wait_read();
          }
        }
      }
    }
    if (store_leader) tc::tma_store_wait_all();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 2 * LM_BN);
}

// ---- operand preparation on the CUDA cores -----------------------------------------------------------
// one thread = 8 consecutive columns of one row
__device__ __forceinline__ void store_split8_dual(__nv_bfloat16* hi, __nv_bfloat16* lo, __nv_bfloat16* rhi,
                                                  __nv_bfloat16* rlo, size_t off, const float* v) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = fmaxf(v[i], 0.f);
  store_split<8>(hi, lo, off, v);
  store_split<8>(rhi, rlo, off, r);
}

// value of positional-encoding column `idx` (0..44) of point p — misc_utils.py:178-184
__device__ __forceinline__ float pe_value(const float* p, int idx) {
  if (idx < 3) return p[idx];
  const int k = (idx - 3) / 6, r = (idx - 3) - k * 6;
  const float arg = __fmul_rn((float)(1 << k), p[r < 3 ? r : r - 3]);
  return r < 3 ? sinf(arg) : cosf(arg);
}

// full chain: X / RX from (feat_2d | feat_3d), PE columns [256, 320) of Y / RY (and of the fp32 feats)
__global__ void __launch_bounds__(256) local_mlp_prep_kernel(const float* __restrict__ f2, const float* __restrict__ f3,
                                                             const float* __restrict__ pts, int64_t rows,
                                                             __nv_bfloat16* x_hi, __nv_bfloat16* x_lo,
                                                             __nv_bfloat16* rx_hi, __nv_bfloat16* rx_lo,
                                                             __nv_bfloat16* y_hi, __nv_bfloat16* y_lo,
                                                             __nv_bfloat16* ry_hi, __nv_bfloat16* ry_lo,
                                                             float* feats_out) {
  constexpr int GX = LM_KX / 8, GP = (LM_KY - LM_C) / 8, G = GX + GP;  // 72 + 8 groups per row
  const int64_t total = rows * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / G;
    const int g = (int)(i - row * G);
    float v[8];
    if (g < GX) {
      const int c0 = g * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        v[j] = c < LM_IN2D ? f2[row * LM_IN2D + c] : (c < LM_IN2D + LM_C ? f3[row * LM_C + (c - LM_IN2D)] : 0.f);
      }
      store_split8_dual(x_hi, x_lo, rx_hi, rx_lo, (size_t)row * LM_KX + c0, v);
    } else {
      const int c0 = (g - GX) * 8;
      const float p[3] = {pts[row * 3], pts[row * 3 + 1], pts[row * 3 + 2]};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < LM_PE) ? pe_value(p, c0 + j) : 0.f;
      store_split8_dual(y_hi, y_lo, ry_hi, ry_lo, (size_t)row * LM_KY + LM_C + c0, v);
      if (feats_out) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < LM_PE) feats_out[row * LM_FEATS + LM_C + c0 + j] = v[j];
      }
    }
  }
}

// texture-modulation MLP alone: Y / RY from caller-provided 301-d features
__global__ void __launch_bounds__(256) local_mlp_prep_feats_kernel(const float* __restrict__ feats, int64_t rows,
                                                                   __nv_bfloat16* y_hi, __nv_bfloat16* y_lo,
                                                                   __nv_bfloat16* ry_hi, __nv_bfloat16* ry_lo) {
  constexpr int G = LM_KY / 8;
  const int64_t total = rows * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / G;
    const int c0 = (int)(i - row * G) * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < LM_FEATS) ? feats[row * LM_FEATS + c0 + j] : 0.f;
    store_split8_dual(y_hi, y_lo, ry_hi, ry_lo, (size_t)row * LM_KY + c0, v);
  }
}

// ---- weight image --------------------------------------------------------------------------------------
// per stage: bf16 hi plane [N][K], bf16 lo plane [N][K] (K-major), then the fp32 bias [N]; zero where padded
struct StageGeom {
  int N, K;
};
__host__ __device__ constexpr StageGeom lm_stage(int s) {
  return s == 0 ? StageGeom{LM_C, LM_KX}
       : s == 1 ? StageGeom{LM_C, LM_KX + LM_C}
       : s == 2 ? StageGeom{2 * LM_C, LM_C}
       : s == 3 ? StageGeom{2 * LM_C, LM_KU}
       : s == 4 ? StageGeom{LM_NR2, LM_KY}
                : StageGeom{2 * LM_C, 2 * LM_KY};
}
static size_t lm_stage_bytes(int s) {
  const StageGeom g = lm_stage(s);
  return (size_t)g.N * g.K * 2 * 2 + (size_t)g.N * 4;
}
static size_t lm_stage_offset(int s) {
  size_t off = 0;
  for (int i = 0; i < s; ++i) off += (lm_stage_bytes(i) + 1023) / 1024 * 1024;
  return off;
}
struct StagePtrs {
  const __nv_bfloat16 *hi, *lo;
  const float* bias;
};
static StagePtrs lm_stage_ptrs(const void* packed, int s) {
  const StageGeom g = lm_stage(s);
  const uint8_t* p = reinterpret_cast<const uint8_t*>(packed) + lm_stage_offset(s);
  StagePtrs r;
  r.hi = reinterpret_cast<const __nv_bfloat16*>(p);
  r.lo = r.hi + (size_t)g.N * g.K;
  r.bias = reinterpret_cast<const float*>(r.lo + (size_t)g.N * g.K);
  return r;
}

// dst[(row0 + r * row_step) * ld + col0 + c] = split(src[r * cols + c])
__global__ void lm_place_block_kernel(const float* __restrict__ src, int rows, int cols, __nv_bfloat16* hi,
                                      __nv_bfloat16* lo, int ld, int row0, int row_step, int col0) {
  const int n = rows * cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / cols, c = i - r * cols;
    __nv_bfloat16 h, l;
    tc::split_bf16(src[i], h, l);
    const size_t o = (size_t)(row0 + r * row_step) * ld + col0 + c;
    hi[o] = h;
    lo[o] = l;
  }
}
__global__ void lm_place_bias_kernel(const float* __restrict__ src, int n, float* dst, int off, int step) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[off + i * step] = src[i];
}

static int lm_make_map(CUtensorMap* tm, const void* base, int cols, int64_t rows, int ld) {
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t str[1] = {(uint64_t)ld * 2};
  const uint32_t box[2] = {LM_BK, LM_BM};
  return make_tensor_map_bf16(tm, base, 2, dims, str, box);
}

struct Operand {  // one activation of the chain in operand form
  __nv_bfloat16 *hi, *lo;
  int cols, ld;
};

template <int EPI>
static int lm_launch_geom(const Operand& a0, const Operand* a1, const StagePtrs& w, StageGeom g, int stage,
                          int64_t rows, LinArgs args, cudaStream_t stream);
template <int EPI>
static int lm_launch(const Operand& a0, const Operand* a1, const StagePtrs& w, int stage, int64_t rows, LinArgs args,
                     cudaStream_t stream) {
  return lm_launch_geom<EPI>(a0, a1, w, lm_stage(stage), stage, rows, args, stream);
}
template <int EPI>
static int lm_launch_geom(const Operand& a0, const Operand* a1, const StagePtrs& w, StageGeom g, int stage,
                          int64_t rows, LinArgs args, cudaStream_t stream) {
  CUtensorMap mA0h, mA0l, mA1h, mA1l, mBh, mBl;
  int rc;
  if ((rc = lm_make_map(&mA0h, a0.hi, a0.cols, rows, a0.ld))) return rc;
  if ((rc = lm_make_map(&mA0l, a0.lo, a0.cols, rows, a0.ld))) return rc;
  const Operand& s1 = a1 ? *a1 : a0;
  if ((rc = lm_make_map(&mA1h, s1.hi, s1.cols, rows, s1.ld))) return rc;
  if ((rc = lm_make_map(&mA1l, s1.lo, s1.cols, rows, s1.ld))) return rc;
  if ((rc = lm_make_map(&mBh, w.hi, g.K, g.N, g.K))) return rc;
  if ((rc = lm_make_map(&mBl, w.lo, g.K, g.N, g.K))) return rc;
  // output maps of the staged TMA-store epilogue (rows past `rows` are clipped by the map)
  CUtensorMap mOa = mA0h, mOb = mA0l;
  if (EPI == LM_EPI_SPLIT || EPI == LM_EPI_RELU || EPI == LM_EPI_LRELU) {
    if ((rc = lm_make_map(&mOa, args.out_hi, args.ldo, rows, args.ldo))) return rc;
    if ((rc = lm_make_map(&mOb, args.out_lo, args.ldo, rows, args.ldo))) return rc;
  } else if (EPI == LM_EPI_F32 || EPI == LM_EPI_F32_LD) {
    const int ld = EPI == LM_EPI_F32 ? LM_C : args.ldo;
    const uint64_t dims[2] = {(uint64_t)ld, (uint64_t)rows};
    const uint64_t str[1] = {(uint64_t)ld * 4};
    const uint32_t box[2] = {32, LM_BM};
    if ((rc = make_tensor_map_f32(&mOa, args.f32_a, 2, dims, str, box))) return rc;
    if ((rc = make_tensor_map_f32(&mOb, EPI == LM_EPI_F32 ? args.f32_b : args.f32_a, 2, dims, str, box))) return rc;
  }
  args.M = (int)rows;
  args.N = g.N;
  args.nkb0 = a0.cols / LM_BK;
  args.nkb1 = a1 ? a1->cols / LM_BK : 0;
  args.bias = w.bias;
  E3_REQUIRE((args.nkb0 + args.nkb1) * LM_BK == g.K, E3_ERR_BAD_ARG, "local mlp: stage %d operand / weight K mismatch", stage);
  auto* fn = tc_linear_kernel<EPI>;
  static thread_local bool attr_set_dev[E3_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_set_dev[device_slot()];
  if (!attr_set) {
    E3_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM_BYTES));
    attr_set = true;
  }
  const int64_t n_tiles = ((rows + LM_BM - 1) / LM_BM) * (g.N / LM_BN);
  const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
  fn<<<grid, LM_THREADS, LM_SMEM_BYTES, stream>>>(mA0h, mA0l, mA1h, mA1l, mBh, mBl, mOa, mOb, args);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

constexpr int64_t LM_WS_BYTES_PER_ROW = (4 * LM_KX + 2 * LM_C + 2 * LM_C + 2 * LM_KU + 4 * LM_KY + 2 * LM_NR2) * 2;

}  // namespace e3

using namespace e3;

extern "C" size_t e3_local_mlp_packed_bytes(void) { return lm_stage_offset(6); }

extern "C" int e3_local_mlp_pack(const e3_local_mlp_weights* w, void* packed, void* stream_) {
  E3_REQUIRE(w && packed, E3_ERR_BAD_ARG, "e3_local_mlp_pack: null argument");
  // the texture-modulation MLP is required; the fusion MLP's 13 tensors come all or none (none = only the
  // `feats_in` form of e3_local_mlp_fwd is meaningful: stages 1-4 stay zero)
  const float* const* all = reinterpret_cast<const float* const*>(w);
  constexpr int kFuse = 13, kAll = (int)(sizeof(e3_local_mlp_weights) / sizeof(float*));
  int n_fuse = 0;
  for (int i = 0; i < kFuse; ++i) n_fuse += all[i] != nullptr;
  E3_REQUIRE(n_fuse == 0 || n_fuse == kFuse, E3_ERR_BAD_ARG, "e3_local_mlp_pack: %d of the 13 fusion tensors given", n_fuse);
  for (int i = kFuse; i < kAll; ++i)
    E3_REQUIRE(all[i] != nullptr, E3_ERR_BAD_ARG, "e3_local_mlp_pack: weight pointer %d is null", i);
  const bool fuse = n_fuse == kFuse;
  cudaStream_t stream = as_stream(stream_);
  E3_CUDA(cudaMemsetAsync(packed, 0, e3_local_mlp_packed_bytes(), stream));
  auto place = [&](int stage, const float* src, int rows, int cols, int row0, int row_step, int col0) {
    const StageGeom g = lm_stage(stage);
    const StagePtrs p = lm_stage_ptrs(packed, stage);
    lm_place_block_kernel<<<(rows * cols + 255) / 256, 256, 0, stream>>>(
        src, rows, cols, const_cast<__nv_bfloat16*>(p.hi), const_cast<__nv_bfloat16*>(p.lo), g.K, row0, row_step, col0);
  };
  auto bias = [&](int stage, const float* src, int n, int off, int step) {
    const StagePtrs p = lm_stage_ptrs(packed, stage);
    lm_place_bias_kernel<<<(n + 255) / 256, 256, 0, stream>>>(src, n, const_cast<float*>(p.bias), off, step);
  };
  const int KIN = LM_IN2D + LM_C;  // 513
  if (fuse) {
  place(0, w->enc_fc0_w, LM_C, KIN, 0, 1, 0);
  bias(0, w->enc_fc0_b, LM_C, 0, 1);
  place(1, w->enc_shortcut_w, LM_C, KIN, 0, 1, 0);
  place(1, w->enc_fc1_w, LM_C, LM_C, 0, 1, LM_KX);
  bias(1, w->enc_fc1_b, LM_C, 0, 1);
  place(2, w->scale0_w, LM_C, LM_C, 0, 1, 0);
  place(2, w->shift0_w, LM_C, LM_C, LM_C, 1, 0);
  bias(2, w->scale0_b, LM_C, 0, 1);
  bias(2, w->shift0_b, LM_C, LM_C, 1);
  place(3, w->scale2_w, LM_C, LM_C, 0, 2, 0);      // even rows read U[:, :256]
  place(3, w->shift2_w, LM_C, LM_C, 1, 2, LM_C);   // odd rows read U[:, 256:]
  bias(3, w->scale2_b, LM_C, 0, 2);
  bias(3, w->shift2_b, LM_C, 1, 2);
  }
  place(4, w->tex_fc0_w, LM_FEATS, LM_FEATS, 0, 1, 0);
  bias(4, w->tex_fc0_b, LM_FEATS, 0, 1);
  place(5, w->tex_shortcut_w, 2 * LM_C, LM_FEATS, 0, 1, 0);
  place(5, w->tex_fc1_w, 2 * LM_C, LM_FEATS, 0, 1, LM_KY);
  bias(5, w->tex_fc1_b, 2 * LM_C, 0, 1);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" size_t e3_local_mlp_workspace_bytes(int64_t rows) {
  if (rows <= 0) return 0;
  const int64_t padded = (rows + LM_BM - 1) / LM_BM * LM_BM;
  return (size_t)padded * LM_WS_BYTES_PER_ROW + 1024;
}

extern "C" int e3_local_mlp_fwd(const void* packed, const float* feat_2d, const float* feat_3d, const float* points,
                                const float* feats_in, int64_t rows, float* alpha, float* beta, float* feats_out,
                                void* workspace, size_t workspace_bytes, void* stream_) {
  E3_REQUIRE(packed && alpha && beta && rows >= 0, E3_ERR_BAD_ARG, "e3_local_mlp_fwd: null argument");
  const bool full = feats_in == nullptr;
  E3_REQUIRE(full ? (feat_2d && feat_3d && points) : !(feat_2d || feat_3d || points || feats_out), E3_ERR_BAD_ARG,
             "e3_local_mlp_fwd: pass either (feat_2d, feat_3d, points) or feats_in");
  if (rows == 0) return E3_OK;
  E3_REQUIRE(workspace && workspace_bytes >= e3_local_mlp_workspace_bytes(LM_BM), E3_ERR_SCRATCH,
             "e3_local_mlp_fwd: workspace too small (%zu bytes; see e3_local_mlp_workspace_bytes)", workspace_bytes);
  cudaStream_t stream = as_stream(stream_);
  // rows are processed in chunks that fit the workspace (each chunk's operands are dead after its six stages)
  int64_t cap = (int64_t)((workspace_bytes - 1024) / (size_t)LM_WS_BYTES_PER_ROW) / LM_BM * LM_BM;
  if (cap > (int64_t)1 << 24) cap = (int64_t)1 << 24;  // tensor-map row counts and int indices stay comfortable
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  ws += (1024 - (reinterpret_cast<uintptr_t>(ws) & 1023)) & 1023;
  auto carve = [&](int cols, int64_t cap_rows) {
    Operand o;
    o.hi = reinterpret_cast<__nv_bfloat16*>(ws);
    o.lo = o.hi + (size_t)cap_rows * cols;
    o.cols = o.ld = cols;
    ws += (size_t)cap_rows * cols * 4;
    return o;
  };
  Operand X = carve(LM_KX, cap), RX = carve(LM_KX, cap), RNET = carve(LM_C, cap), E = carve(LM_C, cap),
          U = carve(LM_KU, cap), Y = carve(LM_KY, cap), RY = carve(LM_KY, cap), RNET2 = carve(LM_NR2, cap);
  Operand RNET2k = RNET2;
  RNET2k.cols = LM_KY;  // stage 6 reads its first 320 columns
  StagePtrs W[6];
  for (int s = 0; s < 6; ++s) W[s] = lm_stage_ptrs(packed, s);
  for (int64_t r0 = 0; r0 < rows; r0 += cap) {
    const int64_t n = rows - r0 < cap ? rows - r0 : cap;
    int rc;
    LinArgs la{};
    if (full) {
      float* fo = feats_out ? feats_out + r0 * LM_FEATS : nullptr;
      const int64_t items = n * ((LM_KX + LM_KY - LM_C) / 8);
      const int64_t blocks = (items + 255) / 256, capb = (int64_t)sm_count() * 16;
      local_mlp_prep_kernel<<<(int)(blocks < capb ? blocks : capb), 256, 0, stream>>>(
          feat_2d + r0 * LM_IN2D, feat_3d + r0 * LM_C, points + r0 * 3, n, X.hi, X.lo, RX.hi, RX.lo, Y.hi, Y.lo,
          RY.hi, RY.lo, fo);
      E3_CUDA(cudaGetLastError());
      la = LinArgs{};
      la.out_hi = RNET.hi, la.out_lo = RNET.lo, la.ldo = RNET.ld;
      if ((rc = lm_launch<LM_EPI_RELU>(RX, nullptr, W[0], 0, n, la, stream))) return rc;
      la = LinArgs{};
      la.out_hi = E.hi, la.out_lo = E.lo, la.ldo = E.ld;
      if ((rc = lm_launch<LM_EPI_SPLIT>(X, &RNET, W[1], 1, n, la, stream))) return rc;
      la = LinArgs{};
      la.out_hi = U.hi, la.out_lo = U.lo, la.ldo = U.ld;
      if ((rc = lm_launch<LM_EPI_LRELU>(E, nullptr, W[2], 2, n, la, stream))) return rc;
      la = LinArgs{};
      la.out_hi = Y.hi, la.out_lo = Y.lo, la.out2_hi = RY.hi, la.out2_lo = RY.lo, la.ldo = Y.ld;
      la.gate_d = feat_3d + r0 * LM_C;
      la.feats_out = fo;
      if ((rc = lm_launch<LM_EPI_SFT>(U, nullptr, W[3], 3, n, la, stream))) return rc;
    } else {
      const int64_t items = n * (LM_KY / 8);
      const int64_t blocks = (items + 255) / 256, capb = (int64_t)sm_count() * 16;
      local_mlp_prep_feats_kernel<<<(int)(blocks < capb ? blocks : capb), 256, 0, stream>>>(
          feats_in + r0 * LM_FEATS, n, Y.hi, Y.lo, RY.hi, RY.lo);
      E3_CUDA(cudaGetLastError());
    }
    la = LinArgs{};
    la.out_hi = RNET2.hi, la.out_lo = RNET2.lo, la.ldo = RNET2.ld;
    if ((rc = lm_launch<LM_EPI_RELU>(RY, nullptr, W[4], 4, n, la, stream))) return rc;
    la = LinArgs{};
    la.f32_a = alpha + r0 * LM_C, la.f32_b = beta + r0 * LM_C;
    if ((rc = lm_launch<LM_EPI_F32>(Y, &RNET2k, W[5], 5, n, la, stream))) return rc;
  }
  return E3_OK;
}


// ---- generic y = x * W^T (+ bias) on the same kernel: the layer-wise SIREN sweeps of the second-order
// (eikonal) gradient use it (e3dge_b200/eikonal.py) -----------------------------------------------------
namespace e3 {
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, int64_t n_vec8,
                                                         __nv_bfloat16* hi, __nv_bfloat16* lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v0 = reinterpret_cast<const float4*>(x)[2 * i], v1 = reinterpret_cast<const float4*>(x)[2 * i + 1];
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    store_split<8>(hi, lo, (size_t)i * 8, v);
  }
}
}  // namespace e3

extern "C" size_t e3_tc_linear_packed_bytes(int n, int k) {
  if (n <= 0 || k <= 0) return 0;
  return (size_t)n * k * 4 + (size_t)n * 4;  // bf16 hi + lo planes [n][k], zero bias [n]
}

extern "C" int e3_tc_linear_pack(const float* w, int n, int k, void* packed, void* stream_) {
  E3_REQUIRE(w && packed && n > 0 && k > 0 && n % LM_BN == 0 && k % LM_BK == 0, E3_ERR_BAD_ARG,
             "e3_tc_linear_pack: needs n %% 128 == 0 and k %% 64 == 0 (got n=%d k=%d)", n, k);
  cudaStream_t stream = as_stream(stream_);
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(packed);
  __nv_bfloat16* lo = hi + (size_t)n * k;
  E3_CUDA(cudaMemsetAsync(lo + (size_t)n * k, 0, (size_t)n * 4, stream));
  lm_place_block_kernel<<<(n * k + 255) / 256, 256, 0, stream>>>(w, n, k, hi, lo, k, 0, 1, 0);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" size_t e3_tc_linear_workspace_bytes(int64_t rows, int k) {
  if (rows <= 0 || k <= 0) return 0;
  return (size_t)((rows + LM_BM - 1) / LM_BM * LM_BM) * k * 4 + 1024;
}

extern "C" int e3_tc_linear_fwd(const void* packed, int n, int k, const float* x, int64_t rows, const float* bias,
                                float* y, void* workspace, size_t workspace_bytes, void* stream_) {
  E3_REQUIRE(packed && x && y && rows >= 0 && n > 0 && k > 0 && n % LM_BN == 0 && k % LM_BK == 0, E3_ERR_BAD_ARG,
             "e3_tc_linear_fwd: bad argument (needs n %% 128 == 0, k %% 64 == 0)");
  if (rows == 0) return E3_OK;
  E3_REQUIRE(rows < ((int64_t)1 << 31) - LM_BM, E3_ERR_BAD_ARG, "e3_tc_linear_fwd: too many rows");
  E3_REQUIRE(workspace && workspace_bytes >= e3_tc_linear_workspace_bytes(rows, k), E3_ERR_SCRATCH,
             "e3_tc_linear_fwd: workspace too small");
  cudaStream_t stream = as_stream(stream_);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  ws += (1024 - (reinterpret_cast<uintptr_t>(ws) & 1023)) & 1023;
  const int64_t padded = (rows + LM_BM - 1) / LM_BM * LM_BM;
  Operand X;
  X.hi = reinterpret_cast<__nv_bfloat16*>(ws);
  X.lo = X.hi + (size_t)padded * k;
  X.cols = X.ld = k;
  const int64_t n_vec8 = rows * k / 8;
  const int64_t blocks = (n_vec8 + 255) / 256, capb = (int64_t)sm_count() * 16;
  split_rows_kernel<<<(int)(blocks < capb ? blocks : capb), 256, 0, stream>>>(x, n_vec8, X.hi, X.lo);
  E3_CUDA(cudaGetLastError());
  StagePtrs W;
  W.hi = reinterpret_cast<const __nv_bfloat16*>(packed);
  W.lo = W.hi + (size_t)n * k;
  W.bias = bias ? bias : reinterpret_cast<const float*>(W.lo + (size_t)n * k);
  LinArgs la{};
  la.f32_a = y;
  la.ldo = n;
  return lm_launch_geom<LM_EPI_F32_LD>(X, nullptr, W, StageGeom{n, k}, -1, rows, la, stream);
}
