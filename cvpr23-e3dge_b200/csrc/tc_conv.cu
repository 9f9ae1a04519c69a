// Tensor-core (tcgen05) implicit-GEMM for the modulated-conv decoder, sm_100a.
//
//   C[pixel, n] = sum_{tap, ci} xs[b, y+dy, x+dx, ci] * Wk[n, tap*Cin + ci]
//
// with xs = x * s[b,:] (the per-sample modulation moved onto the activations, see modconv.cu)
// and both operands split into bf16 hi + lo: three tcgen05.mma passes per k-step
// (hi*hi, hi*lo, lo*hi) accumulate in fp32 in TMEM, which keeps the decoder within ~1e-5 of
// the fp32 reference where plain bf16 operands would sit at ~1e-2 (north_star bar: 1e-3).
//
// Structure (persistent CTAs, one per SM, each walking 128-pixel x 128-channel output tiles; 10 warps):
//   warp 0      TMA producer: 4-D tiled tensor maps over the NHWC bf16 activations
//               (box = 64 ch x bw x bh x bb pixels, 128B swizzle, out-of-bounds = zero fill, which
//               *is* the conv's zero padding) and 2-D maps over the K-major weights; 3-stage
//               full/empty mbarrier ring, 64 KB per stage (A_hi, A_lo, B_hi, B_lo);
//   warp 1      TMEM allocation (2 x 128 columns, ping-pong) + single-thread MMA issue (12 UMMAs
//               128x128x16 per stage), tcgen05.commit releases the stage / signals the epilogue;
//   warps 2..9  epilogue: tcgen05.ld (two warps per 32-lane quarter, half the columns each), demodulation + noise +
//               bias + leaky-ReLU*sqrt(2) (StyledConv, stylesdf_model.py:494-507), fp32 NHWC store.
#include <stdlib.h>

#include "tcgen05.cuh"
#include "modconv.cuh"

namespace e3 {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

static int make_tensor_map_typed(CUtensorMap* tm, CUtensorMapDataType dtype, const void* base, int rank,
                                 const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                 bool swizzle128);
int make_tensor_map_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  return make_tensor_map_typed(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle128);
}
int make_tensor_map_f32(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  return make_tensor_map_typed(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle128);
}
static int make_tensor_map_typed(CUtensorMap* tm, CUtensorMapDataType dtype, const void* base, int rank,
                                 const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                 bool swizzle128) {
  PFN_encodeTiled enc = get_encode_tiled();
  E3_REQUIRE(enc != nullptr, E3_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = enc(tm, dtype, (cuuint32_t)rank, const_cast<void*>(base),
                   gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  E3_REQUIRE(r == CUDA_SUCCESS, E3_ERR_BAD_ARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return E3_OK;
}

// xs = x * s[b,:]  ->  bf16 hi / lo, NHWC
__global__ void __launch_bounds__(256) modulate_split_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ s,
                                                             __nv_bfloat16* __restrict__ hi,
                                                             __nv_bfloat16* __restrict__ lo,
                                                             int64_t n_vec8, int hw, int cin) {
  const int c8n = cin >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec8;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8n) * 8;
    const int64_t pix = i / c8n;
    const int b = (int)(pix / hw);
    const float4 v0 = *reinterpret_cast<const float4*>(x + pix * cin + c);
    const float4 v1 = *reinterpret_cast<const float4*>(x + pix * cin + c + 4);
    const float4 s0 = *reinterpret_cast<const float4*>(s + (size_t)b * cin + c);
    const float4 s1 = *reinterpret_cast<const float4*>(s + (size_t)b * cin + c + 4);
    const float v[8] = {v0.x * s0.x, v0.y * s0.y, v0.z * s0.z, v0.w * s0.w,
                        v1.x * s1.x, v1.y * s1.y, v1.z * s1.z, v1.w * s1.w};
    __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tc::split_bf16(v[j], h[j], l[j]);
    *reinterpret_cast<uint4*>(hi + pix * cin + c) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + pix * cin + c) = *reinterpret_cast<const uint4*>(l);
  }
}

// K-major bf16 hi/lo weights:  plain    Wk[o][tap*cin + ci]        = scale * W[o][ci][tap]
//                              upsample Wk[o][tq*cin + ci]         = scale * W[o][ci][kUpTapOrder[tq]]
//                                       (taps grouped by output parity phase, see tc_upconv_phase_kernel)
// backward (input gradients):  mode 2   Wk[ci][tap*cout + o]       = scale * W[o][ci][8 - tap]  (plain conv)
//                              mode 3   Wk[ci][tap*cout + o]       = scale * W[o][ci][tap]      (up-conv, planar gather)
// conv taps (ky*3 + kx) in parity-phase order: phase (py,px) = (0,0): taps with even ky and kx (4),
// (0,1): even ky, kx = 1 (2), (1,0): ky = 1, even kx (2), (1,1): the centre tap
__constant__ int kUpTapOrder[9] = {0, 2, 6, 8, 1, 7, 3, 5, 4};

__global__ void conv_pack_bf16_kernel(const float* __restrict__ w, int cout, int cin, int upsample,
                                      float scale, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo) {
  const int64_t total = upsample == 4 ? (int64_t)64 * 576 : (int64_t)cout * cin * 9;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int o, ci, tap;
    if (upsample == 4) {
      // x-pair view of a 32 -> 32 conv: rows n' = p_out*32 + co, columns k' = tap'*64 + p_in*32 + ci with
      // tap' = (dy+1)*3 + (pair shift + 1); the entry is W[co][ci][dy+1][kx] for kx = 2*shift + p_in - p_out + 1
      // when that is a tap of the 3x3 kernel, else zero
      const int kq = (int)(idx % 576), np = (int)(idx / 576);
      const int tq = kq >> 6, r = kq & 63, p_in = r >> 5, p_out = np >> 5;
      ci = r & 31, o = np & 31;
      const int kx = 2 * (tq % 3 - 1) + p_in - p_out + 1, ky = tq / 3;
      __nv_bfloat16 h = __float2bfloat16_rn(0.f), l = h;
      if (kx >= 0 && kx <= 2) tc::split_bf16(scale * w[((size_t)o * 32 + ci) * 9 + ky * 3 + kx], h, l);
      hi[idx] = h;
      lo[idx] = l;
      continue;
    }
    if (upsample >= 2) {
      const int k = (int)(idx % ((int64_t)9 * cout));
      ci = (int)(idx / ((int64_t)9 * cout));
      o = k % cout;
      tap = k / cout;
      if (upsample == 2) tap = 8 - tap;
    } else if (upsample) {
      const int k = (int)(idx % ((int64_t)9 * cin));
      o = (int)(idx / ((int64_t)9 * cin));
      ci = k % cin;
      tap = kUpTapOrder[k / cin];
    } else {
      const int k = (int)(idx % ((int64_t)9 * cin));
      o = (int)(idx / ((int64_t)9 * cin));
      ci = k % cin;
      tap = k / cin;
    }
    __nv_bfloat16 h, l;
    tc::split_bf16(scale * w[((size_t)o * cin + ci) * 9 + tap], h, l);
    hi[idx] = h;
    lo[idx] = l;
  }
}

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 3;
constexpr int TC_TILE_BYTES = 128 * 128;            // one operand tile: 128 rows x 128 B
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;   // A_hi, A_lo, B_hi, B_lo
// Epilogue staging: each group of 4 epilogue warps (one per TMEM lane quarter) owns a 128-row x 32-column
// fp32 buffer in the TMA 128B-swizzled layout; one thread of the group hands it to a TMA tensor store.
// (Each thread storing its own row, 32 lanes x 16 B to 32 different lines per instruction, held the
// epilogue at ~8k cycles per tile and stalled on the store queue; full-line TMA writes do not.)
constexpr int TC_OUT_CHUNK_BYTES = TC_BM * 32 * 4;
constexpr int TC_OUT_GROUPS = 2;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_OUT_GROUPS * TC_OUT_CHUNK_BYTES + 256 + 1024;
constexpr int TC_EPI_WARPS = 8;  // two per TMEM lane quarter, each draining half of the tile's columns
constexpr int TC_THREADS = 64 + TC_EPI_WARPS * 32;

struct TcTile {
  int bw, bh, bb, tiles_x, tiles_y, tiles_b;
};

// Persistent: each CTA walks tiles t = blockIdx.x, +gridDim.x, ... (n-tile fastest, so neighbouring
// CTAs share the activation tile in L2).  Two 128-column TMEM accumulators ping-pong: the MMAs of
// tile i+1 run while the epilogue warps drain tile i.
constexpr int TC_ACC_BUFS = 2;

template <int TAPS>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvGemmArgs a,
               const __grid_constant__ TcTile t) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* out_stage = smem + TC_STAGES * TC_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_stage + TC_OUT_GROUPS * TC_OUT_CHUNK_BYTES);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* acc_full = empty + TC_STAGES;
  uint64_t* acc_empty = acc_full + TC_ACC_BUFS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + TC_ACC_BUFS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // narrow layers (N = 64 / 32: the 512^2 / 1024^2 stages of a size-1024 decoder): one n-tile whose MMAs are
  // issued N columns wide and whose weight box holds N rows; accumulator columns past N do not exist
  const int nbox = a.N < TC_BN ? a.N : TC_BN;
  const uint32_t stage_tx = 2 * TC_TILE_BYTES + 2 * nbox * 128;
  const int n_tiles_n = (a.N + TC_BN - 1) / TC_BN;
  const int n_tiles = t.tiles_x * t.tiles_y * t.tiles_b * n_tiles_n;
  const int kpt = a.Cin / TC_BK, nkb = TAPS * kpt;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tensormap(&tmA_hi);
    tc::prefetch_tensormap(&tmA_lo);
    tc::prefetch_tensormap(&tmB_hi);
    tc::prefetch_tensormap(&tmB_lo);
    tc::prefetch_tensormap(&tmOut);
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
#pragma unroll
    for (int s = 0; s < TC_ACC_BUFS; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], TC_EPI_WARPS);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TC_ACC_BUFS * TC_BN);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int tile, int& x0, int& y0, int& b0, int& n0) {
    const int nt = tile % n_tiles_n, mt = tile / n_tiles_n;
    const int tx = mt % t.tiles_x, ty = (mt / t.tiles_x) % t.tiles_y, tb = mt / (t.tiles_x * t.tiles_y);
    x0 = tx * t.bw, y0 = ty * t.bh, b0 = tb * t.bb, n0 = nt * TC_BN;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int x0, y0, b0, n0;
        tile_coords(tile, x0, y0, b0, n0);
        for (int kb = 0; kb < nkb; ++kb) {
          const int tap = kb / kpt, kc = kb - tap * kpt;
          int dx = (TAPS == 9) ? tap % 3 - 1 : 0, dy = (TAPS == 9) ? tap / 3 - 1 : 0, bc = b0;
          if (TAPS == 9 && a.planar) {  // tap (ky,kx) -> plane (ky&1, kx&1), shift (ky>>1, kx>>1)
            const int ky = tap / 3, kx = tap % 3;
            dx = kx >> 1, dy = ky >> 1, bc = ((ky & 1) * 2 + (kx & 1)) * a.B + b0;
          }
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], stage_tx);
          uint8_t* st = smem + stage * TC_STAGE_BYTES;
          tc::tma_load_4d(st, &tmA_hi, &full[stage], kc * TC_BK, x0 + dx, y0 + dy, bc);
          tc::tma_load_4d(st + TC_TILE_BYTES, &tmA_lo, &full[stage], kc * TC_BK, x0 + dx, y0 + dy, bc);
          tc::tma_load_2d(st + 2 * TC_TILE_BYTES, &tmB_hi, &full[stage], tap * a.Cin + kc * TC_BK, n0);
          tc::tma_load_2d(st + 3 * TC_TILE_BYTES, &tmB_lo, &full[stage], tap * a.Cin + kc * TC_BK, n0);
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16_f32(TC_BM, nbox);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1, use = it >> 1;
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);  // epilogue has drained this accumulator
        tc::fence_after_thread_sync();
        const uint32_t dcol = tmem_base + buf * TC_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc::fence_after_thread_sync();
          const uint32_t sb = smem_u32(smem + stage * TC_STAGE_BYTES);
          const uint64_t dA_hi = tc::make_smem_desc_sw128(sb);
          const uint64_t dA_lo = tc::make_smem_desc_sw128(sb + TC_TILE_BYTES);
          const uint64_t dB_hi = tc::make_smem_desc_sw128(sb + 2 * TC_TILE_BYTES);
          const uint64_t dB_lo = tc::make_smem_desc_sw128(sb + 3 * TC_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < TC_BK / 16; ++ks) {
            const uint64_t ah = tc::advance_desc_k(dA_hi, ks), al = tc::advance_desc_k(dA_lo, ks);
            const uint64_t bh = tc::advance_desc_k(dB_hi, ks), bl = tc::advance_desc_k(dB_lo, ks);
            tc::mma_bf16_ss(dcol, ah, bh, idesc, (kb | ks) != 0);
            tc::mma_bf16_ss(dcol, ah, bl, idesc, true);
            tc::mma_bf16_ss(dcol, al, bh, idesc, true);
          }
          tc::mma_commit(&empty[stage]);  // the stage is free once these MMAs have read it
          if (++stage == TC_STAGES) {
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
    const int chalf = (warp - 2) >> 2;  // which half of the tile's 32-column chunks this warp drains
    const int m = q * 32 + lane;
    const int ix = m % t.bw, iy = (m / t.bw) % t.bh, ib = m / (t.bw * t.bh);
    const float nw = (a.mode == 1) ? a.noise_w[0] : 0.f;
    uint8_t* obuf = out_stage + chalf * TC_OUT_CHUNK_BYTES;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      int x0, y0, b0, n0;
      tile_coords(tile, x0, y0, b0, n0);
      const uint32_t buf = it & 1, use = it >> 1;
      const int b = b0 + ib, y = y0 + iy, x = x0 + ix;
      const bool valid = b < a.B;
      const int p = a.pairx ? 2 * (y * a.W + x) : y * a.W + x;  // pair view: first of the row's two pixels
      const float nz0 = (a.mode == 1 && valid) ? nw * a.noise[(size_t)b * a.noise_bstride + p] : 0.f;
      const float nz1 = (a.mode == 1 && valid && a.pairx) ? nw * a.noise[(size_t)b * a.noise_bstride + p + 1] : 0.f;
      mbar_wait(&acc_full[buf], use & 1);
      tc::fence_after_thread_sync();
      constexpr int kChunksPerWarp = TC_BN / 32 / (TC_EPI_WARPS / 4);
#pragma unroll 1
      for (int chunk = chalf * kChunksPerWarp; chunk < (chalf + 1) * kChunksPerWarp; ++chunk) {
        float v[32];
        const bool live = chunk * 32 < nbox;  // (uniform over the warp group)
        if (live) tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TC_BN + chunk * 32, v);
        if (chunk == (chalf + 1) * kChunksPerWarp - 1) {  // this warp's share is read: tell the MMA warp
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (!live) continue;
        const int nb = n0 + chunk * 32;
        const float nz = (a.pairx && (chunk & 1)) ? nz1 : nz0;  // pair view: columns [32,64) are the odd pixel
        if (valid && a.mode == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float tt = fmaf(v[j], a.d[(size_t)b * a.N + nb + j], nz) + a.act_bias[nb + j];
            v[j] = (tt > 0.f ? tt : 0.2f * tt) * 1.41421356237309515f;
          }
        } else if (valid && a.mode == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= a.d[(size_t)b * a.N + nb + j];
        }
        // registers -> swizzled staging buffer of this warp group -> one TMA tensor store (rows of images
        // past the batch are clipped by the tensor map)
        tc::named_bar_sync(2 + chalf, 128);  // the group's previous store has finished reading the buffer
        tc::stage_row32(obuf, m, v);
        fence_proxy_async();
        tc::named_bar_sync(2 + chalf, 128);
        if (q == 0 && lane == 0) {
          tc::tma_store_4d(&tmOut, obuf, nb, x0, y0, b0);
          tc::tma_store_commit_and_wait_read();
        }
      }
    }
    if (q == 0 && lane == 0) tc::tma_store_wait_all();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TC_ACC_BUFS * TC_BN);
}

// ---- CTA-pair variant of the plain conv (tcgen05 cta_group::2) ---------------------------------
// One MMA stream per pair of CTAs: M = 256 = the two CTAs' 128-pixel tiles (consecutive m-tiles), N = NT
// output channels, each CTA holding NT/2 rows of every weight k-block.  Per SM and k-block this halves the
// weight bytes pulled from L2 and read from shared memory; with NT = 256 a k-block also lasts twice as
// long (12 x 128 cycles), so the three stages cover the TMA latency that starves the single-CTA kernel
// (75 % tensor-pipe active with no memory unit above 60 %, profiles/r01l_conv_tc_kernel.txt).  (NT = 128
// with four 48 KB stages was measured too: no faster than the single-CTA kernel, not instantiated.)  Barrier wiring as in the paired render kernel:
// both CTAs' TMA loads are counted on the leader's `full`, tcgen05.commit multicasts `empty` / `acc_full`
// to both, the peer's epilogue warps release the accumulator on the leader's `acc_empty` with a relaxed
// remote arrive.
template <int NT>
struct PairCfg {
  static constexpr int kBHalfBytes = (NT / 2) * 128;                 // one of hi / lo: NT/2 rows x 64 bf16
  static constexpr int kStageBytes = 2 * TC_TILE_BYTES + 2 * kBHalfBytes;
  static constexpr int kStages = NT == 256 ? 3 : 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + TC_OUT_GROUPS * TC_OUT_CHUNK_BYTES + 256 + 1024;
};

template <int TAPS, int NT>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvGemmArgs a,
                    const __grid_constant__ TcTile t) {
  using Cfg = PairCfg<NT>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* out_stage = smem + STAGES * Cfg::kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_stage + TC_OUT_GROUPS * TC_OUT_CHUNK_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + TC_ACC_BUFS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + TC_ACC_BUFS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_tiles_n = a.N / NT;
  const int m_tiles = t.tiles_x * t.tiles_y * t.tiles_b;
  const int n_pair_tiles = ((m_tiles + 1) / 2) * n_tiles_n;
  const int n_pairs = (int)gridDim.x / 2, pair = (int)blockIdx.x / 2;
  const int kpt = a.Cin / TC_BK, nkb = TAPS * kpt;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tensormap(&tmA_hi);
    tc::prefetch_tensormap(&tmA_lo);
    tc::prefetch_tensormap(&tmB_hi);
    tc::prefetch_tensormap(&tmB_lo);
    tc::prefetch_tensormap(&tmOut);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
#pragma unroll
    for (int s = 0; s < TC_ACC_BUFS; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 2 * TC_EPI_WARPS);  // the epilogue warps of both CTAs (leader's barrier is used)
    }
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 1) tc::tmem_alloc_pair(tmem_slot, TC_ACC_BUFS * NT);
  tc::fence_before_thread_sync();
  __syncthreads();
  cluster_sync_all();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;

  // pair tile -> this CTA's pixel box (m-tile 2*pt + rank; past the last one = all out of range: the loads
  // zero-fill, the stores clip) and first channel
  auto tile_coords = [&](int tile, int& x0, int& y0, int& b0, int& n0) {
    const int nt = tile % n_tiles_n, mt = (tile / n_tiles_n) * 2 + (int)rank;
    const int tx = mt % t.tiles_x, ty = (mt / t.tiles_x) % t.tiles_y, tb = mt / (t.tiles_x * t.tiles_y);
    x0 = tx * t.bw, y0 = ty * t.bh, b0 = tb * t.bb, n0 = nt * NT;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = pair; tile < n_pair_tiles; tile += n_pairs) {
        int x0, y0, b0, n0;
        tile_coords(tile, x0, y0, b0, n0);
        for (int kb = 0; kb < nkb; ++kb) {
          const int tap = kb / kpt, kc = kb - tap * kpt;
          int dx = (TAPS == 9) ? tap % 3 - 1 : 0, dy = (TAPS == 9) ? tap / 3 - 1 : 0, bc = b0;
          if (TAPS == 9 && a.planar) {
            const int ky = tap / 3, kx = tap % 3;
            dx = kx >> 1, dy = ky >> 1, bc = ((ky & 1) * 2 + (kx & 1)) * a.B + b0;
          }
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
          const uint32_t lfull = tc::map_to_cta(&full[stage], 0);
          uint8_t* st = smem + stage * Cfg::kStageBytes;
          tc::tma_load_4d_pair(st, &tmA_hi, lfull, kc * TC_BK, x0 + dx, y0 + dy, bc);
          tc::tma_load_4d_pair(st + TC_TILE_BYTES, &tmA_lo, lfull, kc * TC_BK, x0 + dx, y0 + dy, bc);
          const int brow = n0 + (int)rank * (NT / 2);
          tc::tma_load_2d_pair(st + 2 * TC_TILE_BYTES, &tmB_hi, lfull, tap * a.Cin + kc * TC_BK, brow);
          tc::tma_load_2d_pair(st + 2 * TC_TILE_BYTES + Cfg::kBHalfBytes, &tmB_lo, lfull, tap * a.Cin + kc * TC_BK, brow);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      const uint32_t idesc = tc::make_idesc_bf16_f32(256, NT);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = pair; tile < n_pair_tiles; tile += n_pairs, ++it) {
        const uint32_t buf = it & 1, use = it >> 1;
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc::fence_after_thread_sync();
        const uint32_t dcol = tmem_base + buf * NT;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc::fence_after_thread_sync();
          const uint32_t sb = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t dA_hi = tc::make_smem_desc_sw128(sb);
          const uint64_t dA_lo = tc::make_smem_desc_sw128(sb + TC_TILE_BYTES);
          const uint64_t dB_hi = tc::make_smem_desc_sw128(sb + 2 * TC_TILE_BYTES);
          const uint64_t dB_lo = tc::make_smem_desc_sw128(sb + 2 * TC_TILE_BYTES + Cfg::kBHalfBytes);
#pragma unroll
          for (int ks = 0; ks < TC_BK / 16; ++ks) {
            const uint64_t ah = tc::advance_desc_k(dA_hi, ks), al = tc::advance_desc_k(dA_lo, ks);
            const uint64_t bh = tc::advance_desc_k(dB_hi, ks), bl = tc::advance_desc_k(dB_lo, ks);
            tc::mma_bf16_ss_pair(dcol, ah, bh, idesc, (kb | ks) != 0);
            tc::mma_bf16_ss_pair(dcol, ah, bl, idesc, true);
            tc::mma_bf16_ss_pair(dcol, al, bh, idesc, true);
          }
          tc::mma_commit_pair(&empty[stage], 3);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc::mma_commit_pair(&acc_full[buf], 3);
      }
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int ix = m % t.bw, iy = (m / t.bw) % t.bh, ib = m / (t.bw * t.bh);
    const float nw = (a.mode == 1) ? a.noise_w[0] : 0.f;
    uint8_t* obuf = out_stage + chalf * TC_OUT_CHUNK_BYTES;
    uint32_t it = 0;
    for (int tile = pair; tile < n_pair_tiles; tile += n_pairs, ++it) {
      int x0, y0, b0, n0;
      tile_coords(tile, x0, y0, b0, n0);
      const uint32_t buf = it & 1, use = it >> 1;
      const int b = b0 + ib, y = y0 + iy, x = x0 + ix;
      const bool valid = b < a.B;
      const int p = y * a.W + x;
      const float nz = (a.mode == 1 && valid) ? nw * a.noise[(size_t)b * a.noise_bstride + p] : 0.f;
      mbar_wait(&acc_full[buf], use & 1);
      tc::fence_after_thread_sync();
      constexpr int kChunksPerWarp = NT / 32 / (TC_EPI_WARPS / 4);
#pragma unroll 1
      for (int chunk = chalf * kChunksPerWarp; chunk < (chalf + 1) * kChunksPerWarp; ++chunk) {
        float v[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * NT + chunk * 32, v);
        if (chunk == (chalf + 1) * kChunksPerWarp - 1) {  // this warp's share is read: tell the leader's MMA thread
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&acc_empty[buf]);
            else tc::mbar_arrive_cluster_relaxed(tc::map_to_cta(&acc_empty[buf], 0));
          }
        }
        const int nb = n0 + chunk * 32;
        if (valid && a.mode == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float tt = fmaf(v[j], a.d[(size_t)b * a.N + nb + j], nz) + a.act_bias[nb + j];
            v[j] = (tt > 0.f ? tt : 0.2f * tt) * 1.41421356237309515f;
          }
        } else if (valid && a.mode == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= a.d[(size_t)b * a.N + nb + j];
        }
        tc::named_bar_sync(2 + chalf, 128);
        tc::stage_row32(obuf, m, v);
        fence_proxy_async();
        tc::named_bar_sync(2 + chalf, 128);
        if (q == 0 && lane == 0) {
          tc::tma_store_4d(&tmOut, obuf, nb, x0, y0, b0);
          tc::tma_store_commit_and_wait_read();
        }
      }
    }
    if (q == 0 && lane == 0) tc::tma_store_wait_all();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  cluster_sync_all();  // the peer's TMEM / barriers stay valid until both CTAs are done
  if (warp == 1) tc::tmem_dealloc_pair(tmem_base, TC_ACC_BUFS * NT);
}

template <int TAPS, int NT>
static int launch_conv_pair(const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_hi,
                            const CUtensorMap& tmB_lo, const CUtensorMap& tmOut, const ConvGemmArgs& a,
                            const TcTile& t, cudaStream_t stream) {
  auto* fn = tc_conv_pair_kernel<TAPS, NT>;
  static thread_local bool attr_set_dev[E3_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_set_dev[device_slot()];
  if (!attr_set) {
    E3_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<NT>::kSmemBytes));
    attr_set = true;
  }
  const int m_tiles = t.tiles_x * t.tiles_y * t.tiles_b;
  const int n_pair_tiles = ((m_tiles + 1) / 2) * (a.N / NT);
  int pairs = sm_count() / 2;
  if (pairs > n_pair_tiles) pairs = n_pair_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = PairCfg<NT>::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent kernel: no more pairs than can be co-resident as clusters
  static thread_local int max_clusters_dev[E3_MAX_DEVICES] = {};
  int& max_clusters = max_clusters_dev[device_slot()];
  if (!max_clusters) {
    cfg.gridDim = dim3(2 * (sm_count() / 2));
    int n = 0;
    E3_CUDA(cudaOccupancyMaxActiveClusters(&n, fn, &cfg));
    max_clusters = n > 0 ? n : 1;
  }
  if (pairs > max_clusters) pairs = max_clusters;
  cfg.gridDim = dim3(2 * pairs);
  E3_CUDA(cudaLaunchKernelEx(&cfg, fn, tmA_hi, tmA_lo, tmB_hi, tmB_lo, tmOut, a, t));
  return E3_OK;
}

// E3DGE_CONV_PAIR=0 selects the single-CTA kernel (measurement aid)
static bool conv_pairs_enabled() {
  const char* e = getenv("E3DGE_CONV_PAIR");
  return !(e && e[0] == '0');
}

// ---- upsampling conv on the tensor cores: parity-phase formulation ------------------------------
// conv_transpose2d(stride 2, 3x3) output T[P,Q] (P in [0,2H], Q in [0,2W]) splits by the parity of
// (P,Q) into four dense convolutions of the INPUT grid with 4 / 2 / 2 / 1 taps:
//   T[2i+py, 2j+px] = sum_{ky = py (mod 2), kx = px (mod 2)} xs[i - (ky>>1), j - (kx>>1)] * W[ky,kx]
// — the same 9*Cin*Cout MACs per input pixel as the G = xs*W GEMM, but the taps accumulate in TMEM, so
// what reaches HBM is T (4 values per input pixel and channel) instead of G (9), and the consumer is a
// plain 4x4 blur stencil instead of a 49-vector gather.
// Activations are laid out with one zero column after every row and one zero row after every image
// ([B][H+1][W+1][Cin]): a tap shift is then a 1-D shift of the flat pixel index m (the pad supplies the
// zero neighbour across row / image borders, TMA zero-fills m < 0 and m >= Mp), the even phases' extra
// row i = H / column j = W falls out of the same grid, and the tile is just 128 consecutive m.
struct UpPhaseArgs {
  float* t_out;  // [4 phases][Mp][N]
  int Mp, N, Cin, Wp;
};

__global__ void __launch_bounds__(256) modulate_split_padded_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ s,
                                                                    __nv_bfloat16* __restrict__ hi,
                                                                    __nv_bfloat16* __restrict__ lo,
                                                                    int64_t n_vec8, int H, int W, int cin) {
  const int c8n = cin >> 3;
  const int Wp = W + 1, Hp = H + 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec8;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8n) * 8;
    const int64_t pp = i / c8n;  // padded pixel index
    const int jx = (int)(pp % Wp), iy = (int)((pp / Wp) % Hp), b = (int)(pp / ((int64_t)Wp * Hp));
    uint4 h4 = make_uint4(0, 0, 0, 0), l4 = h4;
    if (jx < W && iy < H) {
      const int64_t pix = ((int64_t)b * H + iy) * W + jx;
      const float4 v0 = *reinterpret_cast<const float4*>(x + pix * cin + c);
      const float4 v1 = *reinterpret_cast<const float4*>(x + pix * cin + c + 4);
      const float4 s0 = *reinterpret_cast<const float4*>(s + (size_t)b * cin + c);
      const float4 s1 = *reinterpret_cast<const float4*>(s + (size_t)b * cin + c + 4);
      const float v[8] = {v0.x * s0.x, v0.y * s0.y, v0.z * s0.z, v0.w * s0.w,
                          v1.x * s1.x, v1.y * s1.y, v1.z * s1.z, v1.w * s1.w};
      __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tc::split_bf16(v[j], h[j], l[j]);
      h4 = *reinterpret_cast<const uint4*>(h);
      l4 = *reinterpret_cast<const uint4*>(l);
    }
    *reinterpret_cast<uint4*>(hi + pp * cin + c) = h4;
    *reinterpret_cast<uint4*>(lo + pp * cin + c) = l4;
  }
}

// Same pipeline as tc_conv_kernel (TMA producer warp, single-thread MMA issuer, 8 epilogue warps,
// 3-stage ring, ping-pong TMEM accumulators); a tile = (128 consecutive padded pixels) x (128 output
// channels) x (one parity phase).  Phases cost 4:2:2:1, so the phase of a tile is rotated by the
// CTA's round number (gridDim.x is a multiple of 4 * n-tiles): every CTA walks all four phases.
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_upconv_phase_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                       const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                       const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ UpPhaseArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* out_stage = smem + TC_STAGES * TC_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_stage + TC_OUT_GROUPS * TC_OUT_CHUNK_BYTES);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* acc_full = empty + TC_STAGES;
  uint64_t* acc_empty = acc_full + TC_ACC_BUFS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + TC_ACC_BUFS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbox = a.N < TC_BN ? a.N : TC_BN;  // narrow layers: see tc_conv_kernel
  const uint32_t stage_tx = 2 * TC_TILE_BYTES + 2 * nbox * 128;
  const int n_tiles_n = (a.N + TC_BN - 1) / TC_BN;
  const int m_tiles = (a.Mp + TC_BM - 1) / TC_BM;
  const int n_tiles = 4 * m_tiles * n_tiles_n;
  const int kpt = a.Cin / TC_BK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tensormap(&tmA_hi);
    tc::prefetch_tensormap(&tmA_lo);
    tc::prefetch_tensormap(&tmB_hi);
    tc::prefetch_tensormap(&tmB_lo);
    tc::prefetch_tensormap(&tmOut);
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
#pragma unroll
    for (int s = 0; s < TC_ACC_BUFS; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], TC_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TC_ACC_BUFS * TC_BN);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (phase, first padded pixel, first channel); n fastest, then the phase slot, then m
  auto tile_coords = [&](int tile, int& phase, int& m0, int& n0) {
    const int nt = tile % n_tiles_n, rest = tile / n_tiles_n;
    phase = ((rest & 3) + tile / (int)gridDim.x) & 3;
    m0 = (rest >> 2) * TC_BM;
    n0 = nt * TC_BN;
  };
  // phase p = py*2 + px: number of taps, first tap slot of the packed K axis
  auto phase_taps = [](int p, int& ntaps, int& tq0) {
    ntaps = p == 0 ? 4 : (p == 3 ? 1 : 2);
    tq0 = p == 0 ? 0 : (p == 1 ? 4 : (p == 2 ? 6 : 8));
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase_bit = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int p, m0, n0, ntaps, tq0;
        tile_coords(tile, p, m0, n0);
        phase_taps(p, ntaps, tq0);
        for (int t = 0; t < ntaps; ++t) {
          // flat shift of tap t: (ky>>1) rows and (kx>>1) columns back
          int shift;
          if (p == 0) shift = (t >> 1) * a.Wp + (t & 1);
          else if (p == 1) shift = t * a.Wp;
          else shift = t;  // p == 2: kx = 0, 2;  p == 3: the centre tap
          for (int kc = 0; kc < kpt; ++kc) {
            mbar_wait(&empty[stage], phase_bit ^ 1);
            mbar_arrive_expect_tx(&full[stage], stage_tx);
            uint8_t* st = smem + stage * TC_STAGE_BYTES;
            tc::tma_load_2d(st, &tmA_hi, &full[stage], kc * TC_BK, m0 - shift);
            tc::tma_load_2d(st + TC_TILE_BYTES, &tmA_lo, &full[stage], kc * TC_BK, m0 - shift);
            tc::tma_load_2d(st + 2 * TC_TILE_BYTES, &tmB_hi, &full[stage], (tq0 + t) * a.Cin + kc * TC_BK, n0);
            tc::tma_load_2d(st + 3 * TC_TILE_BYTES, &tmB_lo, &full[stage], (tq0 + t) * a.Cin + kc * TC_BK, n0);
            if (++stage == TC_STAGES) {
              stage = 0;
              phase_bit ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16_f32(TC_BM, nbox);
      uint32_t stage = 0, phase_bit = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        int p, m0, n0, ntaps, tq0;
        tile_coords(tile, p, m0, n0);
        phase_taps(p, ntaps, tq0);
        const int nkb = ntaps * kpt;
        const uint32_t buf = it & 1, use = it >> 1;
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
        tc::fence_after_thread_sync();
        const uint32_t dcol = tmem_base + buf * TC_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase_bit);
          tc::fence_after_thread_sync();
          const uint32_t sb = smem_u32(smem + stage * TC_STAGE_BYTES);
          const uint64_t dA_hi = tc::make_smem_desc_sw128(sb);
          const uint64_t dA_lo = tc::make_smem_desc_sw128(sb + TC_TILE_BYTES);
          const uint64_t dB_hi = tc::make_smem_desc_sw128(sb + 2 * TC_TILE_BYTES);
          const uint64_t dB_lo = tc::make_smem_desc_sw128(sb + 3 * TC_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < TC_BK / 16; ++ks) {
            const uint64_t ah = tc::advance_desc_k(dA_hi, ks), al = tc::advance_desc_k(dA_lo, ks);
            const uint64_t bh = tc::advance_desc_k(dB_hi, ks), bl = tc::advance_desc_k(dB_lo, ks);
            tc::mma_bf16_ss(dcol, ah, bh, idesc, (kb | ks) != 0);
            tc::mma_bf16_ss(dcol, ah, bl, idesc, true);
            tc::mma_bf16_ss(dcol, al, bh, idesc, true);
          }
          tc::mma_commit(&empty[stage]);
          if (++stage == TC_STAGES) {
            stage = 0;
            phase_bit ^= 1;
          }
        }
        tc::mma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    uint8_t* obuf = out_stage + chalf * TC_OUT_CHUNK_BYTES;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      int p, m0, n0;
      tile_coords(tile, p, m0, n0);
      const uint32_t buf = it & 1, use = it >> 1;
      mbar_wait(&acc_full[buf], use & 1);
      tc::fence_after_thread_sync();
      constexpr int kChunksPerWarp = TC_BN / 32 / (TC_EPI_WARPS / 4);
#pragma unroll 1
      for (int chunk = chalf * kChunksPerWarp; chunk < (chalf + 1) * kChunksPerWarp; ++chunk) {
        float v[32];
        const bool live = chunk * 32 < nbox;
        if (live) tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TC_BN + chunk * 32, v);
        if (chunk == (chalf + 1) * kChunksPerWarp - 1) {
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (!live) continue;
        // staging buffer -> TMA tensor store into T[phase] (rows past Mp are clipped by the tensor map)
        tc::named_bar_sync(2 + chalf, 128);
        tc::stage_row32(obuf, row, v);
        fence_proxy_async();
        tc::named_bar_sync(2 + chalf, 128);
        if (q == 0 && lane == 0) {
          tc::tma_store_3d(&tmOut, obuf, n0 + chunk * 32, m0, p);
          tc::tma_store_commit_and_wait_read();
        }
      }
    }
    if (q == 0 && lane == 0) tc::tma_store_wait_all();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TC_ACC_BUFS * TC_BN);
}

// output widths of the tensor-core kernels: whole 128-column tiles, or one narrow tile of 64 / 32 columns
static bool tc_n_supported(int n) { return n > 0 && (n % TC_BN == 0 || n == 64 || n == 32); }

bool tc_upconv_supported(int B, int H, int W, int Cin, int cout) {
  return B > 0 && H > 0 && W > 0 && Cin % TC_BK == 0 && tc_n_supported(cout) &&
         (int64_t)B * (H + 1) * (W + 1) < (1ll << 30);
}
size_t tc_upconv_split_bytes(int B, int H, int W, int Cin) {
  return (size_t)B * (H + 1) * (W + 1) * Cin * 2 * sizeof(__nv_bfloat16);
}
size_t tc_upconv_t_bytes(int B, int H, int W, int cout) {
  return (size_t)4 * B * (H + 1) * (W + 1) * cout * sizeof(float);
}

// x [B,H,W,Cin] fp32 NHWC, s [B,Cin] -> t_out [4][B*(H+1)*(W+1)][cout]; split_scratch holds the padded
// bf16 hi / lo operand halves (tc_upconv_split_bytes).
int tc_upconv_phase_launch(const float* x, const float* s, int B, int H, int W, int Cin, int cout,
                           const void* packed_bf16, void* split_scratch, float* t_out, cudaStream_t stream) {
  E3_REQUIRE(tc_upconv_supported(B, H, W, Cin, cout), E3_ERR_UNSUPPORTED,
             "tensor-core up-conv: unsupported shape B=%d H=%d W=%d Cin=%d cout=%d (needs Cin %% 64 == 0, "
             "cout %% 128 == 0 or cout in {32, 64})", B, H, W, Cin, cout);
  const int64_t Mp = (int64_t)B * (H + 1) * (W + 1);
  __nv_bfloat16* xs_hi = static_cast<__nv_bfloat16*>(split_scratch);
  __nv_bfloat16* xs_lo = xs_hi + Mp * Cin;
  {
    const int64_t nv = Mp * Cin / 8;
    int blocks = (int)((nv + 255) / 256);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    modulate_split_padded_kernel<<<blocks, 256, 0, stream>>>(x, s, xs_hi, xs_lo, nv, H, W, Cin);
    E3_CUDA(cudaGetLastError());
  }
  const int K = 9 * Cin;
  const __nv_bfloat16* w_hi = static_cast<const __nv_bfloat16*>(packed_bf16);
  const __nv_bfloat16* w_lo = w_hi + (size_t)cout * K;
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  const uint64_t adims[2] = {(uint64_t)Cin, (uint64_t)Mp};
  const uint64_t astr[1] = {(uint64_t)Cin * 2};
  const uint32_t abox[2] = {(uint32_t)TC_BK, (uint32_t)TC_BM};
  const uint64_t bdims[2] = {(uint64_t)K, (uint64_t)cout};
  const uint64_t bstr[1] = {(uint64_t)K * 2};
  const uint32_t bbox[2] = {(uint32_t)TC_BK, (uint32_t)(cout < TC_BN ? cout : TC_BN)};
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA_hi, xs_hi, 2, adims, astr, abox))) return rc;
  if ((rc = make_tensor_map_bf16(&tmA_lo, xs_lo, 2, adims, astr, abox))) return rc;
  if ((rc = make_tensor_map_bf16(&tmB_hi, w_hi, 2, bdims, bstr, bbox))) return rc;
  if ((rc = make_tensor_map_bf16(&tmB_lo, w_lo, 2, bdims, bstr, bbox))) return rc;
  static thread_local bool attr_set_dev[E3_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_set_dev[device_slot()];
  if (!attr_set) {
    E3_CUDA(cudaFuncSetAttribute((const void*)tc_upconv_phase_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tmOut;  // T as [phase][Mp][cout] fp32, box = 32 channels x 128 pixels
  const uint64_t odims[3] = {(uint64_t)cout, (uint64_t)Mp, 4};
  const uint64_t ostr[2] = {(uint64_t)cout * 4, (uint64_t)Mp * cout * 4};
  const uint32_t obox[3] = {32, (uint32_t)TC_BM, 1};
  if ((rc = make_tensor_map_f32(&tmOut, t_out, 3, odims, ostr, obox))) return rc;
  UpPhaseArgs a{t_out, (int)Mp, cout, Cin, W + 1};
  const int group = 4 * ((cout + TC_BN - 1) / TC_BN);  // tiles that share one m-tile: all phases x n-tiles
  const int n_tiles = group * (int)((Mp + TC_BM - 1) / TC_BM);
  int grid = sm_count() / group * group;  // multiple of the group: the phase rotation stays a bijection
  if (grid < group) grid = group;
  if (grid > n_tiles) grid = n_tiles;
  tc_upconv_phase_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, tmOut, a);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

bool tc_conv_supported(int B, int H, int W, int Cin, int N) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  return B > 0 && pow2(W) && pow2(H) && W >= 8 && H * W >= 64 && Cin % TC_BK == 0 && tc_n_supported(N);
}

size_t tc_conv_split_bytes(int B, int H, int W, int Cin) {
  return (size_t)B * H * W * Cin * 2 * sizeof(__nv_bfloat16);
}

bool tc_conv_pairx_supported(int B, int H, int W, int Cin, int N) {
  return Cin == 32 && N == 32 && W % 2 == 0 && tc_conv_supported(B, H, W / 2, 64, 64);
}
size_t tc_conv_packed_bf16_bytes(int cout, int cin) {
  if (cout == 32 && cin == 32) return (size_t)64 * 576 * 2 * 2;  // x-pair image of the plain conv (layout 0)
  return (size_t)cout * cin * 9 * 2 * 2;
}

int tc_conv_pack_weight(const float* weight, int cout, int cin, int upsample, float scale,
                        void* packed_bf16, cudaStream_t stream) {
  // a 32 -> 32 plain conv has no K-block of 64 input channels: its tensor-core image is the x-pair view
  if (upsample == 0 && cout == 32 && cin == 32) upsample = 4;
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(packed_bf16);
  const int64_t total = upsample == 4 ? (int64_t)64 * 576 : (int64_t)cout * cin * 9;
  __nv_bfloat16* lo = hi + total;
  int blocks = (int)((total + 255) / 256);
  if (blocks > sm_count() * 32) blocks = sm_count() * 32;
  conv_pack_bf16_kernel<<<blocks, 256, 0, stream>>>(weight, cout, cin, upsample, scale, hi, lo);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

size_t tc_conv_planar_elems(int B, int H, int W, int C) { return (size_t)4 * B * (H + 1) * (W + 1) * C; }

// a: x, s, out, B, H, W, Cin, N, epilogue fields filled by the caller; a.wg is unused here.
int tc_conv_launch(const ConvGemmArgs& a, int taps, const void* packed_bf16, void* split_scratch,
                   cudaStream_t stream) {
  E3_REQUIRE(tc_conv_supported(a.B, a.H, a.W, a.Cin, a.N), E3_ERR_UNSUPPORTED,
             "tensor-core conv: unsupported shape B=%d H=%d W=%d Cin=%d N=%d (needs power-of-two "
             "H, W >= 8, Cin %% 64 == 0, N %% 128 == 0 or N in {32, 64})", a.B, a.H, a.W, a.Cin, a.N);
  const size_t n_act = (size_t)a.B * a.H * a.W * a.Cin;
  __nv_bfloat16* xs_hi = static_cast<__nv_bfloat16*>(split_scratch);
  __nv_bfloat16* xs_lo = xs_hi + n_act;
  {
    const int64_t nv = (int64_t)(n_act / 8);
    int blocks = (int)((nv + 255) / 256);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    modulate_split_kernel<<<blocks, 256, 0, stream>>>(a.x, a.s, xs_hi, xs_lo, nv, a.H * a.W, a.Cin);
    E3_CUDA(cudaGetLastError());
  }
  return tc_conv_launch_presplit(a, taps, packed_bf16, xs_hi, xs_lo, stream);
}

int tc_conv_launch_presplit(const ConvGemmArgs& a, int taps, const void* packed_bf16, const void* xs_hi,
                            const void* xs_lo, cudaStream_t stream) {
  E3_REQUIRE(tc_conv_supported(a.B, a.H, a.W, a.Cin, a.N), E3_ERR_UNSUPPORTED,
             "tensor-core conv: unsupported shape B=%d H=%d W=%d Cin=%d N=%d (needs power-of-two "
             "H, W >= 8, Cin %% 64 == 0, N %% 128 == 0 or N in {32, 64})", a.B, a.H, a.W, a.Cin, a.N);
  E3_REQUIRE(!a.planar || taps == 9, E3_ERR_BAD_ARG, "tensor-core conv: planar operands need 9 taps");
  TcTile t;
  t.bw = a.W < 128 ? a.W : 128;
  t.bh = (128 / t.bw) < a.H ? (128 / t.bw) : a.H;
  t.bb = 128 / (t.bw * t.bh);
  t.tiles_x = a.W / t.bw;
  t.tiles_y = a.H / t.bh;
  t.tiles_b = (a.B + t.bb - 1) / t.bb;

  const int K = taps * a.Cin;
  const __nv_bfloat16* w_hi = static_cast<const __nv_bfloat16*>(packed_bf16);
  const __nv_bfloat16* w_lo = w_hi + (size_t)a.N * K;
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  // planar operand: [4 planes * B][H+1][W+1][Cin]
  const uint64_t aw = a.planar ? a.W + 1 : a.W, ah = a.planar ? a.H + 1 : a.H, ab = a.planar ? 4 * a.B : a.B;
  const uint64_t adims[4] = {(uint64_t)a.Cin, aw, ah, ab};
  const uint64_t astr[3] = {(uint64_t)a.Cin * 2, aw * a.Cin * 2, ah * aw * a.Cin * 2};
  const uint32_t abox[4] = {(uint32_t)TC_BK, (uint32_t)t.bw, (uint32_t)t.bh, (uint32_t)t.bb};
  const uint64_t bdims[2] = {(uint64_t)K, (uint64_t)a.N};
  const uint64_t bstr[1] = {(uint64_t)K * 2};
  const uint32_t bbox[2] = {(uint32_t)TC_BK, (uint32_t)(a.N < TC_BN ? a.N : TC_BN)};
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA_hi, xs_hi, 4, adims, astr, abox))) return rc;
  if ((rc = make_tensor_map_bf16(&tmA_lo, xs_lo, 4, adims, astr, abox))) return rc;
  if ((rc = make_tensor_map_bf16(&tmB_hi, w_hi, 2, bdims, bstr, bbox))) return rc;
  if ((rc = make_tensor_map_bf16(&tmB_lo, w_lo, 2, bdims, bstr, bbox))) return rc;
  CUtensorMap tmOut;  // output [B][H][W][N] fp32, box = 32 channels x the tile's pixel box
  const uint64_t odims[4] = {(uint64_t)a.N, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
  const uint64_t ostr[3] = {(uint64_t)a.N * 4, (uint64_t)a.W * a.N * 4, (uint64_t)a.H * a.W * a.N * 4};
  const uint32_t obox[4] = {32, (uint32_t)t.bw, (uint32_t)t.bh, (uint32_t)t.bb};
  if ((rc = make_tensor_map_f32(&tmOut, a.out, 4, odims, ostr, obox))) return rc;

  // CTA pairs pay when N allows 256-column tiles (measured: 128-column pair tiles are no faster than the
  // single-CTA kernel) and the 256 x 256 tiles fill whole waves of the 74 pairs (conv1 of the 256^2 decoder
  // has 256 of them = 3.5 waves and loses to 1024 single-CTA tiles on 148 SMs)
  if (taps == 9 && a.N % 256 == 0 && conv_pairs_enabled()) {
    const int m_tiles = t.tiles_x * t.tiles_y * t.tiles_b;
    const int n_pair_tiles = ((m_tiles + 1) / 2) * (a.N / 256), pairs = sm_count() / 2;
    const int waves = (n_pair_tiles + pairs - 1) / pairs;
    if (n_pair_tiles * 10 >= waves * pairs * 9) {
      const uint32_t pbox[2] = {(uint32_t)TC_BK, 128};  // the weight maps' box is one CTA's half of the N tile
      CUtensorMap pB_hi, pB_lo;
      if ((rc = make_tensor_map_bf16(&pB_hi, w_hi, 2, bdims, bstr, pbox))) return rc;
      if ((rc = make_tensor_map_bf16(&pB_lo, w_lo, 2, bdims, bstr, pbox))) return rc;
      return launch_conv_pair<9, 256>(tmA_hi, tmA_lo, pB_hi, pB_lo, tmOut, a, t, stream);
    }
  }
  static thread_local bool attr_set_dev[E3_MAX_DEVICES][2] = {};  // function attributes are per device
  bool (&attr_set)[2] = attr_set_dev[device_slot()];
  const int which = taps == 9 ? 1 : 0;
  if (!attr_set[which]) {
    const void* fn = which ? (const void*)tc_conv_kernel<9> : (const void*)tc_conv_kernel<1>;
    E3_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set[which] = true;
  }
  const int n_tiles = t.tiles_x * t.tiles_y * t.tiles_b * ((a.N + TC_BN - 1) / TC_BN);
  dim3 grid(n_tiles < sm_count() ? n_tiles : sm_count());
  if (taps == 9)
    tc_conv_kernel<9><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, tmOut, a, t);
  else
    tc_conv_kernel<1><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, tmOut, a, t);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

}  // namespace e3
