// Tensor-core (tcgen05) variant of the fused StyleSDF volume renderer, sm_100a.
//
// Same contract, inputs and outputs as render_siren.cu (rays -> samples -> 8 x FiLM-SIREN ->
// sdf head -> view layer -> rgb head -> SDF->sigma -> alpha composite, ONE persistent kernel);
// the eight 256x256 hidden-layer contractions run on the 5th-gen tensor cores instead of the
// FFMA pipe:
//
//   * a tile = 128 sample rows (UMMA M = 128 per CTA) = floor(128/S) whole rays;
//   * the hidden state lives in shared memory as the UMMA A operand: bf16 hi + bf16 lo
//     (h = hi + lo to 2^-17), K-major SWIZZLE_128B, 4 k-blocks x 16 KB each — written in place
//     by the epilogue of the previous layer;
//   * per layer D[128x256] (fp32 in TMEM) = h_hi*W_hi^T + h_lo*W_hi^T + h_hi*W_lo^T: the three-product
//     split keeps the result fp32-faithful (measured 5e-5..1.5e-4 end to end vs 2..5e-2 for plain bf16
//     operands; the parity bar is 1e-3);
//   * default (EPI = 7): the two CTAs of a cluster form a PAIR and run their tiles in lockstep through one
//     tcgen05.mma.cta_group::2 stream issued by the leader — 48 MMAs 256x256x16 per layer, M = 128 rows
//     from each CTA, every CTA's ring holding one N-half of each 256 x 64 weight block (2 MB of pre-split,
//     pre-swizzled bf16 tiles streamed from L2 by tensor-map TMA through a 4 x 16 KB full/empty mbarrier
//     ring, both CTAs' bytes counted on the leader's barrier).  One CTA per MMA stream (EPI = 3 / 0, 128x256x16
//     MMAs over paired ring stages) is kept for A/B timing: there the MMAs are paced by the shared-memory
//     pipe (A 4 KB + B 8 KB per MMA + weight TMA writes + epilogue stores);
//   * 16 epilogue warps per CTA: tcgen05.ld -> FiLM (bias folded into beta', table in shared memory) ->
//     2-term 2pi reduction + MUFU sin -> hi/lo split -> swizzled st.shared of the next layer's A operand,
//     the arithmetic packed two channels per instruction (FFMA2 / FADD2); layer 0 (K = 3), the sdf / rgb
//     heads, the view-direction rank-3 update, the transmittance scan and the weighted feature sum stay
//     on the CUDA cores.
//
// Layer pipelining.  Layer l+1 contracts over ALL outputs of layer l, but k-block j of that
// contraction only needs output channels [64j, 64j+64).  The epilogue therefore walks a layer in
// four 64-channel blocks and publishes each one (a_ready[j]; the first and last block in two halves,
// a_half / a_tail); the MMA thread starts layer l+1's k-block j as soon as it lands, accumulating into the
// OTHER half of TMEM (512 columns = two 128x256 fp32 accumulators, ping-pong by layer parity).  Overwriting
// the A operand in place is safe because layer l's MMAs have all completed (d_ready) before its epilogue
// starts.  In a pair every publish goes to the leader's barrier (the peer's warps arrive remotely, relaxed).
//
// Warp roles: warp 0 = TMA weight producer, warp 1 = TMEM owner (+ MMA issuer in the leader), warps 2..17 =
// compute (TMEM lane quarter = warp % 4; the four warps of a quarter share the columns of a block).  The
// sdf->alpha->scan work of a tile overlaps with the view-layer MMAs.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "render_siren.cuh"
#include "tcgen05.cuh"

namespace e3 {

constexpr int TCM = 128;                 // rows per tile
constexpr int TC_RING = 4;               // weight stages
constexpr int NS_RING = 8;               // N-split pairs: the same 64 KB as 8 stages of 64 rows
constexpr int TC_COMPUTE_WARPS = 16;
constexpr int TC_COMPUTE = TC_COMPUTE_WARPS * 32;
constexpr int TC_NTHREADS = 64 + TC_COMPUTE;  // producer warp + MMA warp + compute warps
constexpr int A_KBLOCK_BYTES = TCM * 128;      // 128 rows x 64 bf16
constexpr int A_BYTES = 4 * A_KBLOCK_BYTES;    // one of hi / lo: 64 KB
// timing experiments only (profiles/whatif_render.py): results are numerically WRONG with these set
constexpr uint32_t DBG_SKIP_LO_MMA = 1u << 30;  // issue only the hi*hi products (1/3 of the MMAs)
constexpr uint32_t DBG_NO_SIN = 1u << 31;       // epilogue without the sin evaluation
constexpr uint32_t DBG_NO_WSTREAM = 1u << 29;   // no weight TMA, MMAs read whatever the ring holds

struct SmemTC {
  uint8_t a_hi[A_BYTES];  // also the fp32 [256][128] composite buffer together with a_lo
  uint8_t a_lo[A_BYTES];
  uint8_t ring[TC_RING * TC_TILE_BYTES];
  float film[9][2][SW];   // gamma, beta' = gamma*bias + beta of the current image
  float wsig[SW];
  float z[TCM], dist[TCM], alpha[TCM], wgt[TCM], vis[TCM];
  float sdf_part[4][TCM];   // partial head sums of the 4 column quarters
  float rgb_part[4][3][TCM];
  float ray_o[3][TCM], ray_d[3][TCM];
  uint64_t full[NS_RING], empty[NS_RING];  // TC_RING stages of 16 KB, or NS_RING of 8 KB (N-split pairs)
  uint64_t a_ready[4], d_ready;
  uint64_t d_ready1;  // N-split pairs: accumulator columns of the second N-half (channel blocks 1, 3) are complete
  uint64_t a_half;  // EPI bit 1: channels [0,32) of block 0 are published (the next layer's first MMAs start)
  uint64_t a_tail;  // pairs: channels [192,224) of block 3 are published (half of the last k-block's MMAs start)
  uint32_t tmem_slot;
};
static_assert(sizeof(SmemTC) + 1024 <= 227 * 1024, "shared memory budget");
static_assert(TC_RING == 4 && TC_TILES_PER_LAYER % 4 == 0, "stage pairing of the 256-row weight blocks");

__device__ __forceinline__ void compute_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(TC_COMPUTE) : "memory");
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// byte offset of 16-byte chunk `c16` (8 consecutive k) of row m inside one swizzled k-block
__device__ __forceinline__ uint32_t a_chunk_off(int m, int c16) {
  return (uint32_t)((m >> 3) * 1024 + (m & 7) * 128 + ((c16 ^ (m & 7)) << 4));
}

// 8 consecutive outputs v[0..7] (channels n0..n0+7 of row m) -> next layer's A operand
__device__ __forceinline__ void store_a8(SmemTC& sm, int m, int n0, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair_bf16(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  const uint32_t off = (uint32_t)(n0 >> 6) * A_KBLOCK_BYTES + a_chunk_off(m, (n0 & 63) >> 3);
  *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2: one issue slot for two lanes' worth of work) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2 pk2u(uint32_t a, uint32_t b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// sin of the pair gamma*acc + beta' (same arithmetic as sin_mufu_reduced, two lanes per instruction)
__device__ __forceinline__ void film_sin2(f32x2 acc, f32x2 g, f32x2 bt, float& s0, float& s1, f32x2& arg) {
  const f32x2 kInv = pk2(0.159154943f, 0.159154943f), kMagic = pk2(12582912.0f, 12582912.0f);
  const f32x2 kNegMagic = pk2(-12582912.0f, -12582912.0f), kC1 = pk2(-6.28125f, -6.28125f);
  const f32x2 kC2 = pk2(-1.9353071795864769e-3f, -1.9353071795864769e-3f);
  arg = fma2(g, acc, bt);
  const f32x2 jm = fma2(arg, kInv, kMagic);
  const f32x2 j = add2(jm, kNegMagic);
  f32x2 r = fma2(j, kC1, arg);
  r = fma2(j, kC2, r);
  float r0, r1;
  unpk2(r, r0, r1);
  s0 = __sinf(r0);
  s1 = __sinf(r1);
}
// hi/lo bf16 split of a pair with the subtraction packed
__device__ __forceinline__ void split_pair_bf16_x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const f32x2 h = pk2u(hi << 16, hi & 0xffff0000u);
  const f32x2 d = fma2(h, pk2(-1.f, -1.f), pk2(v0, v1));  // v - h, exact
  float d0, d1;
  unpk2(d, d0, d1);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d1), "f"(d0));
}
template <int EPI>
__device__ __forceinline__ void store_a8_v(SmemTC& sm, int m, int n0, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (EPI & 1) split_pair_bf16_x2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
    else split_pair_bf16(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  }
  const uint32_t off = (uint32_t)(n0 >> 6) * A_KBLOCK_BYTES + a_chunk_off(m, (n0 & 63) >> 3);
  *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// CL = thread-block cluster size.  The weight stream is identical for every tile, so the CTAs of a
// cluster share it: each fetches 1/CL of every 16 KB weight tile and TMA-multicasts it into all CL
// rings (cp.async.bulk ... .multicast::cluster).  A ring stage is recycled when every CTA of the
// cluster has consumed it (tcgen05.commit multicast onto all CL `empty` barriers, count = CL).  This
// divides the L2 -> SM traffic of the stream by CL and multiplies the bytes each SM has in flight,
// which is what bounds the kernel at CL = 1 (profiles/r01e_whatif_render.txt).
// wmap: 2-D tensor map over the pre-swizzled weight stream viewed as rows of 64 bf16 (128 B), box =
// one 128-row tile, no hardware swizzle; use_wmap selects cp.async.bulk.tensor over the 1-D bulk copy.
// STASH: training build that also writes the pre-sin phases for e3_render_bwd (a compile-time switch:
// even predicated off, the per-element address arithmetic would cost the inference build ~3
// instruction slots per element).
// EPI: epilogue code variant of the hidden layers (bit 0: packed f32x2 FiLM / range reduction / split
// and the FiLM rows fetched while the tcgen05.ld is in flight; bit 1: every warp takes 8 channels of
// each 32-channel half of a block instead of 16 contiguous ones, and the first half of block 0 is
// published on its own barrier so that the next layer's MMAs start half a block earlier; bit 2: CTA
// pairs — the two CTAs of a cluster (CL = 2) run their tiles in lockstep through ONE tcgen05.mma
// cta_group::2 stream issued by the leader: M = 256 = 128 rows from each CTA, every CTA's ring holds one
// N-half of each 256 x 64 weight block, so per SM the B-operand reads and the weight stream are halved
// (shared-memory bandwidth is what paces the single-CTA MMAs) and the ring is twice as deep in blocks).
template <int MODE, int CL, bool STASH, int EPI>
// 18 warps: one scheduler holds 5 of them, so 16384 / (5 * 32) = 102 -> 96 registers per thread
__global__ void __launch_bounds__(TC_NTHREADS, 1)
siren_render_tc_kernel(const __grid_constant__ RenderArgs a, const __grid_constant__ CUtensorMap wmap,
                       const __grid_constant__ CUtensorMap wmap64, const int use_wmap) {
  static_assert(EPI == 0 || EPI == 1 || EPI == 3 || EPI == 7 || EPI == 15, "the half-block column mapping needs the packed epilogue");
  constexpr bool PAIR = (EPI & 4) != 0;
  // N-split pairs (EPI bit 3): every 256-column layer is issued as two 128-column halves, interleaved by k-block
  // (units (n0,k0) (n0,k2) (n1,k0) (n1,k2) (n0,k1) (n0,k3) (n1,k1) (n1,k3)), so that the first half — channel
  // blocks 0 and 2 — is complete, and its epilogue running, while the tensor pipe still works on the second
  // one: the per-layer bubble (drain + first block of the epilogue) shrinks from ~2.5k to ~1.3k cycles.
  constexpr bool NS = (EPI & 8) != 0;
  static_assert(!NS || PAIR, "the N-split is built on the CTA-pair MMA stream");
  static_assert(!PAIR || CL == 2, "CTA pairs are clusters of two");
  const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = pair_rank == 0;
  extern __shared__ uint8_t smem_raw[];
  SmemTC& sm = *reinterpret_cast<SmemTC*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS_RING; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], PAIR ? 1 : CL);
    }
    // pairs: the epilogue warps of BOTH CTAs arrive on the leader's barriers
    const uint32_t n_arrive = (PAIR ? 2 : 1) * TC_COMPUTE_WARPS;
#pragma unroll
    for (int j = 0; j < 4; ++j) mbar_init(&sm.a_ready[j], n_arrive);
    mbar_init(&sm.d_ready, 1);
    mbar_init(&sm.d_ready1, 1);
    mbar_init(&sm.a_half, n_arrive);
    mbar_init(&sm.a_tail, n_arrive);
    fence_mbar_init();
  }
  for (int i = tid; i < SW; i += TC_NTHREADS) sm.wsig[i] = a.packed[OFF_WSIG + i];
  if (PAIR) {
    cluster_sync_all();  // both CTAs are resident and their barriers initialised before the paired alloc
    if (warp == 1) tc::tmem_alloc_pair(&sm.tmem_slot, 512);
  } else if (warp == 1) {
    tc::tmem_alloc(&sm.tmem_slot, 512);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every CTA's barriers are initialised before any multicast lands
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = sm.tmem_slot;

  // every CTA runs the same number of tile slots (the cluster shares one weight stream); slots past
  // the last tile are dummies that only keep the pipeline in step
  const int n_my_tiles = (a.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int gemm_layers = a.with_view ? 8 : 7;
  const uint16_t cl_mask = (uint16_t)((1u << CL) - 1);

  if (warp == 0) {
    // ===== TMA producer: 16 weight tiles per layer, in consumption order =====
    if (lane == 0 && !(a.p.flags & DBG_NO_WSTREAM)) {
      const uint8_t* stream = reinterpret_cast<const uint8_t*>(a.packed + OFF_TC_STREAM);
      const int per_tile = gemm_layers * TC_TILES_PER_LAYER;
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        if (NS) {
          // one stage = 64 of this CTA's 128 rows of a weight block (8 KB), in the order the MMA units consume
          // them: (nh, kb) = (0,0) (0,2) (1,0) (1,2) (0,1) (0,3) (1,1) (1,3), W_hi then W_lo each
          for (int l = 0; l < gemm_layers; ++l) {
#pragma unroll 1
            for (int u = 0; u < 8; ++u) {
              const int nh = (u >> 1) & 1, kb = (u & 1) * 2 + (u >> 2);
#pragma unroll 1
              for (int p = 0; p < 2; ++p) {
                mbar_wait(&sm.empty[stage], phase ^ 1);
                if (leader) mbar_arrive_expect_tx(&sm.full[stage], 2 * (TC_TILE_BYTES / 2));
                tc::tma_load_2d_pair(sm.ring + stage * (TC_TILE_BYTES / 2), &wmap64, tc::map_to_cta(&sm.full[stage], 0), 0,
                                     (2 * (l * 8 + kb * 2 + p) + (int)pair_rank) * 128 + nh * 64);
                if (++stage == NS_RING) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
          }
        } else if (PAIR) {
          // one stage = one 256(n) x 64(k) block over the pair: this CTA fetches its n-half, the bytes of
          // both halves are counted on the leader's barrier
          for (int c = 0; c < per_tile / 2; ++c) {
            mbar_wait(&sm.empty[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&sm.full[stage], 2 * TC_TILE_BYTES);
            tc::tma_load_2d_pair(sm.ring + stage * TC_TILE_BYTES, &wmap, tc::map_to_cta(&sm.full[stage], 0), 0,
                                 (2 * c + (int)pair_rank) * 128);
            if (++stage == TC_RING) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        for (int c = 0; c < (PAIR ? 0 : per_tile); ++c) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&sm.full[stage], TC_TILE_BYTES);
          if (CL == 1 && use_wmap) {
            tc::tma_load_2d(sm.ring + stage * TC_TILE_BYTES, &wmap, &sm.full[stage], 0, c * 128);
          } else if (CL == 1) {
            tma_bulk_g2s(sm.ring + stage * TC_TILE_BYTES, stream + (size_t)c * TC_TILE_BYTES,
                         TC_TILE_BYTES, &sm.full[stage]);
          } else {
            constexpr uint32_t slice = TC_TILE_BYTES / CL;
            const uint32_t off = cluster_ctarank() * slice;
            tma_bulk_g2s_multicast(sm.ring + stage * TC_TILE_BYTES + off,
                                   stream + (size_t)c * TC_TILE_BYTES + off, slice, &sm.full[stage], cl_mask);
          }
          if (++stage == TC_RING) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (NS && leader && lane == 0) {
      // ===== CTA pair, N-split: units of 12 MMAs 256 x 128 x 16 =====
      const uint32_t idesc = tc::make_idesc_bf16_f32(256, 128);
      const uint32_t a_hi0 = smem_u32(sm.a_hi), a_lo0 = smem_u32(sm.a_lo);
      uint32_t stage = 0, phase = 0, pa = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        for (int l = 0; l < gemm_layers; ++l) {
#pragma unroll 1
          for (int u = 0; u < 8; ++u) {
            const int nh = (u >> 1) & 1, kb = (u & 1) * 2 + (u >> 2);
            const uint32_t dcol = tmem_base + (uint32_t)(l & 1) * 256 + (uint32_t)nh * 128;
            const bool first = u == 0;  // block 0 arrives in two halves (a_half, then a_ready[0])
            mbar_wait(first ? &sm.a_half : &sm.a_ready[kb], pa);
            tc::fence_after_thread_sync();
            const uint64_t dAh = tc::make_smem_desc_sw128(a_hi0 + kb * A_KBLOCK_BYTES);
            const uint64_t dAl = tc::make_smem_desc_sw128(a_lo0 + kb * A_KBLOCK_BYTES);
            const uint32_t s1 = (stage + 1 == NS_RING) ? 0 : stage + 1;  // W_lo piece (ring of 8: same phase)
            const uint64_t dB0 = tc::make_smem_desc_sw128(smem_u32(sm.ring + stage * (TC_TILE_BYTES / 2)));
            const uint64_t dB1 = tc::make_smem_desc_sw128(smem_u32(sm.ring + s1 * (TC_TILE_BYTES / 2)));
            auto hi_block = [&](int ks) {  // h_hi * W_hi + h_lo * W_hi
              const uint64_t bk = tc::advance_desc_k(dB0, ks);
              tc::mma_bf16_ss_pair(dcol, tc::advance_desc_k(dAh, ks), bk, idesc, (kb | ks) != 0);
              tc::mma_bf16_ss_pair(dcol, tc::advance_desc_k(dAl, ks), bk, idesc, true);
            };
            auto lo_block = [&](int ks) {  // h_hi * W_lo
              tc::mma_bf16_ss_pair(dcol, tc::advance_desc_k(dAh, ks), tc::advance_desc_k(dB1, ks), idesc, true);
            };
            mbar_wait(&sm.full[stage], phase);
            tc::fence_after_thread_sync();
            hi_block(0), hi_block(1);
            if (first) {
              mbar_wait(&sm.a_ready[0], pa);  // second half of block 0
              tc::fence_after_thread_sync();
            }
            hi_block(2), hi_block(3);
            tc::mma_commit_pair(&sm.empty[stage], 3);
            mbar_wait(&sm.full[s1], phase);  // stage is even here: its successor never wraps
            tc::fence_after_thread_sync();
            lo_block(0), lo_block(1), lo_block(2), lo_block(3);
            tc::mma_commit_pair(&sm.empty[s1], 3);
            stage += 2;
            if (stage == NS_RING) {
              stage = 0;
              phase ^= 1;
            }
            if (u == 5) tc::mma_commit_pair(&sm.d_ready, 3);   // half 0 (channel blocks 0, 2) complete
            if (u == 7) tc::mma_commit_pair(&sm.d_ready1, 3);  // half 1 (channel blocks 1, 3) complete
          }
          pa ^= 1;
        }
      }
    }
    if (PAIR && !NS && leader && lane == 0) {
      // ===== CTA pair: one tcgen05.mma cta_group::2 stream over both CTAs' tiles (M = 256) =====
      // A k-block's 12 MMAs use two ring stages (W_hi block, W_lo block).  Blocks 0 and 3 arrive in two
      // halves (a_half / a_tail, then a_ready): the k-steps of the first half are issued for both weight
      // blocks before the second half is waited for, so the head of a layer starts, and the tail after the
      // epilogue's last publish shrinks, by half a block.
      const uint32_t idesc = tc::make_idesc_bf16_f32(256, 256);
      const uint32_t a_hi0 = smem_u32(sm.a_hi), a_lo0 = smem_u32(sm.a_lo);
      uint32_t stage = 0, phase = 0, pa = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        const bool tr = a.trace && blockIdx.x == 0 && t == 1;
        for (int l = 0; l < gemm_layers; ++l) {
          const uint32_t dcol = tmem_base + (uint32_t)(l & 1) * 256;
          for (int kb = 0; kb < 4; ++kb) {
            const bool split = kb == 0 || kb == 3;
            mbar_wait(kb == 0 ? &sm.a_half : (kb == 3 ? &sm.a_tail : &sm.a_ready[kb]), pa);
            if (tr) a.trace[128 + l * 8 + kb] = clock64();
            tc::fence_after_thread_sync();
            const uint64_t dAh = tc::make_smem_desc_sw128(a_hi0 + kb * A_KBLOCK_BYTES);
            const uint64_t dAl = tc::make_smem_desc_sw128(a_lo0 + kb * A_KBLOCK_BYTES);
            const uint64_t dB0 = tc::make_smem_desc_sw128(smem_u32(sm.ring + stage * TC_TILE_BYTES));
            const uint64_t dB1 = tc::make_smem_desc_sw128(smem_u32(sm.ring + (stage + 1) * TC_TILE_BYTES));
            auto hi_block = [&](int ks) {  // h_hi * W_hi + h_lo * W_hi
              const uint64_t bk = tc::advance_desc_k(dB0, ks);
              tc::mma_bf16_ss_pair(dcol, tc::advance_desc_k(dAh, ks), bk, idesc, (kb | ks) != 0);
              tc::mma_bf16_ss_pair(dcol, tc::advance_desc_k(dAl, ks), bk, idesc, true);
            };
            auto lo_block = [&](int ks) {  // h_hi * W_lo
              tc::mma_bf16_ss_pair(dcol, tc::advance_desc_k(dAh, ks), tc::advance_desc_k(dB1, ks), idesc, true);
            };
            mbar_wait(&sm.full[stage], phase);
            tc::fence_after_thread_sync();
            if (split) {
              hi_block(0), hi_block(1);
              mbar_wait(&sm.full[stage + 1], phase);
              tc::fence_after_thread_sync();
              lo_block(0), lo_block(1);
              mbar_wait(&sm.a_ready[kb], pa);  // second half of the block
              tc::fence_after_thread_sync();
              hi_block(2), hi_block(3);
              tc::mma_commit_pair(&sm.empty[stage], 3);
              lo_block(2), lo_block(3);
              tc::mma_commit_pair(&sm.empty[stage + 1], 3);
            } else {
              hi_block(0), hi_block(1), hi_block(2), hi_block(3);
              tc::mma_commit_pair(&sm.empty[stage], 3);
              mbar_wait(&sm.full[stage + 1], phase);
              tc::fence_after_thread_sync();
              lo_block(0), lo_block(1), lo_block(2), lo_block(3);
              tc::mma_commit_pair(&sm.empty[stage + 1], 3);
            }
            if (tr) a.trace[256 + l * 16 + kb * 4 + 3] = clock64();
            stage += 2;
            if (stage == TC_RING) {
              stage = 0;
              phase ^= 1;
            }
          }
          pa ^= 1;
          tc::mma_commit_pair(&sm.d_ready, 3);
        }
      }
    }
    if (!PAIR && lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16_f32(128, 256);
      const uint32_t a_hi0 = smem_u32(sm.a_hi), a_lo0 = smem_u32(sm.a_lo);
      const bool skip_lo = (a.p.flags & DBG_SKIP_LO_MMA) != 0;
      const bool no_w = (a.p.flags & DBG_NO_WSTREAM) != 0;
      uint32_t stage = 0, phase = 0, pa = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        const bool tr = a.trace && blockIdx.x == 0 && t == 1;
        for (int l = 0; l < gemm_layers; ++l) {  // l-th GEMM of the tile = reference layer l+1
          const uint32_t dcol = tmem_base + (uint32_t)(l & 1) * 256;
          for (int kb = 0; kb < 4; ++kb) {
            // k-block kb of this layer's A operand is published
            if ((EPI & 2) && kb == 0) {
              mbar_wait(&sm.a_half, pa);
            } else {
              mbar_wait(&sm.a_ready[kb], pa);
            }
            if (tr) a.trace[128 + l * 8 + kb] = clock64();
            tc::fence_after_thread_sync();
            const uint64_t dAh = tc::make_smem_desc_sw128(a_hi0 + kb * A_KBLOCK_BYTES);
            const uint64_t dAl = tc::make_smem_desc_sw128(a_lo0 + kb * A_KBLOCK_BYTES);
            // Two adjacent ring stages hold the n-halves of one 256 x 64 weight block ((hi,n0),(hi,n1)
            // then (lo,n0),(lo,n1); 16 tiles per layer and 4 stages keep the pairs aligned), so one
            // 128x256x16 UMMA covers both halves: the A operand is read from shared memory once per
            // 256 output columns instead of twice (shared-memory bandwidth bounds this kernel).
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {  // pr 0: W_hi block, pr 1: W_lo block
              if (!no_w) {
                mbar_wait(&sm.full[stage], phase);
                mbar_wait(&sm.full[stage + 1], phase);
              }
              tc::fence_after_thread_sync();
              const uint64_t dB = tc::make_smem_desc_sw128(smem_u32(sm.ring + stage * TC_TILE_BYTES));
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                if ((EPI & 2) && kb == 0 && pr == 0 && ks == 2) {  // second half of block 0
                  mbar_wait(&sm.a_ready[0], pa);
                  tc::fence_after_thread_sync();
                }
                const uint64_t bk = tc::advance_desc_k(dB, ks);
                auto mma = [&](uint64_t da, bool accum) { tc::mma_bf16_ss(dcol, da, bk, idesc, accum); };
                if (pr == 0) {
                  mma(tc::advance_desc_k(dAh, ks), (kb | ks) != 0);
                  if (!skip_lo) mma(tc::advance_desc_k(dAl, ks), true);
                } else if (!skip_lo) {
                  mma(tc::advance_desc_k(dAh, ks), true);
                }
              }
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (no_w) continue;
                if (CL == 1) tc::mma_commit(&sm.empty[stage + h]);
                else tc::mma_commit_multicast(&sm.empty[stage + h], cl_mask);
              }
              if (tr) a.trace[256 + l * 16 + kb * 4 + pr * 2 + 1] = clock64();
              stage += 2;
              if (stage == TC_RING) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          pa ^= 1;  // each a_ready[kb] completes exactly once per layer
          tc::mma_commit(&sm.d_ready);
        }
      }
    }
  } else {
    // ===== compute warps =====
    const e3_render_params& P = a.p;
    const int ct = tid - 64;              // 0..511
    const int q = warp & 3;               // TMEM lane quarter this warp may read
    const int hw = (warp - 2) >> 2;       // which 16-column quarter of each 64-channel block (0..3)
    const int m = q * 32 + lane;          // tile row = TMEM lane
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + hw * 16;
    const int S = (MODE == 0) ? P.n_samples : 1;
    const int HW = (MODE == 0) ? P.height * P.width : a.n_points;
    const float* pk = a.packed;
    float* w0s = &sm.rgb_part[0][0][0];   // [3][256] layer-0 weights (start of a tile)
    float* wvs = &sm.film[0][0][0];       // [6][256] W_dir rows, W_rgb rows (end of a tile, FiLM rows 0..2)
    static_assert(sizeof(sm.rgb_part) >= 3 * SW * sizeof(float) && 3 * 2 * SW == 6 * SW, "aliases");
    uint32_t pd = 0;
    int cur_b = -1;
    const bool no_sin = (P.flags & DBG_NO_SIN) != 0;

    // publish k-block j of the next A operand: generic-proxy stores -> async proxy (UMMA)
    // Pairs: every arrival goes to the LEADER's barrier.  The peer's warps arrive remotely with a relaxed
    // arrive: the release form is a cluster-scope fence (~1.5k cycles on the critical path, measured), and
    // there is nothing left to order — the proxy fence has completed this warp's shared-memory stores, and
    // they are read by the tensor core of the SM that holds them.
    const bool remote = PAIR && !leader;
    const uint32_t ready0 = remote ? tc::map_to_cta(&sm.a_ready[0], 0) : 0;  // a_ready[j] is 8*j bytes on
    const uint32_t half_addr = remote ? tc::map_to_cta(&sm.a_half, 0) : 0;
    const uint32_t tail_addr = remote ? tc::map_to_cta(&sm.a_tail, 0) : 0;
    auto publish_on = [&](uint64_t* bar, uint32_t remote_addr) {
      fence_proxy_async();
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) {
        if (remote) tc::mbar_arrive_cluster_relaxed(remote_addr);
        else mbar_arrive(bar);
      }
    };
    auto publish = [&](int j) { publish_on(&sm.a_ready[j], ready0 + 8u * (uint32_t)j); };
    // first 8 channels of block j are stored: the half-block barriers of blocks 0 (and, in pairs, 3)
    auto publish_first_half = [&](int j) {
      if ((EPI & 2) && j == 0) publish_on(&sm.a_half, half_addr);
      if (PAIR && j == 3) publish_on(&sm.a_tail, tail_addr);
    };
    // first channel of this warp's g8-th group of 8 inside a 64-channel block
    auto col_of = [&](int g8) -> int { return (EPI & 2) ? g8 * 32 + hw * 8 : hw * 16 + g8 * 8; };

    for (int slot = 0; slot < n_my_tiles; ++slot) {
      const int tile_raw = blockIdx.x + slot * gridDim.x;
      const bool dummy = tile_raw >= a.n_tiles;  // keeps this CTA in step with its cluster
      const int tile = dummy ? a.n_tiles - 1 : tile_raw;
      const int b = tile / a.tiles_per_image;
      const int t_in = tile - b * a.tiles_per_image;
      const int unit0 = t_in * a.rays_per_tile;
      const int n_units = dummy ? 0 : min(a.rays_per_tile, HW - unit0);
      const int n_valid = n_units * S;
      const size_t samp0 = ((size_t)b * HW + unit0) * S;
      const bool valid = m < n_valid;
      const int r = valid ? m / S : 0, s = valid ? m - r * S : 0;
      float pre_film[3], pre_w0[2];
      bool film_rows_pending;

      {  // FiLM table of this image -> shared memory (rows gamma, beta'); the previous tile's view-layer
         // epilogue parked W_dir / W_rgb in the rows of layers 0..2, so those come back every tile
        const float* f = a.in.film + (size_t)b * 9 * FILM_ROWS * SW;
        const int n_layers = (b != cur_b) ? 9 : (a.with_view ? 3 : 0);
        for (int i = ct + 3 * 2 * SW; i < n_layers * 2 * SW; i += TC_COMPUTE) {  // layers 3..8: new image only
          const int l = i / (2 * SW), row = (i / SW) & 1, n = i % SW;
          sm.film[l][row][n] = f[(l * FILM_ROWS + (row ? 2 : 0)) * SW + n];
        }
        // rows of layers 0..2 (3 values per thread) and the layer-0 weights [3][256] (they live in the
        // rgb_part array, which is only used at the end of a tile): requested here, stored after the
        // geometry below so that the two latencies overlap
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = ct + k * TC_COMPUTE, l = i / (2 * SW), row = (i / SW) & 1, n = i % SW;
          pre_film[k] = n_layers ? __ldg(f + (l * FILM_ROWS + (row ? 2 : 0)) * SW + n) : 0.f;
        }
        pre_w0[0] = __ldg(pk + OFF_W0N + ct);
        pre_w0[1] = ct < SW ? __ldg(pk + OFF_W0N + TC_COMPUTE + ct) : 0.f;
        film_rows_pending = n_layers != 0;
        cur_b = b;
      }

      // ---- per-row geometry, recomputed by both column halves ----
      float x0 = 0.f, x1 = 0.f, x2 = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f;
      if (MODE == 0) {
        if (valid) {
          const int ray = unit0 + r, py = ray / P.width, px = ray - py * P.width;
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
          if (P.flags & E3_RENDER_STATIC_VIEWDIRS) v0 = dx, v1 = dy, v2 = dz;
          else v0 = rd[0], v1 = rd[1], v2 = rd[2];
          const float vn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(v0, v0), __fmul_rn(v1, v1)), __fmul_rn(v2, v2)));
          v0 = __fdiv_rn(v0, vn), v1 = __fdiv_rn(v1, vn), v2 = __fdiv_rn(v2, vn);
          const float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rd[0], rd[0]), __fmul_rn(rd[1], rd[1])),
                                           __fmul_rn(rd[2], rd[2])));
          const float o[3] = {c2w[3], c2w[7], c2w[11]};
          const float nr = a.in.near[b], fr = a.in.far[b];
          auto z_of = [&](int si) -> float {
            if (a.in.z_jitter) return a.in.z_jitter[samp0 + r * S + si];
            const float t = a.in.t_vals[si];
            return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
          };
          const float z = z_of(s);
          float pw[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) pw[c] = __fadd_rn(o[c], __fmul_rn(rd[c], z));
          x0 = __fmul_rn(pw[0], P.pts_scale), x1 = __fmul_rn(pw[1], P.pts_scale),
          x2 = __fmul_rn(pw[2], P.pts_scale);
          if (hw == 0) {
            float dd;
            if (s + 1 < S) dd = __fsub_rn(z_of(s + 1), z);
            else if (P.flags & E3_RENDER_NO_FORCE_STOP) dd = (S > 1) ? __fsub_rn(z_of(1), z_of(0)) : 0.f;
            else dd = 1e10f;
            dd = __fmul_rn(dd, dn);
            sm.z[m] = z;
            sm.dist[m] = dd;
            if (a.out.dists) a.out.dists[samp0 + m] = dd;
            if (a.out.points) {
              float* op = a.out.points + (samp0 + m) * 3;
              op[0] = pw[0], op[1] = pw[1], op[2] = pw[2];
            }
            if (s == 0) {
#pragma unroll
              for (int c = 0; c < 3; ++c) sm.ray_o[c][r] = o[c], sm.ray_d[c][r] = rd[c];
              const size_t ro = ((size_t)b * HW + ray) * 3;
              if (a.out.rays_o) a.out.rays_o[ro] = o[0], a.out.rays_o[ro + 1] = o[1], a.out.rays_o[ro + 2] = o[2];
              if (a.out.rays_d) a.out.rays_d[ro] = rd[0], a.out.rays_d[ro + 1] = rd[1], a.out.rays_d[ro + 2] = rd[2];
              if (a.out.viewdirs) a.out.viewdirs[ro] = v0, a.out.viewdirs[ro + 1] = v1, a.out.viewdirs[ro + 2] = v2;
            }
          }
        }
      } else if (valid) {
        const float* pp = a.points + (samp0 + m) * 3;
        x0 = __fmul_rn(pp[0], P.pts_scale), x1 = __fmul_rn(pp[1], P.pts_scale),
        x2 = __fmul_rn(pp[2], P.pts_scale);
        if (a.pviewdirs) {
          const float* vv = a.pviewdirs + (samp0 + m) * 3;
          v0 = vv[0], v1 = vv[1], v2 = vv[2];
        }
      }

      if (a.in.local_alpha && a.with_view) {
        // the tile's rows of the local texture modulation (n_valid x 1 KB of alpha and of beta, contiguous)
        // are requested into L2 now; layer 7's epilogue reads them ~70k cycles later
        const char* pa = reinterpret_cast<const char*>(a.in.local_alpha + samp0 * SW);
        const char* pb = reinterpret_cast<const char*>(a.in.local_beta + samp0 * SW);
        for (int line = ct; line < n_valid * 8; line += TC_COMPUTE) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + (size_t)line * 128));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + (size_t)line * 128));
        }
      }
      if (film_rows_pending) {
#pragma unroll
        for (int k = 0; k < 3; ++k) (&sm.film[0][0][0])[ct + k * TC_COMPUTE] = pre_film[k];
      }
      w0s[ct] = pre_w0[0];
      if (ct < SW) w0s[TC_COMPUTE + ct] = pre_w0[1];
      compute_sync();  // FiLM table visible; the previous tile's readers of shared arrays are done

      // rendering.return_feats: hidden states after 0-based layers 0,2,4,6 (volume_renderer.py:179-180)
      auto store_tap = [&](int tap, int n0, const float* v) {
        float4* dst = reinterpret_cast<float4*>(
            a.out.feats_taps + ((size_t)tap * P.batch * HW * S + samp0 + m) * SW + n0);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      };
      const bool taps = (MODE == 0) && a.out.feats_taps && valid;
      // training: pre-sin phases of every FiLM layer for e3_render_bwd ([layer][channel][row])
      float* stash = (STASH && !dummy) ? a.stash + (size_t)tile * STASH_FLOATS_PER_TILE + m : nullptr;
      const bool tr = a.trace && blockIdx.x == 0 && slot == 1 && tid == 64;
      if (tr) a.trace[0] = clock64();
      if (a.trace && blockIdx.x == 0 && slot == 2 && tid == 64) a.trace[68] = clock64();  // next tile's layer 0 starts

      // ---- layer 0 (K = 3) on the CUDA cores: h0 = sin(gamma*(W0 x) + beta'), 4 blocks ----
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          const int n0 = j * 64 + col_of(g8);
          float v[8];
#pragma unroll
          for (int i4 = 0; i4 < 2; ++i4) {
            const int n = n0 + i4 * 4;
            const float4 w0 = *reinterpret_cast<const float4*>(w0s + n);
            const float4 w1 = *reinterpret_cast<const float4*>(w0s + SW + n);
            const float4 w2 = *reinterpret_cast<const float4*>(w0s + 2 * SW + n);
            if (EPI & 1) {
              const float4 g4 = *reinterpret_cast<const float4*>(&sm.film[0][0][n]);
              const float4 b4 = *reinterpret_cast<const float4*>(&sm.film[0][1][n]);
              const f32x2 X0 = pk2(x0, x0), X1 = pk2(x1, x1), X2 = pk2(x2, x2), Z = pk2(0.f, 0.f);
              const f32x2 pa = fma2(pk2(w2.x, w2.y), X2, fma2(pk2(w1.x, w1.y), X1, fma2(pk2(w0.x, w0.y), X0, Z)));
              const f32x2 pb = fma2(pk2(w2.z, w2.w), X2, fma2(pk2(w1.z, w1.w), X1, fma2(pk2(w0.z, w0.w), X0, Z)));
              f32x2 a0, a1;
              film_sin2(pa, pk2(g4.x, g4.y), pk2(b4.x, b4.y), v[i4 * 4], v[i4 * 4 + 1], a0);
              film_sin2(pb, pk2(g4.z, g4.w), pk2(b4.z, b4.w), v[i4 * 4 + 2], v[i4 * 4 + 3], a1);
              if (STASH && stash) {
                float t0, t1, t2, t3;
                unpk2(a0, t0, t1);
                unpk2(a1, t2, t3);
                float* sp = stash + (size_t)n * TCM;
                sp[0] = t0, sp[TCM] = t1, sp[2 * TCM] = t2, sp[3 * TCM] = t3;
              }
            }
            const float a4[4] = {fmaf(w2.x, x2, fmaf(w1.x, x1, w0.x * x0)), fmaf(w2.y, x2, fmaf(w1.y, x1, w0.y * x0)),
                                 fmaf(w2.z, x2, fmaf(w1.z, x1, w0.z * x0)), fmaf(w2.w, x2, fmaf(w1.w, x1, w0.w * x0))};
#pragma unroll
            for (int i = 0; i < ((EPI & 1) ? 0 : 4); ++i) {
              const float arg = fmaf(sm.film[0][0][n + i], a4[i], sm.film[0][1][n + i]);
              if (STASH && stash) stash[(size_t)(n + i) * TCM] = arg;
              v[i4 * 4 + i] = sin_mufu_reduced(arg);
            }
          }
          store_a8_v<EPI>(sm, m, n0, v);
          if (taps) store_tap(0, n0, v);
          if (g8 == 0) publish_first_half(j);
        }
        publish(j);
        if (tr) a.trace[1 + j] = clock64();
      }

      // ---- hidden layers 1..7: TMEM -> FiLM + sin -> next A operand, 64 channels at a time ----
      float sdf_acc = 0.f;
      // accumulator column of channel block j: N-split pairs lay a layer out as [block 0 | block 2 | block 1 | block 3]
      // (each 128-column MMA takes 64 weight rows from either CTA of the pair)
      auto dcb = [](int j) -> uint32_t { return NS ? (uint32_t)((j & 1) * 128 + (j >> 1) * 64) : (uint32_t)(j * 64); };
      for (int l = 1; l < 8; ++l) {
        mbar_wait(&sm.d_ready, pd);
        const uint32_t pd_l = pd;
        pd ^= 1;
        if (tr) a.trace[l * 8] = clock64();
        tc::fence_after_thread_sync();
        const uint32_t dsrc = trow + (uint32_t)((l - 1) & 1) * 256;  // GEMM index l-1 -> TMEM half
        const uint32_t dsrc0 = dsrc - hw * 16;                       // column 0 of that accumulator
        const bool last = (l == 7);
        const bool feed = !last || a.with_view;
        const float* la = nullptr;
        const float* lb = nullptr;
        if (last && a.in.local_alpha && valid) {  // [B,H,W,S,256] / explicit points [B,N,256]
          la = a.in.local_alpha + (samp0 + m) * SW;
          lb = a.in.local_beta + (samp0 + m) * SW;
        }
        // one group = 8 channels of this thread's row: FiLM + sin (pre-activations acc8), heads, taps,
        // hi/lo split and store into the next A operand, first-half publish
        auto finish_group = [&](int j, int g8, int n0, float* v) {
          if (tr && l == 3) a.trace[384 + j * 8 + 1 + g8 * 2] = clock64();
          if (last) {
            // sdf head sees the un-modulated h8 (volume_renderer.py:206-208, 217-220)
#pragma unroll
            for (int i = 0; i < 8; ++i) sdf_acc = fmaf(sm.wsig[n0 + i], v[i], sdf_acc);
            if (MODE == 1 && a.p_h8 && valid) {  // SirenGenerator.forward_generator's output (volume_renderer.py:168-194)
              float4* dst = reinterpret_cast<float4*>(a.p_h8 + (samp0 + m) * SW + n0);
              dst[0] = make_float4(v[0], v[1], v[2], v[3]);
              dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
            if (la) {  // two 16-byte loads per operand (n0 is a multiple of 8, rows are 1 KB apart)
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(la + n0)), a1 = __ldg(reinterpret_cast<const float4*>(la + n0) + 1);
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(lb + n0)), b1 = __ldg(reinterpret_cast<const float4*>(lb + n0) + 1);
              const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
              const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __fadd_rn(__fmul_rn(__fadd_rn(al[i], 1.f), v[i]), be[i]);
            }
          }
          if (taps && (l & 1) == 0) store_tap(l >> 1, n0, v);
          if (feed) store_a8_v<EPI>(sm, m, n0, v);
          if (feed && g8 == 0) publish_first_half(j);
          if (tr && l == 3) a.trace[384 + j * 8 + 2 + g8 * 2] = clock64();
        };
        auto packed_group = [&](int j, int g8, const uint32_t* r8, const float4* g2, const float4* b2) {
          const int n0 = j * 64 + col_of(g8);
          float v[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 g4 = g2[h], b4 = b2[h];
            f32x2 a0, a1;
            film_sin2(pk2u(r8[h * 4], r8[h * 4 + 1]), pk2(g4.x, g4.y), pk2(b4.x, b4.y), v[h * 4], v[h * 4 + 1], a0);
            film_sin2(pk2u(r8[h * 4 + 2], r8[h * 4 + 3]), pk2(g4.z, g4.w), pk2(b4.z, b4.w), v[h * 4 + 2],
                      v[h * 4 + 3], a1);
            if (STASH && stash) {
              float t0, t1, t2, t3;
              unpk2(a0, t0, t1);
              unpk2(a1, t2, t3);
              float* sp = stash + (size_t)(l * SW + n0 + h * 4) * TCM;
              sp[0] = t0, sp[TCM] = t1, sp[2 * TCM] = t2, sp[3 * TCM] = t3;
            }
          }
          finish_group(j, g8, n0, v);
        };
#pragma unroll 1
        for (int jj = 0; jj < 4; ++jj) {
          const int j = NS ? ((jj & 1) * 2 + (jj >> 1)) : jj;  // N-split: blocks 0, 2 (first half), then 1, 3
          if (NS && jj == 2) {
            mbar_wait(&sm.d_ready1, pd_l);
            tc::fence_after_thread_sync();
          }
          if (EPI & 1) {
            uint32_t accr[16];
            float4 fg[4], fb[4];
            auto film_rows = [&](int i) {
              const int n = j * 64 + col_of(i >> 1) + (i & 1) * 4;
              fg[i] = *reinterpret_cast<const float4*>(&sm.film[l][0][n]);
              fb[i] = *reinterpret_cast<const float4*>(&sm.film[l][1][n]);
            };
            if (EPI & 2) {
              // TMEM reads run at ~64 B/clk per SM: sixteen warps asking for 16 columns each wait ~500
              // cycles.  The 8 columns of the first half-block are requested alone, the other 8 land under
              // their processing.
              tc::tmem_ld_32x8_issue(dsrc0 + dcb(j) + hw * 8, accr);
              film_rows(0), film_rows(1);
              tc::tmem_ld_wait8(accr);
              tc::tmem_ld_32x8_issue(dsrc0 + dcb(j) + 32 + hw * 8, accr + 8);
              film_rows(2), film_rows(3);
              if (tr && l == 3) a.trace[384 + j * 8] = clock64();
              packed_group(j, 0, accr, fg, fb);
              tc::tmem_ld_wait8(accr + 8);
              packed_group(j, 1, accr + 8, fg + 2, fb + 2);
            } else {
              tc::tmem_ld_32x16_issue(dsrc + j * 64, accr);
#pragma unroll
              for (int i = 0; i < 4; ++i) film_rows(i);
              tc::tmem_ld_wait16(accr);
              if (tr && l == 3) a.trace[384 + j * 8] = clock64();
              packed_group(j, 0, accr, fg, fb);
              packed_group(j, 1, accr + 8, fg + 2, fb + 2);
            }
          } else {
            float acc[16];
            tc::tmem_ld_32x16(dsrc + j * 64, acc);
            if (tr && l == 3) a.trace[384 + j * 8] = clock64();
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
              const int n0 = j * 64 + col_of(g8);
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float arg = fmaf(sm.film[l][0][n0 + i], acc[g8 * 8 + i], sm.film[l][1][n0 + i]);
                if (STASH && stash) stash[(size_t)(l * SW + n0 + i) * TCM] = arg;
                v[i] = no_sin ? arg * 1e-3f : sin_mufu_reduced(arg);
              }
              finish_group(j, g8, n0, v);
            }
          }
          if (feed) publish(j);
          if (tr && l == 3) a.trace[384 + j * 8 + 5] = clock64();
          if (tr) a.trace[l * 8 + 1 + j] = clock64();
        }
        if (!feed) tc::fence_before_thread_sync();
      }
      sm.sdf_part[hw][m] = sdf_acc;
      compute_sync();
      if (a.with_view) {  // every warp is past layer 7: FiLM rows 0..2 are free until the next tile
        for (int i = ct; i < 3 * SW; i += TC_COMPUTE) {
          wvs[i] = __ldg(pk + OFF_WVDN + i);
          wvs[3 * SW + i] = __ldg(pk + OFF_WRGB + i);
        }
        if (MODE != 0) compute_sync();  // (MODE 0 has the barriers of the alpha / scan steps in between)
      }

      // ---- sdf -> sigma -> alpha -> transmittance scan (overlaps the view-layer MMAs) ----
      if (MODE == 0) {
        if (hw == 0 && valid) {
          const float sd = (sm.sdf_part[0][m] + sm.sdf_part[1][m]) + (sm.sdf_part[2][m] + sm.sdf_part[3][m]) + pk[OFF_HEADB];
          float al;
          if (P.flags & E3_RENDER_NO_SDF) {
            const float sp = (sd > 20.f) ? sd : log1pf(expf(sd));
            al = 1.f - expf(-sp * sm.dist[m]);
          } else {
            const float beta = a.in.sigmoid_beta[0];
            const float sigma = __fdiv_rn(sigmoid_acc(__fdiv_rn(-sd, beta)), beta);
            al = 1.f - expf(-sigma * sm.dist[m]);
          }
          sm.alpha[m] = al;
          if (a.out.sdf) a.out.sdf[samp0 + m] = sd;
        }
        compute_sync();
        if (ct < n_units) {
          const int rr = ct;
          float T = 1.f, wsum = 0.f;
          for (int si = 0; si < S; ++si) {
            const int mm = rr * S + si;
            const float al = sm.alpha[mm];
            float w = __fmul_rn(al, T);
            if ((P.flags & E3_RENDER_FORCE_BACKGROUND) && !(P.flags & E3_RENDER_NO_FORCE_STOP) && si == S - 1)
              w = __fsub_rn(1.f, wsum);
            sm.vis[mm] = T;
            sm.wgt[mm] = w;
            wsum = __fadd_rn(wsum, w);
            T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, al), 1e-10f));
          }
          float depth = 0.f, xs = 0.f, ys = 0.f, zs = 0.f;
          for (int si = 0; si < S; ++si) {
            const int mm = rr * S + si;
            const float w = sm.wgt[mm], z = sm.z[mm];
            depth = fmaf(w, z, depth);
            xs = fmaf(w, __fadd_rn(sm.ray_o[0][rr], __fmul_rn(sm.ray_d[0][rr], z)), xs);
            ys = fmaf(w, __fadd_rn(sm.ray_o[1][rr], __fmul_rn(sm.ray_d[1][rr], z)), ys);
            zs = fmaf(w, __fadd_rn(sm.ray_o[2][rr], __fmul_rn(sm.ray_d[2][rr], z)), zs);
          }
          const size_t pix = (size_t)b * HW + unit0 + rr;
          if (a.out.depth) a.out.depth[pix] = depth;
          if (a.out.mask) a.out.mask[pix] = (depth < P.mask_depth) ? 1.f : 0.f;
          if (a.out.xyz) {
            float* o = a.out.xyz + (size_t)b * 3 * HW + unit0 + rr;
            o[0] = xs, o[HW] = ys, o[2 * (size_t)HW] = zs;
          }
        }
        compute_sync();
        if (hw == 0 && valid) {
          if (a.out.hit_prob) a.out.hit_prob[samp0 + m] = sm.wgt[m];
          if (a.out.visibility) a.out.visibility[samp0 + m] = sm.vis[m];
        }
      } else {
        if (hw == 0 && valid)
          a.p_sdf[samp0 + m] = (sm.sdf_part[0][m] + sm.sdf_part[1][m]) + (sm.sdf_part[2][m] + sm.sdf_part[3][m]) + pk[OFF_HEADB];
      }

      if (a.with_view) {
        // ---- view layer epilogue: + W_dir*viewdir, FiLM, sin; rgb head; weighted feature sum ----
        mbar_wait(&sm.d_ready, pd);
        if (NS) mbar_wait(&sm.d_ready1, pd);
        pd ^= 1;
        if (tr) a.trace[64] = clock64();
        tc::fence_after_thread_sync();
        const uint32_t dsrc = trow + 256u;  // GEMM index 7 -> TMEM half 1
        const float wrow = (MODE == 0 && valid) ? sm.wgt[m] : 0.f;
        float* fbuf = reinterpret_cast<float*>(sm.a_hi);  // [256][128] fp32 over a_hi + a_lo
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        f32x2 C0 = pk2(0.f, 0.f), C1 = C0, C2 = C0;  // packed rgb-head sums (even, odd channels)
        // the TMEM read of block j+1 is in flight while block j is processed (this epilogue is on the
        // tile's critical path: nothing else runs on the SM)
        uint32_t nxt[16];
        if (EPI & 1) tc::tmem_ld_32x16_issue(dsrc + dcb(0), nxt);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          float acc[16];
          if (EPI & 1) {
            tc::tmem_ld_wait16(nxt);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(nxt[i]);
            if (j < 3) tc::tmem_ld_32x16_issue(dsrc + dcb(j + 1), nxt);
          } else {
            tc::tmem_ld_32x16(dsrc + j * 64, acc);
          }
          const int nb = j * 64 + hw * 16;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int n = nb + j4 * 4;
            const float4 d0 = *reinterpret_cast<const float4*>(wvs + n);
            const float4 d1 = *reinterpret_cast<const float4*>(wvs + SW + n);
            const float4 d2 = *reinterpret_cast<const float4*>(wvs + 2 * SW + n);
            const float4 r0 = *reinterpret_cast<const float4*>(wvs + 3 * SW + n);
            const float4 r1 = *reinterpret_cast<const float4*>(wvs + 4 * SW + n);
            const float4 r2 = *reinterpret_cast<const float4*>(wvs + 5 * SW + n);
            const float e0[4] = {d0.x, d0.y, d0.z, d0.w}, e1[4] = {d1.x, d1.y, d1.z, d1.w},
                        e2[4] = {d2.x, d2.y, d2.z, d2.w};
            const float q0[4] = {r0.x, r0.y, r0.z, r0.w}, q1[4] = {r1.x, r1.y, r1.z, r1.w},
                        q2[4] = {r2.x, r2.y, r2.z, r2.w};
            if (EPI & 1) {  // two channels per instruction
              const float4 g4 = *reinterpret_cast<const float4*>(&sm.film[8][0][n]);
              const float4 b4 = *reinterpret_cast<const float4*>(&sm.film[8][1][n]);
              const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
              const f32x2 V0 = pk2(v0, v0), V1 = pk2(v1, v1), V2 = pk2(v2, v2);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int i = 2 * h;
                f32x2 pre = pk2(acc[j4 * 4 + i], acc[j4 * 4 + i + 1]);
                pre = fma2(pk2(e0[i], e0[i + 1]), V0, pre);
                pre = fma2(pk2(e1[i], e1[i + 1]), V1, pre);
                pre = fma2(pk2(e2[i], e2[i + 1]), V2, pre);
                float f0, f1;
                f32x2 arg8;
                film_sin2(pre, pk2(gg[i], gg[i + 1]), pk2(bb[i], bb[i + 1]), f0, f1, arg8);
                if (STASH && stash) {
                  float t0, t1;
                  unpk2(arg8, t0, t1);
                  stash[(size_t)(8 * SW + n + i) * TCM] = t0;
                  stash[(size_t)(8 * SW + n + i + 1) * TCM] = t1;
                }
                const f32x2 ff = pk2(f0, f1);
                C0 = fma2(pk2(q0[i], q0[i + 1]), ff, C0);
                C1 = fma2(pk2(q1[i], q1[i + 1]), ff, C1);
                C2 = fma2(pk2(q2[i], q2[i + 1]), ff, C2);
                acc[j4 * 4 + i] = f0, acc[j4 * 4 + i + 1] = f1;
              }
            }
#pragma unroll
            for (int i = 0; i < ((EPI & 1) ? 0 : 4); ++i) {
              float pre = acc[j4 * 4 + i];
              pre = fmaf(e0[i], v0, pre);
              pre = fmaf(e1[i], v1, pre);
              pre = fmaf(e2[i], v2, pre);
              const float arg8 = fmaf(sm.film[8][0][n + i], pre, sm.film[8][1][n + i]);
              if (STASH && stash) stash[(size_t)(8 * SW + n + i) * TCM] = arg8;
              const float f = sin_mufu_reduced(arg8);
              c0 = fmaf(q0[i], f, c0);
              c1 = fmaf(q1[i], f, c1);
              c2 = fmaf(q2[i], f, c2);
              acc[j4 * 4 + i] = f;
            }
          }
          if (MODE == 1) {
            if (a.p_feat && valid) {
              float4* dst = reinterpret_cast<float4*>(a.p_feat + (samp0 + m) * SW + nb);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
            }
          } else {
            // stage w*f for the per-ray sum (all view-layer MMAs have completed — d_ready — so the
            // A-operand region is free; each thread owns its (n, m) slots, no cross-thread hazard)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = nb + i;
              fbuf[n * TCM + (m ^ (n & 31))] = wrow * acc[i];
            }
          }
        }
        if (EPI & 1) {
          float lo, hi;
          unpk2(C0, lo, hi), c0 = lo + hi;
          unpk2(C1, lo, hi), c1 = lo + hi;
          unpk2(C2, lo, hi), c2 = lo + hi;
        }
        sm.rgb_part[hw][0][m] = c0;
        sm.rgb_part[hw][1][m] = c1;
        sm.rgb_part[hw][2][m] = c2;
        tc::fence_before_thread_sync();
        if (tr) a.trace[66] = clock64();
        compute_sync();
        if (tr) a.trace[67] = clock64();

        if (MODE == 0) {
          // rgb head: one thread per (colour, sample) closes the four column-quarter partial sums, writes the
          // raw value and leaves w * sigmoid(raw) in place for the per-ray sums below (the 24 sigmoids of a
          // ray used to run one after the other in a single thread)
          if (ct < 3 * TCM) {
            const int c = ct >> 7, mm = ct & (TCM - 1);
            const float raw = (sm.rgb_part[0][c][mm] + sm.rgb_part[1][c][mm]) +
                              (sm.rgb_part[2][c][mm] + sm.rgb_part[3][c][mm]) + pk[OFF_HEADB + 1 + c];
            if (mm < n_valid) {
              if (a.out.raw_rgb) a.out.raw_rgb[(samp0 + mm) * 3 + c] = raw;
              sm.rgb_part[0][c][mm] = __fmul_rn(sm.wgt[mm], sigmoid_acc(raw));
            }
          }
          if (a.out.features) {
            const int n = ct & (SW - 1);  // one output channel per thread pair: even / odd rays
            const float* row = fbuf + n * TCM;
            const int sw = n & 31;
            float* o = a.out.features + ((size_t)b * SW + n) * HW + unit0;
            for (int rr = ct >> 8; rr < n_units; rr += 2) {
              float acc0 = 0.f, acc1 = 0.f;  // two chains: the shared-memory latency overlaps
              int si = 0;
              for (; si + 1 < S; si += 2) {
                acc0 += row[(rr * S + si) ^ sw];
                acc1 += row[(rr * S + si + 1) ^ sw];
              }
              if (si < S) acc0 += row[(rr * S + si) ^ sw];
              o[rr] = acc0 + acc1;
            }
          }
          if (a.out.thumb_rgb) {
            compute_sync();
            if (ct < 3 * n_units) {
              const int c = ct / n_units, rr = ct - c * n_units;
              float accum = 0.f;
              for (int si = 0; si < S; ++si) accum = __fadd_rn(accum, sm.rgb_part[0][c][rr * S + si]);
              a.out.thumb_rgb[((size_t)b * 3 + c) * HW + unit0 + rr] = -1.f + 2.f * accum;
            }
          }
        } else if (a.p_rgb && hw == 0 && valid) {
          float* o = a.p_rgb + (samp0 + m) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            o[c] = (sm.rgb_part[0][c][m] + sm.rgb_part[1][c][m]) + (sm.rgb_part[2][c][m] + sm.rgb_part[3][c][m]) +
                   pk[OFF_HEADB + 1 + c];
        }
      }
      if (tr) a.trace[65] = clock64();
      compute_sync();  // shared memory is reused by the next tile
    }
  }

  tc::fence_before_thread_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA exits while a peer may still multicast into its ring
  if (warp == 1) {
    if (PAIR) tc::tmem_dealloc_pair(tmem_base, 512);
    else tc::tmem_dealloc(tmem_base, 512);
  }
}

template <int MODE, int CL, bool STASH = false, int EPI = 0>
static int launch_tc_variant(const RenderArgs& a, cudaStream_t stream) {
  static thread_local bool attr_set_dev[E3_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_set_dev[device_slot()];
  const int smem_bytes = (int)sizeof(SmemTC) + 1024;
  auto* fn = siren_render_tc_kernel<MODE, CL, STASH, EPI>;
  if (!attr_set) {
    E3_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(TC_NTHREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent kernel: never more CTAs than can be co-resident as whole clusters (GPCs of 16/18/20
  // SMs cannot all be tiled by clusters of 4)
  static thread_local int max_clusters_dev[E3_MAX_DEVICES] = {};
  int& max_clusters = max_clusters_dev[device_slot()];
  if (!max_clusters) {
    cfg.gridDim = dim3((sm_count() / CL) * CL);
    int n = 0;
    E3_CUDA(cudaOccupancyMaxActiveClusters(&n, fn, &cfg));
    max_clusters = n > 0 ? n : 1;
  }
  int clusters = (a.n_tiles + CL - 1) / CL;
  if (clusters > max_clusters) clusters = max_clusters;
  cfg.gridDim = dim3(clusters * CL);
  // weight stream as a 2-D tensor: [8 layers * 16 tiles * 128 rows][64 bf16]
  CUtensorMap wmap;
  const uint64_t wdims[2] = {64, (uint64_t)8 * TC_TILES_PER_LAYER * 128};
  const uint64_t wstr[1] = {128};
  const uint32_t wbox[2] = {64, 128};
  int rc = make_tensor_map_bf16(&wmap, a.packed + OFF_TC_STREAM, 2, wdims, wstr, wbox, /*swizzle128=*/false);
  if (rc) return rc;
  CUtensorMap wmap64;  // N-split pairs: 64-row pieces of the same stream
  const uint32_t wbox64[2] = {64, 64};
  rc = make_tensor_map_bf16(&wmap64, a.packed + OFF_TC_STREAM, 2, wdims, wstr, wbox64, /*swizzle128=*/false);
  if (rc) return rc;
  static int use_wmap = -1;
  if (use_wmap < 0) {
    const char* e = getenv("E3DGE_RENDER_WSTREAM");  // measurement aid: "bulk" | "tensor"
    use_wmap = (e && e[0] == 'b') ? 0 : 1;
  }
  E3_CUDA(cudaLaunchKernelEx(&cfg, fn, a, wmap, wmap64, use_wmap));
  return E3_OK;
}

// Cluster size of the shared weight stream; E3DGE_RENDER_CLUSTER=1|2|4 overrides (measurement aid).
static int render_cluster_size() {
  static int cached = 0;
  if (!cached) {
    const char* e = getenv("E3DGE_RENDER_CLUSTER");
    const int v = e ? atoi(e) : 0;
    cached = (v == 1 || v == 2 || v == 4) ? v : 1;
  }
  return cached;
}

int launch_render_tc(const RenderArgs& a_in, int mode, cudaStream_t stream) {
  if (a_in.n_tiles <= 0) return E3_OK;
  RenderArgs a = a_in;
#ifdef E3_TRACE  // measurement build only (profiles/trace_render.py): clock64 samples through a raw device pointer
  if (const char* tp = getenv("E3DGE_RENDER_TRACE_PTR")) a.trace = reinterpret_cast<unsigned long long*>(strtoull(tp, nullptr, 0));
#else
  a.trace = nullptr;
#endif
  const int cl = render_cluster_size();
  static int epi = -1;
  if (epi < 0) {
    // epilogue / MMA-issue variant (template parameter EPI): 7 = CTA pairs, whole 256-column layers (default),
    // 15 = CTA pairs with N-split layers (measured slower: 2.83 vs 2.37 ms, profiles/r02_ab_render_nsplit.txt —
    // the 128-column MMAs re-read the A operand twice as often and the shared-memory pipe paces them),
    // 3 = one CTA per MMA stream, 0 = the scalar epilogue the what-if flags and the multicast weight-stream clusters
    // (E3DGE_RENDER_CLUSTER) apply to
    const char* e = getenv("E3DGE_RENDER_EPI");
    epi = e ? atoi(e) : 7;
    if (epi != 0 && epi != 3 && epi != 15) epi = 7;
  }
  // every variant computes each value with the same operations, but the order of the accumulations (MMA
  // issue order, head sums per thread-to-column mapping) differs: the training forward (stash) is the
  // same variant as the inference forward, so that the two agree bit for bit
  if (a.stash) {
    if (epi == 0) return mode == 0 ? launch_tc_variant<0, 1, true, 0>(a, stream) : launch_tc_variant<1, 1, true, 0>(a, stream);
    if (epi == 3) return mode == 0 ? launch_tc_variant<0, 1, true, 3>(a, stream) : launch_tc_variant<1, 1, true, 3>(a, stream);
    if (epi == 7) return mode == 0 ? launch_tc_variant<0, 2, true, 7>(a, stream) : launch_tc_variant<1, 2, true, 7>(a, stream);
    return mode == 0 ? launch_tc_variant<0, 2, true, 15>(a, stream) : launch_tc_variant<1, 2, true, 15>(a, stream);
  }
  if (epi == 15) return mode == 0 ? launch_tc_variant<0, 2, false, 15>(a, stream) : launch_tc_variant<1, 2, false, 15>(a, stream);
  if (epi == 7) return mode == 0 ? launch_tc_variant<0, 2, false, 7>(a, stream) : launch_tc_variant<1, 2, false, 7>(a, stream);
  if (epi == 3) return mode == 0 ? launch_tc_variant<0, 1, false, 3>(a, stream) : launch_tc_variant<1, 1, false, 3>(a, stream);
  if (mode == 0) {
    if (cl == 4) return launch_tc_variant<0, 4>(a, stream);
    if (cl == 2) return launch_tc_variant<0, 2>(a, stream);
    return launch_tc_variant<0, 1>(a, stream);
  }
  return launch_tc_variant<1, 1>(a, stream);
}

}  // namespace e3
