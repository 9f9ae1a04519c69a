// Shared device/host helpers for the e3dge_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/e3dge_b200.h"

namespace e3 {

// ---- host-side error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define E3_REQUIRE(cond, code, ...)          \
  do {                                       \
    if (!(cond)) {                           \
      ::e3::set_error(__VA_ARGS__);          \
      return (code);                         \
    }                                        \
  } while (0)

#define E3_CUDA(call)                                        \
  do {                                                       \
    int _rc = ::e3::check_cuda((call), #call);               \
    if (_rc != 0) return _rc;                                \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();
// slot of the calling thread's current device in per-device caches (function attributes, occupancy)
constexpr int E3_MAX_DEVICES = 64;
int device_slot();

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src,
                                             uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Multicast variant: this CTA fetches `bytes` once and the hardware writes them to the same shared
// memory offset of every CTA in `cta_mask`, signalling complete_tx on each destination's mbarrier.
__device__ __forceinline__ void tma_bulk_g2s_multicast(void* smem_dst, const void* gmem_src,
                                                       uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], "
      "%2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// sin(x) accurate to <= 1.5 ulp for |x| < 2^15: 3-term Cody-Waite reduction to
// [-pi/4, pi/4] + minimax polynomials.  FiLM pre-activations reach |x| ~ 1e2, where
// sin.approx (MUFU) does not hold the 1e-3 end-to-end parity bar.
__device__ __forceinline__ float sin_accurate(float a) {
  float j = fmaf(a, 0.636619747f, 12582912.0f);  // rint(a * 2/pi) via the 1.5*2^23 trick
  const int q = __float_as_int(j);
  j -= 12582912.0f;
  float r = fmaf(j, -1.57079601e+00f, a);
  r = fmaf(j, -3.13916473e-07f, r);
  r = fmaf(j, -5.39030253e-15f, r);
  const float s = r * r;
  float ps = fmaf(-1.9515295891e-4f, s, 8.3321608736e-3f);
  ps = fmaf(ps, s, -1.6666654611e-1f);
  ps = fmaf(ps * s, r, r);
  float pc = fmaf(2.443315711809948e-5f, s, -1.388731625493765e-3f);
  pc = fmaf(pc, s, 4.166664568298827e-2f);
  pc = fmaf(pc, s, -0.5f);
  pc = fmaf(pc, s, 1.0f);
  float res = (q & 1) ? pc : ps;
  return (q & 2) ? -res : res;
}

// sin(x) with |abs error| <= 1.2e-7 for |x| < 2^11: reduction to [-pi/2, pi/2] by a 3-term
// Cody-Waite split of pi, one odd degree-9 polynomial, sign from the parity of the quotient.
// ~13 instructions; used in the tensor-core renderer's epilogue where sin is the critical path.
__device__ __forceinline__ float sin_fast_accurate(float a) {
  const float jm = fmaf(a, 0.318309886f, 12582912.0f);  // rint(a / pi) + 1.5 * 2^23
  const float j = jm - 12582912.0f;
  float r = fmaf(j, -3.140625f, a);
  r = fmaf(j, -9.67502593994140625e-4f, r);
  r = fmaf(j, -1.509957990978376e-7f, r);
  const float s = r * r;
  float p = fmaf(2.612235134960028e-06f, s, -1.981260050515031e-04f);
  p = fmaf(p, s, 8.333111242781598e-03f);
  p = fmaf(p, s, -1.666666179743073e-01f);
  const float res = fmaf(s * r, p, r);
  return __int_as_float(__float_as_int(res) ^ (__float_as_int(jm) << 31));
}

// sin(x) = sin.approx (MUFU) of x reduced to [-pi, pi] by a 2-term Cody-Waite split of 2*pi:
// 6.28125 has 9 significant bits, so j*6.28125 and x - j*6.28125 are exact for |x| < 2^12; the second
// term carries the remainder of 2*pi rounded to fp32 (representation error 6e-11 * j).  On [-pi, pi]
// the hardware approximation has max abs error 2^-21.41 (3.6e-7, CUDA math API); the raw __sinf range
// reduction would add |x| * 6e-8.  6 instructions (FFMA, FADD, 2 FFMA, FMUL.RZ, MUFU.SIN), the
// transcendental on the otherwise idle MUFU pipe.  Used by the tensor-core renderer, whose operands
// are quantised to hi+lo bf16 (2^-17 relative, 20x coarser than this error); the exact-fp32 renderer
// keeps the polynomial version.
__device__ __forceinline__ float reduce_2pi(float a) {
  const float jm = fmaf(a, 0.159154943f, 12582912.0f);  // rint(a / 2pi) + 1.5 * 2^23
  const float j = jm - 12582912.0f;
  const float r = fmaf(j, -6.28125f, a);
  return fmaf(j, -1.9353071795864769e-3f, r);
}
__device__ __forceinline__ float sin_mufu_reduced(float a) { return __sinf(reduce_2pi(a)); }

// x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi), for a pair, packed as the UMMA operands want
// them (element 0 in the low half): 2 F2FP + 2 bit extractions + 2 FADD per pair.
__device__ __forceinline__ void split_pair_bf16(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v1 - h1), "f"(v0 - h0));
}

__device__ __forceinline__ float warp_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
#endif  // __CUDACC__

}  // namespace e3
