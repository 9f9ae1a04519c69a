// Backward of the modulated-conv decoder (channels-last fp32 activations), sm_100a.
//
// What autograd gives the reference through Decoder.forward (stylesdf_model.py:742-797) when the
// E3DGE runners back-propagate image losses into the encoders (trainer.py:881-900, generator frozen at :1569): gradients with
// respect to the layer input (-> the renderer's feature map) and the per-layer latent
// (ModulatedConv2d.modulation, stylesdf_model.py:319).  The generator weights are frozen on this
// path: no weight, noise-strength or bias gradients.
//
// With the forward formulation of modconv.cu,
//     y = act( d[b,o] * C[b,o,p] + nw*noise[p] + bias[o] ),   C = conv(x * s[b,:], W*scale),
// the adjoint per layer is
//     da  = dy * act'(y)                               (gate on the sign of the saved output)
//     dd  = sum_p da * C,  C recovered from y: (pre - nw*noise - bias)/d     -> no stashed C
//     dxs = conv^T(da * d)        the SAME implicit-GEMM kernels with a transposed weight image:
//                                 plain conv: 3x3 conv with flipped taps (pack layout 2);
//                                 up-conv: the 4x4 blur adjoint written parity-planar, then nine
//                                 dense shifted reads instead of a stride-2 gather (pack layout 3)
//     dx  = dxs * s,   ds = sum_p dxs * x
//     dlatent = (ds - s * scale^2 * sum_o dd_o d_o^3 wsq[o,:]) . mod_w / sqrt(512)
// Channel sums over pixels are two-stage (per-chunk partials, then a fixed-order sum): no atomics,
// bit-reproducible.
#include <cuda_bf16.h>

#include "modconv.cuh"
#include "tcgen05.cuh"

namespace e3 {
namespace {

constexpr float kSqrt2 = 1.41421356237309515f;
constexpr int RED_CHUNKS = 64;  // pixel chunks per image of the two-stage channel sums

struct ActBwdArgs {
  const float* dy;  // [B,HW,cout]
  const float* y;
  const float* d;         // [B,cout]
  const float* noise;     // [HW] (+ b*noise_bstride)
  int64_t noise_bstride;
  const float* noise_w;
  const float* act_bias;  // NULL: bare modulated conv (y = d*C)
  float* da;
  float* partial;  // NULL or [B][RED_CHUNKS][cout]
  int B, HW, cout;
};

// sum of one float4 accumulator per thread over the threads that share a channel group, via smem
__device__ __forceinline__ void reduce_store4(float4 acc, int c4, int ps, int c4n, int n_ps, float* smem,
                                              float* dst) {
  float4* sm4 = reinterpret_cast<float4*>(smem);
  if (ps < n_ps) sm4[ps * c4n + c4] = acc;
  __syncthreads();
  if (ps == 0) {
    float4 t = sm4[c4];
    for (int k = 1; k < n_ps; ++k) {
      const float4 v = sm4[k * c4n + c4];
      t.x += v.x, t.y += v.y, t.z += v.z, t.w += v.w;
    }
    *reinterpret_cast<float4*>(dst + c4 * 4) = t;
  }
}

__global__ void __launch_bounds__(256) act_bwd_kernel(const __grid_constant__ ActBwdArgs a) {
  __shared__ __align__(16) float red[256 * 4];
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int c4n = a.cout >> 2, n_ps = 256 / c4n;
  const int c4 = threadIdx.x % c4n, ps = threadIdx.x / c4n;
  const int per = (a.HW + RED_CHUNKS - 1) / RED_CHUNKS;
  const int p0 = chunk * per, p1 = min(a.HW, p0 + per);
  const bool linear = a.act_bias == nullptr;
  const float nw = linear ? 0.f : a.noise_w[0];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ps < n_ps) {
    const float4 dv = *reinterpret_cast<const float4*>(a.d + (size_t)b * a.cout + c4 * 4);
    const float inv_d[4] = {1.f / dv.x, 1.f / dv.y, 1.f / dv.z, 1.f / dv.w};
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (!linear) {
      const float4 bv = *reinterpret_cast<const float4*>(a.act_bias + c4 * 4);
      bias[0] = bv.x, bias[1] = bv.y, bias[2] = bv.z, bias[3] = bv.w;
    }
    for (int p = p0 + ps; p < p1; p += n_ps) {
      const size_t off = ((size_t)b * a.HW + p) * a.cout + c4 * 4;
      const float4 g4 = *reinterpret_cast<const float4*>(a.dy + off);
      const float4 y4 = *reinterpret_cast<const float4*>(a.y + off);
      const float g[4] = {g4.x, g4.y, g4.z, g4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w};
      float da[4], cc[4];
      if (linear) {
#pragma unroll
        for (int i = 0; i < 4; ++i) da[i] = g[i], cc[i] = yv[i] * inv_d[i];
      } else {
        const float nz = nw * a.noise[(size_t)b * a.noise_bstride + p];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool pos = yv[i] > 0.f;
          da[i] = g[i] * (pos ? kSqrt2 : 0.2f * kSqrt2);
          const float pre = yv[i] * (pos ? (1.f / kSqrt2) : (1.f / (0.2f * kSqrt2)));
          cc[i] = (pre - nz - bias[i]) * inv_d[i];
        }
      }
      *reinterpret_cast<float4*>(a.da + off) = make_float4(da[0], da[1], da[2], da[3]);
      acc.x = fmaf(da[0], cc[0], acc.x), acc.y = fmaf(da[1], cc[1], acc.y);
      acc.z = fmaf(da[2], cc[2], acc.z), acc.w = fmaf(da[3], cc[3], acc.w);
    }
  }
  if (a.partial) reduce_store4(acc, c4, ps, c4n, n_ps, red, a.partial + ((size_t)b * RED_CHUNKS + chunk) * a.cout);
}

// g [B,HW,C] (in: dL/d(x*s); out: dL/dx = g*s), partial[b][chunk][c] = sum_p g_in * x
__global__ void __launch_bounds__(256) modgrad_kernel(float* __restrict__ g, const float* __restrict__ x,
                                                      const float* __restrict__ s, float* __restrict__ partial,
                                                      int HW, int C) {
  __shared__ __align__(16) float red[256 * 4];
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int c4n = C >> 2, n_ps = 256 / c4n;
  const int c4 = threadIdx.x % c4n, ps = threadIdx.x / c4n;
  const int per = (HW + RED_CHUNKS - 1) / RED_CHUNKS;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ps < n_ps) {
    const float4 sv = *reinterpret_cast<const float4*>(s + (size_t)b * C + c4 * 4);
    for (int p = p0 + ps; p < p1; p += n_ps) {
      const size_t off = ((size_t)b * HW + p) * C + c4 * 4;
      const float4 g4 = *reinterpret_cast<const float4*>(g + off);
      const float4 x4 = *reinterpret_cast<const float4*>(x + off);
      acc.x = fmaf(g4.x, x4.x, acc.x), acc.y = fmaf(g4.y, x4.y, acc.y);
      acc.z = fmaf(g4.z, x4.z, acc.z), acc.w = fmaf(g4.w, x4.w, acc.w);
      *reinterpret_cast<float4*>(g + off) = make_float4(g4.x * sv.x, g4.y * sv.y, g4.z * sv.z, g4.w * sv.w);
    }
  }
  reduce_store4(acc, c4, ps, c4n, n_ps, red, partial + ((size_t)b * RED_CHUNKS + chunk) * C);
}

__global__ void chunk_sum_kernel(const float* __restrict__ partial, float* __restrict__ out, int B, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int b = idx / C, c = idx - b * C;
  float acc = 0.f;
  for (int k = 0; k < RED_CHUNKS; ++k) acc += partial[((size_t)b * RED_CHUNKS + k) * C + c];
  out[idx] = acc;
}

// Adjoint of the 4x4 [1,3,3,1] blur (gain 4, pad (1,1)) that follows conv_transpose2d stride 2
// (stylesdf_model.py:283-291, 331-346), times the demodulation, written parity-planar and split:
//   dT[b,t,u,o] = d[b,o] * sum_{i,j} da[b, t-i+1, u-j+1, o] * kb[i]*kb[j],  t,u in [0, 2H]
//   planar index [(t&1)*2 + (u&1)][b][t>>1][u>>1][o]; entries with t > 2H or u > 2W are zero.
// SPLIT: bf16 hi / lo halves for the tensor-core GEMM; otherwise one fp32 image (`hi` reinterpreted)
// for the CUDA-core GEMM.
template <bool SPLIT>
__global__ void __launch_bounds__(256) upconv_dT_kernel(const float* __restrict__ da, const float* __restrict__ d,
                                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                        int B, int H, int W, int cout) {
  const int c8n = cout >> 3, OH = 2 * H, OW = 2 * W;
  const int64_t total = (int64_t)4 * B * (H + 1) * (W + 1) * c8n;
  const float kb[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c8n) * 8;
    int64_t r = idx / c8n;
    const int l = (int)(r % (W + 1));
    r /= (W + 1);
    const int k = (int)(r % (H + 1));
    r /= (H + 1);
    const int b = (int)(r % B), plane = (int)(r / B);
    const int t = 2 * k + (plane >> 1), u = 2 * l + (plane & 1);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (t <= OH && u <= OW) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int Y = t - i + 1;
        if (Y < 0 || Y >= OH) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int X = u - j + 1;
          if (X < 0 || X >= OW) continue;
          const float wgt = kb[i] * kb[j];
          const float* src = da + (((size_t)b * OH + Y) * OW + X) * cout + c;
          const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
          acc[0] = fmaf(wgt, v0.x, acc[0]), acc[1] = fmaf(wgt, v0.y, acc[1]);
          acc[2] = fmaf(wgt, v0.z, acc[2]), acc[3] = fmaf(wgt, v0.w, acc[3]);
          acc[4] = fmaf(wgt, v1.x, acc[4]), acc[5] = fmaf(wgt, v1.y, acc[5]);
          acc[6] = fmaf(wgt, v1.z, acc[6]), acc[7] = fmaf(wgt, v1.w, acc[7]);
        }
      }
    }
    const size_t o = ((((size_t)plane * B + b) * (H + 1) + k) * (W + 1) + l) * cout + c;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= d[(size_t)b * cout + c + i];
    if (SPLIT) {
      __align__(16) __nv_bfloat16 h8[8], l8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) tc::split_bf16(acc[i], h8[i], l8[i]);
      *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(h8);
      *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(l8);
    } else {
      float* out32 = reinterpret_cast<float*>(hi) + o;
      *reinterpret_cast<float4*>(out32) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(out32 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// ToRGB backward: rgb[b,c,p] = sum_i scale*w[c,i]*s[b,i]*x[b,p,i] + ...   (stylesdf_model.py:531-541)
//   t_i = scale * sum_c w[c,i] * drgb[b,c,p];  dx[b,p,i] = t_i * s[b,i];  partial = sum_p x * t
__global__ void __launch_bounds__(256) torgb_bwd_kernel(const float* __restrict__ drgb, const float* __restrict__ x,
                                                        const float* __restrict__ w, const float* __restrict__ s,
                                                        float* __restrict__ dx, float* __restrict__ partial, int HW,
                                                        int cin) {
  __shared__ __align__(16) float red[256 * 4];
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int c4n = cin >> 2, n_ps = 256 / c4n;
  const int c4 = threadIdx.x % c4n, ps = threadIdx.x / c4n;
  const int per = (HW + RED_CHUNKS - 1) / RED_CHUNKS;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  const float scale = rsqrtf((float)cin);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ps < n_ps) {
    float4 wc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      wc[c] = *reinterpret_cast<const float4*>(w + (size_t)c * cin + c4 * 4);
      wc[c].x *= scale, wc[c].y *= scale, wc[c].z *= scale, wc[c].w *= scale;
    }
    const float4 sv = *reinterpret_cast<const float4*>(s + (size_t)b * cin + c4 * 4);
    for (int p = p0 + ps; p < p1; p += n_ps) {
      const float g0 = drgb[((size_t)b * 3 + 0) * HW + p], g1 = drgb[((size_t)b * 3 + 1) * HW + p],
                  g2 = drgb[((size_t)b * 3 + 2) * HW + p];
      float4 t;
      t.x = fmaf(g2, wc[2].x, fmaf(g1, wc[1].x, g0 * wc[0].x));
      t.y = fmaf(g2, wc[2].y, fmaf(g1, wc[1].y, g0 * wc[0].y));
      t.z = fmaf(g2, wc[2].z, fmaf(g1, wc[1].z, g0 * wc[0].z));
      t.w = fmaf(g2, wc[2].w, fmaf(g1, wc[1].w, g0 * wc[0].w));
      const size_t off = ((size_t)b * HW + p) * cin + c4 * 4;
      const float4 x4 = *reinterpret_cast<const float4*>(x + off);
      acc.x = fmaf(t.x, x4.x, acc.x), acc.y = fmaf(t.y, x4.y, acc.y);
      acc.z = fmaf(t.z, x4.z, acc.z), acc.w = fmaf(t.w, x4.w, acc.w);
      *reinterpret_cast<float4*>(dx + off) = make_float4(t.x * sv.x, t.y * sv.y, t.z * sv.z, t.w * sv.w);
    }
  }
  reduce_store4(acc, c4, ps, c4n, n_ps, red, partial + ((size_t)b * RED_CHUNKS + chunk) * cin);
}

// dlatent[b,:] = ( ds[b,:] - s[b,:] * scale^2 * sum_o dd[b,o] d[b,o]^3 wsq[o,:] ) . mod_w / sqrt(512)
// (adjoint of mod_style_kernel + demod_kernel, modconv.cu; stylesdf_model.py:319-326)
// grid (8, B): every block rebuilds the cin-vector (cheap, keeps one launch) and produces 64 of the
// 512 latent entries, 4 partial sums per entry.
__global__ void __launch_bounds__(256) styles_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ dd,
                                                         const float* __restrict__ s, const float* __restrict__ d,
                                                         const float* __restrict__ wsq, const float* __restrict__ mod_w,
                                                         int cin, int cout, float scale2, float* __restrict__ dlatent) {
  extern __shared__ float sm[];  // e[cout] | dst[cin] | part[4][64]
  float* e = sm;
  float* dst = sm + cout;
  float* part = dst + cin;
  const int b = blockIdx.y, k0 = blockIdx.x * 64;
  if (dd) {
    for (int o = threadIdx.x; o < cout; o += blockDim.x) {
      const float dv = d[(size_t)b * cout + o];
      e[o] = dd[(size_t)b * cout + o] * dv * dv * dv;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cin; i += blockDim.x) {
    float v = ds[(size_t)b * cin + i];
    if (dd) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float* wp = wsq + i;
      int o = 0;
#pragma unroll 2
      for (; o + 4 <= cout; o += 4) {
        a0 = fmaf(e[o], wp[(size_t)o * cin], a0);
        a1 = fmaf(e[o + 1], wp[(size_t)(o + 1) * cin], a1);
        a2 = fmaf(e[o + 2], wp[(size_t)(o + 2) * cin], a2);
        a3 = fmaf(e[o + 3], wp[(size_t)(o + 3) * cin], a3);
      }
      for (; o < cout; ++o) a0 = fmaf(e[o], wp[(size_t)o * cin], a0);
      v -= s[(size_t)b * cin + i] * scale2 * ((a0 + a1) + (a2 + a3));
    }
    dst[i] = v;
  }
  __syncthreads();
  const int kk = threadIdx.x & 63, grp = threadIdx.x >> 6;  // 4 groups split the cin range
  const int per = (cin + 3) / 4, i0 = grp * per, i1 = min(cin, i0 + per);
  float a0 = 0.f, a1 = 0.f;
  int i = i0;
  for (; i + 2 <= i1; i += 2) {
    a0 = fmaf(dst[i], mod_w[(size_t)i * 512 + k0 + kk], a0);
    a1 = fmaf(dst[i + 1], mod_w[(size_t)(i + 1) * 512 + k0 + kk], a1);
  }
  if (i < i1) a0 = fmaf(dst[i], mod_w[(size_t)i * 512 + k0 + kk], a0);
  part[grp * 64 + kk] = a0 + a1;
  __syncthreads();
  if (grp == 0)
    dlatent[(size_t)b * 512 + k0 + kk] =
        ((part[kk] + part[64 + kk]) + (part[128 + kk] + part[192 + kk])) * 0.04419417382415922f;
}

int grid_cap(int64_t blocks) {
  const int64_t cap = (int64_t)sm_count() * 32;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace
}  // namespace e3

using namespace e3;

// scratch = [da: B*OH*OW*cout fp32] [partials: B*64*max(cin,cout) fp32] [bf16 operand halves]
extern "C" size_t e3_styled_conv_bwd_scratch_bytes(int batch, int h, int w, int cin, int cout, int upsample) {
  const size_t opix = (size_t)batch * h * w * (upsample ? 4 : 1);
  const size_t da = align256(opix * cout * sizeof(float));
  const size_t part = align256((size_t)batch * RED_CHUNKS * (cin > cout ? cin : cout) * sizeof(float));
  const size_t split = upsample ? tc_conv_planar_elems(batch, h, w, cout) * 2 * sizeof(__nv_bfloat16)
                                : tc_conv_split_bytes(batch, h, w, cout);
  return da + part + align256(split) + 256;
}

extern "C" int e3_styled_conv3x3_bwd(const float* dy, const float* y, const float* x, const void* wpacked_bwd,
                                     const float* s, const float* d, const float* noise,
                                     int64_t noise_batch_stride, const float* noise_w, const float* act_bias,
                                     float* dx, float* ds, float* dd, int batch, int h, int w, int cin, int cout,
                                     int upsample, void* scratch, size_t scratch_bytes, uint32_t flags,
                                     void* stream) {
  E3_REQUIRE(batch >= 0 && h > 0 && w > 0 && batch <= 65535, E3_ERR_BAD_ARG, "e3_styled_conv3x3_bwd: bad shape");
  E3_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && cin <= 1024 && cout <= 1024, E3_ERR_UNSUPPORTED,
             "e3_styled_conv3x3_bwd: needs cin, cout multiples of 16 and <= 1024 (got cin=%d cout=%d)", cin, cout);
  if (batch == 0) return E3_OK;
  E3_REQUIRE(dy && y && x && wpacked_bwd && s && d && dx && ds, E3_ERR_BAD_ARG, "e3_styled_conv3x3_bwd: null argument");
  E3_REQUIRE(!act_bias || (noise && noise_w), E3_ERR_BAD_ARG,
             "e3_styled_conv3x3_bwd: noise and noise_w are required unless act_bias is NULL");
  const size_t need = e3_styled_conv_bwd_scratch_bytes(batch, h, w, cin, cout, upsample);
  E3_REQUIRE(scratch && scratch_bytes >= need, E3_ERR_SCRATCH, "e3_styled_conv3x3_bwd: scratch %zu < %zu bytes",
             scratch_bytes, need);
  cudaStream_t st = as_stream(stream);
  const int oh = upsample ? 2 * h : h, ow = upsample ? 2 * w : w;
  char* base = reinterpret_cast<char*>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  float* da = reinterpret_cast<float*>(base);
  float* partial = reinterpret_cast<float*>(base + align256((size_t)batch * oh * ow * cout * sizeof(float)));
  char* split = reinterpret_cast<char*>(partial) +
                align256((size_t)batch * RED_CHUNKS * (cin > cout ? cin : cout) * sizeof(float));

  // 1. activation gate, demodulation-factor gradient
  ActBwdArgs ab{dy, y, d, noise, noise_batch_stride, noise_w, act_bias, da, dd ? partial : nullptr, batch, oh * ow, cout};
  act_bwd_kernel<<<dim3(RED_CHUNKS, batch), 256, 0, st>>>(ab);
  E3_CUDA(cudaGetLastError());
  if (dd) {
    chunk_sum_kernel<<<(batch * cout + 255) / 256, 256, 0, st>>>(partial, dd, batch, cout);
    E3_CUDA(cudaGetLastError());
  }

  // 2. dL/d(x*s) = conv^T(da * d): the forward implicit-GEMM kernels on a transposed weight image
  ConvGemmArgs a{};
  a.s = d, a.wg = static_cast<const float*>(wpacked_bwd), a.out = dx;
  a.B = batch, a.H = h, a.W = w, a.Cin = cout, a.N = cin;
  a.mode = 0;
  const bool tcore = !(flags & E3_CONV_FP32_CUDA_CORES) && tc_conv_supported(batch, h, w, cout, cin);
  E3_REQUIRE(tcore || !(flags & E3_CONV_TENSOR_CORES), E3_ERR_UNSUPPORTED,
             "e3_styled_conv3x3_bwd: E3_CONV_TENSOR_CORES requested for an unsupported shape");
  const void* wbf16 = conv_packed_bf16_part(wpacked_bwd, cout, cin);
  int rc;
  if (!upsample) {
    a.x = da;
    if (tcore) {
      if ((rc = tc_conv_launch(a, 9, wbf16, split, st))) return rc;
    } else if ((rc = conv_gemm_ffma_launch(a, 9, st))) {
      return rc;
    }
  } else {
    // the planar image holds 2 bf16 halves or 1 fp32 copy: the same bytes either way
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(split);
    __nv_bfloat16* lo = hi + tc_conv_planar_elems(batch, h, w, cout);
    const int64_t total = (int64_t)tc_conv_planar_elems(batch, h, w, cout) / 8;
    a.planar = 1;
    if (tcore) {
      upconv_dT_kernel<true><<<grid_cap((total + 255) / 256), 256, 0, st>>>(da, d, hi, lo, batch, h, w, cout);
      E3_CUDA(cudaGetLastError());
      if ((rc = tc_conv_launch_presplit(a, 9, wbf16, hi, lo, st))) return rc;
    } else {
      upconv_dT_kernel<false><<<grid_cap((total + 255) / 256), 256, 0, st>>>(da, d, hi, lo, batch, h, w, cout);
      E3_CUDA(cudaGetLastError());
      a.x = reinterpret_cast<const float*>(split);
      if ((rc = conv_gemm_ffma_launch(a, 9, st))) return rc;
    }
  }

  // 3. dx = dxs * s (in place), ds = sum_p dxs * x
  modgrad_kernel<<<dim3(RED_CHUNKS, batch), 256, 0, st>>>(dx, x, s, partial, h * w, cin);
  E3_CUDA(cudaGetLastError());
  chunk_sum_kernel<<<(batch * cin + 255) / 256, 256, 0, st>>>(partial, ds, batch, cin);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" size_t e3_torgb_bwd_scratch_bytes(int batch, int cin) {
  return (size_t)batch * RED_CHUNKS * cin * sizeof(float) + 256;
}

extern "C" int e3_torgb_bwd(const float* drgb, const float* x, const float* weight, const float* s, float* dx,
                            float* ds, int batch, int h, int w, int cin, void* scratch, size_t scratch_bytes,
                            void* stream) {
  E3_REQUIRE(batch >= 0 && h > 0 && w > 0 && cin > 0 && cin % 4 == 0 && cin <= 1024 && batch <= 65535,
             E3_ERR_BAD_ARG, "e3_torgb_bwd: bad shape (cin %% 4 == 0, cin <= 1024 required)");
  if (batch == 0) return E3_OK;
  E3_REQUIRE(drgb && x && weight && s && dx && ds, E3_ERR_BAD_ARG, "e3_torgb_bwd: null argument");
  E3_REQUIRE(scratch && scratch_bytes >= e3_torgb_bwd_scratch_bytes(batch, cin), E3_ERR_SCRATCH,
             "e3_torgb_bwd: scratch too small");
  float* partial = reinterpret_cast<float*>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  cudaStream_t st = as_stream(stream);
  torgb_bwd_kernel<<<dim3(RED_CHUNKS, batch), 256, 0, st>>>(drgb, x, weight, s, dx, partial, h * w, cin);
  E3_CUDA(cudaGetLastError());
  chunk_sum_kernel<<<(batch * cin + 255) / 256, 256, 0, st>>>(partial, ds, batch, cin);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" int e3_modconv_styles_bwd(const float* ds, const float* dd, const float* s, const float* d,
                                     const float* wsq, const float* mod_w, int batch, int cin, int cout, int ksize,
                                     float* dlatent, void* stream) {
  E3_REQUIRE(batch >= 0 && cin > 0 && cin <= 4096 && cout >= 0 && cout <= 4096 && ksize > 0, E3_ERR_BAD_ARG,
             "e3_modconv_styles_bwd: bad shape");
  if (batch == 0) return E3_OK;
  E3_REQUIRE(ds && mod_w && dlatent, E3_ERR_BAD_ARG, "e3_modconv_styles_bwd: null argument");
  E3_REQUIRE(!dd || (s && d && wsq && cout > 0), E3_ERR_BAD_ARG,
             "e3_modconv_styles_bwd: the demodulation term needs s, d, wsq and cout");
  const float scale2 = 1.f / (float)(cin * ksize * ksize);
  styles_bwd_kernel<<<dim3(8, batch), 256, (size_t)(cin + cout + 256) * sizeof(float), as_stream(stream)>>>(
      ds, dd, s, d, wsq, mod_w, cin, cout, scale2, dlatent);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}
