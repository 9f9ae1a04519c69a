// Modulated-conv StyleGAN2 decoder kernels (channels-last fp32 activations).
//
// Replaces ModulatedConv2d / StyledConv / ToRGB / NoiseInjection / Blur / Upsample of
//   project/models/stylesdf_model.py:263-362, 365-466, 469-541, 96-165
// (arithmetic spec: SURVEY.md Appendix A.7).
//
// Formulation.  The reference builds per-sample weights W'[b,o,i,k] = scale*W[o,i,k]*s[b,i]
// * d[b,o] and runs a grouped conv with groups = batch.  Here the modulation is moved onto
// the activations and the demodulation onto the output:
//     y[b,o] = d[b,o] * conv(x[b,i] * s[b,i], scale*W[o,i])
// so one shared weight matrix serves the whole batch (a plain implicit GEMM), and
// d[b,o] = rsqrt(scale^2 * sum_i s[b,i]^2 * sum_k W[o,i,k]^2 + 1e-8) is a tiny GEMV.
// The upsampling conv (conv_transpose2d stride 2, then the 4x4 [1,3,3,1] blur) is computed
// at its minimum FLOP count: a 1x1-style GEMM  G[pix, (ky,kx,o)] = sum_i xs[pix,i] *
// W[o,i,ky,kx]  followed by a fused col2im + blur + noise + bias + leaky-ReLU gather.
//
// Two arithmetic back ends behind the same entry points (flags E3_CONV_*):
//   * tensor cores (tc_conv.cu): tcgen05 split-bf16 implicit GEMM, TMA-fed — the fast path for
//     the decoder's real shapes (power-of-two maps, Cin % 64 == 0, N % 128 == 0);
//   * CUDA cores (this file): exact-fp32 FFMA implicit GEMM for every other shape, and the
//     numerical baseline the tensor-core path is tested against.
#include "modconv.cuh"
#include "tcgen05.cuh"

namespace e3 {

constexpr float kSqrt2 = 1.41421356237309515f;

// ---- weight preparation -----------------------------------------------------------------
__global__ void weight_sq_kernel(const float* __restrict__ w, int n_oi, int kk, float* __restrict__ wsq) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_oi) return;
  float acc = 0.f;
  for (int k = 0; k < kk; ++k) {
    const float v = w[(size_t)idx * kk + k];
    acc = fmaf(v, v, acc);
  }
  wsq[idx] = acc;
}

// plain:    Wg[tap][ci][o]        = scale * W[o][ci][tap]
// upsample: Wg[ci][tap*cout + o]  = scale * W[o][ci][tap]
// mode 2 (backward of plain): Wg[tap][o][ci] = scale * W[o][ci][8 - tap]; mode 3: = scale * W[o][ci][tap]
__global__ void conv_pack_kernel(const float* __restrict__ w, int cout, int cin, int upsample,
                                 float scale, float* __restrict__ wg) {
  const int64_t total = (int64_t)cout * cin * 9;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int o, ci, tap;
    if (upsample >= 2) {
      ci = (int)(idx % cin);
      o = (int)((idx / cin) % cout);
      tap = (int)(idx / ((int64_t)cout * cin));
      if (upsample == 2) tap = 8 - tap;
    } else if (upsample) {
      o = (int)(idx % cout);
      tap = (int)((idx / cout) % 9);
      ci = (int)(idx / ((int64_t)cout * 9));
    } else {
      o = (int)(idx % cout);
      ci = (int)((idx / cout) % cin);
      tap = (int)(idx / ((int64_t)cout * cin));
    }
    wg[idx] = scale * w[((size_t)o * cin + ci) * 9 + tap];
  }
}

// s[b,i] = (mod_w[i,:] . latent[b,:]) / sqrt(512) + mod_b[i]   (EqualLinear, stylesdf_model.py:234-244)
__global__ void __launch_bounds__(256) mod_style_kernel(const float* __restrict__ latent,
                                                        int64_t latent_stride,
                                                        const float* __restrict__ mod_w,
                                                        const float* __restrict__ mod_b, int cin,
                                                        float* __restrict__ s) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= cin) return;
  const float* lat = latent + (size_t)b * latent_stride;
  float acc = 0.f;
  for (int j = lane; j < 512; j += 32) acc = fmaf(mod_w[(size_t)i * 512 + j], lat[j], acc);
  acc = warp_sum(acc);
  if (lane == 0) s[(size_t)b * cin + i] = acc * 0.04419417382415922f + mod_b[i];  // 1/sqrt(512)
}

// d[b,o] = rsqrt(scale^2 * sum_i wsq[o,i] * s[b,i]^2 + 1e-8)   (stylesdf_model.py:321-326)
__global__ void __launch_bounds__(256) demod_kernel(const float* __restrict__ wsq,
                                                    const float* __restrict__ s, int cin, int cout,
                                                    float scale2, float* __restrict__ d) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + warp;
  if (o >= cout) return;
  float acc = 0.f;
  for (int i = lane; i < cin; i += 32) {
    const float sv = s[(size_t)b * cin + i];
    acc = fmaf(wsq[(size_t)o * cin + i], sv * sv, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) d[(size_t)b * cout + o] = rsqrtf(fmaf(scale2, acc, 1e-8f));
}

// ---- implicit-GEMM conv on the FFMA pipe ---------------------------------------------------
// C[m][n] = sum_{tap,ci} (x[b, y+dy, x+dx, ci] * s[b,ci]) * Wg[tap][ci][n],  m = (b,y,x)
// 128x128 tile, BK = 16, 256 threads, 8x8 register tile, register-staged double buffering.
constexpr int CG_BM = 128, CG_BN = 128, CG_BK = 16, CG_LDA = CG_BM + 4;

template <int TAPS>
__global__ void __launch_bounds__(256, 2) conv_gemm_ffma_kernel(const __grid_constant__ ConvGemmArgs a) {
  __shared__ __align__(16) float As[2][CG_BK][CG_LDA];
  __shared__ __align__(16) float Bs[2][CG_BK][CG_BN];
  const int tid = threadIdx.x;
  const int tn = tid & 15, tm = tid >> 4;
  const int HW = a.H * a.W;
  const int64_t M = (int64_t)a.B * HW;
  const int64_t m0 = (int64_t)blockIdx.x * CG_BM;
  const int n0 = blockIdx.y * CG_BN;
  const int k_chunks_per_tap = a.Cin / CG_BK;
  const int n_chunks = TAPS * k_chunks_per_tap;

  // A-gather role: one pixel row per thread, 8 channels
  const int ar = tid & 127, ah = tid >> 7;
  const int64_t am = m0 + ar;
  const bool a_row_ok = am < M;
  int ab = 0, ay = 0, ax = 0;
  if (a_row_ok) {
    ab = (int)(am / HW);
    const int p = (int)(am - (int64_t)ab * HW);
    ay = p / a.W;
    ax = p - ay * a.W;
  }
  float4 ra[2], rb[2];

  auto load_chunk = [&](int kc) {
    const int tap = kc / k_chunks_per_tap;
    const int ci0 = (kc - tap * k_chunks_per_tap) * CG_BK + ah * 8;
    int yy = ay, xx = ax;
    if (TAPS == 9) {
      yy += tap / 3 - 1;
      xx += tap % 3 - 1;
    }
    ra[0] = ra[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (TAPS == 9 && a.planar) {
      // up-conv backward: x is the fp32 parity-planar gradient [py][px][B][H+1][W+1][Cin], already
      // multiplied by the demodulation; tap (ky,kx) reads plane (ky&1, kx&1) at (y + (ky>>1), x + (kx>>1))
      if (a_row_ok) {
        const int ky = tap / 3, kx = tap % 3;
        const int plane = (ky & 1) * 2 + (kx & 1);
        const float4* src = reinterpret_cast<const float4*>(
            a.x + ((((size_t)plane * a.B + ab) * (a.H + 1) + ay + (ky >> 1)) * (a.W + 1) + ax + (kx >> 1)) * a.Cin + ci0);
        ra[0] = src[0];
        ra[1] = src[1];
      }
    } else if (a_row_ok && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
      const float4* src =
          reinterpret_cast<const float4*>(a.x + (((size_t)ab * a.H + yy) * a.W + xx) * a.Cin + ci0);
      const float4* sp = reinterpret_cast<const float4*>(a.s + (size_t)ab * a.Cin + ci0);
      const float4 v0 = src[0], v1 = src[1], s0 = sp[0], s1 = sp[1];
      ra[0] = make_float4(v0.x * s0.x, v0.y * s0.y, v0.z * s0.z, v0.w * s0.w);
      ra[1] = make_float4(v1.x * s1.x, v1.y * s1.y, v1.z * s1.z, v1.w * s1.w);
    }
    const float* wrow = a.wg + ((size_t)tap * a.Cin + (kc - tap * k_chunks_per_tap) * CG_BK) * a.N;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = tid + 256 * j, row = f >> 5, c4 = (f & 31) * 4;
      rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + c4 < a.N) rb[j] = *reinterpret_cast<const float4*>(wrow + (size_t)row * a.N + n0 + c4);
    }
  };
  auto store_chunk = [&](int buf) {
    const float v[8] = {ra[0].x, ra[0].y, ra[0].z, ra[0].w, ra[1].x, ra[1].y, ra[1].z, ra[1].w};
#pragma unroll
    for (int k = 0; k < 8; ++k) As[buf][ah * 8 + k][ar] = v[k];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = tid + 256 * j, row = f >> 5, c4 = (f & 31) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][row][c4]) = rb[j];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int kc = 0; kc < n_chunks; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < n_chunks) load_chunk(kc + 1);
#pragma unroll
    for (int k = 0; k < CG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tn * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kc + 1 < n_chunks) {
      store_chunk(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const float nw = (a.mode == 1) ? a.noise_w[0] : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + tm * 8 + i;
    if (m >= M) continue;
    const int b = (int)(m / HW);
    const int p = (int)(m - (int64_t)b * HW);
    const float nz = (a.mode == 1) ? nw * a.noise[(size_t)b * a.noise_bstride + p] : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tn * 4;
      if (n >= a.N) continue;
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      if (a.mode == 1) {
        const float4 dv = *reinterpret_cast<const float4*>(a.d + (size_t)b * a.N + n);
        const float4 bv = *reinterpret_cast<const float4*>(a.act_bias + n);
        const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float t = fmaf(v[q], dd[q], nz) + bb[q];
          v[q] = (t > 0.f ? t : 0.2f * t) * kSqrt2;
        }
      } else if (a.mode == 2) {
        const float4 dv = *reinterpret_cast<const float4*>(a.d + (size_t)b * a.N + n);
        v[0] *= dv.x, v[1] *= dv.y, v[2] *= dv.z, v[3] *= dv.w;
      }
      *reinterpret_cast<float4*>(a.out + (size_t)m * a.N + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// ---- col2im (stride-2 transposed conv scatter, as a gather) + 4x4 blur + StyledConv epilogue ----
// T[P,Q,o]  = sum_{ky,kx : P-ky, Q-kx even, in range} G[(P-ky)/2, (Q-kx)/2, ky, kx, o]   (P,Q in [0,2H])
// out[Y,X,o] = lrelu( d * sum_{a,c<4} kb[a] kb[c] T[Y+a-1, X+c-1, o] + noise + bias ) * sqrt2,
// kb = [1,3,3,1]/4  (blur kernel outer([1,3,3,1])/64 * 4, pad (1,1); stylesdf_model.py:283-291,339-346)
struct Col2imArgs {
  const float* g;  // [B,H,W,9,cout]
  float* y;        // [B,2H,2W,cout]
  int B, H, W, cout;
  const float* d;
  const float* noise;
  int64_t noise_bstride;
  const float* noise_w;
  const float* act_bias;  // NULL: bare modulated conv (d * blur(convT)), no noise / bias / act
};

// One thread = a 2x2 output block (rows 2y,2y+1; cols 2x,2x+1) x 4 channels.  Along one axis the
// pair of outputs (2y, 2y+1) depends on input rows y-1, y, y+1 through 7 (row, tap) terms:
//   out[2y]   = .25 G[y-1,k1] + .75 G[y-1,k2] + .75 G[y,k0] + .75 G[y,k1] + .25 G[y,k2] + .25 G[y+1,k0]
//   out[2y+1] = .25 G[y-1,k2] + .25 G[y,k0] + .75 G[y,k1] + .75 G[y,k2] + .75 G[y+1,k0] + .25 G[y+1,k1]
// (transposed-conv scatter P = 2*iy + k, then blur taps [1,3,3,1]/4 at P = Y-1..Y+2), so the block
// reads 7x7 = 49 G vectors instead of 4 x 36.
__device__ __forceinline__ float c2i_coef(int phase, int d, int k) {
  // d = input offset + 1 (0..2), k = conv tap (0..2)
  const float c[2][3][3] = {{{0.f, .25f, .75f}, {.75f, .75f, .25f}, {.25f, 0.f, 0.f}},
                            {{0.f, 0.f, .25f}, {.25f, .75f, .75f}, {.75f, .25f, 0.f}}};
  return c[phase][d][k];
}

__global__ void __launch_bounds__(256) col2im_blur_act_kernel(const __grid_constant__ Col2imArgs a) {
  const int c4n = a.cout >> 2;
  const int OW = 2 * a.W, OH = 2 * a.H;
  const int64_t total = (int64_t)a.B * a.H * a.W * c4n;
  const bool linear = a.act_bias == nullptr;
  const float nw = linear ? 0.f : a.noise_w[0];
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx % c4n) * 4;
    int64_t t = idx / c4n;
    const int x = (int)(t % a.W);
    t /= a.W;
    const int y = (int)(t % a.H);
    const int b = (int)(t / a.H);
    float4 acc[2][2];
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) acc[py][px] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = y + dy - 1;
      if (iy < 0 || iy >= a.H) continue;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float cy0 = c2i_coef(0, dy, ky), cy1 = c2i_coef(1, dy, ky);
        if (cy0 == 0.f && cy1 == 0.f) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int ix = x + dx - 1;
          if (ix < 0 || ix >= a.W) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float cx0 = c2i_coef(0, dx, kx), cx1 = c2i_coef(1, dx, kx);
            if (cx0 == 0.f && cx1 == 0.f) continue;
            const float4 gv = *reinterpret_cast<const float4*>(
                a.g + ((((size_t)b * a.H + iy) * a.W + ix) * 9 + ky * 3 + kx) * a.cout + o);
            const float w00 = cy0 * cx0, w01 = cy0 * cx1, w10 = cy1 * cx0, w11 = cy1 * cx1;
            if (w00 != 0.f) acc[0][0].x = fmaf(w00, gv.x, acc[0][0].x), acc[0][0].y = fmaf(w00, gv.y, acc[0][0].y),
                            acc[0][0].z = fmaf(w00, gv.z, acc[0][0].z), acc[0][0].w = fmaf(w00, gv.w, acc[0][0].w);
            if (w01 != 0.f) acc[0][1].x = fmaf(w01, gv.x, acc[0][1].x), acc[0][1].y = fmaf(w01, gv.y, acc[0][1].y),
                            acc[0][1].z = fmaf(w01, gv.z, acc[0][1].z), acc[0][1].w = fmaf(w01, gv.w, acc[0][1].w);
            if (w10 != 0.f) acc[1][0].x = fmaf(w10, gv.x, acc[1][0].x), acc[1][0].y = fmaf(w10, gv.y, acc[1][0].y),
                            acc[1][0].z = fmaf(w10, gv.z, acc[1][0].z), acc[1][0].w = fmaf(w10, gv.w, acc[1][0].w);
            if (w11 != 0.f) acc[1][1].x = fmaf(w11, gv.x, acc[1][1].x), acc[1][1].y = fmaf(w11, gv.y, acc[1][1].y),
                            acc[1][1].z = fmaf(w11, gv.z, acc[1][1].z), acc[1][1].w = fmaf(w11, gv.w, acc[1][1].w);
          }
        }
      }
    }
    const float4 dv = *reinterpret_cast<const float4*>(a.d + (size_t)b * a.cout + o);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!linear) bv = *reinterpret_cast<const float4*>(a.act_bias + o);
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const int Y = 2 * y + py, X = 2 * x + px;
        const float4 s4 = acc[py][px];
        float v[4] = {s4.x * dv.x, s4.y * dv.y, s4.z * dv.z, s4.w * dv.w};
        if (!linear) {
          const float nz = nw * a.noise[(size_t)b * a.noise_bstride + (size_t)Y * OW + X];
          v[0] = fmaf(s4.x, dv.x, nz) + bv.x, v[1] = fmaf(s4.y, dv.y, nz) + bv.y;
          v[2] = fmaf(s4.z, dv.z, nz) + bv.z, v[3] = fmaf(s4.w, dv.w, nz) + bv.w;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) v[qq] = (v[qq] > 0.f ? v[qq] : 0.2f * v[qq]) * kSqrt2;
        }
        *reinterpret_cast<float4*>(a.y + (((size_t)b * OH + Y) * OW + X) * a.cout + o) =
            make_float4(v[0], v[1], v[2], v[3]);
      }
  }
}

// ---- 4x4 blur + StyledConv epilogue over the parity-phase transposed-conv output ---------------
// T[P,Q] (P in [0,2H], Q in [0,2W]) lives as four phase planes on the padded grid (tc_conv.cu):
//   T[2i+py, 2j+px, o] = t[(py*2+px)][(b*(H+1) + i)*(W+1) + j][o]
// out[Y,X,o] = lrelu( d * sum_{a,c<4} kb[a] kb[c] T[Y+a-1, X+c-1, o] + noise + bias ) * sqrt2, kb = [1,3,3,1]/4.
// One thread = 2x2 input pixels = a 4x4 output block x 4 channels: T rows 2y-1 .. 2y+5 = (odd,y-1) (even,y)
// (odd,y) (even,y+1) (odd,y+1) (even,y+2) (odd,y+2), same along x: 49 vector reads for 64 outputs.
struct UpBlurArgs {
  const float* t;  // [4][Mp][cout]
  float* y;        // [B,2H,2W,cout]
  int B, H, W, cout;
  const float* d;
  const float* noise;
  int64_t noise_bstride;
  const float* noise_w;
  const float* act_bias;  // NULL: bare modulated conv
  // inference fusion: when xs_hi is set, y is not written; the epilogue stores the NEXT plain conv's
  // tensor-core operands xs = out * next_s[b,:] as bf16 hi / lo ([B,2H,2W,cout] each) instead
  const float* next_s;
  __nv_bfloat16* xs_hi;
  __nv_bfloat16* xs_lo;
};

// weight of T row r (0..6, relative to 2y-1) in output row 2y+k (k = 0..3): kb[r - k], kb = [1,3,3,1]/4
__device__ __forceinline__ float blur_w(int r, int k) {
  const int i = r - k;
  return (i == 0 || i == 3) ? .25f : ((i == 1 || i == 2) ? .75f : 0.f);
}

__global__ void __launch_bounds__(256) upconv_blur_act_kernel(const __grid_constant__ UpBlurArgs a) {
  const int c4n = a.cout >> 2;
  const int OW = 2 * a.W, OH = 2 * a.H, Wp = a.W + 1, Hp = a.H + 1;
  const int H2 = (a.H + 1) >> 1, W2 = (a.W + 1) >> 1;
  const int64_t Mp = (int64_t)a.B * Hp * Wp;
  const int64_t total = (int64_t)a.B * H2 * W2 * c4n;
  const bool linear = a.act_bias == nullptr;
  const float nw = linear ? 0.f : a.noise_w[0];
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx % c4n) * 4;
    int64_t t = idx / c4n;
    const int x = (int)(t % W2) * 2;
    t /= W2;
    const int y = (int)(t % H2) * 2;
    const int b = (int)(t / H2);
    float4 acc[4][4];
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) acc[ky][kx] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const int pr = (r & 1) ^ 1;            // rows alternate odd, even, odd, ...
      const int i = y + ((r + 1) >> 1) - 1;  // y-1, y, y, y+1, y+1, y+2, y+2
      if (i < 0 || i >= a.H + (pr ? 0 : 1)) continue;
      // the row's seven vectors are requested together (predicated, no branches between the loads): the
      // kernel is bound by the number of loads in flight, not by bytes
      float4 tv[7];
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        const int pc = (c & 1) ^ 1;
        const int j = x + ((c + 1) >> 1) - 1;
        const bool ok = j >= 0 && j < a.W + (pc ? 0 : 1);
        const float4* src = reinterpret_cast<const float4*>(
            a.t + ((size_t)(pr * 2 + pc) * Mp + ((size_t)b * Hp + i) * Wp + (ok ? j : 0)) * a.cout + o);
        tv[c] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // separable: the row's four horizontal blur outputs first (4 taps each), then into the <= 4 output
      // rows this T row feeds: 64 + 16..64 FMAs per row instead of up to 256
      float4 hr[4];
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        hr[kx] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          const float k = blur_w(c, kx);
          if (k != 0.f) {
            hr[kx].x = fmaf(k, tv[c].x, hr[kx].x), hr[kx].y = fmaf(k, tv[c].y, hr[kx].y);
            hr[kx].z = fmaf(k, tv[c].z, hr[kx].z), hr[kx].w = fmaf(k, tv[c].w, hr[kx].w);
          }
        }
      }
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const float k = blur_w(r, ky);
        if (k != 0.f) {
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            acc[ky][kx].x = fmaf(k, hr[kx].x, acc[ky][kx].x), acc[ky][kx].y = fmaf(k, hr[kx].y, acc[ky][kx].y);
            acc[ky][kx].z = fmaf(k, hr[kx].z, acc[ky][kx].z), acc[ky][kx].w = fmaf(k, hr[kx].w, acc[ky][kx].w);
          }
        }
      }
    }
    const float4 dv = *reinterpret_cast<const float4*>(a.d + (size_t)b * a.cout + o);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!linear) bv = *reinterpret_cast<const float4*>(a.act_bias + o);
    float4 ns = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.xs_hi) ns = *reinterpret_cast<const float4*>(a.next_s + (size_t)b * a.cout + o);
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int Y = 2 * y + ky, X = 2 * x + kx;
        if (Y >= OH || X >= OW) continue;  // odd H / W: the last block is half empty
        const float4 s4 = acc[ky][kx];
        float v[4] = {s4.x * dv.x, s4.y * dv.y, s4.z * dv.z, s4.w * dv.w};
        if (!linear) {
          const float nz = nw * a.noise[(size_t)b * a.noise_bstride + (size_t)Y * OW + X];
          v[0] = fmaf(s4.x, dv.x, nz) + bv.x, v[1] = fmaf(s4.y, dv.y, nz) + bv.y;
          v[2] = fmaf(s4.z, dv.z, nz) + bv.z, v[3] = fmaf(s4.w, dv.w, nz) + bv.w;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) v[qq] = (v[qq] > 0.f ? v[qq] : 0.2f * v[qq]) * kSqrt2;
        }
        const size_t at = (((size_t)b * OH + Y) * OW + X) * a.cout + o;
        if (a.xs_hi) {  // same arithmetic as modulate_split_kernel on the stored fp32 value
          const float m[4] = {v[0] * ns.x, v[1] * ns.y, v[2] * ns.z, v[3] * ns.w};
          __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) tc::split_bf16(m[qq], h[qq], l[qq]);
          *reinterpret_cast<uint2*>(a.xs_hi + at) = *reinterpret_cast<const uint2*>(h);
          *reinterpret_cast<uint2*>(a.xs_lo + at) = *reinterpret_cast<const uint2*>(l);
        } else {
          *reinterpret_cast<float4*>(a.y + at) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
  }
}

// ---- ToRGB: 1x1 modulated conv (no demod) + bias + FIR-upsampled skip ----------------------
struct ToRgbArgs {
  const float* x;     // [B,H,W,cin]
  const float* w;     // [3,cin]
  const float* s;     // [B,cin]
  const float* bias;  // [3]
  const float* skip;  // NULL | [B,3,H/2,W/2] | [B,3,H,W]
  int upsample_skip;
  float* rgb;  // [B,3,H,W]
  int B, H, W, cin;
};

// Eight lanes per pixel: the pixel's cin fp32 channels are contiguous (NHWC) and are read as float4, one
// full 128-byte line per 8-lane group and step (every load coalesced, many in flight per lane); the three
// modulated weight rows sit in shared memory; the three dot products close with a 3-step shuffle
// reduction inside the group.  (One thread per pixel, the earlier layout, ran at 2.9 TB/s: each load
// instruction touched 32 different lines.)
__global__ void __launch_bounds__(256) torgb_kernel(const __grid_constant__ ToRgbArgs a) {
  extern __shared__ float ws[];  // [3][cin] = scale * W[c,i] * s[b,i]
  const int b = blockIdx.y, HW = a.H * a.W;
  const float scale = rsqrtf((float)a.cin);  // 1/sqrt(cin*1*1)
  for (int i = threadIdx.x; i < 3 * a.cin; i += blockDim.x)
    ws[i] = scale * a.w[i] * a.s[(size_t)b * a.cin + (i % a.cin)];
  __syncthreads();
  const float kb[4] = {0.25f, 0.75f, 0.75f, 0.25f};  // flipped == itself (symmetric)
  const int gl = threadIdx.x & 7;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, n_groups = (gridDim.x * blockDim.x) >> 3;
  const int n_iter = (HW + n_groups - 1) / n_groups;  // the same for every lane: shuffles stay converged
  for (int it = 0; it < n_iter; ++it) {
    const int p = group + it * n_groups;
    const bool live = p < HW;
    const float* xp = a.x + ((size_t)b * HW + (live ? p : 0)) * a.cin;
    float acc[3] = {0.f, 0.f, 0.f};
    if (live) {
#pragma unroll 4
      for (int c = gl * 4; c < a.cin; c += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xp + c));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 w = *reinterpret_cast<const float4*>(ws + k * a.cin + c);
          acc[k] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[k]))));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 4);
      acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
      acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
    }
    if (!live || gl >= 3) continue;
    const int c = gl;  // lanes 0..2 of the group finish one colour channel each
    const int Y = p / a.W, X = p - Y * a.W;
    float v = (c == 0 ? acc[0] : (c == 1 ? acc[1] : acc[2])) + a.bias[c];
    if (a.skip) {
      if (a.upsample_skip) {
        // upfirdn2d(skip, outer(kb,kb), up=2, pad=(2,1)): U[2i]=skip[i]; P[y]=U[y-2]
        const int h2 = a.H >> 1, w2 = a.W >> 1;
        const float* sp = a.skip + ((size_t)b * 3 + c) * h2 * w2;
        float up = 0.f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const int uy = Y + ky - 2;
          if (uy < 0 || (uy & 1) || (uy >> 1) >= h2) continue;
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            const int ux = X + kx - 2;
            if (ux < 0 || (ux & 1) || (ux >> 1) >= w2) continue;
            up = fmaf(sp[(size_t)(uy >> 1) * w2 + (ux >> 1)], kb[3 - ky] * kb[3 - kx], up);
          }
        }
        v += up;
      } else {
        v += a.skip[((size_t)b * 3 + c) * HW + p];
      }
    }
    a.rgb[((size_t)b * 3 + c) * HW + p] = v;
  }
}

// ---- image-parallel inversion record ----------------------------------------
// One thread-block cluster of RECORD_PARTS CTAs per image (blockIdx.y = image, blockIdx.x = slice): every
// slice copies its share of the latents and reduces its share of the squared / absolute error; the slices'
// partial sums meet in CTA 0 of the cluster through distributed shared memory and are added in slice
// order — no atomics, the record is bit-reproducible.
constexpr int RECORD_PARTS = 8;
constexpr int RECORD_THREADS = 512;
__global__ void __launch_bounds__(RECORD_THREADS) pack_record_kernel(const float* __restrict__ w_plus,
                                                          const float* __restrict__ w_dec,
                                                          int n_latent, const float* __restrict__ image,
                                                          const float* __restrict__ target,
                                                          int64_t image_numel,
                                                          float* __restrict__ record) {
  const int b = blockIdx.y, part = blockIdx.x, parts = gridDim.x;
  const int rec_len = 2304 + n_latent * 512 + 2;
  float* rec = record + (size_t)b * rec_len;
  const int tid = part * blockDim.x + threadIdx.x, nthr = parts * blockDim.x;
  for (int i = tid; i < 2304; i += nthr) rec[i] = w_plus[(size_t)b * 2304 + i];
  for (int i = tid; i < n_latent * 512; i += nthr) rec[2304 + i] = w_dec[(size_t)b * n_latent * 512 + i];
  float se = 0.f, ae = 0.f;
  if (image && target) {
    const float* im = image + (size_t)b * image_numel;
    const float* tg = target + (size_t)b * image_numel;
    for (int64_t i = tid; i < image_numel; i += nthr) {
      const float dlt = im[i] - tg[i];
      se = fmaf(dlt, dlt, se);
      ae += fabsf(dlt);
    }
  }
  __shared__ float red[2][RECORD_THREADS / 32];
  __shared__ float part_sum[2];
  se = warp_sum(se), ae = warp_sum(ae);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[0][warp] = se, red[1][warp] = ae;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s2 = 0.f, a2 = 0.f;
    for (int i = 0; i < RECORD_THREADS / 32; ++i) s2 += red[0][i], a2 += red[1][i];
    part_sum[0] = s2, part_sum[1] = a2;
  }
  if (parts == 1) {  // launched without a cluster (no image): the metric slots are zero
    if (threadIdx.x == 0) rec[rec_len - 2] = 0.f, rec[rec_len - 1] = 0.f;
    return;
  }
  cluster_sync_all();  // every slice's partial sums are in its shared memory
  if (part == 0 && threadIdx.x == 0) {
    float s2 = 0.f, a2 = 0.f;
    for (int r = 0; r < parts; ++r) {
      const uint32_t addr = tc::map_to_cta(part_sum, (uint32_t)r);
      float ps, pa;
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(ps) : "r"(addr));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pa) : "r"(addr + 4));
      s2 += ps, a2 += pa;
    }
    rec[rec_len - 2] = s2 / (float)image_numel;
    rec[rec_len - 1] = a2 / (float)image_numel;
  }
  cluster_sync_all();  // no slice exits while CTA 0 may still read its shared memory
}

int conv_gemm_ffma_launch(const ConvGemmArgs& a, int taps, cudaStream_t stream) {
  E3_REQUIRE(a.Cin % CG_BK == 0 && a.N % 4 == 0 && (!a.planar || taps == 9), E3_ERR_UNSUPPORTED,
             "CUDA-core conv: needs Cin %% 16 == 0 and N %% 4 == 0 (got Cin=%d N=%d)", a.Cin, a.N);
  const int64_t M = (int64_t)a.B * a.H * a.W;
  dim3 grid((unsigned)((M + CG_BM - 1) / CG_BM), (a.N + CG_BN - 1) / CG_BN);
  if (taps == 9) conv_gemm_ffma_kernel<9><<<grid, 256, 0, stream>>>(a);
  else conv_gemm_ffma_kernel<1><<<grid, 256, 0, stream>>>(a);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

static int grid_cap(int64_t blocks) {
  const int64_t cap = (int64_t)sm_count() * 32;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace e3

using namespace e3;

extern "C" int e3_modconv_weight_sq(const float* weight, int cout, int cin, int ksize, float* wsq,
                                    void* stream) {
  E3_REQUIRE(weight && wsq && cout > 0 && cin > 0 && ksize > 0, E3_ERR_BAD_ARG,
             "e3_modconv_weight_sq: bad argument");
  const int n = cout * cin;
  weight_sq_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(weight, n, ksize * ksize, wsq);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" int e3_modconv_styles(const float* latent, int64_t latent_stride, const float* mod_w,
                                 const float* mod_b, const float* wsq, int batch, int cin, int cout,
                                 int ksize, float* s, float* d, void* stream) {
  E3_REQUIRE(latent && mod_w && mod_b && s, E3_ERR_BAD_ARG, "e3_modconv_styles: null argument");
  E3_REQUIRE(batch >= 0 && cin > 0 && batch <= 65535, E3_ERR_BAD_ARG, "e3_modconv_styles: bad shape");
  E3_REQUIRE(!d || (wsq && cout > 0 && ksize > 0), E3_ERR_BAD_ARG,
             "e3_modconv_styles: demodulation needs wsq, cout, ksize");
  if (batch == 0) return E3_OK;
  mod_style_kernel<<<dim3((cin + 7) / 8, batch), 256, 0, as_stream(stream)>>>(latent, latent_stride,
                                                                           mod_w, mod_b, cin, s);
  if (d) {
    const float scale2 = 1.f / (float)(cin * ksize * ksize);
    demod_kernel<<<dim3((cout + 7) / 8, batch), 256, 0, as_stream(stream)>>>(wsq, s, cin, cout,
                                                                          scale2, d);
  }
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

// packed image: [fp32 GEMM-major (cout*cin*9 floats)] [bf16 hi K-major] [bf16 lo K-major]
extern "C" size_t e3_conv_packed_bytes(int cout, int cin) {
  return (size_t)cout * cin * 9 * sizeof(float) + tc_conv_packed_bf16_bytes(cout, cin);
}
static inline const void* packed_bf16_part(const void* packed, int cout, int cin) {
  return static_cast<const char*>(packed) + (size_t)cout * cin * 9 * sizeof(float);
}
namespace e3 {
const void* conv_packed_bf16_part(const void* packed, int cout, int cin) { return packed_bf16_part(packed, cout, cin); }
}

extern "C" int e3_conv_pack_weight(const float* weight, int cout, int cin, int upsample,
                                   void* packed, void* stream) {
  E3_REQUIRE(weight && packed && cout > 0 && cin > 0, E3_ERR_BAD_ARG, "e3_conv_pack_weight: bad argument");
  E3_REQUIRE(upsample >= 0 && upsample <= 3, E3_ERR_BAD_ARG, "e3_conv_pack_weight: layout %d outside 0..3", upsample);
  const float scale = 1.f / sqrtf((float)(cin * 9));
  conv_pack_kernel<<<grid_cap(((int64_t)cout * cin * 9 + 255) / 256), 256, 0, as_stream(stream)>>>(
      weight, cout, cin, upsample, scale, static_cast<float*>(packed));
  E3_CUDA(cudaGetLastError());
  return tc_conv_pack_weight(weight, cout, cin, upsample, scale,
                             const_cast<void*>(packed_bf16_part(packed, cout, cin)), as_stream(stream));
}

// rows of the x-pair view: s2[b] = [s[b] | s[b]], d2[b] = [d[b] | d[b]], bias2 = [bias | bias] (32 -> 64 columns)
__global__ void dup_pair_kernel(const float* __restrict__ s, const float* __restrict__ d,
                                const float* __restrict__ bias, int batch, float* s2, float* d2, float* b2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch * 64) {
    const int b = i >> 6, c = i & 31;
    s2[i] = s[b * 32 + c];
    d2[i] = d[b * 32 + c];
  }
  if (i < 64 && bias) b2[i] = bias[i & 31];
}

static bool use_tensor_cores(uint32_t flags, int batch, int h, int w, int cin, int n) {
  if (flags & E3_CONV_FP32_CUDA_CORES) return false;
  return tc_conv_supported(batch, h, w, cin, n);
}

// scratch, plain conv:  [xs_hi | xs_lo: B*H*W*cin bf16 each]
// upsampling conv:      [G: B*H*W*9*cout fp32 (CUDA-core path)  |  T: 4*B*(H+1)*(W+1)*cout fp32 (tensor-core
//                        path, always smaller)] [xs_hi | xs_lo on the zero-padded grid B*(H+1)*(W+1)*cin]
extern "C" size_t e3_styled_conv_scratch_bytes(int batch, int h, int w, int cin, int cout,
                                               int upsample) {
  // (+ the duplicated s / d / bias rows of the x-pair view of a 32 -> 32 conv)
  if (!upsample) return tc_conv_split_bytes(batch, h, w, cin) + 256 + ((size_t)2 * batch + 1) * 64 * sizeof(float) + 256;
  size_t g = (size_t)batch * h * w * 9 * cout * sizeof(float);
  const size_t t = tc_upconv_t_bytes(batch, h, w, cout);
  if (t > g) g = t;
  return g + tc_upconv_split_bytes(batch, h, w, cin) + 256;
}

static int check_conv_shapes(const char* who, int batch, int h, int w, int cin, int cout) {
  E3_REQUIRE(batch >= 0 && h > 0 && w > 0, E3_ERR_BAD_ARG, "%s: bad shape", who);
  E3_REQUIRE(cin % 16 == 0 && cout % 4 == 0, E3_ERR_UNSUPPORTED,
             "%s: needs cin %% 16 == 0 and cout %% 4 == 0 (got cin=%d cout=%d)", who, cin, cout);
  return E3_OK;
}

extern "C" int e3_styled_conv3x3_fwd(const float* x, const void* wpacked, const float* s,
                                     const float* d, const float* noise, int64_t noise_batch_stride,
                                     const float* noise_w, const float* act_bias, float* y, int batch,
                                     int h, int w, int cin, int cout, void* scratch,
                                     size_t scratch_bytes, uint32_t flags, void* stream) {
  int rc = check_conv_shapes("e3_styled_conv3x3_fwd", batch, h, w, cin, cout);
  if (rc) return rc;
  if (batch == 0) return E3_OK;
  E3_REQUIRE(x && wpacked && s && d && y, E3_ERR_BAD_ARG, "e3_styled_conv3x3_fwd: null argument");
  E3_REQUIRE(!act_bias || (noise && noise_w), E3_ERR_BAD_ARG,
             "e3_styled_conv3x3_fwd: noise and noise_w are required unless act_bias is NULL");
  ConvGemmArgs a{};
  a.x = x, a.s = s, a.wg = static_cast<const float*>(wpacked), a.out = y;
  a.B = batch, a.H = h, a.W = w, a.Cin = cin, a.N = cout;
  a.mode = act_bias ? 1 : 2, a.d = d, a.noise = noise, a.noise_bstride = noise_batch_stride;
  a.noise_w = noise_w, a.act_bias = act_bias;
  const bool tcore = use_tensor_cores(flags, batch, h, w, cin, cout);
  const bool pairx = !tcore && !(flags & E3_CONV_FP32_CUDA_CORES) && tc_conv_pairx_supported(batch, h, w, cin, cout);
  E3_REQUIRE(tcore || pairx || !(flags & E3_CONV_TENSOR_CORES), E3_ERR_UNSUPPORTED,
             "e3_styled_conv3x3_fwd: E3_CONV_TENSOR_CORES requested for an unsupported shape");
  if (pairx) {
    // 32 -> 32 at 1024^2: two x-adjacent pixels form one 64-channel operand row (same memory), the 3x3 kernel
    // becomes a block-structured 3 x 3-pair-tap kernel (half of its 64 x 64 blocks' entries are zero)
    E3_REQUIRE(scratch && scratch_bytes >= e3_styled_conv_scratch_bytes(batch, h, w, cin, cout, 0),
               E3_ERR_SCRATCH, "e3_styled_conv3x3_fwd: scratch too small");
    char* split = reinterpret_cast<char*>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    float* dup = reinterpret_cast<float*>(
        ((uintptr_t)(split + tc_conv_split_bytes(batch, h, w, cin)) + 255) & ~(uintptr_t)255);
    float *s2 = dup, *d2 = dup + (size_t)batch * 64, *b2 = dup + (size_t)2 * batch * 64;
    dup_pair_kernel<<<(batch * 64 + 255) / 256 + 1, 256, 0, as_stream(stream)>>>(s, d, act_bias, batch, s2, d2, b2);
    E3_CUDA(cudaGetLastError());
    a.s = s2, a.d = d2, a.act_bias = act_bias ? b2 : nullptr;
    a.W = w / 2, a.Cin = 64, a.N = 64, a.pairx = 1;
    return tc_conv_launch(a, 9, packed_bf16_part(wpacked, cout, cin), split, as_stream(stream));
  }
  if (tcore) {
    E3_REQUIRE(scratch && scratch_bytes >= e3_styled_conv_scratch_bytes(batch, h, w, cin, cout, 0),
               E3_ERR_SCRATCH, "e3_styled_conv3x3_fwd: scratch too small");
    void* split = reinterpret_cast<void*>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    return tc_conv_launch(a, 9, packed_bf16_part(wpacked, cout, cin), split, as_stream(stream));
  }
  const int64_t M = (int64_t)batch * h * w;
  dim3 grid((unsigned)((M + CG_BM - 1) / CG_BM), (cout + CG_BN - 1) / CG_BN);
  conv_gemm_ffma_kernel<9><<<grid, 256, 0, as_stream(stream)>>>(a);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

static int up_fwd_impl(const float* x, const void* wpacked, const float* s, const float* d, const float* noise,
                       int64_t noise_batch_stride, const float* noise_w, const float* act_bias, float* y,
                       const float* next_s, void* xs_hi, void* xs_lo, int batch, int h, int w, int cin,
                       int cout, void* scratch, size_t scratch_bytes, uint32_t flags, void* stream) {
  int rc = check_conv_shapes("e3_styled_conv3x3_up_fwd", batch, h, w, cin, cout);
  if (rc) return rc;
  if (batch == 0) return E3_OK;
  E3_REQUIRE(x && wpacked && s && d && (y || xs_hi), E3_ERR_BAD_ARG, "e3_styled_conv3x3_up_fwd: null argument");
  E3_REQUIRE(!act_bias || (noise && noise_w), E3_ERR_BAD_ARG,
             "e3_styled_conv3x3_up_fwd: noise and noise_w are required unless act_bias is NULL");
  const size_t need = e3_styled_conv_scratch_bytes(batch, h, w, cin, cout, 1);
  E3_REQUIRE(scratch && scratch_bytes >= need, E3_ERR_SCRATCH,
             "e3_styled_conv3x3_up_fwd: scratch %zu < %zu bytes", scratch_bytes, need);
  ConvGemmArgs a{};
  a.x = x, a.s = s, a.wg = static_cast<const float*>(wpacked), a.out = static_cast<float*>(scratch);
  a.B = batch, a.H = h, a.W = w, a.Cin = cin, a.N = 9 * cout;
  a.mode = 0;
  const bool tcore = !(flags & E3_CONV_FP32_CUDA_CORES) && tc_upconv_supported(batch, h, w, cin, cout);
  E3_REQUIRE(tcore || !(flags & E3_CONV_TENSOR_CORES), E3_ERR_UNSUPPORTED,
             "e3_styled_conv3x3_up_fwd: E3_CONV_TENSOR_CORES requested for an unsupported shape");
  E3_REQUIRE(tcore || !xs_hi, E3_ERR_UNSUPPORTED,
             "e3_styled_conv3x3_up_fwd_split: the fused operand split exists on the tensor-core path only");
  if (tcore) {
    // four parity-phase convolutions accumulate the transposed conv in TMEM -> T, then blur + epilogue
    size_t t_bytes = tc_upconv_t_bytes(batch, h, w, cout);
    const size_t g_bytes = (size_t)batch * h * w * 9 * cout * sizeof(float);
    if (g_bytes > t_bytes) t_bytes = g_bytes;
    char* after_t = static_cast<char*>(scratch) + t_bytes;
    void* split = reinterpret_cast<void*>(((uintptr_t)after_t + 255) & ~(uintptr_t)255);
    float* t_out = static_cast<float*>(scratch);
    rc = tc_upconv_phase_launch(x, s, batch, h, w, cin, cout, packed_bf16_part(wpacked, cout, cin), split,
                                t_out, as_stream(stream));
    if (rc) return rc;
    UpBlurArgs u{};
    u.t = t_out, u.y = y, u.B = batch, u.H = h, u.W = w, u.cout = cout;
    u.d = d, u.noise = noise, u.noise_bstride = noise_batch_stride, u.noise_w = noise_w, u.act_bias = act_bias;
    u.next_s = next_s, u.xs_hi = static_cast<__nv_bfloat16*>(xs_hi), u.xs_lo = static_cast<__nv_bfloat16*>(xs_lo);
    const int64_t total = (int64_t)batch * ((h + 1) / 2) * ((w + 1) / 2) * (cout / 4);
    upconv_blur_act_kernel<<<grid_cap((total + 255) / 256), 256, 0, as_stream(stream)>>>(u);
    E3_CUDA(cudaGetLastError());
    return E3_OK;
  }
  {
    const int64_t M = (int64_t)batch * h * w;
    dim3 grid((unsigned)((M + CG_BM - 1) / CG_BM), (a.N + CG_BN - 1) / CG_BN);
    conv_gemm_ffma_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(a);
    E3_CUDA(cudaGetLastError());
  }
  Col2imArgs c{};
  c.g = static_cast<const float*>(scratch), c.y = y, c.B = batch, c.H = h, c.W = w, c.cout = cout;
  c.d = d, c.noise = noise, c.noise_bstride = noise_batch_stride, c.noise_w = noise_w;
  c.act_bias = act_bias;
  const int64_t total = (int64_t)batch * h * w * (cout / 4);  // one thread per 2x2 output block x 4 ch
  col2im_blur_act_kernel<<<grid_cap((total + 255) / 256), 256, 0, as_stream(stream)>>>(c);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" int e3_styled_conv3x3_up_fwd(const float* x, const void* wpacked, const float* s,
                                        const float* d, const float* noise,
                                        int64_t noise_batch_stride, const float* noise_w,
                                        const float* act_bias, float* y, int batch, int h, int w,
                                        int cin, int cout, void* scratch, size_t scratch_bytes,
                                        uint32_t flags, void* stream) {
  E3_REQUIRE(y || batch == 0, E3_ERR_BAD_ARG, "e3_styled_conv3x3_up_fwd: null argument");
  return up_fwd_impl(x, wpacked, s, d, noise, noise_batch_stride, noise_w, act_bias, y, nullptr, nullptr, nullptr,
                     batch, h, w, cin, cout, scratch, scratch_bytes, flags, stream);
}

extern "C" int e3_styled_conv_pair_fusable(int batch, int h, int w, int cin, int cout, uint32_t flags) {
  if (flags & E3_CONV_FP32_CUDA_CORES) return 0;
  return tc_upconv_supported(batch, h, w, cin, cout) && tc_conv_supported(batch, 2 * h, 2 * w, cout, cout) ? 1 : 0;
}

extern "C" int e3_styled_conv3x3_up_fwd_split(const float* x, const void* wpacked, const float* s,
                                              const float* d, const float* noise,
                                              int64_t noise_batch_stride, const float* noise_w,
                                              const float* act_bias, const float* next_s, void* xs_hi,
                                              void* xs_lo, int batch, int h, int w, int cin, int cout,
                                              void* scratch, size_t scratch_bytes, uint32_t flags,
                                              void* stream) {
  E3_REQUIRE((next_s && xs_hi && xs_lo) || batch == 0, E3_ERR_BAD_ARG,
             "e3_styled_conv3x3_up_fwd_split: null argument");
  return up_fwd_impl(x, wpacked, s, d, noise, noise_batch_stride, noise_w, act_bias, nullptr, next_s, xs_hi, xs_lo,
                     batch, h, w, cin, cout, scratch, scratch_bytes, flags, stream);
}

extern "C" int e3_styled_conv3x3_fwd_presplit(const void* xs_hi, const void* xs_lo, const void* wpacked,
                                              const float* d, const float* noise,
                                              int64_t noise_batch_stride, const float* noise_w,
                                              const float* act_bias, float* y, int batch, int h, int w,
                                              int cin, int cout, uint32_t flags, void* stream) {
  int rc = check_conv_shapes("e3_styled_conv3x3_fwd_presplit", batch, h, w, cin, cout);
  if (rc) return rc;
  if (batch == 0) return E3_OK;
  E3_REQUIRE(xs_hi && xs_lo && wpacked && d && y, E3_ERR_BAD_ARG, "e3_styled_conv3x3_fwd_presplit: null argument");
  E3_REQUIRE(!act_bias || (noise && noise_w), E3_ERR_BAD_ARG,
             "e3_styled_conv3x3_fwd_presplit: noise and noise_w are required unless act_bias is NULL");
  E3_REQUIRE(use_tensor_cores(flags, batch, h, w, cin, cout), E3_ERR_UNSUPPORTED,
             "e3_styled_conv3x3_fwd_presplit: pre-split operands exist on the tensor-core path only");
  ConvGemmArgs a{};
  a.out = y;
  a.B = batch, a.H = h, a.W = w, a.Cin = cin, a.N = cout;
  a.mode = act_bias ? 1 : 2, a.d = d, a.noise = noise, a.noise_bstride = noise_batch_stride;
  a.noise_w = noise_w, a.act_bias = act_bias;
  return tc_conv_launch_presplit(a, 9, packed_bf16_part(wpacked, cout, cin), xs_hi, xs_lo, as_stream(stream));
}

extern "C" int e3_torgb_fwd(const float* x, const float* weight, const float* s, const float* bias,
                            const float* skip, int upsample_skip, float* rgb, int batch, int h, int w,
                            int cin, void* stream) {
  E3_REQUIRE(batch >= 0 && h > 0 && w > 0 && cin > 0 && cin % 4 == 0 && batch <= 65535,
             E3_ERR_BAD_ARG, "e3_torgb_fwd: bad shape (cin %% 4 == 0 required)");
  E3_REQUIRE(!(skip && upsample_skip) || (h % 2 == 0 && w % 2 == 0), E3_ERR_BAD_ARG,
             "e3_torgb_fwd: upsampled skip needs even output size");
  if (batch == 0) return E3_OK;
  E3_REQUIRE(x && weight && s && bias && rgb, E3_ERR_BAD_ARG, "e3_torgb_fwd: null argument");
  ToRgbArgs a{x, weight, s, bias, skip, upsample_skip, rgb, batch, h, w, cin};
  int bx = (h * w + 31) / 32;  // 32 pixels per block of 256 threads and pass
  const int cap = (sm_count() * 8 + batch - 1) / batch;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  torgb_kernel<<<dim3(bx, batch), 256, 3 * cin * sizeof(float), as_stream(stream)>>>(a);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" int e3_pack_inversion_record(const float* w_plus, const float* w_dec, int n_latent,
                                        const float* image, const float* target, int batch,
                                        int64_t image_numel, float* record, void* stream) {
  E3_REQUIRE(w_plus && w_dec && record && n_latent > 0 && batch >= 0, E3_ERR_BAD_ARG,
             "e3_pack_inversion_record: bad argument");
  E3_REQUIRE((image == nullptr) == (target == nullptr), E3_ERR_BAD_ARG,
             "e3_pack_inversion_record: image and target come together");
  if (batch == 0) return E3_OK;
  const int parts = image ? RECORD_PARTS : 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(parts, batch);
  cfg.blockDim = dim3(RECORD_THREADS);
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = parts;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  E3_CUDA(cudaLaunchKernelEx(&cfg, pack_record_kernel, w_plus, w_dec, n_latent, image, target,
                             image_numel > 0 ? image_numel : (int64_t)1, record));
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}
