// Pixel-aligned feature query of the E3DGE local branch, sm_100a.
//
// Replaces HGPIFuNetGAN.query(points, calibs, ..., im_feat=F) as the E3DGE runner calls it
// (project/trainers/E3DGE/e3dge_full_runner.py:219-226, 271-278; vendor/pifu/lib/model/HGPIFuGANNet.py:
// 85-150): perspective projection of every sample point into a reference view
// (vendor/pifu/lib/geometry.py:108-135), y flip + in-image test, bilinear zero-padded grid_sample
// (align_corners=False; geometry.py:64-80, project/models/op/grid_sample_gradfix.py:29-36) of a C-channel map.
//
// The reference gathers from an NCHW map into [B,C,N] and the caller permutes to [B,N,C] right away
// (e3dge_full_runner.py:229-230): one strided 4-byte read per (point, channel, tap) and a transposing
// copy of a 100 MB tensor per image.  Here the map is channels-last and the output is written as
// [B,N,C] directly: 8 lanes per point, each tap a run of contiguous 128-byte lines, every output line
// written once in full.  HBM-bound on the output (C*4 bytes per point); the map (H*W*C*4 bytes per
// image) stays in L2.
#include "common.cuh"

namespace e3 {

struct LocalQueryArgs {
  const float* feat;    // [B,H,W,C] channels-last
  const float* points;  // element (b, k, n) at points[b*pb + k*pk + n*pn]
  int64_t pb, pk, pn;
  const float* calibs;  // [B][calib_stride] row-major 3x4 (or 4x4) projection, first 12 floats used
  int calib_stride;
  int B, N, H, W, C;
  float* feats;            // [B,N,C] or NULL (projection only)
  float* proj_xy;          // [B,2,N] or NULL
  float* depth;            // [B,1,N] or NULL
  unsigned char* in_img;   // [B,N] or NULL
};

__global__ void __launch_bounds__(256) local_query_kernel(const __grid_constant__ LocalQueryArgs a) {
  const int gl = threadIdx.x & 7;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) >> 3;
  const int64_t total = (int64_t)a.B * a.N;
  // "look at -z" switch of geometry.perspective: decided by point 0 of image 0 for the whole batch
  float zsign;
  {
    const float* c = a.calibs;
    const float hz = __fadd_rn(fmaf(c[10], a.points[2 * a.pk], fmaf(c[9], a.points[a.pk], __fmul_rn(c[8], a.points[0]))), c[11]);
    zsign = hz < 0.f ? -1.f : 1.f;
  }
  for (int64_t g = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3); g < total; g += n_groups) {
    const int b = (int)(g / a.N), n = (int)(g - (int64_t)b * a.N);
    const float* c = a.calibs + (size_t)b * a.calib_stride;
    const float* p = a.points + b * a.pb + n * a.pn;
    const float px = p[0], py = p[a.pk], pz = p[2 * a.pk];
    // trans + rot @ p  (baddbmm: the product sum first, then the translation)
    const float hx = __fadd_rn(fmaf(c[2], pz, fmaf(c[1], py, __fmul_rn(c[0], px))), c[3]);
    const float hy = __fadd_rn(fmaf(c[6], pz, fmaf(c[5], py, __fmul_rn(c[4], px))), c[7]);
    const float hz = __fadd_rn(fmaf(c[10], pz, fmaf(c[9], py, __fmul_rn(c[8], px))), c[11]);
    const float z = hz * zsign;
    const float x = __fdiv_rn(hx, z), y = -__fdiv_rn(hy, z);
    if (gl == 0) {
      if (a.proj_xy) {
        a.proj_xy[((size_t)b * 2) * a.N + n] = x;
        a.proj_xy[((size_t)b * 2 + 1) * a.N + n] = y;
      }
      if (a.depth) a.depth[(size_t)b * a.N + n] = z;
      if (a.in_img) a.in_img[(size_t)b * a.N + n] = (x >= -1.f && x <= 1.f && y >= -1.f && y <= 1.f) ? 1 : 0;
    }
    if (!a.feats) continue;
    // grid_sample, bilinear, zeros, align_corners=False
    const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(x, 1.f), (float)a.W), 1.f), 0.5f);
    const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(y, 1.f), (float)a.H), 1.f), 0.5f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float tx = ix - fx0, ty = iy - fy0;
    // NaN / far-away coordinates: no tap is in range (the float comparisons below are false)
    const bool x0ok = fx0 >= 0.f && fx0 <= (float)(a.W - 1), x1ok = fx0 + 1.f >= 0.f && fx0 + 1.f <= (float)(a.W - 1);
    const bool y0ok = fy0 >= 0.f && fy0 <= (float)(a.H - 1), y1ok = fy0 + 1.f >= 0.f && fy0 + 1.f <= (float)(a.H - 1);
    const int x0 = x0ok ? (int)fx0 : 0, x1 = x1ok ? (int)fx0 + 1 : 0, y0 = y0ok ? (int)fy0 : 0, y1 = y1ok ? (int)fy0 + 1 : 0;
    const float w00 = (x0ok && y0ok) ? (1.f - tx) * (1.f - ty) : 0.f, w01 = (x1ok && y0ok) ? tx * (1.f - ty) : 0.f;
    const float w10 = (x0ok && y1ok) ? (1.f - tx) * ty : 0.f, w11 = (x1ok && y1ok) ? tx * ty : 0.f;
    const float* fb = a.feat + (size_t)b * a.H * a.W * a.C;
    const float* t00 = fb + ((size_t)y0 * a.W + x0) * a.C;
    const float* t01 = fb + ((size_t)y0 * a.W + x1) * a.C;
    const float* t10 = fb + ((size_t)y1 * a.W + x0) * a.C;
    const float* t11 = fb + ((size_t)y1 * a.W + x1) * a.C;
    float* o = a.feats + (size_t)g * a.C;
#pragma unroll 2
    for (int ch = gl * 4; ch < a.C; ch += 32) {
      const float4 v00 = __ldg(reinterpret_cast<const float4*>(t00 + ch));
      const float4 v01 = __ldg(reinterpret_cast<const float4*>(t01 + ch));
      const float4 v10 = __ldg(reinterpret_cast<const float4*>(t10 + ch));
      const float4 v11 = __ldg(reinterpret_cast<const float4*>(t11 + ch));
      float4 r;
      r.x = fmaf(v11.x, w11, fmaf(v10.x, w10, fmaf(v01.x, w01, v00.x * w00)));
      r.y = fmaf(v11.y, w11, fmaf(v10.y, w10, fmaf(v01.y, w01, v00.y * w00)));
      r.z = fmaf(v11.z, w11, fmaf(v10.z, w10, fmaf(v01.z, w01, v00.z * w00)));
      r.w = fmaf(v11.w, w11, fmaf(v10.w, w10, fmaf(v01.w, w01, v00.w * w00)));
      __stcs(reinterpret_cast<float4*>(o + ch), r);  // streamed: the output is not re-read by this kernel
    }
  }
}

// Adjoint of the gather with respect to the feature map (the hourglass filter of netLocal trains through it;
// the reference routes this through op/grid_sample_gradfix.py): d_map[b, tap] += w_tap * d_feats[b, n] for the
// four taps of every point, vector atomics (red.global.add.v4.f32) into a zeroed channels-last gradient map.
// Points and calibration receive no gradient (the sample positions are detached on this path).
__global__ void __launch_bounds__(256) local_query_bwd_kernel(const __grid_constant__ LocalQueryArgs a,
                                                              const float* __restrict__ d_feats, float* d_map) {
  const int gl = threadIdx.x & 7;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) >> 3;
  const int64_t total = (int64_t)a.B * a.N;
  float zsign;
  {
    const float* c = a.calibs;
    const float hz = __fadd_rn(fmaf(c[10], a.points[2 * a.pk], fmaf(c[9], a.points[a.pk], __fmul_rn(c[8], a.points[0]))), c[11]);
    zsign = hz < 0.f ? -1.f : 1.f;
  }
  for (int64_t g = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3); g < total; g += n_groups) {
    const int b = (int)(g / a.N), n = (int)(g - (int64_t)b * a.N);
    const float* c = a.calibs + (size_t)b * a.calib_stride;
    const float* p = a.points + b * a.pb + n * a.pn;
    const float px = p[0], py = p[a.pk], pz = p[2 * a.pk];
    const float hx = __fadd_rn(fmaf(c[2], pz, fmaf(c[1], py, __fmul_rn(c[0], px))), c[3]);
    const float hy = __fadd_rn(fmaf(c[6], pz, fmaf(c[5], py, __fmul_rn(c[4], px))), c[7]);
    const float hz = __fadd_rn(fmaf(c[10], pz, fmaf(c[9], py, __fmul_rn(c[8], px))), c[11]);
    const float z = hz * zsign;
    const float x = __fdiv_rn(hx, z), y = -__fdiv_rn(hy, z);
    const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(x, 1.f), (float)a.W), 1.f), 0.5f);
    const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(y, 1.f), (float)a.H), 1.f), 0.5f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float tx = ix - fx0, ty = iy - fy0;
    const bool x0ok = fx0 >= 0.f && fx0 <= (float)(a.W - 1), x1ok = fx0 + 1.f >= 0.f && fx0 + 1.f <= (float)(a.W - 1);
    const bool y0ok = fy0 >= 0.f && fy0 <= (float)(a.H - 1), y1ok = fy0 + 1.f >= 0.f && fy0 + 1.f <= (float)(a.H - 1);
    const int x0 = x0ok ? (int)fx0 : 0, x1 = x1ok ? (int)fx0 + 1 : 0, y0 = y0ok ? (int)fy0 : 0, y1 = y1ok ? (int)fy0 + 1 : 0;
    const float w00 = (x0ok && y0ok) ? (1.f - tx) * (1.f - ty) : 0.f, w01 = (x1ok && y0ok) ? tx * (1.f - ty) : 0.f;
    const float w10 = (x0ok && y1ok) ? (1.f - tx) * ty : 0.f, w11 = (x1ok && y1ok) ? tx * ty : 0.f;
    float* fb = d_map + (size_t)b * a.H * a.W * a.C;
    float* t00 = fb + ((size_t)y0 * a.W + x0) * a.C;
    float* t01 = fb + ((size_t)y0 * a.W + x1) * a.C;
    float* t10 = fb + ((size_t)y1 * a.W + x0) * a.C;
    float* t11 = fb + ((size_t)y1 * a.W + x1) * a.C;
    const float* gi = d_feats + (size_t)g * a.C;
    for (int ch = gl * 4; ch < a.C; ch += 32) {
      const float4 gv = __ldcs(reinterpret_cast<const float4*>(gi + ch));
      auto add = [&](float* dst, float w) {
        if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(dst + ch), make_float4(gv.x * w, gv.y * w, gv.z * w, gv.w * w));
      };
      add(t00, w00), add(t01, w01), add(t10, w10), add(t11, w11);
    }
  }
}

}  // namespace e3

using namespace e3;

extern "C" int e3_local_feature_query_bwd(const float* d_feats, const float* points, int64_t pts_batch_stride,
                                          int64_t pts_coord_stride, int64_t pts_point_stride, const float* calibs,
                                          int calib_stride, int batch, int n_points, int h, int w, int c,
                                          float* d_feat_nhwc, void* stream) {
  E3_REQUIRE(batch >= 0 && n_points >= 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, E3_ERR_BAD_ARG,
             "e3_local_feature_query_bwd: bad shape (channels %% 4 == 0 required)");
  E3_REQUIRE(calib_stride >= 12, E3_ERR_BAD_ARG, "e3_local_feature_query_bwd: calib_stride must be >= 12");
  if (batch == 0) return E3_OK;
  E3_REQUIRE(d_feat_nhwc, E3_ERR_BAD_ARG, "e3_local_feature_query_bwd: null output");
  E3_CUDA(cudaMemsetAsync(d_feat_nhwc, 0, (size_t)batch * h * w * c * sizeof(float), as_stream(stream)));
  if (n_points == 0) return E3_OK;
  E3_REQUIRE(d_feats && points && calibs, E3_ERR_BAD_ARG, "e3_local_feature_query_bwd: null argument");
  LocalQueryArgs a{nullptr, points, pts_batch_stride, pts_coord_stride, pts_point_stride, calibs, calib_stride,
                   batch, n_points, h, w, c, nullptr, nullptr, nullptr, nullptr};
  const int64_t groups = (int64_t)batch * n_points;
  int64_t blocks = (groups * 8 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  local_query_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a, d_feats, d_feat_nhwc);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" int e3_local_feature_query(const float* feat_nhwc, const float* points, int64_t pts_batch_stride,
                                      int64_t pts_coord_stride, int64_t pts_point_stride, const float* calibs,
                                      int calib_stride, int batch, int n_points, int h, int w, int c,
                                      float* feats, float* proj_xy, float* depth, unsigned char* in_img,
                                      void* stream) {
  E3_REQUIRE(batch >= 0 && n_points >= 0 && h > 0 && w > 0 && c > 0, E3_ERR_BAD_ARG,
             "e3_local_feature_query: bad shape");
  E3_REQUIRE(c % 4 == 0, E3_ERR_UNSUPPORTED, "e3_local_feature_query: channels %% 4 == 0 required (got %d)", c);
  E3_REQUIRE(calib_stride >= 12, E3_ERR_BAD_ARG, "e3_local_feature_query: calib_stride must be >= 12 (3x4 rows)");
  if (batch == 0 || n_points == 0) return E3_OK;
  E3_REQUIRE(points && calibs && (!feats || feat_nhwc), E3_ERR_BAD_ARG, "e3_local_feature_query: null argument");
  LocalQueryArgs a{feat_nhwc, points, pts_batch_stride, pts_coord_stride, pts_point_stride, calibs, calib_stride,
                   batch, n_points, h, w, c, feats, proj_xy, depth, in_img};
  const int64_t groups = (int64_t)batch * n_points;
  int64_t blocks = (groups * 8 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  local_query_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}
