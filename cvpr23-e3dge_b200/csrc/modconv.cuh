// Shared between the CUDA-core (modconv.cu) and tensor-core (tc_conv.cu) conv paths.
#pragma once
#include "common.cuh"

namespace e3 {

// C[m][n] = sum_{tap,ci} (x[b, y+dy, x+dx, ci] * s[b,ci]) * W[tap][ci][n],  m = (b,y,x)
struct ConvGemmArgs {
  const float* x;   // [B,H,W,Cin] NHWC
  const float* s;   // [B,Cin]
  const float* wg;  // fp32 GEMM-major weights [TAPS][Cin][N] (CUDA-core path)
  float* out;       // [M][N]
  int B, H, W, Cin, N;
  // epilogue: mode 0 raw store; 1 lrelu(d*acc + noise_w*noise + bias)*sqrt2; 2 d*acc
  int mode;
  const float* d;         // [B,N]
  const float* noise;     // [H*W] (+ b*noise_bstride)
  int64_t noise_bstride;
  const float* noise_w;   // [1]
  const float* act_bias;  // [N]
  // 1: `x` is the parity-planar transposed-conv gradient [py][px][B][H+1][W+1][Cin] and tap (ky,kx)
  // reads plane (ky&1, kx&1) at (y + (ky>>1), x + (kx>>1)) — the stride-2 gather of the up-conv
  // backward as nine dense shifted reads (tensor-core path only, operands pre-split by the caller)
  int planar;
  // 1: x-pair view of a 32 -> 32 conv (tensor-core path): the operands are [B][H][W/2][2 pixels x 32 ch] with
  // W here = W/2, Cin = N = 64 and a block-structured weight image (tc_conv_pack_weight layout 4); the
  // epilogue's only difference is the noise, which belongs to pixel 2*x + (column >> 5)
  int pairx;
};

bool tc_conv_supported(int B, int H, int W, int Cin, int N);
size_t tc_conv_split_bytes(int B, int H, int W, int Cin);
int tc_conv_pack_weight(const float* weight, int cout, int cin, int upsample, float scale,
                        void* packed_bf16, cudaStream_t stream);
// 32 -> 32 plain convs run on the tensor cores through the x-pair view (ConvGemmArgs::pairx)
bool tc_conv_pairx_supported(int B, int H, int W, int Cin, int N);
size_t tc_conv_packed_bf16_bytes(int cout, int cin);
int tc_conv_launch(const ConvGemmArgs& a, int taps, const void* packed_bf16, void* split_scratch,
                   cudaStream_t stream);
// same without the modulate+split pass: xs_hi / xs_lo are the caller's bf16 operand halves
int tc_conv_launch_presplit(const ConvGemmArgs& a, int taps, const void* packed_bf16, const void* xs_hi,
                            const void* xs_lo, cudaStream_t stream);
// upsampling conv as four parity-phase convolutions of a zero-padded flat pixel grid (tc_conv.cu):
// T [4][B*(H+1)*(W+1)][cout] = conv_transpose2d(x*s, W, stride 2) by output parity
bool tc_upconv_supported(int B, int H, int W, int Cin, int cout);
size_t tc_upconv_split_bytes(int B, int H, int W, int Cin);
size_t tc_upconv_t_bytes(int B, int H, int W, int cout);
int tc_upconv_phase_launch(const float* x, const float* s, int B, int H, int W, int Cin, int cout,
                           const void* packed_bf16, void* split_scratch, float* t_out, cudaStream_t stream);
// exact-fp32 implicit GEMM on the CUDA cores (modconv.cu); needs Cin % 16 == 0 and N % 4 == 0
int conv_gemm_ffma_launch(const ConvGemmArgs& a, int taps, cudaStream_t stream);
size_t tc_conv_planar_elems(int B, int H, int W, int C);  // elements of one planar operand half
const void* conv_packed_bf16_part(const void* packed, int cout, int cin);

}  // namespace e3
