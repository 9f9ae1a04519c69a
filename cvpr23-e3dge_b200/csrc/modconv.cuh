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
};

bool tc_conv_supported(int B, int H, int W, int Cin, int N);
size_t tc_conv_split_bytes(int B, int H, int W, int Cin);
int tc_conv_pack_weight(const float* weight, int cout, int cin, int upsample, float scale,
                        void* packed_bf16, cudaStream_t stream);
int tc_conv_launch(const ConvGemmArgs& a, int taps, const void* packed_bf16, void* split_scratch,
                   cudaStream_t stream);

}  // namespace e3
