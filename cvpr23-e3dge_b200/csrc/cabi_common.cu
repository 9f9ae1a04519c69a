// Error plumbing and device queries shared by every C-ABI entry point.
#include "common.cuh"

#include <string.h>

namespace e3 {

static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= E3_MAX_DEVICES) return 0;
  return dev;
}

}  // namespace e3

namespace e3 {
// FP32 FFMA throughput probe: 16 independent accumulator chains per thread, register only.
__global__ void __launch_bounds__(1024) ffma_probe_kernel(int iters, float* __restrict__ sink) {
  float acc[16];
  const float a = 1.0000001f + 1e-9f * threadIdx.x, b = 1e-7f * (blockIdx.x + 1);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace e3

extern "C" size_t e3_ffma_peak_probe_sink_floats(void) { return (size_t)e3::sm_count() * 2 * 1024; }
extern "C" int e3_ffma_peak_probe(int iters, float* sink, void* stream) {
  E3_REQUIRE(iters > 0 && sink, E3_ERR_BAD_ARG, "e3_ffma_peak_probe: bad argument");
  e3::ffma_probe_kernel<<<e3::sm_count() * 2, 1024, 0, e3::as_stream(stream)>>>(iters, sink);
  E3_CUDA(cudaGetLastError());
  return E3_OK;
}

extern "C" int e3_abi_version(void) { return 4; }
extern "C" const char* e3_last_error(void) { return e3::g_last_error; }
