// Roofline denominators measured in the run that quotes them (bench.py): the FP64 tensor-core
// (DMMA.8x8x4) issue rate and a streaming read bandwidth.  Not on the product path.
#include "common.cuh"

namespace agf {

// 8 independent accumulator tiles per warp: the DMMA pipe stays full (37.1 TFLOP/s on B200 with
// 4-16 warps per SM, tools/microbench/fp64_peak.cu).
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c[i][0] = threadIdx.x;
    c[i][1] = i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) read_peak_kernel(const float4* __restrict__ src, int64_t n_vec, float* out) {
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n_vec; i += 4 * stride) {
    const float4 v0 = __ldg(src + i), v1 = __ldg(src + i + stride), v2 = __ldg(src + i + 2 * stride),
                 v3 = __ldg(src + i + 3 * stride);
    acc += v0.x + v0.y + v0.z + v0.w + v1.x + v1.y + v1.z + v1.w + v2.x + v2.y + v2.z + v2.w + v3.x + v3.y + v3.z + v3.w;
  }
  for (; i < n_vec; i += stride) {
    const float4 v = __ldg(src + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) out[0] = acc;  // keeps the loads alive
}

}  // namespace agf

extern "C" int agf_probe_dmma(int32_t iters, double* sink, int64_t* flop_out, void* stream) {
  using namespace agf;
  AGF_REQUIRE(sink && iters > 0, "agf_probe_dmma: sink must hold sm_count*8*256 doubles");
  const int blocks = sm_count() * 8;
  dmma_peak_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sink, iters, 1.0000001, 0.9999999);
  AGF_CUDA_TRY(cudaGetLastError());
  // per warp and iteration: 8 DMMA.8x8x4 of 2*8*8*4 flop
  if (flop_out) *flop_out = (int64_t)blocks * 8 * (int64_t)iters * 8 * 512;
  return AGF_OK;
}

extern "C" int agf_probe_read(const void* src, int64_t bytes, float* sink, void* stream) {
  using namespace agf;
  AGF_REQUIRE(src && sink && bytes >= 16 && (reinterpret_cast<uintptr_t>(src) & 15) == 0, "agf_probe_read: bad buffer");
  read_peak_kernel<<<sm_count() * 4, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(src), bytes / 16, sink);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}
