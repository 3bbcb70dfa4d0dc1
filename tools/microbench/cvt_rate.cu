// Microbenchmark: throughput of f32->f64 conversion (F2F.F64.F32) vs an integer bit-expansion.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double cvt_int(float x) {
  unsigned b = __float_as_uint(x), a = b & 0x7fffffffu;
  unsigned hi = (b & 0x80000000u) | ((a >> 3) + 0x38000000u);
  if (a < 0x00800000u) hi = b & 0x80000000u;
  return __hiloint2double((int)hi, (int)(b << 29));
}
template <int MODE>
__global__ void k(const float* in, double* out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x + 32 * i];
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double d;
      if (MODE == 0) d = (double)v[i];
      else if (MODE == 1) d = cvt_int(v[i]);
      else d = 1.0;
      if (MODE == 3) { acc[i] += d; } else {
        // keep the conversion live and loop-variant without adding FP64 work: xor the float bits
        v[i] = __uint_as_float(__float_as_uint(v[i]) ^ (unsigned)(__double2hiint(d) & 1) ^ (unsigned)it);
        acc[i] = d;
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
float run(const float* in, double* out, int iters, int blocks, int threads) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(in, out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* in; double* out; cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0x3f, 4096 * 4); cudaMalloc(&out, 8 * 148 * 8 * 1024);
  int iters = 20000, blocks = p.multiProcessorCount * 4, threads = 256;
  double n = 8.0 * iters * blocks * threads;
  double clk = p.clockRate * 1e3;
  float a = run<0>(in, out, iters, blocks, threads), b = run<1>(in, out, iters, blocks, threads);
  printf("{\"f2f_f64_f32_lanes_per_clk_per_sm\": %.2f, \"int_expand_lanes_per_clk_per_sm\": %.2f, \"ms\": [%.3f, %.3f]}\n",
         n / (a * 1e-3) / clk / p.multiProcessorCount, n / (b * 1e-3) / clk / p.multiProcessorCount, a, b);
  return 0;
}
