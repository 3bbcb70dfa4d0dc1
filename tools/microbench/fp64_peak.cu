// Microbenchmark: FP64 throughput on B200 via DFMA, DMMA m8n8k4, DMMA m16n8k8, DMMA m16n8k16,
// and streaming HBM read bandwidth. Feeds the roofline denominators used in DESIGN.md.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double c[NT][2];
#pragma unroll
  for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_dmma1688(double* out, int iters, double a, double b) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_dmma16816(double* out, int iters, double a, double b) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_read(const float4* __restrict__ p, size_t n, float* out) {
  float s = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 v = __ldg(p + i);
    s += v.x + v.y + v.z + v.w;
  }
  if (s == 123.456f) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
  for (int warps : {4, 8, 16}) {
    int threads = warps * 32; int blocks = sms * 2;
    {
      float ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 8 * iters * (double)threads * blocks;
      printf(", \"dfma_w%d_tflops\": %.2f", warps, fl / ms / 1e9);
    }
    {
      float ms = time_ms([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)warps * blocks;
      printf(", \"dmma884_w%d_tflops\": %.2f", warps, fl / ms / 1e9);
    }
    {
      float ms = time_ms([&] { k_dmma1688<8><<<blocks, threads>>>(out, iters / 2, 1.0000001, 1e-9); });
      double fl = 2.0 * 16 * 8 * 8 * 8 * (iters / 2) * (double)warps * blocks;
      printf(", \"dmma1688_w%d_tflops\": %.2f", warps, fl / ms / 1e9);
    }
    {
      float ms = time_ms([&] { k_dmma16816<8><<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
      double fl = 2.0 * 16 * 8 * 16 * 8 * (iters / 4) * (double)warps * blocks;
      printf(", \"dmma16816_w%d_tflops\": %.2f", warps, fl / ms / 1e9);
    }
  }
  {
    size_t bytes = (size_t)8 << 30; float4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    float* o2; CK(cudaMalloc(&o2, 4));
    for (int bps : {4, 8, 16}) {
      float ms = time_ms([&] { k_read<<<sms * bps, 256>>>(buf, bytes / 16, o2); });
      printf(", \"hbm_read_b%d_gbs\": %.1f", bps, bytes / ms / 1e6);
    }
  }
  printf("}\n");
  return 0;
}
