// Prototype of the int8 (Ozaki-sliced) Gram on the 5th-generation tensor cores: tcgen05.mma kind::i8 with
// int32 accumulators in TMEM, operands in the canonical MN-major no-swizzle shared-memory layout.
//
//   G_level[l][x][y] = sum_{s + t = l} sum_k D_s[k][x] * D_t[k][y]        x < 128 (M), y < 96 (N), l < 5
// D_s: the s-th signed 8-bit digit plane of the scaled group sums (tools/ozaki/ozaki_numerics.py).
// 15 MMAs (M=128, N=96, K=32) per block of 32 rows, five accumulators of 96 TMEM columns (480 <= 512).
//
// This program (1) checks the result against the CPU for random digits, (2) measures the MMA issue rate
// of one SM.  Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o i8_syrk_probe i8_syrk_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int kSlices = 5;
constexpr int kM = 128, kN = 96, kK = 32;       // one tcgen05.mma.kind::i8
constexpr int kGroupBytes = 8 * kM;             // 8 k-rows x 128 columns = 1024 bytes (8 core matrices)
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(kN >> 3) << 17) |
                            ((uint32_t)(kM >> 4) << 24);  // D s32, A/B s8, both MN-major, N = 96, M = 128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical MN-major, no swizzle, 8-bit: core matrix = 8 k-rows x 16 bytes; MN blocks SBO apart, k groups LBO apart
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, layout type SWIZZLE_NONE
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc),
      "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (long long spin = 0; spin < (1ll << 26); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

__global__ void __launch_bounds__(192, 1) probe_kernel(const int8_t* __restrict__ digits, int k_rows, int32_t* __restrict__ out,
                                                       int iters, long long* cycles, int* fail) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  const int n_groups = k_rows / 8;
  const size_t plane_bytes = (size_t)n_groups * kGroupBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // digits[s][k][x] (row-major) -> canonical layout
  for (size_t i = threadIdx.x; i < (size_t)kSlices * k_rows * kM; i += blockDim.x) {
    const int s = (int)(i / ((size_t)k_rows * kM));
    const int rem = (int)(i - (size_t)s * k_rows * kM);
    const int k = rem / kM, x = rem - k * kM;
    smem[s * plane_bytes + (size_t)(k >> 3) * kGroupBytes + (x >> 4) * 128 + (k & 7) * 16 + (x & 15)] = (unsigned char)digits[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 4 && lane == 0) {
    const long long t0 = clock64();
    const uint32_t base = smem_u32(smem);
    for (int it = 0; it < iters; ++it) {
      for (int kb = 0; kb < k_rows / kK; ++kb) {
        uint32_t started = (it > 0 || kb > 0) ? 0x1Fu : 0u;  // bit l: accumulator l already holds a product
#pragma unroll
        for (int s = 0; s < kSlices; ++s) {
#pragma unroll
          for (int t = 0; t < kSlices - s; ++t) {
            const int l = s + t;
            const uint64_t da = make_desc(base + (uint32_t)(s * plane_bytes) + (uint32_t)kb * 4 * kGroupBytes, kGroupBytes, 128);
            const uint64_t db = make_desc(base + (uint32_t)(t * plane_bytes) + (uint32_t)kb * 4 * kGroupBytes, kGroupBytes, 128);
            mma_i8(tmem_base + (uint32_t)(l * kN), da, db, (started >> l) & 1u);
            started |= 1u << l;
          }
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    if (!mbar_wait_bounded(&done_bar, 0)) *fail = 1;
    cycles[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4 && *fail == 0) {
    // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (row x = 32 w + lane), 32 columns at a time
    for (int l = 0; l < kSlices; ++l) {
      for (int c0 = 0; c0 < kN; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(l * kN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int x = warp * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) out[((size_t)l * kM + x) * kN + c0 + j] = (int32_t)r[j];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

int main(int argc, char** argv) {
  const int k_rows = argc > 1 ? atoi(argv[1]) : 96;  // multiple of 32; 5 planes x k_rows x 128 bytes of shared memory
  const int iters = argc > 2 ? atoi(argv[2]) : 2000;
  if (k_rows % 32 || k_rows <= 0 || (size_t)kSlices * k_rows * kM > 200 * 1024) { printf("bad k_rows\n"); return 1; }
  std::vector<int8_t> h((size_t)kSlices * k_rows * kM);
  srand(7);
  for (auto& v : h) v = (int8_t)(rand() % 256 - 128);
  int8_t* d_digits; int32_t* d_out; long long* d_cyc; int* d_fail;
  CK(cudaMalloc(&d_digits, h.size()));
  CK(cudaMalloc(&d_out, sizeof(int32_t) * kSlices * kM * kN));
  CK(cudaMalloc(&d_cyc, sizeof(long long)));
  CK(cudaMalloc(&d_fail, sizeof(int)));
  CK(cudaMemcpy(d_digits, h.data(), h.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_fail, 0, sizeof(int)));
  CK(cudaMemset(d_out, 0xff, sizeof(int32_t) * kSlices * kM * kN));
  const size_t smem = (size_t)kSlices * k_rows * kM;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // correctness: one pass
  probe_kernel<<<1, 192, smem>>>(d_digits, k_rows, d_out, 1, d_cyc, d_fail);
  CK(cudaDeviceSynchronize());
  int fail = 0;
  CK(cudaMemcpy(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost));
  if (fail) { printf("FAILED: the MMA commit never arrived\n"); return 2; }
  std::vector<int32_t> got((size_t)kSlices * kM * kN);
  CK(cudaMemcpy(got.data(), d_out, got.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int l = 0; l < kSlices; ++l)
    for (int x = 0; x < kM; ++x)
      for (int y = 0; y < kN; ++y) {
        long long ref = 0;
        for (int s = 0; s <= l; ++s) {
          const int t = l - s;
          for (int k = 0; k < k_rows; ++k)
            ref += (long long)h[((size_t)s * k_rows + k) * kM + x] * (long long)h[((size_t)t * k_rows + k) * kM + y];
        }
        if (ref != got[((size_t)l * kM + x) * kN + y]) {
          if (bad < 5) printf("mismatch level %d x %d y %d: got %d want %lld\n", l, x, y, got[((size_t)l * kM + x) * kN + y], ref);
          ++bad;
        }
      }
  printf("correctness: %lld mismatches of %d entries (k_rows %d)\n", bad, kSlices * kM * kN, k_rows);
  // throughput of one SM, and of all SMs at once
  int sms = 0, clock_khz = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  probe_kernel<<<1, 192, smem>>>(d_digits, k_rows, d_out, iters, d_cyc, d_fail);
  CK(cudaDeviceSynchronize());
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
  const double mmas = (double)iters * (k_rows / kK) * 15.0;
  const double ops = mmas * 2.0 * kM * kN * kK;
  printf("one SM: %.0f MMAs (M128 N96 K32) in %lld cycles = %.1f cycles per MMA = %.0f int8 op/clk/SM\n", mmas, cyc,
         cyc / mmas, ops / cyc);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  probe_kernel<<<sms, 192, smem>>>(d_digits, k_rows, d_out, iters, d_cyc, d_fail);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("all %d SMs: %.3f ms -> %.1f dense int8 TOP/s; 15 MMAs cover 32 rows = 10.67 frames: %.2e frames/s of MMA issue\n", sms, ms,
         ops * sms / (ms * 1e-3) / 1e12, (double)iters * (k_rows / kK) * 32.0 / 3.0 * sms / (ms * 1e-3));
  return bad ? 3 : 0;
}
