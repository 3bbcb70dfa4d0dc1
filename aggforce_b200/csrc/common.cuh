// Shared device/host helpers for libagf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/agf_b200.h"

#ifndef __CUDA_ARCH__
#define AGF_HOST_ONLY 1
#endif

namespace agf {

// ---------------------------------------------------------------- host: error handling
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define AGF_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return agf::cuda_fail(_e, #expr); \
  } while (0)

#define AGF_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) {                   \
      agf::set_error(__VA_ARGS__);   \
      return AGF_E_INVALID;          \
    }                                \
  } while (0)

int sm_count();

// ---------------------------------------------------------------- device: mbarrier + TMA bulk
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Orders prior generic-proxy shared-memory accesses before subsequent async-proxy (TMA) ones.
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Arrive without the release fence (no MEMBAR.ALL.CTA).  Only for use right after a bar.sync that
// already ordered the producers' shared-memory writes (BAR.SYNC drains pending STS).
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0,
// both addresses 16-byte aligned).  SASS: UBLKCP.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// One chunk is moved as several bulk copies of at most this many bytes (all completing on the
// same mbarrier): the TMA unit overlaps independent copies, a single large one is latency bound.
#ifndef AGF_BULK_PIECE
#define AGF_BULK_PIECE 4096u
#endif
constexpr uint32_t kBulkPiece = AGF_BULK_PIECE;

// ---------------------------------------------------------------- device: FP64 tensor core
// D(8x8) += A(8x4, row) * B(4x8, col).  Fragment ownership (lane = 4*g + q):
//   a : A[g][q]      b : B[q][g]      c0,c1 : C[g][2q], C[g][2q+1]
// SASS: DMMA.8x8x4 (measured 37.1 TFLOP/s on B200, profiles/r01_fp64_hbm_microbench.json).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <typename T>
__device__ __forceinline__ double to_f64(T v) {
  return static_cast<double>(v);
}

// Pads a row length (in doubles) so that DMMA fragment loads -- lane (g,q) reading
// element [k0+q][x0+g] or [k0+g][x0+q] of a [k][x] panel -- are bank-conflict free:
// needs stride % 16 in {4, 12}.
__host__ __device__ inline int panel_stride(int cols) {
  int s = cols;
  while ((s % 16) != 4 && (s % 16) != 12) ++s;
  return s;
}

}  // namespace agf
