// Pieces shared by the two int8 (Ozaki-sliced) Gram kernels: tcgen05.mma kind::i8 issue, shared-memory
// descriptors of the canonical MN-major no-swizzle layout, the fixed-point digit split.
#pragma once
#include "common.cuh"

namespace agf {

// Instruction descriptor: D int32, A and B signed 8-bit, both MN-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// Canonical MN-major, no swizzle, 8-bit operands: core matrix = 8 k-rows x 16 bytes (16 MN positions of one
// k-row are contiguous); groups of 8 k-rows `kgroup_bytes` apart (LBO field), blocks of 16 MN positions
// `mnblock_bytes` apart (SBO field); descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc_mn_i8(uint32_t addr, uint32_t kgroup_bytes, uint32_t mnblock_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((kgroup_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((mnblock_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ void umma_i8_issue(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// byte `b` of four words -> one word (byte i from word i)
__device__ __forceinline__ uint32_t gather_bytes(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int b) {
  const uint32_t sel = (uint32_t)b | ((uint32_t)(4 + b) << 4);  // result byte 0 = w0.b, byte 1 = w1.b
  const uint32_t t01 = __byte_perm(w0, w1, sel), t23 = __byte_perm(w2, w3, sel);
  return __byte_perm(t01, t23, 0x5410);
}

// Scale of a column: values below 2^(E-1) fit the 39-bit fixed point; E leaves 2-4x headroom over the sample.
__device__ __forceinline__ int column_exponent(unsigned long long max_bits) {
  const double m = __longlong_as_double((long long)max_bits);
  if (!(m > 0.0) || !(m < 1.0e300)) return -900;
  int e = ilogb(m) + 3;
  return e < -900 ? -900 : (e > 900 ? 900 : e);
}

// Fixed point in ONE instruction: t = fma(v, 2^(39-E), kI8Magic) holds q + 0x8080808080 in its low 40 bits
// (round to nearest); the upper 24 bits equal kI8HiExpect's exactly when |q| is in range.  Digit s (most
// significant first) = byte 4 - s of the low 40 bits with the top bit flipped (offset binary -> two's complement).
constexpr double kI8Magic = 6755399441055744.0 + 551911719040.0;  // 1.5 * 2^52 + 0x8080808080
constexpr uint32_t kI8HiExpect = 0x43380000u;                     // upper word of 1.5 * 2^52 (low byte: digit 0)

}  // namespace agf
