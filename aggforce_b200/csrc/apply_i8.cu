// Kernel (d), Blackwell tensor-core path for LARGE dense maps (n_cg > 64, float32 input): the map application
// through int8 digit planes on tcgen05 -- the counterpart of gram_i8t.cu for
//     out[t][b][d] = sum_u W[b][u] S[t][u][d]          (src/aggforce/util.py:119-124 of the reference, in float64)
// with S the sums of the sites sharing a unique matrix column u.  The FP64 DMMA GEMM it replaces
// (agf_map_apply_ws) runs at 0.5-0.75 of the 37 TFLOP/s FP64 tensor roof; the int8 pipe is 40x wider.
//
//   * S -> digits exactly as for the Gram (i8_digits.cuh: per-column power-of-two scale 2^E_u from a sample,
//     39-bit fixed point q, five signed 8-bit digits), stored with the columns as the CONTRACTION dimension
//     (K-major core matrices): one (block of 32 frames = 96 (xyz, frame) rows, slab of 32 columns) is 15 360
//     contiguous bytes = the B operand of one MMA k-step for all five planes, one TMA bulk copy.
//   * W -> digits once per call: W'[b][u] = W[b][u] 2^E_u (the column scale moves into the weights), a
//     power-of-two scale 2^F_b per bead, 39-bit fixed point p, five digits, laid out per (block of 128 beads,
//     slab of 32 columns) as 20 480 contiguous bytes = the A operand.
//   * persistent GEMM (one CTA per SM): unit = (frame block, bead block); per k-step 8 tcgen05.mma.kind::i8 --
//     the weight plane A_s against the stack of sum planes B_0 .. B_{4-s} (N = 96 (5 - s), halves when > 256) --
//     accumulate the 15 plane products with s + t <= 4 EXACTLY in int32 in five TMEM accumulators (one per
//     level l = s + t); epilogue: out = 2^(F_b - 14) sum_l 2^(-8 l) acc_l in float64, stored, squared and summed.
//   * frames holding a non-finite or out-of-range value are taken out of the planes (scrub) and computed by
//     i8a_leftover_kernel in float64 with the reference's NaN protocol (map/core.py:219-240), after the GEMM.
// Error: the dropped products (s + t >= 5) and the two roundings are below 2^-38 of (largest |W'| of the bead)
// x (column scale) per term -- about 1e-11 of the result for maps without catastrophic cancellation (bar 1e-6).
#include <stdlib.h>

#include "i8_digits.cuh"

namespace agf {

constexpr int kA_M = 128, kA_N = 96;
constexpr int kA_AHalf = kT_Slices * (kA_M / 8) * 128;  // one half (16 columns) of a k-slab, five planes: 10 240
constexpr int kA_BHalf = kT_Slices * (kA_N / 8) * 128;  // 7 680
constexpr int kA_AStage = 2 * kA_AHalf, kA_BStage = 2 * kA_BHalf;
constexpr int kA_StageBytes = kA_AStage + kA_BStage;     // 35 840
constexpr int kA_Stages = 6;
constexpr int kA_Threads = 192;  // load warp, MMA warp, four epilogue warps
constexpr int kA_MaxUcol = 8192;
constexpr double kA_NanRtol = 1e-5;  // np.allclose default, as in apply.cu

// Instruction descriptor: D int32, A and B signed 8-bit, both K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_i8_k(int m, int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// Canonical K-major no-swizzle 8-bit operand: core matrix = 8 MN rows x 16 bytes of K; the two 16-byte halves of a
// k-step `khalf_bytes` apart (LBO field), groups of 8 MN rows `group_bytes` apart (SBO field).
__device__ __forceinline__ uint64_t umma_desc_k_i8(uint32_t addr, uint32_t khalf_bytes, uint32_t group_bytes) {
  return umma_desc_mn_i8(addr, khalf_bytes, group_bytes);  // same fields, see i8.cuh
}

// ------------------------------------------------------------------------------------------------ weights
// row scale: max_u |W[b][u]| 2^E_u  (thread = bead, block row = slice of the columns)
__global__ void __launch_bounds__(128) i8a_rowmax_kernel(const double* __restrict__ umat_t, int n_ucol, int n_cg,
                                                         const int32_t* __restrict__ exps,
                                                         unsigned long long* __restrict__ rowmax_bits) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_cg) return;
  double best = 0.0;
  for (int u = blockIdx.y; u < n_ucol; u += gridDim.y)
    best = fmax(best, fabs(ldexp(__ldg(umat_t + (int64_t)u * n_cg + b), __ldg(exps + u))));
  atomicMax(rowmax_bits + b, (unsigned long long)__double_as_longlong(best));
}

// digits of W': thread = (bead, 16 columns) -> one 16-byte row of a core matrix per plane
__global__ void __launch_bounds__(128) i8a_wdigits_kernel(const double* __restrict__ umat_t, int n_ucol, int n_cg, int n_ks,
                                                          const int32_t* __restrict__ exps,
                                                          const unsigned long long* __restrict__ rowmax_bits,
                                                          double* __restrict__ fpow2, unsigned char* __restrict__ wdig) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;  // < n_mb * 128
  const int ub = blockIdx.y;                            // block of 16 columns
  int fexp = -900;
  if (b < n_cg) {
    const double m = __longlong_as_double((long long)rowmax_bits[b]);
    if (m > 0.0 && m < 1.0e300) fexp = ilogb(m) + 2;  // |W'| < 2^(F - 1): the fixed point cannot overflow
  }
  if (ub == 0) fpow2[b] = ldexp(1.0, fexp - 14);
  uint32_t words[kT_Slices][4];
#pragma unroll
  for (int s = 0; s < kT_Slices; ++s)
#pragma unroll
    for (int w = 0; w < 4; ++w) words[s][w] = 0u;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int u = ub * 16 + i;
    double v = 0.0;
    if (b < n_cg && u < n_ucol) v = ldexp(__ldg(umat_t + (int64_t)u * n_cg + b), __ldg(exps + u) + 39 - fexp);
    const double t = v + kI8Magic;  // round to nearest: q + 0x8080808080 in the low 40 bits
    const uint32_t lo = (uint32_t)__double2loint(t) ^ 0x80808080u, hi = (uint32_t)__double2hiint(t) ^ 0x80u;
    const int sh = 8 * (i & 3);
    words[0][i >> 2] |= (hi & 0xFFu) << sh;
    words[1][i >> 2] |= ((lo >> 24) & 0xFFu) << sh;
    words[2][i >> 2] |= ((lo >> 16) & 0xFFu) << sh;
    words[3][i >> 2] |= ((lo >> 8) & 0xFFu) << sh;
    words[4][i >> 2] |= (lo & 0xFFu) << sh;
  }
  const int mb = b >> 7, bl = b & 127, ks = ub >> 1, j = ub & 1;
#pragma unroll
  for (int s = 0; s < kT_Slices; ++s) {
    unsigned char* dst = wdig + (((((size_t)mb * n_ks + ks) * 2 + j) * kT_Slices + s) * (kA_M / 8) + (bl >> 3)) * 128 + (bl & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(words[s][0], words[s][1], words[s][2], words[s][3]);
  }
}

// ------------------------------------------------------------------------------------------------ GEMM
struct I8aGemmParams {
  const unsigned char* wdig;
  const unsigned char* xdig;
  int32_t n_ks, n_mb, n_rb, n_cg;
  int64_t n_frames;      // frames of the slab
  const double* fpow2;   // [n_mb * 128] 2^(F_b - 14)
  void* out;             // first frame of the slab
  double* sumsq;
};

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TO>
__global__ void __launch_bounds__(kA_Threads, 1) i8a_gemm_kernel(const __grid_constant__ I8aGemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kA_Stages * kA_StageBytes);
  uint64_t* full = bars;                   // [kA_Stages] TMA -> MMA
  uint64_t* empty = full + kA_Stages;      // [kA_Stages] MMA -> TMA (tcgen05.commit)
  uint64_t* acc_full = empty + kA_Stages;  // MMA -> epilogue
  uint64_t* acc_empty = acc_full + 1;      // epilogue (4 warps) -> MMA
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kA_Stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const int n_units = p.n_rb * p.n_mb;  // frame block major: the bead blocks of one frame block run side by side

  if (warp == 0) {
    // ------------------------------------------------ load warp: one thread, four bulk copies per k-step
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int rb = u / p.n_mb, mb = u - rb * p.n_mb;
        const unsigned char* a_src = p.wdig + (size_t)mb * p.n_ks * kA_AStage;
        const unsigned char* b_src = p.xdig + (size_t)rb * p.n_ks * kA_BStage;
        for (int ks = 0; ks < p.n_ks; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1u);
          unsigned char* dst = stages + (size_t)stage * kA_StageBytes;
          mbar_expect_tx(&full[stage], kA_StageBytes);
          tma_bulk_g2s(dst, a_src, kA_AHalf, &full[stage]);
          tma_bulk_g2s(dst + kA_AHalf, a_src + kA_AHalf, kA_AHalf, &full[stage]);
          tma_bulk_g2s(dst + kA_AStage, b_src, kA_BHalf, &full[stage]);
          tma_bulk_g2s(dst + kA_AStage + kA_BHalf, b_src + kA_BHalf, kA_BHalf, &full[stage]);
          a_src += kA_AStage;
          b_src += kA_BStage;
          if (++stage == kA_Stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA warp: one thread issues
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t sbase = smem_u32(stages);
      bool first_unit = true;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        if (!first_unit) {  // the epilogue warps have drained the accumulators of the previous unit
          mbar_wait(acc_empty, acc_phase);
          acc_phase ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        first_unit = false;
        for (int ks = 0; ks < p.n_ks; ++ks) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = sbase + (uint32_t)stage * kA_StageBytes;
          const uint32_t b0 = a0 + kA_AStage;
          const uint32_t fresh = ks > 0 ? 1u : 0u;
#pragma unroll
          for (int s = 0; s < kT_Slices; ++s) {
            // weight plane s against the stack of sum planes 0 .. 4 - s (12 row groups each, contiguous)
            const uint64_t da = umma_desc_k_i8(a0 + s * (kA_M / 8) * 128, kA_AHalf, 128);
            const int width = kA_N * (kT_Slices - s);
            const uint32_t acc = s == 0 ? fresh : 1u;  // s = 0 touches every level first
            if (width > 256) {
              const int half = width / 2;
              umma_i8_issue(tmem_base + (uint32_t)(s * kA_N), da, umma_desc_k_i8(b0, kA_BHalf, 128),
                            umma_idesc_i8_k(kA_M, half), acc);
              umma_i8_issue(tmem_base + (uint32_t)(s * kA_N + half), da, umma_desc_k_i8(b0 + half * 16, kA_BHalf, 128),
                            umma_idesc_i8_k(kA_M, half), acc);
            } else {
              umma_i8_issue(tmem_base + (uint32_t)(s * kA_N), da, umma_desc_k_i8(b0, kA_BHalf, 128),
                            umma_idesc_i8_k(kA_M, width), acc);
            }
          }
          umma_commit(&empty[stage]);  // arrives when the tensor core has finished reading the stage
          if (++stage == kA_Stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps: warp w may touch TMEM lanes 32 (w % 4) ...
    const int quarter = warp & 3;
    uint32_t acc_phase = 0;
    double sq = 0.0;
    TO* out = reinterpret_cast<TO*>(p.out);
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int rb = u / p.n_mb, mb = u - rb * p.n_mb;
      const int b = mb * kA_M + quarter * 32 + lane;
      const double fb = __ldg(p.fpow2 + b);
      mbar_wait(acc_full, acc_phase);
      acc_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = 0; c0 < kA_N; c0 += 16) {  // 16 frames of one xyz component
        // the five levels of 16 columns: loads issued back to back, ONE wait; the levels are combined in integers
        // first (hi = 2^8 acc_0 + acc_1, lo = 2^16 acc_2 + 2^8 acc_3 + acc_4: exact in int64), so that two
        // int64 -> float64 conversions per element replace five -- the accumulators are free again sooner
        uint32_t r[kT_Slices][16];
#pragma unroll
        for (int l = 0; l < kT_Slices; ++l) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(l * kA_N + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[l][0]), "=r"(r[l][1]), "=r"(r[l][2]), "=r"(r[l][3]), "=r"(r[l][4]), "=r"(r[l][5]), "=r"(r[l][6]),
                "=r"(r[l][7]), "=r"(r[l][8]), "=r"(r[l][9]), "=r"(r[l][10]), "=r"(r[l][11]), "=r"(r[l][12]), "=r"(r[l][13]),
                "=r"(r[l][14]), "=r"(r[l][15])
              : "r"(taddr));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        double g[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const long long hi = (long long)(int32_t)r[0][i] * 256 + (long long)(int32_t)r[1][i];
          const long long lo = ((long long)(int32_t)r[2][i] * 256 + (long long)(int32_t)r[3][i]) * 256 + (long long)(int32_t)r[4][i];
          g[i] = (double)hi * (1.0 / 256.0) + (double)lo * (1.0 / 4294967296.0);
        }
        if (c0 + 16 >= kA_N) {  // last column group read: the accumulators may be overwritten
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        if (b < p.n_cg) {
          const int d = c0 >> 5;
          const int64_t f0 = (int64_t)rb * kT_ChunkFrames + (c0 & 31);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (f0 + i < p.n_frames) {
              const TO v = static_cast<TO>(g[i] * fb);
              out[((f0 + i) * p.n_cg + b) * 3 + d] = v;
              sq = fma((double)v, (double)v, sq);
            }
          }
        }
      }
    }
    if (p.sumsq != nullptr) {
      sq = warp_sum_f64(sq);
      if (lane == 0 && sq != 0.0) atomicAdd(p.sumsq, sq);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// ------------------------------------------------------------------------------------------------ leftover frames
// One CTA per listed frame, float64, with the NaN protocol of apply.cu's kernels: in nan_mode 1 a NaN input
// counts as 0, nan_flags[0] reports it and nan_flags[1] whether a result would move beyond
// atol + 1e-5 |result| if the NaNs were -1 instead; in nan_mode 0 NaN propagates (0 * NaN = NaN).
template <typename TO>
__global__ void __launch_bounds__(256) i8a_leftover_kernel(const float* __restrict__ x, int n_sites,
                                                           const int32_t* __restrict__ ucol_ptr,
                                                           const int32_t* __restrict__ ucol_sites, int n_ucol,
                                                           const double* __restrict__ umat_t, int n_cg,
                                                           const int32_t* __restrict__ count,
                                                           const int32_t* __restrict__ frames, TO* __restrict__ out,
                                                           double* sumsq, int nan_mode, double nan_atol, int32_t* nan_flags) {
  constexpr int XB = 512;
  __shared__ double xg[XB][3];
  __shared__ double xn[XB][3];
  __shared__ int s_has_nan;
  const int n = *count;
  double sq = 0.0;
  bool saw_nan = false, bad = false;
  for (int li = blockIdx.x; li < n; li += gridDim.x) {
    const int64_t t = frames[li];
    const float* fr = x + t * (int64_t)n_sites * 3;
    for (int cb = 0; cb < n_cg; cb += blockDim.x) {
      const int c = cb + threadIdx.x;
      double acc[3] = {0.0, 0.0, 0.0}, nw[3] = {0.0, 0.0, 0.0};
      bool any_nan = false;
      for (int u0 = 0; u0 < n_ucol; u0 += XB) {
        __syncthreads();
        if (threadIdx.x == 0) s_has_nan = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < XB * 3; i += blockDim.x) {
          const int uu = i / 3, d = i - uu * 3;
          double v = 0.0, nc = 0.0;
          if (u0 + uu < n_ucol) {
            for (int m = __ldg(ucol_ptr + u0 + uu); m < __ldg(ucol_ptr + u0 + uu + 1); ++m) {
              double f = (double)__ldg(fr + 3 * __ldg(ucol_sites + m) + d);
              if (nan_mode && f != f) {
                f = 0.0;
                nc += 1.0;
              }
              v += f;
            }
          }
          xg[uu][d] = v;
          xn[uu][d] = nc;
          if (nc != 0.0) s_has_nan = 1;
        }
        __syncthreads();
        const bool has_nan = s_has_nan != 0;
        any_nan |= has_nan;
        if (c < n_cg) {
          const int lim = min(XB, n_ucol - u0);
          for (int uu = 0; uu < lim; ++uu) {
            const double w = __ldg(umat_t + (int64_t)(u0 + uu) * n_cg + c);
#pragma unroll
            for (int d = 0; d < 3; ++d) acc[d] = fma(w, xg[uu][d], acc[d]);
            if (has_nan) {
#pragma unroll
              for (int d = 0; d < 3; ++d) nw[d] = fma(w, xn[uu][d], nw[d]);
            }
          }
        }
      }
      saw_nan |= any_nan;
      if (c < n_cg) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const TO v = static_cast<TO>(acc[d]);
          out[(t * n_cg + c) * 3 + d] = v;
          sq = fma((double)v, (double)v, sq);
          if (any_nan) bad |= fabs(nw[d]) > nan_atol + kA_NanRtol * fabs(acc[d] - nw[d]);
        }
      }
    }
  }
  if (sumsq != nullptr) {
    sq = warp_sum_f64(sq);
    if ((threadIdx.x & 31) == 0 && sq != 0.0) atomicAdd(sumsq, sq);
  }
  if (nan_mode) {
    if (saw_nan) atomicOr(nan_flags, 1);
    if (bad) atomicOr(nan_flags + 1, 1);
  }
}

struct I8aLayout {
  size_t colmax, exps, scales, pow2, rowmax, fpow2, count, leftover, flags, wdig, xdig, total;
  int64_t slab;
  int n_kpad, n_ks, n_mb;
};

static I8aLayout i8a_layout(int n_ucol, int n_cg, int64_t n_frames) {
  I8aLayout L;
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  L.n_kpad = (n_ucol + kT_PanelCols - 1) / kT_PanelCols * kT_PanelCols;
  L.n_ks = L.n_kpad / 32;
  L.n_mb = (n_cg + kA_M - 1) / kA_M;
  const size_t np = (size_t)L.n_kpad, mp = (size_t)L.n_mb * kA_M;
  L.colmax = 0;
  L.exps = up(np * 8 * kT_MaxSampleGroups);
  L.scales = L.exps + up(np * 4);
  L.pow2 = L.scales + up(np * 8);
  L.rowmax = L.pow2 + up(np * 8);
  L.fpow2 = L.rowmax + up(mp * 8);
  L.count = L.fpow2 + up(mp * 8);
  L.leftover = L.count + 1024;
  const int64_t rounded = (n_frames + kT_ChunkFrames - 1) / kT_ChunkFrames * kT_ChunkFrames;
  L.slab = rounded < kT_SlabFrames ? rounded : kT_SlabFrames;
  L.flags = L.leftover + up((size_t)n_frames * 4);
  L.wdig = L.flags + up((size_t)L.slab * 4);
  L.xdig = L.wdig + (size_t)L.n_mb * L.n_ks * kA_AStage;
  L.total = L.xdig + (size_t)(L.slab / kT_ChunkFrames) * L.n_ks * kA_BStage;
  return L;
}

template <typename TO>
static int i8a_run(const float* x, int64_t n_frames, int32_t n_sites, const int32_t* ucol_ptr, const int32_t* ucol_sites,
                   int32_t n_ucol, const double* umat_t, int32_t n_cg, TO* out, double* sumsq, int nan_mode, double nan_atol,
                   int32_t* nan_flags, char* ws, cudaStream_t s) {
  const I8aLayout L = i8a_layout(n_ucol, n_cg, n_frames);
  double* gmax = reinterpret_cast<double*>(ws + L.colmax);  // [sample groups][padded columns]
  int32_t* exps = reinterpret_cast<int32_t*>(ws + L.exps);
  double* scales = reinterpret_cast<double*>(ws + L.scales);
  double* pow2 = reinterpret_cast<double*>(ws + L.pow2);
  unsigned long long* rowmax = reinterpret_cast<unsigned long long*>(ws + L.rowmax);
  double* fpow2 = reinterpret_cast<double*>(ws + L.fpow2);
  int32_t* count = reinterpret_cast<int32_t*>(ws + L.count);
  int32_t* leftover = reinterpret_cast<int32_t*>(ws + L.leftover);
  int32_t* flags = reinterpret_cast<int32_t*>(ws + L.flags);
  unsigned char* wdig = reinterpret_cast<unsigned char*>(ws + L.wdig);
  unsigned char* xdig = reinterpret_cast<unsigned char*>(ws + L.xdig);
  const int n_xb = L.n_kpad / 16;
  AGF_CUDA_TRY(cudaMemsetAsync(ws, 0, L.leftover, s));
  {
    const I8tSamplePlan sp = i8t_sample_plan(n_frames);
    i8t_sample_kernel<<<dim3((n_ucol + 127) / 128, sp.groups), 128, 0, s>>>(x, n_frames, sp.stride, n_sites, ucol_ptr, ucol_sites,
                                                                          n_ucol, L.n_kpad, gmax);
    AGF_CUDA_TRY(cudaGetLastError());
    i8t_scale_kernel<<<(L.n_kpad + 7) / 8, 256, 0, s>>>(gmax, sp.groups, n_ucol, L.n_kpad, L.n_kpad, exps, scales, pow2);
    AGF_CUDA_TRY(cudaGetLastError());
    i8a_rowmax_kernel<<<dim3((n_cg + 127) / 128, 32), 128, 0, s>>>(umat_t, n_ucol, n_cg, exps, rowmax);
    AGF_CUDA_TRY(cudaGetLastError());
    i8a_wdigits_kernel<<<dim3(L.n_mb, n_xb), 128, 0, s>>>(umat_t, n_ucol, n_cg, L.n_ks, exps, rowmax, fpow2, wdig);
    AGF_CUDA_TRY(cudaGetLastError());
  }
  const size_t gemm_smem = (size_t)kA_Stages * kA_StageBytes + 128;
  void (*digits_kernel)(I8tDigitsParams) = i8t_digits_kernel<kApplyLayout>;  // 3 CTAs per SM (2 and 4 measured the same)
  AGF_CUDA_TRY(cudaFuncSetAttribute(digits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kT_TileBytes));
  AGF_CUDA_TRY(cudaFuncSetAttribute(i8a_gemm_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
  const int sms = sm_count();
  int digit_ctas_per_sm = 1;
  AGF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&digit_ctas_per_sm, digits_kernel, 256,
                                                             2 * kT_TileBytes));
  if (digit_ctas_per_sm < 1) digit_ctas_per_sm = 1;
  for (int64_t f0 = 0; f0 < n_frames; f0 += L.slab) {
    I8tDigitsParams d;
    memset(&d, 0, sizeof(d));
    d.n_frames = n_frames - f0 < L.slab ? n_frames - f0 : L.slab;
    d.forces = x + f0 * (int64_t)n_sites * 3;
    d.n_sites = n_sites;
    d.n_red = n_ucol;
    d.n_xb = n_xb;
    const int n_fb = (int)((d.n_frames + kT_ChunkFrames - 1) / kT_ChunkFrames);
    d.n_groups = n_fb * (kT_ChunkFrames / kT_ItemFrames);
    d.col_ptr = ucol_ptr;
    d.col_sites = ucol_sites;
    d.scales = scales;
    d.digits = xdig;
    d.flags = flags;
    const int n_flags = n_fb * kT_ChunkFrames;
    AGF_CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)n_flags * 4, s));
    const int64_t n_items = (int64_t)(n_xb * 16 / kT_PanelCols) * d.n_groups;
    const int64_t want = (int64_t)sms * digit_ctas_per_sm;
    digits_kernel<<<(int)(n_items < want ? n_items : want), 256, 2 * kT_TileBytes, s>>>(d);
    AGF_CUDA_TRY(cudaGetLastError());
    i8t_scrub_kernel<kApplyLayout><<<(n_flags + 255) / 256 < sms ? (n_flags + 255) / 256 : sms, 256, 0, s>>>(
        flags, n_flags, d.n_frames, f0, n_xb, xdig, count, leftover);
    AGF_CUDA_TRY(cudaGetLastError());
    I8aGemmParams g;
    memset(&g, 0, sizeof(g));
    g.wdig = wdig;
    g.xdig = xdig;
    g.n_ks = L.n_ks;
    g.n_mb = L.n_mb;
    g.n_rb = n_fb;
    g.n_cg = n_cg;
    g.n_frames = d.n_frames;
    g.fpow2 = fpow2;
    g.out = out + f0 * (int64_t)n_cg * 3;
    g.sumsq = sumsq;
    const int64_t units = (int64_t)n_fb * L.n_mb;
    i8a_gemm_kernel<TO><<<(int)(units < sms ? units : sms), kA_Threads, gemm_smem, s>>>(g);
    AGF_CUDA_TRY(cudaGetLastError());
  }
  i8a_leftover_kernel<TO><<<sms, 256, 0, s>>>(x, n_sites, ucol_ptr, ucol_sites, n_ucol, umat_t, n_cg, count, leftover, out, sumsq,
                                               nan_mode, nan_atol, nan_flags);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}

}  // namespace agf

extern "C" size_t agf_map_apply_i8_workspace_bytes(int32_t n_sites, int32_t n_ucol, int32_t n_cg, int64_t n_frames) {
  using namespace agf;
  if (n_sites < 1 || n_ucol < 1 || n_ucol > kA_MaxUcol || n_cg < 1 || n_frames < 1 || n_frames >= ((int64_t)1 << 31)) return 0;
  return i8a_layout(n_ucol, n_cg, n_frames).total;
}

extern "C" int agf_map_apply_i8(const void* points, int in_dtype, int64_t n_frames, int32_t n_sites, const int32_t* ucol_ptr,
                                const int32_t* ucol_sites, int32_t n_ucol, const double* umat_t, int32_t n_cg, void* out,
                                int out_dtype, double* sumsq, int nan_mode, double nan_atol, int32_t* nan_flags,
                                void* workspace, size_t workspace_bytes, void* stream) {
  using namespace agf;
  AGF_REQUIRE(points && ucol_ptr && ucol_sites && umat_t && out && workspace, "agf_map_apply_i8: null pointer");
  AGF_REQUIRE(in_dtype == AGF_F32, "agf_map_apply_i8: float32 input only (float64 input takes agf_map_apply_ws)");
  AGF_REQUIRE(out_dtype == AGF_F32 || out_dtype == AGF_F64, "agf_map_apply_i8: bad output dtype");
  AGF_REQUIRE(nan_mode == 0 || nan_flags != nullptr, "agf_map_apply_i8: nan_mode 1 needs nan_flags");
  const size_t need = agf_map_apply_i8_workspace_bytes(n_sites, n_ucol, n_cg, n_frames);
  AGF_REQUIRE(need != 0, "agf_map_apply_i8: needs 1 <= n_ucol <= %d, n_cg >= 1, 1 <= n_frames < 2^31", kA_MaxUcol);
  AGF_REQUIRE(workspace_bytes >= need, "agf_map_apply_i8: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  AGF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) % 16) == 0, "agf_map_apply_i8: workspace must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float* x = reinterpret_cast<const float*>(points);
  char* ws = reinterpret_cast<char*>(workspace);
  if (out_dtype == AGF_F64)
    return i8a_run<double>(x, n_frames, n_sites, ucol_ptr, ucol_sites, n_ucol, umat_t, n_cg, reinterpret_cast<double*>(out), sumsq,
                           nan_mode, nan_atol, nan_flags, ws, s);
  return i8a_run<float>(x, n_frames, n_sites, ucol_ptr, ucol_sites, n_ucol, umat_t, n_cg, reinterpret_cast<float*>(out), sumsq,
                        nan_mode, nan_atol, nan_flags, ws, s);
}
