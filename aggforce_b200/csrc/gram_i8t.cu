// Kernel (a), Blackwell tensor-core path for LARGE reduced problems (n_red > 97): the linear Gram through
// int8 digit planes on tcgen05, tiled over the Gram (the n_red <= 97 variant in gram_i8.cu keeps the whole
// Gram in one accumulator set and builds its digits in shared memory).
//
// Same result as agf_gram_linear_ws (src/aggforce/qp/qplinear.py:66-71 of the reference: P = (R C)' (R C) in
// float64) for float32 forces.  Two kernels per slab of frames:
//
//   1. i8t_digits_kernel: group sums (float64, exact) -> 39-bit fixed point with a per-column power-of-two
//      scale -> five signed 8-bit digit planes, written to the workspace in the layout the tensor core reads:
//          D[chunk][plane s][x-block][k-group][k-row][16 columns]       one byte per entry
//      chunk = 32 frames of ONE xyz component (any order of the contraction rows gives the same Gram),
//      x-block = 16 reduced columns, k-group = 8 frames: one x-block of a chunk and plane is 512 contiguous
//      bytes = four 8 x 16 core matrices of the canonical MN-major no-swizzle UMMA layout, and any window of
//      x-blocks is one contiguous span -> one 1-D TMA bulk copy per operand plane, no tensor map.
//   2. i8t_syrk_kernel (persistent, one CTA per SM): work unit = (slice of <= p.slice_chunks chunks, tile of
//      128 x 96 Gram elements on or above the diagonal).  Warp 0 streams the unit's operand planes through a
//      6-stage ring (per chunk: 5 x 4 KB for the 128 rows, 5 x 3 KB for the 96 columns), warp 1 issues
//      tcgen05.mma.kind::i8 (M 128, N 96, K 32): the 15 plane products with s + t <= 4, accumulated EXACTLY
//      in int32 in five TMEM accumulators (one per level l = s + t, 480 of 512 columns), warps 2-5 read the
//      accumulators back (tcgen05.ld), recombine the levels in float64,
//          G[x][y] += 2^(E_x + E_y - 14) sum_l 2^(-8 l) acc_l[x][y],
//      and add the upper-triangle elements to the Gram.  Units are ordered slice-major so that the CTAs
//      running at one time read the same ~50 MB of digits: the operands come from L2, not HBM.
// Frames holding a value outside the fixed-point range (or a non-finite one) are left out of the planes and
// added in float64 by i8t_leftover_kernel, so the result does not depend on the scale sample.
#include <stdlib.h>

#include "i8_digits.cuh"

namespace agf {

constexpr int kT_M = 128, kT_N = 96;  // one tile; K = 32 contraction rows per MMA
constexpr int kT_APlane = (kT_M / 16) * kT_XbBytes;      // 4096
constexpr int kT_BPlane = (kT_N / 16) * kT_XbBytes;      // 3072
constexpr int kT_StageBytes = kT_Slices * (kT_APlane + kT_BPlane);  // 35 840
constexpr int kT_Stages = 6;
constexpr int kT_SliceChunks = 256;     // 8 192 contraction rows per unit (int32 headroom allows 768 chunks)
constexpr int kT_SyrkThreads = 192;     // load warp, MMA warp, four epilogue warps
constexpr int kT_MaxRed = 8192;

int i8t_pad(int n_red) {
  const int a = (n_red + kT_M - 1) / kT_M * kT_M, b = (n_red + kT_N - 1) / kT_N * kT_N;
  return a > b ? a : b;
}

// ------------------------------------------------------------------------------------------------
struct I8tSyrkParams {
  const unsigned char* digits;
  int32_t n_chunks;  // chunks of the slab (3 per 32 frames)
  int32_t n_red, n_xb, n_mb, n_nb, n_tiles, n_slices, slice_chunks;
  int32_t n_batch;              // independent Grams over the same frames (beads of a featurised fit), 1 otherwise
  int64_t digits_batch_stride;  // bytes between the digit buffers of consecutive batch members
  int64_t gram_batch_stride;    // doubles between their Grams; pow2 is [n_batch][n_xb * 16]
  const double* pow2;  // [n_pad] 2^(E_x - 7)
  double* gram;
};

// tile index -> (row block of 128, column block of 96); a row block keeps the column blocks that reach its
// first row or beyond (elements with y >= x exist)
__device__ __forceinline__ void i8t_tile(int t, int n_nb, int& mi, int& nj) {
  mi = 0;
  for (;;) {
    const int nj0 = (kT_M * mi) / kT_N;  // first column block with 96 nj + 95 >= 128 mi
    const int cnt = n_nb - nj0;
    if (t < cnt) {
      nj = nj0 + t;
      return;
    }
    t -= cnt;
    ++mi;
  }
}

// unit -> (slice, batch member, tile): slice-major, so that the CTAs running at one time share a slice of the digits
__device__ __forceinline__ void i8t_unit(const I8tSyrkParams& p, int u, int& ks, int& bm, int& mi, int& nj) {
  const int per_slice = p.n_tiles * p.n_batch;
  ks = u / per_slice;
  const int r = u - ks * per_slice;
  bm = r / p.n_tiles;
  i8t_tile(r - bm * p.n_tiles, p.n_nb, mi, nj);
}

__global__ void __launch_bounds__(kT_SyrkThreads, 1) i8t_syrk_kernel(const __grid_constant__ I8tSyrkParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kT_Stages * kT_StageBytes);
  uint64_t* full = bars;                  // [kT_Stages] TMA -> MMA
  uint64_t* empty = full + kT_Stages;     // [kT_Stages] MMA -> TMA (tcgen05.commit)
  uint64_t* acc_full = empty + kT_Stages; // MMA -> epilogue
  uint64_t* acc_empty = acc_full + 1;     // epilogue (4 warps) -> MMA
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kT_Stages; ++i) {
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
  const int n_units = p.n_slices * p.n_tiles * p.n_batch;

  if (warp == 0) {
    // ------------------------------------------------ load warp: one thread, ten bulk copies per chunk
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const size_t plane_bytes = (size_t)p.n_xb * kT_XbBytes;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        int ks, bm, mi, nj;
        i8t_unit(p, u, ks, bm, mi, nj);
        const int c0 = ks * p.slice_chunks;
        const int c1 = c0 + p.slice_chunks < p.n_chunks ? c0 + p.slice_chunks : p.n_chunks;
        const unsigned char* base = p.digits + (size_t)bm * p.digits_batch_stride + (size_t)c0 * kT_Slices * plane_bytes;
        const unsigned char* a_src = base + (size_t)mi * (kT_M / 16) * kT_XbBytes;
        const unsigned char* b_src = base + (size_t)nj * (kT_N / 16) * kT_XbBytes;
        for (int c = c0; c < c1; ++c) {
          mbar_wait(&empty[stage], phase ^ 1u);
          unsigned char* dst = stages + (size_t)stage * kT_StageBytes;
          mbar_expect_tx(&full[stage], kT_StageBytes);
#pragma unroll
          for (int s = 0; s < kT_Slices; ++s) {
            tma_bulk_g2s(dst + s * kT_APlane, a_src + s * plane_bytes, kT_APlane, &full[stage]);
            tma_bulk_g2s(dst + kT_Slices * kT_APlane + s * kT_BPlane, b_src + s * plane_bytes, kT_BPlane, &full[stage]);
          }
          a_src += kT_Slices * plane_bytes;
          b_src += kT_Slices * plane_bytes;
          if (++stage == kT_Stages) {
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
        const int ks = u / (p.n_tiles * p.n_batch);
        const int c0 = ks * p.slice_chunks;
        const int c1 = c0 + p.slice_chunks < p.n_chunks ? c0 + p.slice_chunks : p.n_chunks;
        if (!first_unit) {  // the epilogue warps have drained the accumulators of the previous unit
          mbar_wait(acc_empty, acc_phase);
          acc_phase ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        first_unit = false;
        for (int c = c0; c < c1; ++c) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = sbase + (uint32_t)stage * kT_StageBytes;
          const uint32_t b0 = a0 + kT_Slices * kT_APlane;
          // A_s against the STACK of column planes B_0 .. B_{4-s} (contiguous in the stage, 6 x-blocks each): one
          // MMA of N = 96 (5 - s) writes the accumulators of the levels s .. 4, which are adjacent in TMEM.
          // N <= 256: the three widest are issued as two halves.  8 MMAs instead of 15 per chunk -- the operand
          // reads from shared memory drop from 105 KB to 77 KB (the kernel is bound by them, not by the math).
          const uint32_t fresh = c > c0 ? 1u : 0u;
#pragma unroll
          for (int s = 0; s < kT_Slices; ++s) {
            const uint64_t da = umma_desc_mn_i8(a0 + s * kT_APlane, 128, kT_XbBytes);
            const int width = kT_N * (kT_Slices - s);
            const uint32_t acc = s == 0 ? fresh : 1u;  // s = 0 touches every level first
            if (width > 256) {
              const int half = width / 2;
              umma_i8_issue(tmem_base + (uint32_t)(s * kT_N), da, umma_desc_mn_i8(b0, 128, kT_XbBytes),
                            umma_idesc_i8(kT_M, half), acc);
              umma_i8_issue(tmem_base + (uint32_t)(s * kT_N + half), da,
                            umma_desc_mn_i8(b0 + (half / 16) * kT_XbBytes, 128, kT_XbBytes), umma_idesc_i8(kT_M, half), acc);
            } else {
              umma_i8_issue(tmem_base + (uint32_t)(s * kT_N), da, umma_desc_mn_i8(b0, 128, kT_XbBytes),
                            umma_idesc_i8(kT_M, width), acc);
            }
          }
          umma_commit(&empty[stage]);  // arrives when the tensor core has finished reading the stage
          if (++stage == kT_Stages) {
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
    double* patch = reinterpret_cast<double*>(smem + (size_t)kT_Stages * kT_StageBytes + 128) + quarter * 512;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      int ks, bm, mi, nj;
      i8t_unit(p, u, ks, bm, mi, nj);
      const double* pow2 = p.pow2 + (size_t)bm * p.n_xb * 16;
      double* gram = p.gram + (int64_t)bm * p.gram_batch_stride;
      const double px = __ldg(pow2 + mi * kT_M + quarter * 32 + lane);  // row < n_pad always
      mbar_wait(acc_full, acc_phase);
      acc_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = 0; c0 < kT_N; c0 += 16) {
        // the five levels of 16 columns: loads issued back to back, ONE wait; the levels are combined in integers
        // first (hi = 2^8 acc_0 + acc_1, lo = 2^16 acc_2 + 2^8 acc_3 + acc_4: exact in int64), so that two
        // int64 -> float64 conversions per element replace five -- the accumulators are free again sooner
        uint32_t r[kT_Slices][16];
#pragma unroll
        for (int l = 0; l < kT_Slices; ++l) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(l * kT_N + c0);
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
        if (c0 + 16 >= kT_N) {  // last column group read: the accumulators may be overwritten
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        const int y0 = nj * kT_N + c0;
        if (y0 + 15 < mi * kT_M + quarter * 32 || y0 >= p.n_red) continue;  // warp-uniform: nothing on or above the diagonal
        // A thread holds 16 columns of ONE row: adds straight from here would touch 32 sectors per warp
        // instruction.  Scale, transpose through the warp's 32 x 16 patch of shared memory (xor-swizzled: no
        // bank conflicts either way), then two rows x 16 columns per instruction = 8 sectors.
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) patch[lane * 16 + (i ^ (lane & 15))] = g[i] * px * __ldg(pow2 + y0 + i);
        __syncwarp();
        const int c = lane & 15, y = y0 + c;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int r = 2 * k + (lane >> 4);
          const int xr = mi * kT_M + quarter * 32 + r;
          if (y >= xr && y < p.n_red && xr < p.n_red)
            atomicAdd(gram + (int64_t)xr * p.n_red + y, patch[r * 16 + (c ^ (r & 15))]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// Frames the fixed-point pass declined: exact float64 rank-3 updates of the upper triangle.
__global__ void __launch_bounds__(256) i8t_leftover_kernel(const float* __restrict__ forces, int n_sites,
                                                           const int32_t* __restrict__ col_ptr,
                                                           const int32_t* __restrict__ col_sites, int n_red,
                                                           const int32_t* __restrict__ count,
                                                           const int32_t* __restrict__ frames, double* __restrict__ gram) {
  extern __shared__ double lv[];  // [3][n_red]
  const int n = *count;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const float* fr = forces + (int64_t)frames[i] * (int64_t)n_sites * 3;
    __syncthreads();
    for (int x = threadIdx.x; x < n_red; x += blockDim.x) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int m = col_ptr[x]; m < col_ptr[x + 1]; ++m) {
        const float* q = fr + 3 * col_sites[m];
        s0 += (double)q[0];
        s1 += (double)q[1];
        s2 += (double)q[2];
      }
      lv[x] = s0;
      lv[n_red + x] = s1;
      lv[2 * n_red + x] = s2;
    }
    __syncthreads();
    for (int64_t e = threadIdx.x; e < (int64_t)n_red * n_red; e += blockDim.x) {
      const int x = (int)(e / n_red), y = (int)(e - (int64_t)x * n_red);
      if (y >= x) atomicAdd(gram + e, lv[x] * lv[y] + lv[n_red + x] * lv[n_red + y] + lv[2 * n_red + x] * lv[2 * n_red + y]);
    }
  }
}

int i8t_launch_syrk(const I8tSyrkLaunch& l, cudaStream_t s) {
  // stages, 128 bytes of barriers + TMEM address, four 32 x 16 float64 transposition patches
  const size_t smem = (size_t)kT_Stages * kT_StageBytes + 128 + 4 * 512 * sizeof(double);
  AGF_CUDA_TRY(cudaFuncSetAttribute(i8t_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  I8tSyrkParams q;
  memset(&q, 0, sizeof(q));
  q.digits = l.digits;
  q.n_chunks = l.n_chunks;
  q.n_red = l.n_red;
  q.n_xb = i8t_pad(l.n_red) / 16;
  q.n_mb = (l.n_red + kT_M - 1) / kT_M;
  q.n_nb = (l.n_red + kT_N - 1) / kT_N;
  for (int mi = 0; mi < q.n_mb; ++mi) q.n_tiles += q.n_nb - (kT_M * mi) / kT_N;
  q.n_batch = l.n_batch;
  q.digits_batch_stride = l.digits_batch_stride;
  q.gram_batch_stride = l.gram_batch_stride;
  q.pow2 = l.pow2;
  q.gram = l.gram;
  q.slice_chunks = l.slice_chunks >= 1 && l.slice_chunks <= 768 ? l.slice_chunks : kT_SliceChunks;
  if (const char* e = getenv("AGF_I8T_SLICE_CHUNKS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 768) q.slice_chunks = v;  // 768 chunks = 24 576 rows: the int32 accumulators' limit
  }
  q.n_slices = (q.n_chunks + q.slice_chunks - 1) / q.slice_chunks;
  const int64_t units = (int64_t)q.n_slices * q.n_tiles * q.n_batch;
  if (units == 0) return AGF_OK;
  AGF_REQUIRE(units < ((int64_t)1 << 31), "tiled Gram: too many work units");
  const int sms = sm_count();
  i8t_syrk_kernel<<<(int)(units < sms ? units : sms), kT_SyrkThreads, smem, s>>>(q);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}

struct I8tLayout {
  size_t colmax, exps, scales, pow2, count, leftover, flags, digits, total;
  int64_t slab;
};

static I8tLayout i8t_layout(int n_red, int64_t n_frames) {
  I8tLayout L;
  const size_t n_pad = (size_t)i8t_pad(n_red);
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  L.colmax = 0;
  L.exps = up(n_pad * 8 * kT_MaxSampleGroups);
  L.scales = L.exps + up(n_pad * 4);
  L.pow2 = L.scales + up(n_pad * 8);
  L.count = L.pow2 + up(n_pad * 8);
  L.leftover = L.count + 1024;
  const int64_t rounded = (n_frames + kT_ChunkFrames - 1) / kT_ChunkFrames * kT_ChunkFrames;
  L.slab = rounded < kT_SlabFrames ? rounded : kT_SlabFrames;
  L.flags = L.leftover + up((size_t)n_frames * 4);
  L.digits = L.flags + up((size_t)L.slab * 4);
  L.total = L.digits + (size_t)(L.slab / kT_ChunkFrames) * 3 * kT_Slices * (n_pad / 16) * kT_XbBytes;
  return L;
}

}  // namespace agf

extern "C" size_t agf_gram_linear_i8t_workspace_bytes(int32_t n_sites, int32_t n_red, int64_t n_frames) {
  using namespace agf;
  if (n_sites < 1 || n_red <= 97 || n_red > kT_MaxRed || n_frames < 1 || n_frames >= ((int64_t)1 << 31)) return 0;
  return i8t_layout(n_red, n_frames).total;
}

extern "C" int agf_gram_linear_i8t(const void* forces, int dtype, int64_t n_frames, int32_t n_sites, const int32_t* col_ptr,
                                   const int32_t* col_sites, int32_t n_red, double* gram, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  using namespace agf;
  AGF_REQUIRE(forces && col_ptr && col_sites && gram && workspace, "agf_gram_linear_i8t: null pointer");
  AGF_REQUIRE(dtype == AGF_F32, "agf_gram_linear_i8t: float32 forces only (float64 input takes agf_gram_linear_ws)");
  const size_t need = agf_gram_linear_i8t_workspace_bytes(n_sites, n_red, n_frames);
  AGF_REQUIRE(need != 0, "agf_gram_linear_i8t: needs 97 < n_red <= %d and 1 <= n_frames < 2^31 (got %d, %lld)", kT_MaxRed,
              n_red, (long long)n_frames);
  AGF_REQUIRE(workspace_bytes >= need, "agf_gram_linear_i8t: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  AGF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) % 16) == 0, "agf_gram_linear_i8t: workspace must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const I8tLayout L = i8t_layout(n_red, n_frames);
  char* ws = reinterpret_cast<char*>(workspace);
  double* gmax = reinterpret_cast<double*>(ws + L.colmax);  // [sample groups][padded columns]
  int32_t* exps = reinterpret_cast<int32_t*>(ws + L.exps);
  double* scales = reinterpret_cast<double*>(ws + L.scales);
  double* pow2 = reinterpret_cast<double*>(ws + L.pow2);
  int32_t* count = reinterpret_cast<int32_t*>(ws + L.count);
  int32_t* leftover = reinterpret_cast<int32_t*>(ws + L.leftover);
  int32_t* flags = reinterpret_cast<int32_t*>(ws + L.flags);
  unsigned char* digits = reinterpret_cast<unsigned char*>(ws + L.digits);
  const float* f = reinterpret_cast<const float*>(forces);
  const int n_pad = i8t_pad(n_red), n_xb = n_pad / 16;
  AGF_CUDA_TRY(cudaMemsetAsync(ws, 0, L.leftover, s));
  {
    const I8tSamplePlan sp = i8t_sample_plan(n_frames);
    i8t_sample_kernel<<<dim3((n_red + 127) / 128, sp.groups), 128, 0, s>>>(f, n_frames, sp.stride, n_sites, col_ptr, col_sites,
                                                                          n_red, n_pad, gmax);
    AGF_CUDA_TRY(cudaGetLastError());
    i8t_scale_kernel<<<(n_pad + 7) / 8, 256, 0, s>>>(gmax, sp.groups, n_red, n_pad, n_pad, exps, scales, pow2);
    AGF_CUDA_TRY(cudaGetLastError());
  }
  void (*digits_kernel)(I8tDigitsParams) = i8t_digits_kernel<kGramLayout>;  // 3 CTAs per SM (2 and 4 measured the same)
  AGF_CUDA_TRY(cudaFuncSetAttribute(digits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kT_TileBytes));
  const int sms = sm_count();
  int digit_ctas_per_sm = 1;
  AGF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&digit_ctas_per_sm, digits_kernel, 256, 2 * kT_TileBytes));
  if (digit_ctas_per_sm < 1) digit_ctas_per_sm = 1;
  for (int64_t f0 = 0; f0 < n_frames; f0 += L.slab) {
    I8tDigitsParams d;
    memset(&d, 0, sizeof(d));
    d.n_frames = n_frames - f0 < L.slab ? n_frames - f0 : L.slab;
    d.forces = f + f0 * (int64_t)n_sites * 3;
    d.n_sites = n_sites;
    d.n_red = n_red;
    d.n_xb = n_xb;
    const int n_fb = (int)((d.n_frames + kT_ChunkFrames - 1) / kT_ChunkFrames);
    d.n_groups = n_fb * (kT_ChunkFrames / kT_ItemFrames);
    d.col_ptr = col_ptr;
    d.col_sites = col_sites;
    d.scales = scales;
    d.digits = digits;
    d.flags = flags;
    const int n_flags = n_fb * kT_ChunkFrames;
    AGF_CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)n_flags * 4, s));
    const int64_t n_items = (int64_t)((n_xb * 16 + kT_PanelCols - 1) / kT_PanelCols) * d.n_groups;
    const int64_t want = (int64_t)sms * digit_ctas_per_sm;  // all resident at once: equal item ranges = equal work
    digits_kernel<<<(int)(n_items < want ? n_items : want), 256, 2 * kT_TileBytes, s>>>(d);
    AGF_CUDA_TRY(cudaGetLastError());
    i8t_scrub_kernel<kGramLayout><<<(n_flags + 255) / 256 < sms ? (n_flags + 255) / 256 : sms, 256, 0, s>>>(flags, n_flags, d.n_frames, f0, n_xb,
                                                                                            digits, count, leftover);
    AGF_CUDA_TRY(cudaGetLastError());
    I8tSyrkLaunch q;
    q.digits = digits;
    q.n_chunks = 3 * n_fb;
    q.n_red = n_red;
    q.slice_chunks = 0;
    q.n_batch = 1;
    q.digits_batch_stride = 0;
    q.pow2 = pow2;
    q.gram = gram;
    q.gram_batch_stride = 0;
    const int rc = i8t_launch_syrk(q, s);
    if (rc) return rc;
  }
  const size_t lsmem = (size_t)3 * n_red * sizeof(double);
  AGF_CUDA_TRY(cudaFuncSetAttribute(i8t_leftover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
  i8t_leftover_kernel<<<sms, 256, lsmem, s>>>(f, n_sites, col_ptr, col_sites, n_red, count, leftover, gram);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}
