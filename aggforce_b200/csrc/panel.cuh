// Packed-panel FP64 tensor-core GEMM core shared by the large Gram (gram.cu), the featurised
// Gram (featgram.cu) and the large dense map application (apply.cu).
//
// A "panel" is the shared-memory image of a DMMA operand slab: 24 k-rows x 128 columns of f64
// with a row stride of 132 doubles (conflict-free fragment loads), 25 344 contiguous bytes.  Pack
// kernels write panels to an HBM workspace ONCE (group sums, f32->f64 promotion, feature
// evaluation); the GEMM CTAs then do nothing but
//     producer warp : 1-D TMA bulk copies of an (A, B) panel pair per k-chunk into a 4-stage ring
//                     (full / empty mbarriers),
//     8 MMA warps   : C[128 x 128] += A^T B  with DMMA.8x8x4, warp w owning tile rows {w, 15-w}
//                     x 16 tile columns, accumulators in registers for the CTA's whole k range.
// Rows {w, 15-w} give every warp 17 tiles of a DIAGONAL block's upper triangle, so diagonal
// blocks of a SYRK cost 17/32 of a full block instead of computing the mirrored half; narrow last
// blocks are handled by the same per-row column masks.
#pragma once
#include "common.cuh"

namespace agf {

constexpr int kPanelCols = 128;
constexpr int kPanelStride = 132;
constexpr int kPanelKF = 8;                   // frames per k-chunk of a Gram panel
constexpr int kPanelRows = 3 * kPanelKF;      // k-rows per panel
constexpr int kPanelElems = kPanelRows * kPanelStride;
constexpr uint32_t kPanelBytes = kPanelElems * sizeof(double);
constexpr int kPanelStages = 4;
constexpr int kPanelMmaWarps = 8;
constexpr int kPanelThreads = (kPanelMmaWarps + 1) * 32;
constexpr size_t kPanelSmem = 128 + (size_t)kPanelStages * 2 * kPanelBytes;

// The sequence of (A, B) panel pairs one CTA contracts over: local k-chunk j uses the panels at
// a + j * a_step and b + j * b_step (in doubles).  `same`: B is A (diagonal SYRK block).
struct PanelStream {
  const double* a;
  int64_t a_step;
  const double* b;
  int64_t b_step;
  int64_t n;
  bool same;
};

// Column mask of tile row r of a block with `ct_rows` row tiles against a block with `ct_cols`
// column tiles; `diag`: only columns >= r (upper triangle).
__device__ __forceinline__ uint32_t panel_row_mask(int r, int ct_rows, int ct_cols, bool diag) {
  if (r >= ct_rows) return 0u;
  uint32_t m = (ct_cols >= 16) ? 0xffffu : ((1u << ct_cols) - 1u);
  if (diag) m &= ~((1u << r) - 1u);
  return m;
}

// All kPanelThreads threads call this.  Returns true for MMA warps (which then own acc for tile
// rows r0 = warp, r1 = 15 - warp), false for the producer warp.
__device__ __forceinline__ bool panel_mainloop(unsigned char* smem, const PanelStream& st, uint32_t m0, uint32_t m1,
                                               double (&acc)[2][16][2]) {
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kPanelStages;
  double* stages = reinterpret_cast<double*>(smem + 128);  // [kPanelStages][2][kPanelElems]
  if (threadIdx.x == 0) {
    for (int i = 0; i < kPanelStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kPanelMmaWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kPanelMmaWarps) {
    if (lane == 0) {
      for (int64_t j = 0; j < st.n; ++j) {
        const int stage = (int)(j % kPanelStages);
        mbar_wait(&empty[stage], (uint32_t)(((j / kPanelStages) & 1) ^ 1));
        char* dst = reinterpret_cast<char*>(stages + (size_t)stage * 2 * kPanelElems);
        const char* src_a = reinterpret_cast<const char*>(st.a + j * st.a_step);
        const char* src_b = reinterpret_cast<const char*>(st.b + j * st.b_step);
        mbar_expect_tx(&full[stage], st.same ? kPanelBytes : 2 * kPanelBytes);
        for (uint32_t off = 0; off < kPanelBytes; off += kBulkPiece) {
          const uint32_t piece = kPanelBytes - off < kBulkPiece ? kPanelBytes - off : kBulkPiece;
          tma_bulk_g2s(dst + off, src_a + off, piece, &full[stage]);
          if (!st.same) tma_bulk_g2s(dst + kPanelBytes + off, src_b + off, piece, &full[stage]);
        }
      }
    }
    return false;
  }
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
  const bool all = (m0 & m1) == 0xffffu;
  const uint32_t many = m0 | m1;
  const int off0 = q * kPanelStride + g + warp * 8, off1 = q * kPanelStride + g + (15 - warp) * 8;
  const int offb = q * kPanelStride + g;
  for (int64_t j = 0; j < st.n; ++j) {
    const int stage = (int)(j % kPanelStages);
    mbar_wait(&full[stage], (uint32_t)((j / kPanelStages) & 1));
    const double* pa = stages + (size_t)stage * 2 * kPanelElems;
    const double* pb = st.same ? pa : pa + kPanelElems;
    if (all) {
#pragma unroll 2
      for (int kk = 0; kk < kPanelRows / 4; ++kk) {
        const double a0 = pa[off0 + kk * 4 * kPanelStride], a1 = pa[off1 + kk * 4 * kPanelStride];
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
          const double b = pb[offb + kk * 4 * kPanelStride + cc * 8];
          dmma884(acc[0][cc][0], acc[0][cc][1], a0, b);
          dmma884(acc[1][cc][0], acc[1][cc][1], a1, b);
        }
      }
    } else if (many) {
#pragma unroll 1
      for (int kk = 0; kk < kPanelRows / 4; ++kk) {
        const double a0 = pa[off0 + kk * 4 * kPanelStride], a1 = pa[off1 + kk * 4 * kPanelStride];
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
          if ((many >> cc) & 1u) {
            const double b = pb[offb + kk * 4 * kPanelStride + cc * 8];
            if ((m0 >> cc) & 1u) dmma884(acc[0][cc][0], acc[0][cc][1], a0, b);
            if ((m1 >> cc) & 1u) dmma884(acc[1][cc][0], acc[1][cc][1], a1, b);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }
  return true;
}

// ---------------------------------------------------------------- host launchers (gram.cu)
// Batched SYRK over packed panels:  gram[b] (+=) sum_chunks P^T P  on the upper block triangle.
// Workspace layout: ws[chunk][batch][block][kPanelElems].
int launch_panel_syrk(const double* ws, int64_t n_chunks, int32_t n, int32_t batch, double* gram,
                      cudaStream_t stream);

inline int panel_blocks(int n) { return (n + kPanelCols - 1) / kPanelCols; }

}  // namespace agf
