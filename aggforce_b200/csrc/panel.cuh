// Packed-panel FP64 tensor-core GEMM core shared by the large Gram (gram.cu), the featurised
// Gram (featgram.cu) and the large dense map application (apply.cu).
//
// A "panel" is the shared-memory image of a DMMA operand slab: 24 k-rows x 128 columns of f64
// with a row stride of 132 doubles (conflict-free fragment loads), 25 344 contiguous bytes.  Pack
// kernels write panels to an HBM workspace ONCE (group sums, f32->f64 promotion, feature
// evaluation); the GEMM CTAs then do nothing but
//     producer warp : 1-D TMA bulk copies of an (A, B) panel pair per k-chunk into a 4-stage ring
//                     (full / empty mbarriers),
//     16 MMA warps  : C[128 x 128] += A^T B  with DMMA.8x8x4, each warp owning one tile row x 16 tile
//                     columns, accumulators in registers for the CTA's whole k range.  Four warps
//                     per scheduler sub-partition: with two, the fixed issue latencies around every
//                     DMMA left the FP64 tensor pipe idle 29 % of the time (ncu, r01_syrk_cfg3).
// Each sub-partition owns tile rows {s, 7-s, 8+s, 15-s}: 34 tiles of a DIAGONAL block's upper
// triangle each, so diagonal blocks of a SYRK cost 17/32 of a full block instead of computing the
// mirrored half; narrow last blocks are handled by the same per-row column masks.
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
constexpr int kPanelMmaWarps = 16;
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

// Tile row owned by MMA warp w.  The four warps of a scheduler sub-partition (w & 3) own rows
// {s, 7-s, 8+s, 15-s}: 34 tiles of a diagonal block's upper triangle for every sub-partition.
__device__ __forceinline__ int panel_warp_row(int w) {
  const int s = w & 3, k = w >> 2;
  return k == 0 ? s : (k == 1 ? 7 - s : (k == 2 ? 8 + s : 15 - s));
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// One stage for a warp whose active column tiles are [LO, 16): software pipeline over the
// 6 x (16 - LO) (k-step, column tile) fragments.  B fragments live in a ring of kRing registers and
// are fetched kRing-2 DMMAs ahead of their use; the slot refilled before DMMA f was last read by
// DMMA f-2 (an LDS whose destination is still a source of a not-yet-dispatched DMMA would stall the
// warp's issue).  The A fragment of the next k-step is fetched half a k-step ahead.
template <int LO>
__device__ __forceinline__ void panel_sweep(uint32_t sa, uint32_t sb, double (&acc)[16][2]) {
  constexpr int W = 16 - LO, kSteps = kPanelRows / 4, kFrags = kSteps * W;
  constexpr int kRing = 8, kAhead = kRing - 2;
  double b[kRing], a[2];
#pragma unroll
  for (int f = 0; f < kAhead && f < kFrags; ++f)
    b[f] = lds_f64(sb + (uint32_t)((((f / W) * 4 * kPanelStride) + (LO + f % W) * 8) * 8));
  a[0] = lds_f64(sa);
#pragma unroll
  for (int f = 0; f < kFrags; ++f) {
    const int kk = f / W, cc = LO + f % W;
    if (f + kAhead < kFrags) {
      const int fn = f + kAhead;
      b[fn % kRing] = lds_f64(sb + (uint32_t)((((fn / W) * 4 * kPanelStride) + (LO + fn % W) * 8) * 8));
    }
    if ((f % W) == W / 2 && kk + 1 < kSteps) a[(kk + 1) & 1] = lds_f64(sa + (uint32_t)((kk + 1) * 4 * kPanelStride * 8));
    dmma884(acc[cc][0], acc[cc][1], a[kk & 1], b[f % kRing]);
  }
}

// All kPanelThreads threads call this.  TRI: instantiate the triangular (diagonal block) sweeps.  Returns true for MMA warps (which then hold in acc the
// 16 column tiles of tile row panel_warp_row(warp), masked by m), false for the producer warp.
template <bool TRI>
__device__ __forceinline__ bool panel_mainloop(unsigned char* smem, const PanelStream& st, uint32_t m,
                                               double (&acc)[16][2]) {
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
  for (int c = 0; c < 16; ++c) acc[c][0] = acc[c][1] = 0.0;
  // active columns [lo, 16) -> pipelined sweep specialised on lo; anything else (narrow last
  // blocks) -> generic masked sweep
  int lo = m ? __ffs((int)m) - 1 : 16;
  if (m != (0xffffu & ~((1u << lo) - 1u)) || (!TRI && lo != 0)) lo = -1;
  const int offa = q * kPanelStride + g + panel_warp_row(warp) * 8;
  const int offb = q * kPanelStride + g;
  for (int64_t j = 0; j < st.n; ++j) {
    const int stage = (int)(j % kPanelStages);
    mbar_wait(&full[stage], (uint32_t)((j / kPanelStages) & 1));
    const double* pa = stages + (size_t)stage * 2 * kPanelElems;
    const double* pb = st.same ? pa : pa + kPanelElems;
    const uint32_t sa = smem_u32(pa) + (uint32_t)offa * 8u, sb = smem_u32(pb) + (uint32_t)offb * 8u;
    if (lo == 0) {
      panel_sweep<0>(sa, sb, acc);
    } else if (TRI && lo > 0) {
      switch (lo) {
        case 1: panel_sweep<1>(sa, sb, acc); break;
        case 2: panel_sweep<2>(sa, sb, acc); break;
        case 3: panel_sweep<3>(sa, sb, acc); break;
        case 4: panel_sweep<4>(sa, sb, acc); break;
        case 5: panel_sweep<5>(sa, sb, acc); break;
        case 6: panel_sweep<6>(sa, sb, acc); break;
        case 7: panel_sweep<7>(sa, sb, acc); break;
        case 8: panel_sweep<8>(sa, sb, acc); break;
        case 9: panel_sweep<9>(sa, sb, acc); break;
        case 10: panel_sweep<10>(sa, sb, acc); break;
        case 11: panel_sweep<11>(sa, sb, acc); break;
        case 12: panel_sweep<12>(sa, sb, acc); break;
        case 13: panel_sweep<13>(sa, sb, acc); break;
        case 14: panel_sweep<14>(sa, sb, acc); break;
        case 15: panel_sweep<15>(sa, sb, acc); break;
        default: break;  // lo == 16: no active tile
      }
    } else if (m) {
#pragma unroll 2
      for (int kk = 0; kk < kPanelRows / 4; ++kk) {
        const double a0 = pa[offa + kk * 4 * kPanelStride];
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
          if ((m >> cc) & 1u) {
            const double b = pb[offb + kk * 4 * kPanelStride + cc * 8];
            dmma884(acc[cc][0], acc[cc][1], a0, b);
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
