// Kernel (a): linear force second-moment Gram  P = sum_{t,d} v v^T,  v = group-summed forces.
//
// Replaces src/aggforce/qp/qplinear.py:66-71 of the reference (two dgemm calls on a
// transposed copy of the force array).  Here the constraint-group sum, the f32->f64
// promotion and the SYRK are one kernel:
//   * frames stream global -> shared memory as 1-D TMA bulk copies (frame_pipe.cuh);
//   * each chunk is converted to an f64 panel  panel[k=(t,d)][x=reduced column]  with a
//     bank-conflict-free stride;
//   * 8 warps run DMMA.8x8x4 on the upper block-triangle, every warp owning a balanced,
//     table-driven list of 8x8 output tiles whose accumulators stay in registers for the
//     whole frame range of the CTA (split-K over frames across CTAs);
//   * accumulators are added to the global Gram with f64 RED atomics at the end.
// Reduced columns are tiled in blocks of 128; for n_red <= 128 (cln025: 97) a single CTA
// shape covers the whole matrix, for larger systems CTAs enumerate block pairs (I <= J).
#include "frame_pipe.cuh"

namespace agf {

constexpr int kGramThreads = 256;
constexpr int kGramWarps = 8;
constexpr int kBlockCols = 128;  // reduced columns per block
constexpr int kMaxSlots = 32;

struct WarpPlan {
  uint8_t nslots;
  uint8_t row[kMaxSlots];  // 8-row tile index within the row block
  uint8_t col[kMaxSlots];  // 8-col tile index within the col block
};

struct ShapePlan {
  WarpPlan warp[kGramWarps];
};

// shape ids: 0 diag(full), 1 off(full x full), 2 off(full x last), 3 diag(last)
struct GramParams {
  const void* forces;
  int64_t n_frames;
  int32_t n_sites;
  const int32_t* col_ptr;
  const int32_t* col_sites;
  int32_t n_red;
  int32_t n_blocks;    // ceil(n_red / 128)
  int32_t n_pairs;     // n_blocks (n_blocks + 1) / 2
  int32_t k_splits;    // CTAs per block pair
  int32_t stride;      // panel row stride in doubles (bank-conflict-free, see panel_stride)
  double* gram;
  ChunkSchedule sch;
  ShapePlan plan[4];
};

static void build_plan(ShapePlan& sp, int row_tiles, int col_tiles, bool diag) {
  // tiles in strip order (row major); diag shapes keep only col >= row
  int total = 0;
  for (int r = 0; r < row_tiles; ++r) total += col_tiles - (diag ? r : 0);
  int base = total / kGramWarps, extra = total % kGramWarps;
  int w = 0, filled = 0;
  int quota = base + (w < extra ? 1 : 0);
  for (int i = 0; i < kGramWarps; ++i) sp.warp[i].nslots = 0;
  for (int r = 0; r < row_tiles; ++r) {
    for (int c = (diag ? r : 0); c < col_tiles; ++c) {
      while (w < kGramWarps - 1 && filled >= quota) {
        ++w;
        filled = 0;
        quota = base + (w < extra ? 1 : 0);
      }
      WarpPlan& wp = sp.warp[w];
      wp.row[wp.nslots] = (uint8_t)r;
      wp.col[wp.nslots] = (uint8_t)c;
      ++wp.nslots;
      ++filled;
    }
  }
}

template <typename T>
__device__ __forceinline__ void fill_panel_staged(const T* __restrict__ raw, int nf, int kf, int n_sites,
                                                  const int32_t* __restrict__ col_ptr,
                                                  const int32_t* __restrict__ col_sites, int col0, int n_red,
                                                  int width_pad, double* __restrict__ panel, int stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < kf; t += kGramWarps) {
    const T* fr = raw + (int64_t)t * n_sites * 3;
    for (int x = lane; x < width_pad; x += 32) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      int col = col0 + x;
      if (t < nf && col < n_red) {
        int b = __ldg(col_ptr + col), e = __ldg(col_ptr + col + 1);
        for (int m = b; m < e; ++m) {
          const T* p = fr + 3 * __ldg(col_sites + m);
          s0 += to_f64(p[0]);
          s1 += to_f64(p[1]);
          s2 += to_f64(p[2]);
        }
      }
      double* dst = panel + (t * 3) * stride + x;
      dst[0] = s0;
      dst[stride] = s1;
      dst[2 * stride] = s2;
    }
  }
}

template <typename T>
__device__ __forceinline__ void fill_panel_global(const T* __restrict__ forces, int64_t t0, int nf, int kf,
                                                  int n_sites, const int32_t* __restrict__ col_ptr,
                                                  const int32_t* __restrict__ col_sites, int col0, int n_red,
                                                  int width_pad, double* __restrict__ panel, int stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < kf; t += kGramWarps) {
    const T* fr = forces + (t0 + t) * (int64_t)n_sites * 3;
    for (int x = lane; x < width_pad; x += 32) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      int col = col0 + x;
      if (t < nf && col < n_red) {
        int b = __ldg(col_ptr + col), e = __ldg(col_ptr + col + 1);
        for (int m = b; m < e; ++m) {
          const T* p = fr + 3 * __ldg(col_sites + m);
          s0 += to_f64(__ldg(p));
          s1 += to_f64(__ldg(p + 1));
          s2 += to_f64(__ldg(p + 2));
        }
      }
      double* dst = panel + (t * 3) * stride + x;
      dst[0] = s0;
      dst[stride] = s1;
      dst[2 * stride] = s2;
    }
  }
}

// One k-sweep of a chunk: KROWS = 3*KF panel rows.
template <int SLOTS, int KROWS>
__device__ __forceinline__ void mma_sweep(const double* __restrict__ pa, const double* __restrict__ pb, int stride,
                                          const uint32_t (&desc)[SLOTS], int nslots, double (&acc)[SLOTS][2]) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const double* a_base = pa + q * stride + g;
  const double* b_base = pb + q * stride + g;
#pragma unroll 2
  for (int kk = 0; kk < KROWS / 4; ++kk) {
    const double* ak = a_base + kk * 4 * stride;
    const double* bk = b_base + kk * 4 * stride;
    double a = 0.0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if (s < nslots) {
        uint32_t d = desc[s];
        if (d & 0x80000000u) a = ak[d & 0xffu];
        double b = bk[(d >> 8) & 0xffu];
        dmma884(acc[s][0], acc[s][1], a, b);
      }
    }
  }
}

template <typename T, int KF, int SLOTS, bool STAGED>
__global__ void __launch_bounds__(kGramThreads, STAGED ? 2 : 1) gram_kernel(const __grid_constant__ GramParams p) {
  constexpr int KROWS = 3 * KF;
  constexpr int STAGES = 2;
  extern __shared__ __align__(128) unsigned char smem[];

  // ---- which block pair / k-split am I
  const int pair = blockIdx.x % p.n_pairs;
  const int ksplit = blockIdx.x / p.n_pairs;
  int bi = 0, bj = 0;
  {
    int rem = pair;
    int rowlen = p.n_blocks;
    while (rem >= rowlen) {
      rem -= rowlen;
      --rowlen;
      ++bi;
    }
    bj = bi + rem;
  }
  const bool diag = (bi == bj);
  const int last = p.n_blocks - 1;
  const int shape = diag ? (bi == last ? 3 : 0) : (bj == last ? 2 : 1);
  const int col0_i = bi * kBlockCols, col0_j = bj * kBlockCols;
  const int width_i = min(kBlockCols, p.n_red - col0_i), width_j = min(kBlockCols, p.n_red - col0_j);
  const int wpad_i = (width_i + 7) & ~7, wpad_j = (width_j + 7) & ~7;
  const int stride = p.stride;

  // ---- smem carve-up
  double* panel_i = reinterpret_cast<double*>(smem);
  double* panel_j = diag ? panel_i : panel_i + KROWS * stride;
  size_t off = (size_t)(STAGED ? 1 : 2) * KROWS * stride * sizeof(double);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + off);
  off += 64;
  T* raw = reinterpret_cast<T*>(smem + off);

  const int warp = threadIdx.x >> 5;
  const WarpPlan& wp = p.plan[shape].warp[warp];
  const int nslots = wp.nslots;
  uint32_t desc[SLOTS];
  {
    int prev_row = -1;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      uint32_t d = 0;
      if (s < nslots) {
        int r = wp.row[s], c = wp.col[s];
        d = (uint32_t)(r * 8) | ((uint32_t)(c * 8) << 8);
        if (r != prev_row) d |= 0x80000000u;
        prev_row = r;
      }
      desc[s] = d;
    }
  }
  double acc[SLOTS][2];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) acc[s][0] = acc[s][1] = 0.0;

  const T* forces = reinterpret_cast<const T*>(p.forces);
  const int64_t n_chunks = p.sch.n_chunks;
  const int64_t first = ksplit, step = p.k_splits;

  if constexpr (STAGED) {
    FrameStager<T, STAGES> st;
    st.init(raw, full, forces, (int64_t)p.n_sites * 3, p.sch);
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s) st.issue(first + (int64_t)s * step, s);
    }
    int stage = 0;
    for (int64_t c = first; c < n_chunks; c += step) {
      const int nf = p.sch.count(c);
      if (nf > 0) {
        st.wait(c, stage);
        fill_panel_staged<T>(st.stage_ptr(stage), nf, KF, p.n_sites, p.col_ptr, p.col_sites, col0_i, p.n_red,
                             wpad_i, panel_i, stride);
      }
      __syncthreads();  // panel complete, raw stage free
      if (threadIdx.x == 0) st.issue(c + (int64_t)STAGES * step, stage);
      if (nf > 0) mma_sweep<SLOTS, KROWS>(panel_i, panel_i, stride, desc, nslots, acc);
      __syncthreads();  // panel free
      stage = (stage + 1) % STAGES;
    }
  } else {
    for (int64_t c = first; c < n_chunks; c += step) {
      const int nf = p.sch.count(c);
      if (nf > 0) {
        const int64_t t0 = p.sch.start(c);
        fill_panel_global<T>(forces, t0, nf, KF, p.n_sites, p.col_ptr, p.col_sites, col0_i, p.n_red, wpad_i,
                             panel_i, stride);
        if (!diag)
          fill_panel_global<T>(forces, t0, nf, KF, p.n_sites, p.col_ptr, p.col_sites, col0_j, p.n_red, wpad_j,
                               panel_j, stride);
      }
      __syncthreads();
      if (nf > 0) mma_sweep<SLOTS, KROWS>(panel_i, panel_j, stride, desc, nslots, acc);
      __syncthreads();
    }
  }

  // ---- epilogue: RED.ADD.F64 into the global Gram (upper block triangle)
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    if (s < nslots) {
      int r = (int)(desc[s] & 0xffu), c = (int)((desc[s] >> 8) & 0xffu);
      int i = col0_i + r + g;
      int j = col0_j + c + 2 * q;
      if (i < p.n_red) {
        double* row = p.gram + (int64_t)i * p.n_red;
        if (j < p.n_red) atomicAdd(row + j, acc[s][0]);
        if (j + 1 < p.n_red) atomicAdd(row + j + 1, acc[s][1]);
      }
    }
  }
}

__global__ void symmetrize_kernel(double* g, int n) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * n) return;
  int i = (int)(idx / n), j = (int)(idx % n);
  if (i > j) g[idx] = g[(int64_t)j * n + i];
}

template <typename T, int KF>
static int launch_gram(GramParams& p, cudaStream_t stream) {
  const bool staged = (p.n_blocks == 1);
  const int stride = staged ? panel_stride((p.n_red + 7) & ~7) : panel_stride(kBlockCols);
  p.stride = stride;
  const int sms = sm_count();
  p.sch = make_schedule(p.forces, p.n_frames, (int64_t)p.n_sites * 3 * sizeof(T), KF);
  if (staged) {
    size_t stage_bytes = ((size_t)KF * p.n_sites * 3 * sizeof(T) + 15) / 16 * 16;
    size_t smem = (size_t)3 * KF * stride * sizeof(double) + 64 + 2 * stage_bytes;
    if (smem <= 113 * 1024) {
      int64_t want = p.sch.n_chunks;
      p.k_splits = (int32_t)(want < 2 * sms ? (want < 1 ? 1 : want) : 2 * sms);
      auto kern = gram_kernel<T, KF, 17, true>;
      AGF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<p.k_splits * p.n_pairs, kGramThreads, smem, stream>>>(p);
      AGF_CUDA_TRY(cudaGetLastError());
      return AGF_OK;
    }
    // frames too wide to stage two chunks: fall through to the gather variant
  }
  {
    p.stride = panel_stride(kBlockCols);
    const int stride = p.stride;
    size_t smem = (size_t)2 * 3 * KF * stride * sizeof(double) + 64;
    int64_t ctas_wanted = (int64_t)sms * 2;
    int64_t ks = (ctas_wanted + p.n_pairs - 1) / p.n_pairs;
    if (ks > p.sch.n_chunks) ks = p.sch.n_chunks;
    if (ks < 1) ks = 1;
    p.k_splits = (int32_t)ks;
    auto kern = gram_kernel<T, KF, 32, false>;
    AGF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<p.k_splits * p.n_pairs, kGramThreads, smem, stream>>>(p);
    AGF_CUDA_TRY(cudaGetLastError());
  }
  return AGF_OK;
}

}  // namespace agf

extern "C" int agf_gram_linear(const void* forces, int dtype, int64_t n_frames, int32_t n_sites,
                               const int32_t* col_ptr, const int32_t* col_sites, int32_t n_red, double* gram,
                               void* stream) {
  using namespace agf;
  AGF_REQUIRE(forces && col_ptr && col_sites && gram, "agf_gram_linear: null pointer");
  AGF_REQUIRE(n_frames >= 0 && n_sites > 0 && n_red > 0, "agf_gram_linear: bad sizes T=%lld n=%d n_red=%d",
              (long long)n_frames, n_sites, n_red);
  AGF_REQUIRE(dtype == AGF_F32 || dtype == AGF_F64, "agf_gram_linear: dtype must be AGF_F32 or AGF_F64");
  if (n_frames == 0) return AGF_OK;
  GramParams p;
  memset(&p, 0, sizeof(p));
  p.forces = forces;
  p.n_frames = n_frames;
  p.n_sites = n_sites;
  p.col_ptr = col_ptr;
  p.col_sites = col_sites;
  p.n_red = n_red;
  p.gram = gram;
  p.n_blocks = (n_red + kBlockCols - 1) / kBlockCols;
  p.n_pairs = p.n_blocks * (p.n_blocks + 1) / 2;
  const int full_tiles = kBlockCols / 8;
  const int last_tiles = ((n_red - (p.n_blocks - 1) * kBlockCols) + 7) / 8;
  build_plan(p.plan[0], full_tiles, full_tiles, true);
  build_plan(p.plan[1], full_tiles, full_tiles, false);
  build_plan(p.plan[2], full_tiles, last_tiles, false);
  build_plan(p.plan[3], last_tiles, last_tiles, true);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == AGF_F32) return launch_gram<float, 16>(p, s);
  return launch_gram<double, 8>(p, s);
}

extern "C" int agf_symmetrize(double* gram, int32_t n, void* stream) {
  using namespace agf;
  AGF_REQUIRE(gram && n > 0, "agf_symmetrize: bad arguments");
  int64_t total = (int64_t)n * n;
  int blocks = (int)((total + 255) / 256);
  symmetrize_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gram, n);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}
