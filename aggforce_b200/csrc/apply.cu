// Kernel (d): map application  out[t,c,:] = sum_f M[c,f] X[t,f,:], fused with the NaN probe /
// NaN protocol and the residual sum(out^2).
//
// Replaces src/aggforce/util.py:119-124 (einsum), src/aggforce/map/core.py:13-16,219-240
// (NaN probe + protocol) and src/aggforce/agg.py:291-297 (residual) of the reference.
//
// Three paths, chosen by the host from the structure of M:
//   * sparse rows (slice / uniform maps: a few non-zeros per bead): sector-granular gather,
//     only the referenced sites are read from HBM;
//   * dense, small (n_cg <= 64 and the unique-column matrix fits in shared memory): frames
//     stream through a TMA bulk-copy ring fed by a producer warp; consumer warps build DMMA
//     A-fragments (8 frames x 4 unique columns of one xyz component) straight from the raw
//     f32 stage -- the constraint-group sum happens in the fragment build -- and multiply by
//     the register/shared resident map;
//   * dense, large: tiled DFMA fallback with the map read through L2.
// All arithmetic is f64 (the reference promotes f32 points x f64 matrix to f64).
#include "frame_pipe.cuh"

namespace agf {

constexpr double kNanRtol = 1e-5;  // numpy.allclose default used at map/core.py:230

template <typename TO>
__device__ __forceinline__ void store_out(TO* p, double v) {
  *p = static_cast<TO>(v);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------ sparse
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) apply_sparse_kernel(const TI* __restrict__ x, int64_t n_frames, int n_sites,
                                                           const int32_t* __restrict__ row_ptr,
                                                           const int32_t* __restrict__ row_sites,
                                                           const double* __restrict__ row_w, int n_cg,
                                                           TO* __restrict__ out, double* sumsq, int nan_mode,
                                                           double nan_atol, int32_t* nan_flags) {
  const int64_t total = n_frames * n_cg;
  double sq = 0.0;
  bool saw_nan = false, bad = false;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / n_cg;
    const int c = (int)(idx - t * n_cg);
    const TI* fr = x + t * (int64_t)n_sites * 3;
    double a0 = 0, a1 = 0, a2 = 0, w0 = 0, w1 = 0, w2 = 0;
    const int b = __ldg(row_ptr + c), e = __ldg(row_ptr + c + 1);
    for (int m = b; m < e; ++m) {
      const TI* p = fr + 3 * __ldg(row_sites + m);
      const double w = __ldg(row_w + m);
      double v0 = to_f64(__ldg(p)), v1 = to_f64(__ldg(p + 1)), v2 = to_f64(__ldg(p + 2));
      if (nan_mode) {
        if (v0 != v0) { v0 = 0; w0 += w; saw_nan = true; }
        if (v1 != v1) { v1 = 0; w1 += w; saw_nan = true; }
        if (v2 != v2) { v2 = 0; w2 += w; saw_nan = true; }
      }
      a0 = fma(w, v0, a0);
      a1 = fma(w, v1, a1);
      a2 = fma(w, v2, a2);
    }
    if (nan_mode) {
      bad |= fabs(w0) > nan_atol + kNanRtol * fabs(a0 - w0);
      bad |= fabs(w1) > nan_atol + kNanRtol * fabs(a1 - w1);
      bad |= fabs(w2) > nan_atol + kNanRtol * fabs(a2 - w2);
    }
    TO* o = out + idx * 3;
    store_out(o, a0);
    store_out(o + 1, a1);
    store_out(o + 2, a2);
    // the residual is taken on the values as stored (f32 maps store f32)
    double r0 = (double)static_cast<TO>(a0), r1 = (double)static_cast<TO>(a1), r2 = (double)static_cast<TO>(a2);
    sq += r0 * r0 + r1 * r1 + r2 * r2;
  }
  if (sumsq) {
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) atomicAdd(sumsq, sq);
  }
  if (nan_mode) {
    if (saw_nan) atomicOr(nan_flags, 1);
    if (bad) atomicOr(nan_flags + 1, 1);
  }
}

// ------------------------------------------------------------------------------------ dense small
// Work item = 8 consecutive frames ("octet").  A producer warp streams octets through a ring
// of shared-memory stages with 1-D TMA bulk copies; each of the 8 consumer warps owns whole
// octets (stage j -> warp j % 8), so there is no block-wide barrier anywhere.  For its octet a
// warp runs, per 4 unique columns, 3 x NT DMMAs (xyz components x bead tiles): lane (g, q)
// builds the A element "frame g, unique column x0+q" for all three components straight from
// the raw stage (constraint-group sum + f64 promotion in registers), B fragments come from the
// shared-memory copy of the map.  NaNs are detected on the OUTPUT (0 * NaN = NaN poisons the
// whole row) and only then is the octet redone with the masking protocol.
struct DenseSmallParams {
  const void* x;
  int64_t n_frames;
  int32_t n_sites;
  const int32_t* ucol_ptr;
  const int32_t* ucol_sites;
  int32_t n_ucol;
  int32_t nnz;
  const double* umat_t;  // [n_ucol, n_cg]
  int32_t n_cg;
  void* out;
  double* sumsq;
  int32_t nan_mode;
  double nan_atol;
  int32_t* nan_flags;
  ChunkSchedule sch;
  int32_t n_stages;
};

constexpr int kApplyConsumers = 8;
constexpr int kApplyThreads = (kApplyConsumers + 1) * 32;
constexpr int kOctet = 8;
constexpr int kMaxStages = 16;

template <typename TI>
__device__ __forceinline__ void group_value3(const TI* __restrict__ fr, const int32_t* __restrict__ s_ptr,
                                             const int32_t* __restrict__ s_sites, int x, int n_ucol, bool mask_nan,
                                             double (&v)[3], double (&cnt)[3]) {
  v[0] = v[1] = v[2] = 0.0;
  cnt[0] = cnt[1] = cnt[2] = 0.0;
  if (x >= n_ucol) return;
  for (int m = s_ptr[x]; m < s_ptr[x + 1]; ++m) {
    const TI* p = fr + 3 * s_sites[m];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double f = to_f64(p[d]);
      if (mask_nan && f != f) {
        f = 0.0;
        cnt[d] += 1.0;
      }
      v[d] += f;
    }
  }
}

template <typename TI, typename TO, int NT>
__global__ void __launch_bounds__(kApplyThreads, 1) apply_dense_small_kernel(const __grid_constant__ DenseSmallParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int n_cg_pad = NT * 8;
  const int su = panel_stride(n_cg_pad);
  const int xpad = (p.n_ucol + 3) & ~3;
  // carve-up: U^T panel | member table | CSR | barriers | raw ring
  double* s_u = reinterpret_cast<double*>(smem);
  size_t off = (size_t)xpad * su * sizeof(double);
  int4* s_tab = reinterpret_cast<int4*>(smem + off);  // up to 4 member offsets (site*3) per unique column, -1 = none
  off += (size_t)xpad * sizeof(int4);
  int32_t* s_ptr = reinterpret_cast<int32_t*>(smem + off);
  off += (size_t)(p.n_ucol + 1) * 4;
  int32_t* s_sites = reinterpret_cast<int32_t*>(smem + off);
  off += (size_t)p.nnz * 4;
  off = (off + 15) / 16 * 16;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + off);
  uint64_t* empty = full + kMaxStages;
  off += 2 * kMaxStages * sizeof(uint64_t);
  off = (off + 127) / 128 * 128;
  TI* raw = reinterpret_cast<TI*>(smem + off);
  const int64_t frame_elems = (int64_t)p.n_sites * 3;
  const int64_t stage_elems = ((int64_t)kOctet * frame_elems * (int64_t)sizeof(TI) + 15) / 16 * 16 / (int64_t)sizeof(TI);
  const int n_stages = p.n_stages;

  for (int i = threadIdx.x; i < xpad * su; i += blockDim.x) {
    const int xx = i / su, c = i - xx * su;
    s_u[i] = (xx < p.n_ucol && c < p.n_cg) ? p.umat_t[(int64_t)xx * p.n_cg + c] : 0.0;
  }
  for (int i = threadIdx.x; i <= p.n_ucol; i += blockDim.x) s_ptr[i] = p.ucol_ptr[i];
  for (int i = threadIdx.x; i < p.nnz; i += blockDim.x) s_sites[i] = p.ucol_sites[i];
  __shared__ int s_generic;  // some unique column has more than 4 member sites -> CSR walk
  if (threadIdx.x == 0) s_generic = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < xpad; i += blockDim.x) {
    int m[4] = {-1, -1, -1, -1};
    if (i < p.n_ucol) {
      const int b = p.ucol_ptr[i], e = p.ucol_ptr[i + 1];
      for (int k = 0; k < 4 && b + k < e; ++k) m[k] = 3 * p.ucol_sites[b + k];
      if (e - b > 4) s_generic = 1;
    }
    s_tab[i] = make_int4(m[0], m[1], m[2], m[3]);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TI* base = reinterpret_cast<const TI*>(p.x);
  const int64_t first = blockIdx.x, step = gridDim.x;
  const int64_t n_mine = p.sch.n_chunks > first ? (p.sch.n_chunks - first + step - 1) / step : 0;

  if (warp == kApplyConsumers) {
    // ---------------- producer warp
    for (int64_t j = 0; j < n_mine; ++j) {
      const int64_t c = first + j * step;
      const int stage = (int)(j % n_stages);
      const uint32_t parity = (uint32_t)((j / n_stages) & 1);
      mbar_wait(&empty[stage], parity ^ 1u);
      const int nf = p.sch.count(c);
      const int64_t bytes = (int64_t)nf * frame_elems * (int64_t)sizeof(TI);
      const TI* src = base + p.sch.start(c) * frame_elems;
      TI* dst = raw + (int64_t)stage * stage_elems;
      const bool bulk = c != 0 && bytes > 0 && (bytes % 16) == 0 && (reinterpret_cast<uintptr_t>(src) % 16) == 0;
      if (bulk) {
        if (lane == 0) {
          fence_proxy_async();
          mbar_expect_tx(&full[stage], (uint32_t)bytes);
          uint32_t done = 0;
          while (done < (uint32_t)bytes) {
            const uint32_t piece = (uint32_t)bytes - done < kBulkPiece ? (uint32_t)bytes - done : kBulkPiece;
            tma_bulk_g2s(reinterpret_cast<char*>(dst) + done, reinterpret_cast<const char*>(src) + done, piece,
                         &full[stage]);
            done += piece;
          }
        }
      } else {
        const int64_t n = (int64_t)nf * frame_elems;
        for (int64_t i = lane; i < n; i += 32) dst[i] = src[i];
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
      }
      __syncwarp();
    }
    return;
  }

  // ---------------- consumer warps
  const int g = lane >> 2, q = lane & 3;
  const bool nan_mode = p.nan_mode != 0;
  const bool generic_groups = s_generic != 0;
  TO* out = reinterpret_cast<TO*>(p.out);
  double sq = 0.0;
  bool saw_nan = false, bad = false;
  // A waiter can tell the current mbarrier phase from the previous one only, so at most n_stages
  // warps may consume (then the stage a warp waits for is never two phases ahead of its fill).
  const int n_cons = n_stages < kApplyConsumers ? n_stages : kApplyConsumers;
  if (warp >= n_cons) return;
  for (int64_t j = warp; j < n_mine; j += n_cons) {
    const int64_t c = first + j * step;
    const int stage = (int)(j % n_stages);
    const uint32_t parity = (uint32_t)((j / n_stages) & 1);
    mbar_wait(&full[stage], parity);
    const int nf = p.sch.count(c);
    if (nf > 0) {
      const int64_t t0 = p.sch.start(c);
      const bool valid = g < nf;
      const TI* fr = raw + (int64_t)stage * stage_elems + (int64_t)(valid ? g : 0) * frame_elems;
      double acc[3][NT][2];
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[d][n][0] = acc[d][n][1] = 0.0;
      if (!generic_groups) {
#pragma unroll 2
        for (int x0 = 0; x0 < xpad; x0 += 4) {
          const int4 mem = s_tab[x0 + q];
          double a0 = 0.0, a1 = 0.0, a2 = 0.0;
          if (mem.x >= 0) { a0 = to_f64(fr[mem.x]); a1 = to_f64(fr[mem.x + 1]); a2 = to_f64(fr[mem.x + 2]); }
          if (mem.y >= 0) { a0 += to_f64(fr[mem.y]); a1 += to_f64(fr[mem.y + 1]); a2 += to_f64(fr[mem.y + 2]); }
          if (mem.z >= 0) { a0 += to_f64(fr[mem.z]); a1 += to_f64(fr[mem.z + 1]); a2 += to_f64(fr[mem.z + 2]); }
          if (mem.w >= 0) { a0 += to_f64(fr[mem.w]); a1 += to_f64(fr[mem.w + 1]); a2 += to_f64(fr[mem.w + 2]); }
          if (!valid) a0 = a1 = a2 = 0.0;
          const double* bp = s_u + (x0 + q) * su + g;
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const double b = bp[n * 8];
            dmma884(acc[0][n][0], acc[0][n][1], a0, b);
            dmma884(acc[1][n][0], acc[1][n][1], a1, b);
            dmma884(acc[2][n][0], acc[2][n][1], a2, b);
          }
        }
      } else {
        for (int x0 = 0; x0 < xpad; x0 += 4) {
          double v[3], cnt[3];
          group_value3<TI>(fr, s_ptr, s_sites, x0 + q, p.n_ucol, false, v, cnt);
          if (!valid) v[0] = v[1] = v[2] = 0.0;
          const double* bp = s_u + (x0 + q) * su + g;
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const double b = bp[n * 8];
            dmma884(acc[0][n][0], acc[0][n][1], v[0], b);
            dmma884(acc[1][n][0], acc[1][n][1], v[1], b);
            dmma884(acc[2][n][0], acc[2][n][1], v[2], b);
          }
        }
      }
      if (nan_mode) {
        bool has_nan = false;
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
          for (int n = 0; n < NT; ++n) has_nan |= (acc[d][n][0] != acc[d][n][0]) | (acc[d][n][1] != acc[d][n][1]);
        if (__any_sync(0xffffffffu, has_nan)) {
          // rare path: redo the octet with NaN -> 0 and accumulate the weight mass on NaN entries
          double nw[3][NT][2];
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int n = 0; n < NT; ++n) acc[d][n][0] = acc[d][n][1] = nw[d][n][0] = nw[d][n][1] = 0.0;
          bool any_cnt = false;
          for (int x0 = 0; x0 < xpad; x0 += 4) {
            double v[3], cnt[3];
            group_value3<TI>(fr, s_ptr, s_sites, x0 + q, p.n_ucol, true, v, cnt);
            if (!valid) v[0] = v[1] = v[2] = cnt[0] = cnt[1] = cnt[2] = 0.0;
            any_cnt |= (cnt[0] + cnt[1] + cnt[2]) != 0.0;
            const double* bp = s_u + (x0 + q) * su + g;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              const double b = bp[n * 8];
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                dmma884(acc[d][n][0], acc[d][n][1], v[d], b);
                dmma884(nw[d][n][0], nw[d][n][1], cnt[d], b);
              }
            }
          }
          saw_nan |= __any_sync(0xffffffffu, any_cnt);
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              bad |= fabs(nw[d][n][0]) > p.nan_atol + kNanRtol * fabs(acc[d][n][0] - nw[d][n][0]);
              bad |= fabs(nw[d][n][1]) > p.nan_atol + kNanRtol * fabs(acc[d][n][1] - nw[d][n][1]);
            }
        }
      }
      // C[g][2q], C[g][2q+1] of component d: frame t0+g, beads n*8+2q (+1) -> 6 consecutive outputs
      if (valid) {
        TO* orow = out + (t0 + g) * (int64_t)p.n_cg * 3;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const int cb = n * 8 + 2 * q;
          if (cb < p.n_cg) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              store_out(orow + (int64_t)cb * 3 + d, acc[d][n][0]);
              const double r = (double)static_cast<TO>(acc[d][n][0]);
              sq += r * r;
            }
          }
          if (cb + 1 < p.n_cg) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              store_out(orow + (int64_t)(cb + 1) * 3 + d, acc[d][n][1]);
              const double r = (double)static_cast<TO>(acc[d][n][1]);
              sq += r * r;
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }
  if (p.sumsq) {
    sq = warp_sum(sq);
    if (lane == 0) atomicAdd(p.sumsq, sq);
  }
  if (nan_mode) {
    if (saw_nan) atomicOr(p.nan_flags, 1);
    if (bad) atomicOr(p.nan_flags + 1, 1);
  }
}

// ------------------------------------------------------------------------------------ dense large (fallback)
// CTA = 8 frames (24 rows); thread = one bead c (strided); unique columns processed in
// blocks of 64 whose group sums are staged in shared memory as xg[x][row].
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) apply_dense_big_kernel(const TI* __restrict__ x, int64_t n_frames, int n_sites,
                                                              const int32_t* __restrict__ ucol_ptr,
                                                              const int32_t* __restrict__ ucol_sites, int n_ucol,
                                                              const double* __restrict__ umat_t, int n_cg,
                                                              TO* __restrict__ out, double* sumsq, int nan_mode,
                                                              double nan_atol, int32_t* nan_flags) {
  constexpr int FR = 8, ROWS = FR * 3, XB = 64;
  __shared__ double xg[XB][ROWS];
  __shared__ double xn[XB][ROWS];
  __shared__ int s_has_nan;
  double sq = 0.0;
  bool saw_nan = false, bad = false;
  const int64_t n_groups = (n_frames + FR - 1) / FR;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t t0 = grp * FR;
    const int nf = (int)min((int64_t)FR, n_frames - t0);
    for (int cb = 0; cb < n_cg; cb += blockDim.x) {
      const int c = cb + threadIdx.x;
      double acc[ROWS], nw[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] = nw[r] = 0.0;
      bool any_nan_blocks = false;
      for (int xb = 0; xb < n_ucol; xb += XB) {
        __syncthreads();
        if (threadIdx.x == 0) s_has_nan = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < XB * ROWS; i += blockDim.x) {
          const int xx = i / ROWS, r = i - xx * ROWS;
          const int t = r / 3, d = r - t * 3;
          double v = 0.0, nc = 0.0;
          if (xb + xx < n_ucol && t < nf) {
            const TI* fr = x + (t0 + t) * (int64_t)n_sites * 3 + d;
            const int b = __ldg(ucol_ptr + xb + xx), e = __ldg(ucol_ptr + xb + xx + 1);
            for (int m = b; m < e; ++m) {
              double f = to_f64(__ldg(fr + 3 * __ldg(ucol_sites + m)));
              if (nan_mode && f != f) {
                f = 0.0;
                nc += 1.0;
              }
              v += f;
            }
          }
          xg[xx][r] = v;
          xn[xx][r] = nc;
          if (nc != 0.0) s_has_nan = 1;
        }
        __syncthreads();
        const bool has_nan = s_has_nan != 0;
        any_nan_blocks |= has_nan;
        if (c < n_cg) {
          const int lim = min(XB, n_ucol - xb);
          for (int xx = 0; xx < lim; ++xx) {
            const double u = __ldg(umat_t + (int64_t)(xb + xx) * n_cg + c);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] = fma(u, xg[xx][r], acc[r]);
            if (has_nan) {
#pragma unroll
              for (int r = 0; r < ROWS; ++r) nw[r] = fma(u, xn[xx][r], nw[r]);
            }
          }
        }
      }
      saw_nan |= any_nan_blocks;
      if (c < n_cg) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const int t = r / 3, d = r - t * 3;
          if (t < nf) {
            store_out(out + ((t0 + t) * (int64_t)n_cg + c) * 3 + d, acc[r]);
            double v = (double)static_cast<TO>(acc[r]);
            sq += v * v;
            if (any_nan_blocks) bad |= fabs(nw[r]) > nan_atol + kNanRtol * fabs(acc[r] - nw[r]);
          }
        }
      }
    }
  }
  if (sumsq) {
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) atomicAdd(sumsq, sq);
  }
  if (nan_mode) {
    if (saw_nan) atomicOr(nan_flags, 1);
    if (bad) atomicOr(nan_flags + 1, 1);
  }
}

template <typename TI, typename TO, int NT>
static int launch_small(DenseSmallParams& p, cudaStream_t stream) {
  p.sch = make_schedule(p.x, p.n_frames, (int64_t)p.n_sites * 3 * sizeof(TI), kOctet);
  const int su = panel_stride(NT * 8);
  const int xpad = (p.n_ucol + 3) & ~3;
  size_t off = (size_t)xpad * su * sizeof(double) + (size_t)xpad * 16 + (size_t)(p.n_ucol + 1) * 4 + (size_t)p.nnz * 4;
  off = (off + 15) / 16 * 16 + 2 * kMaxStages * sizeof(uint64_t);
  off = (off + 127) / 128 * 128;
  size_t stage_bytes = ((size_t)kOctet * p.n_sites * 3 * sizeof(TI) + 15) / 16 * 16;
  const size_t budget = 224 * 1024;
  if (off + 4 * stage_bytes > budget) return 1;  // does not fit: caller uses the fallback
  int n_stages = (int)((budget - off) / stage_bytes);
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  p.n_stages = n_stages;
  size_t smem = off + (size_t)n_stages * stage_bytes;
  auto kern = apply_dense_small_kernel<TI, TO, NT>;
  AGF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t ctas = sm_count();
  int64_t want = (p.sch.n_chunks + kApplyConsumers - 1) / kApplyConsumers;
  if (ctas > want) ctas = want;
  if (ctas < 1) ctas = 1;
  kern<<<(int)ctas, kApplyThreads, smem, stream>>>(p);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}

template <typename TI, typename TO>
static int dispatch_small(DenseSmallParams& p, cudaStream_t stream) {
  const int nt = (p.n_cg + 7) / 8;
  switch (nt) {
    case 1: return launch_small<TI, TO, 1>(p, stream);
    case 2: return launch_small<TI, TO, 2>(p, stream);
    case 3: return launch_small<TI, TO, 3>(p, stream);
    case 4: return launch_small<TI, TO, 4>(p, stream);
    case 5:
    case 6: return launch_small<TI, TO, 6>(p, stream);
    case 7:
    case 8: return launch_small<TI, TO, 8>(p, stream);
    default: return 1;
  }
}

template <typename TI, typename TO>
static int apply_typed(const void* points, int64_t n_frames, int32_t n_sites, const int32_t* ucol_ptr,
                       const int32_t* ucol_sites, int32_t n_ucol, int32_t nnz, const double* umat_t, int32_t n_cg,
                       void* out, double* sumsq, int nan_mode, double nan_atol, int32_t* nan_flags,
                       cudaStream_t stream) {
  DenseSmallParams p;
  memset(&p, 0, sizeof(p));
  p.x = points;
  p.n_frames = n_frames;
  p.n_sites = n_sites;
  p.ucol_ptr = ucol_ptr;
  p.ucol_sites = ucol_sites;
  p.n_ucol = n_ucol;
  p.nnz = nnz;
  p.umat_t = umat_t;
  p.n_cg = n_cg;
  p.out = out;
  p.sumsq = sumsq;
  p.nan_mode = nan_mode;
  p.nan_atol = nan_atol;
  p.nan_flags = nan_flags;
  int rc = 1;
  if (n_cg <= 64 && nnz >= 0) rc = dispatch_small<TI, TO>(p, stream);
  if (rc <= 0) return rc;
  int64_t groups = (n_frames + 7) / 8;
  int blocks = (int)(groups < (int64_t)sm_count() * 4 ? groups : (int64_t)sm_count() * 4);
  if (blocks < 1) blocks = 1;
  apply_dense_big_kernel<TI, TO><<<blocks, 256, 0, stream>>>(
      reinterpret_cast<const TI*>(points), n_frames, n_sites, ucol_ptr, ucol_sites, n_ucol, umat_t, n_cg,
      reinterpret_cast<TO*>(out), sumsq, nan_mode, nan_atol, nan_flags);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}

}  // namespace agf

extern "C" int agf_map_apply(const void* points, int in_dtype, int64_t n_frames, int32_t n_sites,
                             const int32_t* ucol_ptr, const int32_t* ucol_sites, int32_t n_ucol, int32_t nnz,
                             const double* umat_t, int32_t n_cg, void* out, int out_dtype, double* sumsq,
                             int nan_mode, double nan_atol, int32_t* nan_flags, void* stream) {
  using namespace agf;
  AGF_REQUIRE(points && ucol_ptr && ucol_sites && umat_t && out, "agf_map_apply: null pointer");
  AGF_REQUIRE(n_frames >= 0 && n_sites > 0 && n_ucol > 0 && n_cg > 0 && nnz > 0, "agf_map_apply: bad sizes");
  AGF_REQUIRE(nan_mode == 0 || nan_flags != nullptr, "agf_map_apply: nan_mode 1 needs nan_flags");
  AGF_REQUIRE((in_dtype == AGF_F32 || in_dtype == AGF_F64) && (out_dtype == AGF_F32 || out_dtype == AGF_F64),
              "agf_map_apply: bad dtype");
  if (n_frames == 0) return AGF_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define AGF_APPLY(TI, TO)                                                                                      \
  return apply_typed<TI, TO>(points, n_frames, n_sites, ucol_ptr, ucol_sites, n_ucol, nnz, umat_t, n_cg, out, \
                             sumsq, nan_mode, nan_atol, nan_flags, s)
  if (in_dtype == AGF_F32 && out_dtype == AGF_F64) AGF_APPLY(float, double);
  if (in_dtype == AGF_F32 && out_dtype == AGF_F32) AGF_APPLY(float, float);
  if (in_dtype == AGF_F64 && out_dtype == AGF_F64) AGF_APPLY(double, double);
  AGF_APPLY(double, float);
#undef AGF_APPLY
}

extern "C" int agf_map_apply_sparse(const void* points, int in_dtype, int64_t n_frames, int32_t n_sites,
                                    const int32_t* row_ptr, const int32_t* row_sites, const double* row_weights,
                                    int32_t n_cg, void* out, int out_dtype, double* sumsq, int nan_mode,
                                    double nan_atol, int32_t* nan_flags, void* stream) {
  using namespace agf;
  AGF_REQUIRE(points && row_ptr && row_sites && row_weights && out, "agf_map_apply_sparse: null pointer");
  AGF_REQUIRE(n_frames >= 0 && n_sites > 0 && n_cg > 0, "agf_map_apply_sparse: bad sizes");
  AGF_REQUIRE(nan_mode == 0 || nan_flags != nullptr, "agf_map_apply_sparse: nan_mode 1 needs nan_flags");
  AGF_REQUIRE((in_dtype == AGF_F32 || in_dtype == AGF_F64) && (out_dtype == AGF_F32 || out_dtype == AGF_F64),
              "agf_map_apply_sparse: bad dtype");
  if (n_frames == 0) return AGF_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int64_t total = n_frames * n_cg;
  int64_t want = (total + 255) / 256;
  int blocks = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
#define AGF_SP(TI, TO)                                                                                          \
  apply_sparse_kernel<TI, TO><<<blocks, 256, 0, s>>>(reinterpret_cast<const TI*>(points), n_frames, n_sites,    \
                                                     row_ptr, row_sites, row_weights, n_cg,                     \
                                                     reinterpret_cast<TO*>(out), sumsq, nan_mode, nan_atol,     \
                                                     nan_flags)
  if (in_dtype == AGF_F32 && out_dtype == AGF_F64) AGF_SP(float, double);
  else if (in_dtype == AGF_F32 && out_dtype == AGF_F32) AGF_SP(float, float);
  else if (in_dtype == AGF_F64 && out_dtype == AGF_F64) AGF_SP(double, double);
  else AGF_SP(double, float);
#undef AGF_SP
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}
