// Digit planes for the tiled int8 tensor-core kernels (gram_i8t.cu: Gram, apply_i8.cu: dense map application):
// column scales from a sample of the frames, group sums -> 39-bit fixed point -> five signed 8-bit digit planes
// written to a workspace in the shared-memory layout the tensor core reads, and the scrub pass that takes
// out-of-range / non-finite frames out of the planes again (they are handled in float64 by the caller).
#pragma once
#include "i8.cuh"

namespace agf {

constexpr int kT_Slices = 5;
constexpr int kT_ChunkFrames = 32;                       // one chunk = one MMA k-block
constexpr int kT_XbBytes = (kT_ChunkFrames / 8) * 128;   // x-block of one chunk and plane: 512 B
constexpr int kT_SlabFrames = 16384;    // frames whose digits are resident at a time
constexpr int kT_SampleFrames = 1024;

constexpr int kT_PanelCols = 128;                        // digits kernel: columns per work item
constexpr int kT_ItemFrames = 8;                         // ... and frames (one k-group, one warp each)
constexpr int kT_TileXb = kT_ItemFrames * 16 + 16;       // padded x-block stride in the staging tile (bank spread)
constexpr int kT_TileBytes = 3 * kT_Slices * (kT_PanelCols / 16) * kT_TileXb;  // 17 280, two of them per CTA


// ------------------------------------------------------------------------------------------------
// Column scales from a strided sample of the frames.  Thread = column, block row = one GROUP of sampled frames
// (strided over the whole array); every group leaves its maximum |group sum| per column in gmax[group][column].
// The scale of a column is taken from the UPPER QUARTILE of its positive group maxima (i8t_scale_kernel): a few
// stray huge values raise only their own groups' maxima and cannot coarsen the fixed point of the column (their
// frames go to the float64 pass like any other out-of-range frame), and a column that is active in part of the
// trajectory only is scaled by its active groups.
constexpr int kT_MaxSampleGroups = 64;

static __global__ void __launch_bounds__(128) i8t_sample_kernel(const float* __restrict__ forces, int64_t n_frames, int64_t stride,
                                                                int n_sites, const int32_t* __restrict__ col_ptr,
                                                                const int32_t* __restrict__ col_sites, int n_red, int n_pad,
                                                                double* __restrict__ gmax) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n_red) return;
  const int b = __ldg(col_ptr + x), e = __ldg(col_ptr + x + 1);
  double best = 0.0;
  for (int64_t t = (int64_t)blockIdx.y * stride; t < n_frames; t += (int64_t)gridDim.y * stride) {
    const float* fr = forces + t * (int64_t)n_sites * 3;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int m = b; m < e; ++m) {
      const float* q = fr + 3 * __ldg(col_sites + m);
      s0 += (double)__ldg(q);
      s1 += (double)__ldg(q + 1);
      s2 += (double)__ldg(q + 2);
    }
    const double m = fmax(fabs(s0), fmax(fabs(s1), fabs(s2)));
    if (m < 1.0e300) best = fmax(best, m);
  }
  gmax[(size_t)blockIdx.y * n_pad + x] = best;
}

// Upper quartile of the positive entries of gmax[0 .. n_groups)[x]; 0 when there is none.  Whole warp: lane l
// holds groups l and l + 32, ranks them against all others with shuffles (ties broken by group index).
__device__ __forceinline__ double i8t_group_quartile(const double* __restrict__ gmax, int n_groups, int n_pad, int x) {
  const int lane = threadIdx.x & 31;
  double v[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int g = lane + 32 * h;
    const double m = g < n_groups ? gmax[(size_t)g * n_pad + x] : 0.0;
    v[h] = m > 0.0 ? m : 0.0;  // NaN and negatives count as "no positive maximum"
  }
  const int n = __popc(__ballot_sync(0xffffffffu, v[0] > 0.0)) + __popc(__ballot_sync(0xffffffffu, v[1] > 0.0));
  if (n == 0) return 0.0;
  int rank[2] = {0, 0};  // number of positive entries ordered before mine
#pragma unroll
  for (int oh = 0; oh < 2; ++oh) {
    for (int ol = 0; ol < 32; ++ol) {
      const double o = __shfl_sync(0xffffffffu, v[oh], ol);
      const int og = ol + 32 * oh;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        rank[h] += (o > 0.0 && (o < v[h] || (o == v[h] && og < lane + 32 * h))) ? 1 : 0;
    }
  }
  const int want = (3 * (n - 1) + 2) / 4;
  double pick = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (v[h] > 0.0 && rank[h] == want) pick = v[h];
  // exactly one lane holds the wanted rank
  const uint32_t who = __ballot_sync(0xffffffffu, pick > 0.0);
  return __shfl_sync(0xffffffffu, pick, __ffs(who) - 1);
}

// Group maxima of 8-16 sampled frames sit between 2 and 3 sigma for bell-shaped data: values below 4-8x their
// upper quartile (2^(E-1) with E = ilogb + 4, i.e. beyond 10 sigma) fit the 39-bit fixed point, whose step is then
// about 1e-10 of a typical value.  (One bit less headroom puts the limit at 5-9 sigma: a frame in a thousand of a
// 5 000-atom system would take the float64 pass.)
__device__ __forceinline__ int column_exponent_robust(double q) {
  if (!(q > 0.0) || !(q < 1.0e300)) return -900;  // nothing but zeros (or nothing finite) in the sample
  int e = ilogb(q) + 4;
  return e < -900 ? -900 : (e > 900 ? 900 : e);
}

// n_pad: columns per row of gmax / exps / scales / pow2; a column x is live when x % n_period < n_red (batched
// column sets -- the beads of a featurised fit -- repeat with period n_period; n_period = n_pad otherwise).
static __global__ void __launch_bounds__(256) i8t_scale_kernel(const double* __restrict__ gmax, int n_groups, int n_red, int n_pad,
                                                               int n_period, int32_t* __restrict__ exps,
                                                               double* __restrict__ scales, double* __restrict__ pow2) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // warp = column
  if (x >= n_pad) return;
  const bool live = x % n_period < n_red;
  const int e = live ? column_exponent_robust(i8t_group_quartile(gmax, n_groups, n_pad, x)) : 0;
  if ((threadIdx.x & 31) != 0) return;
  exps[x] = e;
  scales[x] = live ? ldexp(1.0, 39 - e) : 0.0;
  pow2[x] = ldexp(1.0, e - 7);  // G[x][y] = pow2[x] pow2[y] sum_l 2^(-8 l) acc_l[x][y]
}

// Sampling plan: about kT_SampleFrames frames in groups of at least 8.
struct I8tSamplePlan {
  int64_t stride;
  int groups;
};
static inline I8tSamplePlan i8t_sample_plan(int64_t n_frames) {
  I8tSamplePlan sp;
  sp.stride = n_frames > kT_SampleFrames ? n_frames / kT_SampleFrames : 1;
  const int64_t n_sample = (n_frames + sp.stride - 1) / sp.stride;
  int64_t g = n_sample / 8;
  sp.groups = (int)(g < 1 ? 1 : (g > kT_MaxSampleGroups ? kT_MaxSampleGroups : g));
  return sp;
}

// ------------------------------------------------------------------------------------------------
struct I8tDigitsParams {
  const float* forces;   // first frame of the slab
  int64_t n_frames;      // frames in the slab
  int32_t n_sites, n_red, n_xb;
  int32_t n_groups;      // groups of kT_ItemFrames frames, rounded up to whole chunks
  const int32_t* col_ptr;
  const int32_t* col_sites;
  const double* scales;  // [n_pad]
  unsigned char* digits;
  int32_t* flags;        // [n_groups * kT_ItemFrames]: frame holds a value outside the fixed-point range
};

// Work item = (column pass of 128 columns, group of 8 frames); the items are dealt to the CTAs in equal
// contiguous ranges, pass-major, so a CTA keeps its columns' member lists while it walks over frame groups.
// Warp = frame, lane q = column quad 4 q .. 4 q + 3 of the pass: the four digits of a plane form one word.
// The digits go through a double-buffered staging tile and leave as 128-byte rows (one k-group of one x-block).
//
// Two workspace layouts (16 columns x of one frame row are always 16 contiguous bytes, 8 frames one 128-byte core
// matrix):
//   kGramLayout   D[chunk = (frame block of 32, xyz)][plane][x-block of 16][k-group of 8 frames][8][16]
//                 -- the columns are the MN dimension of the Gram's operands (MN-major core matrices);
//   kApplyLayout  D[frame block of 32][k-slab of 32 x][half of 16 x][plane][row group: xyz * 4 + k-group][8][16]
//                 -- the columns are the CONTRACTION dimension of the map application (K-major core matrices),
//                 96 rows (xyz, frame) per block; one (frame block, k-slab) is 15 360 contiguous bytes.
enum I8tLayoutKind { kGramLayout = 0, kApplyLayout = 1 };

template <int LAYOUT>
__device__ __forceinline__ size_t i8t_row_offset(int n_xb, int fb, int d, int s, int gxb, int kg) {
  if (LAYOUT == kGramLayout)
    return (((size_t)(fb * 3 + d) * kT_Slices + s) * n_xb + gxb) * kT_XbBytes + kg * 128;
  // n_xb / 2 k-slabs per frame block
  return (((((size_t)fb * (n_xb >> 1) + (gxb >> 1)) * 2 + (gxb & 1)) * kT_Slices + s) * 12 + d * 4 + kg) * 128;
}

template <int LAYOUT>
__global__ void __launch_bounds__(256, 3) i8t_digits_kernel(const __grid_constant__ I8tDigitsParams p) {
  extern __shared__ __align__(16) unsigned char tiles[];  // 2 x [xyz][plane][x-block of the pass][kT_TileXb]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t frame_elems = (int64_t)p.n_sites * 3;
  const int n_pass = (p.n_xb * 16 + kT_PanelCols - 1) / kT_PanelCols;
  const int64_t n_items = (int64_t)n_pass * p.n_groups;
  const int64_t lo = n_items * blockIdx.x / gridDim.x, hi = n_items * (blockIdx.x + 1) / gridDim.x;
  int cur_pass = -1, buf = 0;
  int cb[4] = {0, 0, 0, 0}, cn[4] = {0, 0, 0, 0}, cs0[4] = {0, 0, 0, 0}, cs1[4] = {0, 0, 0, 0};
  double csc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t item = lo; item < hi; ++item, buf ^= 1) {
    const int pass = (int)(item / p.n_groups), fg = (int)(item - (int64_t)pass * p.n_groups);
    if (pass != cur_pass) {
      cur_pass = pass;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int x = pass * kT_PanelCols + 4 * lane + cc;
        cb[cc] = cn[cc] = cs0[cc] = cs1[cc] = 0;
        csc[cc] = 0.0;
        if (x < p.n_red) {
          cb[cc] = __ldg(p.col_ptr + x);
          cn[cc] = __ldg(p.col_ptr + x + 1) - cb[cc];
          if (cn[cc] > 0) cs0[cc] = 3 * __ldg(p.col_sites + cb[cc]);
          if (cn[cc] > 1) cs1[cc] = 3 * __ldg(p.col_sites + cb[cc] + 1);
          csc[cc] = __ldg(p.scales + x);
        }
      }
    }
    const int64_t gf = (int64_t)fg * kT_ItemFrames + warp;
    const bool live = gf < p.n_frames;
    const float* fr = p.forces + (live ? gf : p.n_frames - 1) * frame_elems;
    // the first two members of the lane's four columns: 24 independent loads in flight at once (member lists
    // walked one after the other would be as many dependent round trips to DRAM)
    float a[4][3], b[4][3];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const bool h0 = cn[cc] > 0, h1 = cn[cc] > 1;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        a[cc][d] = h0 ? __ldg(fr + cs0[cc] + d) : 0.f;
        b[cc][d] = h1 ? __ldg(fr + cs1[cc] + d) : 0.f;
      }
    }
    uint32_t lo_w[3][4], hi_w[3][4];
    uint32_t range = 0;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      double v0 = (double)a[cc][0] + (double)b[cc][0], v1 = (double)a[cc][1] + (double)b[cc][1],
             v2 = (double)a[cc][2] + (double)b[cc][2];
      for (int m = 2; m < cn[cc]; ++m) {
        const float* q = fr + 3 * __ldg(p.col_sites + cb[cc] + m);
        v0 += (double)__ldg(q);
        v1 += (double)__ldg(q + 1);
        v2 += (double)__ldg(q + 2);
      }
      const double sc = live ? csc[cc] : 0.0;
      const double t0 = fma(v0, sc, kI8Magic), t1 = fma(v1, sc, kI8Magic), t2 = fma(v2, sc, kI8Magic);
      lo_w[0][cc] = (uint32_t)__double2loint(t0);
      hi_w[0][cc] = (uint32_t)__double2hiint(t0);
      lo_w[1][cc] = (uint32_t)__double2loint(t1);
      hi_w[1][cc] = (uint32_t)__double2hiint(t1);
      lo_w[2][cc] = (uint32_t)__double2loint(t2);
      hi_w[2][cc] = (uint32_t)__double2hiint(t2);
      range |= (hi_w[0][cc] ^ kI8HiExpect) | (hi_w[1][cc] ^ kI8HiExpect) | (hi_w[2][cc] ^ kI8HiExpect);
    }
    if (__any_sync(0xffffffffu, (range & 0xFFFFFF00u) != 0) && lane == 0) p.flags[gf] = 1;
    const uint32_t tbase = smem_u32(tiles) + (uint32_t)buf * kT_TileBytes;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const uint32_t dst = tbase + (uint32_t)((d * kT_Slices) * (kT_PanelCols / 16) + (lane >> 2)) * kT_TileXb +
                           (uint32_t)(warp * 16 + (lane & 3) * 4);
      constexpr uint32_t ps = (kT_PanelCols / 16) * kT_TileXb;  // plane stride inside the tile
      sts_u32(dst + 0 * ps, gather_bytes(hi_w[d][0], hi_w[d][1], hi_w[d][2], hi_w[d][3], 0) ^ 0x80808080u);
      sts_u32(dst + 1 * ps, gather_bytes(lo_w[d][0], lo_w[d][1], lo_w[d][2], lo_w[d][3], 3) ^ 0x80808080u);
      sts_u32(dst + 2 * ps, gather_bytes(lo_w[d][0], lo_w[d][1], lo_w[d][2], lo_w[d][3], 2) ^ 0x80808080u);
      sts_u32(dst + 3 * ps, gather_bytes(lo_w[d][0], lo_w[d][1], lo_w[d][2], lo_w[d][3], 1) ^ 0x80808080u);
      sts_u32(dst + 4 * ps, gather_bytes(lo_w[d][0], lo_w[d][1], lo_w[d][2], lo_w[d][3], 0) ^ 0x80808080u);
    }
    __syncthreads();  // also: everyone has left the copy-out of the item before the previous one (same buffer)
    // tile -> workspace: rows of 128 bytes (k-group fg % 4 of x-block gxb, chunk (fg / 4, xyz), plane)
    const unsigned char* tile = tiles + (size_t)buf * kT_TileBytes;
    const int fb = fg >> 2, kg = fg & 3;
    for (int idx = threadIdx.x; idx < 3 * kT_Slices * (kT_PanelCols / 16) * kT_ItemFrames; idx += blockDim.x) {
      const int q = idx & (kT_ItemFrames - 1);
      const int xb = (idx >> 3) & (kT_PanelCols / 16 - 1), ds = idx >> 6;
      const int gxb = pass * (kT_PanelCols / 16) + xb;
      if (gxb >= p.n_xb) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(tile + (size_t)(ds * (kT_PanelCols / 16) + xb) * kT_TileXb + q * 16);
      const int d = ds / kT_Slices, s = ds - d * kT_Slices;
      *reinterpret_cast<uint4*>(p.digits + i8t_row_offset<LAYOUT>(p.n_xb, fb, d, s, gxb, kg) + q * 16) = v;
    }
  }
}

// Flagged frames: clear their rows in every plane (the float64 pass adds them), list the ones that exist.
template <int LAYOUT>
__global__ void __launch_bounds__(256) i8t_scrub_kernel(const int32_t* __restrict__ flags, int n_flags, int64_t n_frames,
                                                        int64_t frame0, int n_xb, unsigned char* __restrict__ digits,
                                                        int32_t* __restrict__ leftover_count, int32_t* __restrict__ leftover) {
  const int lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < n_flags; base += n_warps * 32) {
    const int mine = base + lane < n_flags ? flags[base + lane] : 0;
    uint32_t mask = __ballot_sync(0xffffffffu, mine != 0);
    while (mask) {
      const int f = base + __ffs(mask) - 1;
      mask &= mask - 1;
      const int fb = f / kT_ChunkFrames, r = f - fb * kT_ChunkFrames;
      for (int idx = lane; idx < 3 * kT_Slices * n_xb; idx += 32) {
        const int gxb = idx % n_xb, ds = idx / n_xb;
        const int d = ds / kT_Slices, s = ds - d * kT_Slices;
        unsigned char* dst = digits + i8t_row_offset<LAYOUT>(n_xb, fb, d, s, gxb, r >> 3) + (r & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
      if (lane == 0 && f < n_frames) {
        const int slot = atomicAdd(leftover_count, 1);
        leftover[slot] = (int32_t)(frame0 + f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The tiled SYRK over digit planes in kGramLayout (defined in gram_i8t.cu): gram[bm] (+=, upper triangle) for
// n_batch independent column sets of n_red columns each, whose digit buffers lie digits_batch_stride bytes apart.
struct I8tSyrkLaunch {
  const unsigned char* digits;
  int32_t n_chunks;            // chunks (32 frames of one xyz component) in the buffer
  int32_t n_red;
  int32_t slice_chunks;        // chunks per work unit, 0 = default (256; at most 768)
  int32_t n_batch;
  int64_t digits_batch_stride;
  const double* pow2;          // [n_batch][i8t_pad(n_red)]: 2^(E - 7) per column
  double* gram;                // [n_batch][n_red][n_red]
  int64_t gram_batch_stride;   // doubles
};
int i8t_launch_syrk(const I8tSyrkLaunch& l, cudaStream_t s);
int i8t_pad(int n_red);  // columns per buffer row: whole 128-column row blocks and whole 96-column blocks

}  // namespace agf
