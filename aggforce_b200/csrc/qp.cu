// Equality-constrained QP of qp_linear_map for SMALL reduced problems, solved on the device.
//
// Replaces src/aggforce/qp/qplinear.py:76-86 of the reference: P += l2 * C'C, then for every bead
//     min 1/2 x'Px   s.t.  (cmap C) x = e_bead
// through qpsolvers/OSQP, one solve per bead with the same P.  The problem has no inequality and no
// linear term, so the minimiser is  X = P^-1 A' (A P^-1 A')^-1  for all beads at once (SURVEY 8c,
// 8f-1).  For n_red <= 128 the whole thing fits one CTA: P lives in shared memory, one Cholesky
// factorisation, warp-per-right-hand-side triangular solves, a tiny second Cholesky for the Schur
// complement (the forward substitutions ride along the factorisation as passenger rows).  Gram -> solve -> force application then never leaves the device (the benchmarked
// step used to stop at the host here); the host solver stays the fallback (status != 0) and the
// `backend="qpsolvers"` path.
#include "common.cuh"

namespace agf {

constexpr int kQpMaxRed = 128;
constexpr int kQpMaxCg = 32;
constexpr int kQpPerRow = 4;                                       // threads sharing one matrix row
constexpr int kQpThreads = (kQpMaxRed + kQpMaxCg) * kQpPerRow;     // 640

struct QpParams {
  const double* gram;      // [n, n], upper triangle valid
  const double* diag_add;  // [n] or null
  const double* a_mat;     // [m, n]
  int32_t n, m;
  const int32_t* x_index;  // [n]: x_out[c, x_index[p]] = x[c][p]
  double* x_out;           // [m, n]
  const int32_t* u_index;  // [n] or null: u_out[u_index[p], c] = x[c][p] (negative: skip)
  double* u_out;           // [*, m]
  int32_t* status;
};

// Left-looking Cholesky of the leading n x n block of `a` (row stride ld), carried through `rows`
// >= n rows: rows n.. are "passengers" that receive the same column operations, i.e. on return
// a[i][0:n] = (L^-1 a_i)' for i >= n -- the forward substitution of those right-hand sides for free.
// Four adjacent lanes share a row (partial dot products, two shuffles); two block barriers per
// column.  invd[k] = 1 / L[k][k].  Returns false (uniformly) at the first non-positive pivot.
__device__ bool cta_cholesky_rows(double* a, int n, int rows, int ld, double* tmp, double* invd) {
  const int row = threadIdx.x / kQpPerRow, part = threadIdx.x % kQpPerRow;
  for (int k = 0; k < n; ++k) {
    const bool active = row >= k && row < rows;
    const double* ri = a + row * ld;
    double s = 0.0;
    if (active) {
      const double* rk = a + k * ld;
      for (int j = part; j < k; j += kQpPerRow) s = fma(ri[j], rk[j], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);  // all lanes take part: the four lanes of a row are adjacent
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (active && part == 0) {
      const double v = ri[k] - s;
      tmp[row] = v;
      if (row == k) tmp[rows] = (v > 0.0 && v < 1.0e300) ? 1.0 / sqrt(v) : 0.0;
    }
    __syncthreads();
    const double inv = tmp[rows];
    if (inv == 0.0) return false;
    if (part == 0 && row >= k && row < rows) {
      a[row * ld + k] = tmp[row] * inv;  // row == k: v / sqrt(v) = sqrt(v)
      if (row == k) invd[k] = inv;
    }
    __syncthreads();
  }
  return true;
}

__global__ void __launch_bounds__(kQpThreads, 1) qp_small_kernel(const __grid_constant__ QpParams p) {
  extern __shared__ double qsm[];
  const int n = p.n, m = p.m, rows = n + m;
  const int ld = (n + 1) | 1, lds = (m + 1) | 1;  // odd strides: column walks hit distinct banks
  double* M = qsm;                 // [n + m][ld]  P, then L; rows n.. : A, then W = A L^-T
  double* V = M + rows * ld;       // [m][ld]      V = S^-1 W, then X
  double* S = V + m * ld;          // [m][lds]     Schur complement W W', then its factor
  double* tmp = S + m * lds;       // [rows + 1]
  double* invd = tmp + rows + 1;   // [n]
  double* invs = invd + n;         // [m]
  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;

  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx - i * n;
    const double v = j >= i ? __ldg(p.gram + (int64_t)i * n + j) : __ldg(p.gram + (int64_t)j * n + i);
    M[i * ld + j] = (i == j && p.diag_add) ? v + __ldg(p.diag_add + i) : v;
  }
  for (int idx = tid; idx < m * n; idx += nt) {
    const int c = idx / n, i = idx - c * n;
    M[(n + c) * ld + i] = __ldg(p.a_mat + idx);
  }
  __syncthreads();
  if (!cta_cholesky_rows(M, n, rows, ld, tmp, invd)) {
    if (tid == 0) *p.status = 1;
    return;
  }
  const double* W = M + n * ld;  // W[c][j] = (L^-1 a_c)[j]
  for (int idx = tid; idx < m * m; idx += nt) {
    const int c = idx / m, d = idx - c * m;
    double s = 0.0;
    for (int j = 0; j < n; ++j) s = fma(W[c * ld + j], W[d * ld + j], s);
    S[c * lds + d] = s;  // A P^-1 A' = W W'
  }
  __syncthreads();
  if (!cta_cholesky_rows(S, m, m, lds, tmp, invs)) {
    if (tid == 0) *p.status = 2;
    return;
  }
  // V[:, j] = S^-1 W[:, j]: one thread per column j, m x m triangular solves in registers
  for (int j = tid; j < n; j += nt) {
#pragma unroll 1
    for (int c = 0; c < m; ++c) {
      double s = W[c * ld + j];
      for (int d = 0; d < c; ++d) s = fma(-S[c * lds + d], V[d * ld + j], s);
      V[c * ld + j] = s * invs[c];
    }
#pragma unroll 1
    for (int c = m - 1; c >= 0; --c) {
      double s = V[c * ld + j];
      for (int d = c + 1; d < m; ++d) s = fma(-S[d * lds + c], V[d * ld + j], s);
      V[c * ld + j] = s * invs[c];
    }
  }
  __syncthreads();
  // X[c] = L^-T V[c]: backward substitution, one warp per bead
  for (int c = warp; c < m; c += nw) {
    double* b = V + c * ld;
    for (int k = n - 1; k >= 0; --k) {
      const double xk = b[k] * invd[k];
      __syncwarp();
      if (lane == 0) b[k] = xk;
      const double* lk = M + k * ld;
      for (int i = lane; i < k; i += 32) b[i] = fma(-lk[i], xk, b[i]);
      __syncwarp();
    }
  }
  __syncthreads();
  bool bad = false;
  for (int idx = tid; idx < m * n; idx += nt) {
    const int c = idx / n, i = idx - c * n;
    const double x = V[c * ld + i];
    if (!(fabs(x) < 1.0e300)) bad = true;
    p.x_out[(int64_t)c * n + __ldg(p.x_index + i)] = x;
    if (p.u_out) {
      const int u = __ldg(p.u_index + i);
      if (u >= 0) p.u_out[(int64_t)u * m + c] = x;
    }
  }
  if (bad) *p.status = 3;
  // the host solver's acceptance test (qp/solver.py): the equality constraints must hold to 1e-6
  for (int idx = tid; idx < m * m; idx += nt) {
    const int c = idx / m, d = idx - c * m;
    double r = c == d ? -1.0 : 0.0;
    for (int i = 0; i < n; ++i) r = fma(__ldg(p.a_mat + (int64_t)d * n + i), V[c * ld + i], r);
    if (!(fabs(r) <= 1.0e-6)) *p.status = 4;
  }
}

}  // namespace agf

extern "C" int agf_qp_equality_small(const double* gram, int32_t n_red, const double* diag_add, const double* a_mat,
                                     int32_t n_cg, const int32_t* x_index, double* x_out, const int32_t* u_index,
                                     double* u_out, int32_t* status, void* stream) {
  using namespace agf;
  AGF_REQUIRE(gram && a_mat && x_index && x_out && status, "agf_qp_equality_small: null pointer");
  AGF_REQUIRE((u_out == nullptr) == (u_index == nullptr), "agf_qp_equality_small: u_out and u_index go together");
  AGF_REQUIRE(n_red > 0 && n_red <= kQpMaxRed && n_cg > 0 && n_cg <= kQpMaxCg && n_cg <= n_red,
              "agf_qp_equality_small: needs 0 < n_cg <= %d, n_cg <= n_red <= %d (got n_red %d, n_cg %d)", kQpMaxCg,
              kQpMaxRed, n_red, n_cg);
  QpParams p;
  p.gram = gram;
  p.diag_add = diag_add;
  p.a_mat = a_mat;
  p.n = n_red;
  p.m = n_cg;
  p.x_index = x_index;
  p.x_out = x_out;
  p.u_index = u_index;
  p.u_out = u_out;
  p.status = status;
  const size_t ld = (size_t)((n_red + 1) | 1), lds = (size_t)((n_cg + 1) | 1);
  const size_t smem = sizeof(double) * ((size_t)(n_red + 2 * n_cg) * ld + (size_t)n_cg * lds + (size_t)(n_red + n_cg + 1) +
                                        (size_t)n_red + (size_t)n_cg);
  AGF_CUDA_TRY(cudaFuncSetAttribute(qp_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qp_small_kernel<<<1, kQpThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}

extern "C" int agf_qp_equality_small_supported(int32_t n_red, int32_t n_cg) {
  return n_red > 0 && n_red <= agf::kQpMaxRed && n_cg > 0 && n_cg <= agf::kQpMaxCg && n_cg <= n_red;
}
