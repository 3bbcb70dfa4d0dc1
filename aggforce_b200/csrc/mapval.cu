// Validation projections on random Gaussian force fields.
//
// Replaces src/aggforce/jaxmapval.py:365-401 (sq_gaussian_energies + jacrev) and the per-sample
// loops of random_force_proj / random_residual_shift (jaxmapval.py:227-237, 309-319).  The
// reference draws one Gaussian offset per sample, differentiates
//     E_s(x) = sum_{i,j} exp(-((|x_j - x_i|^2 - offset_s) / width)^2)
// (the FULL squared-distance matrix: both orders of every pair, zero diagonal) with JAX for every
// sample, and reduces  <F, G_s>  or  mean((F - G_s)^2).  Closed form of the force field:
//     G_s[t, i, :] = -4 sum_j g'_s(s_ij) (x_i - x_j),   g'_s(s) = -2 (s - offset_s)/width^2 * exp(-((s - offset_s)/width)^2).
// Here ALL samples are evaluated in one pass over the frames: a CTA stages a frame's coordinates
// and forces in shared memory (f64), every thread owns a fixed subset of samples and keeps
//     ip[s] = sum_{t,i} F . G_s        gg[s] = sum_{t,i} |G_s|^2
// in registers for the CTA's whole frame range; one atomicAdd per (CTA, sample) at the end.
// Compute bound (n_sites^2 f64 exponentials per frame and sample); the mapped arrays are tiny.
#include "common.cuh"

namespace agf {

constexpr int kMvThreads = 256;
constexpr int kMvMaxPerThread = 8;  // samples per thread per launch: <= 2048 samples per launch

struct MapvalParams {
  const void* coords;
  const void* forces;
  int64_t n_frames;
  int32_t n_sites;
  const double* offsets;
  int32_t n_samples;
  double inv_width;
  double* out;  // [n_samples, 2] (+=)
};

template <typename T>
__global__ void __launch_bounds__(kMvThreads) field_moments_kernel(const __grid_constant__ MapvalParams p) {
  extern __shared__ double sm[];
  double* x = sm;                   // [n][3]
  double* f = sm + 3 * p.n_sites;   // [n][3]
  const int n = p.n_sites;
  double off[kMvMaxPerThread], ip[kMvMaxPerThread], gg[kMvMaxPerThread];
  int mine = 0;
  for (int s = threadIdx.x; s < p.n_samples && mine < kMvMaxPerThread; s += kMvThreads) {
    off[mine] = __ldg(p.offsets + s);
    ip[mine] = 0.0;
    gg[mine] = 0.0;
    ++mine;
  }
  const T* coords = reinterpret_cast<const T*>(p.coords);
  const T* forces = reinterpret_cast<const T*>(p.forces);
  for (int64_t t = blockIdx.x; t < p.n_frames; t += gridDim.x) {
    __syncthreads();
    for (int e = threadIdx.x; e < 3 * n; e += kMvThreads) {
      x[e] = to_f64(__ldg(coords + t * 3 * n + e));
      f[e] = to_f64(__ldg(forces + t * 3 * n + e));
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < kMvMaxPerThread; ++m) {
      if (m >= mine) break;
      double a_ip = 0.0, a_gg = 0.0;
      for (int i = 0; i < n; ++i) {
        const double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
        double gx = 0.0, gy = 0.0, gz = 0.0;
        for (int j = 0; j < n; ++j) {
          if (j == i) continue;
          const double dx = xi - x[3 * j], dy = yi - x[3 * j + 1], dz = zi - x[3 * j + 2];
          const double z = (dx * dx + dy * dy + dz * dz - off[m]) * p.inv_width;
          const double w = z * exp(-z * z);  // g' = -2 inv_width * w ; G = -4 g' d = 8 inv_width * w * d
          gx = fma(w, dx, gx);
          gy = fma(w, dy, gy);
          gz = fma(w, dz, gz);
        }
        const double c = 8.0 * p.inv_width;
        gx *= c, gy *= c, gz *= c;
        a_ip += gx * f[3 * i] + gy * f[3 * i + 1] + gz * f[3 * i + 2];
        a_gg += gx * gx + gy * gy + gz * gz;
      }
      ip[m] += a_ip;
      gg[m] += a_gg;
    }
  }
  for (int m = 0; m < mine; ++m) {
    const int s = threadIdx.x + m * kMvThreads;
    atomicAdd(p.out + 2 * s, ip[m]);
    atomicAdd(p.out + 2 * s + 1, gg[m]);
  }
}

// One force field, written out: thread per (frame, site).
template <typename T, typename O>
__global__ void __launch_bounds__(256) field_forces_kernel(const T* __restrict__ coords, int64_t n_frames, int n,
                                                           double offset, double inv_width, O* __restrict__ out) {
  const int64_t total = n_frames * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / n;
    const int i = (int)(idx - t * n);
    const T* fr = coords + t * 3 * n;
    const double xi = to_f64(__ldg(fr + 3 * i)), yi = to_f64(__ldg(fr + 3 * i + 1)), zi = to_f64(__ldg(fr + 3 * i + 2));
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      const double dx = xi - to_f64(__ldg(fr + 3 * j)), dy = yi - to_f64(__ldg(fr + 3 * j + 1)),
                   dz = zi - to_f64(__ldg(fr + 3 * j + 2));
      const double z = (dx * dx + dy * dy + dz * dz - offset) * inv_width;
      const double w = z * exp(-z * z);
      gx = fma(w, dx, gx);
      gy = fma(w, dy, gy);
      gz = fma(w, dz, gz);
    }
    const double c = 8.0 * inv_width;
    out[idx * 3] = static_cast<O>(c * gx);
    out[idx * 3 + 1] = static_cast<O>(c * gy);
    out[idx * 3 + 2] = static_cast<O>(c * gz);
  }
}

}  // namespace agf

extern "C" int agf_gauss_field_moments(const void* coords, const void* forces, int dtype, int64_t n_frames,
                                       int32_t n_sites, const double* offsets, int32_t n_samples, double width,
                                       double* out, void* stream) {
  using namespace agf;
  AGF_REQUIRE(coords && forces && offsets && out, "agf_gauss_field_moments: null pointer");
  AGF_REQUIRE(dtype == AGF_F32 || dtype == AGF_F64, "agf_gauss_field_moments: bad dtype");
  AGF_REQUIRE(n_frames >= 0 && n_sites > 0 && n_samples >= 0 && width > 0, "agf_gauss_field_moments: bad sizes");
  AGF_REQUIRE(n_sites <= 2048, "agf_gauss_field_moments: at most 2048 (mapped) sites, got %d", n_sites);
  if (n_frames == 0 || n_samples == 0) return AGF_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)6 * n_sites * sizeof(double);
  if (smem > 48 * 1024) {
    AGF_CUDA_TRY(cudaFuncSetAttribute(field_moments_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AGF_CUDA_TRY(cudaFuncSetAttribute(field_moments_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int per_launch = kMvThreads * kMvMaxPerThread;
  for (int32_t s0 = 0; s0 < n_samples; s0 += per_launch) {
    MapvalParams p;
    p.coords = coords;
    p.forces = forces;
    p.n_frames = n_frames;
    p.n_sites = n_sites;
    p.offsets = offsets + s0;
    p.n_samples = (n_samples - s0) < per_launch ? (n_samples - s0) : per_launch;
    p.inv_width = 1.0 / width;
    p.out = out + 2 * (int64_t)s0;
    const int64_t cap = (int64_t)sm_count() * 4;
    const int blocks = (int)(n_frames < cap ? n_frames : cap);
    if (dtype == AGF_F32) field_moments_kernel<float><<<blocks, kMvThreads, smem, s>>>(p);
    else field_moments_kernel<double><<<blocks, kMvThreads, smem, s>>>(p);
    AGF_CUDA_TRY(cudaGetLastError());
  }
  return AGF_OK;
}

extern "C" int agf_sq_gaussian_forces(const void* coords, int dtype, int64_t n_frames, int32_t n_sites,
                                      double offset, double width, void* out, int out_dtype, void* stream) {
  using namespace agf;
  AGF_REQUIRE(coords && out, "agf_sq_gaussian_forces: null pointer");
  AGF_REQUIRE((dtype == AGF_F32 || dtype == AGF_F64) && (out_dtype == AGF_F32 || out_dtype == AGF_F64),
              "agf_sq_gaussian_forces: bad dtype");
  AGF_REQUIRE(n_frames >= 0 && n_sites > 0 && width > 0, "agf_sq_gaussian_forces: bad sizes");
  if (n_frames == 0) return AGF_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t want = (n_frames * n_sites + 255) / 256;
  const int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
  const double iw = 1.0 / width;
#define AGF_LAUNCH(T, O)                                                                                  \
  field_forces_kernel<T, O><<<blocks, 256, 0, s>>>(reinterpret_cast<const T*>(coords), n_frames, n_sites, \
                                                   offset, iw, reinterpret_cast<O*>(out))
  if (dtype == AGF_F32 && out_dtype == AGF_F32) AGF_LAUNCH(float, float);
  else if (dtype == AGF_F32) AGF_LAUNCH(float, double);
  else if (out_dtype == AGF_F32) AGF_LAUNCH(double, float);
  else AGF_LAUNCH(double, double);
#undef AGF_LAUNCH
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}
