// Gaussian augmentation for joptgauss_map (config 5): one pass that builds the augmented
// coordinate and force arrays of  y = A x + eps,  eps ~ N(0, var I).
//
// Replaces src/aggforce/trajectory/core.py:353-390 (AugmentedTrajectory._augment) together with the
// JAX augmenter it calls (src/aggforce/trajectory/jaxgausstraj.py:232-284: sample + autodiff
// log-gradients), whose closed form is (cf. simplegausstraj.py:108-110)
//     grad_y log g = -(y - A x) / var = -eps / var,      grad_x log g = A^T eps / var
//     coords_aug = [x ; A x + eps]      forces_aug = [F + kbt A^T eps / var ; -kbt eps / var].
// eps = sqrt(var) z with z either injected by the caller (tests: parity needs the same draw) or
// drawn in-kernel from Philox keyed by (seed, global frame, bead, draw id), so slabs of frames and
// ranks reproduce the same noise without communicating.  Arithmetic in f64, stored in the array
// dtype.  HBM bound: 24 n_fg bytes read, 24 (n_fg + n_cg) written per frame (f32).
#include <stdlib.h>

#include "philox.cuh"

namespace agf {

struct AugmentParams {
  const void* coords;
  const void* forces;
  int64_t n_frames;
  int32_t n_sites, n_cg;
  const int32_t* bead_ptr;  // CSR rows of A (bead -> sites)
  const int32_t* bead_sites;
  const double* bead_w;
  const int32_t* site_ptr;  // CSR rows of A^T (site -> beads)
  const int32_t* site_beads;
  const double* site_w;
  double sd, kbt_over_var;
  const void* noise;  // [n_frames, n_cg, 3] standard normals in the array dtype, or null
  uint64_t seed;
  uint32_t draw;
  int64_t frame0;
  void* out_coords;  // [n_frames, n_sites + n_cg, 3] or null
  void* out_forces;
};

template <typename T>
__device__ __forceinline__ void bead_noise(const AugmentParams& p, int64_t t, int c, double (&eps)[3]) {
  if (p.noise) {
    const T* z = reinterpret_cast<const T*>(p.noise) + (t * p.n_cg + c) * 3;
    eps[0] = p.sd * to_f64(__ldg(z));
    eps[1] = p.sd * to_f64(__ldg(z + 1));
    eps[2] = p.sd * to_f64(__ldg(z + 2));
  } else {
    float z[3];
    normal3(p.seed, (uint64_t)(p.frame0 + t), (uint32_t)c, p.draw, z);
    eps[0] = p.sd * (double)z[0];
    eps[1] = p.sd * (double)z[1];
    eps[2] = p.sd * (double)z[2];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) augment_kernel(const __grid_constant__ AugmentParams p) {
  const int n_all = p.n_sites + p.n_cg;
  const int64_t total = p.n_frames * n_all;
  const T* coords = reinterpret_cast<const T*>(p.coords);
  const T* forces = reinterpret_cast<const T*>(p.forces);
  T* oc = reinterpret_cast<T*>(p.out_coords);
  T* of = reinterpret_cast<T*>(p.out_forces);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / n_all;
    const int s = (int)(idx - t * n_all);
    const int64_t in_frame = t * (int64_t)p.n_sites * 3;
    if (s < p.n_sites) {
      if (oc) {
        const T* x = coords + in_frame + 3 * s;
        oc[idx * 3] = __ldg(x);
        oc[idx * 3 + 1] = __ldg(x + 1);
        oc[idx * 3 + 2] = __ldg(x + 2);
      }
      if (of) {
        const T* f = forces + in_frame + 3 * s;
        double v[3] = {to_f64(__ldg(f)), to_f64(__ldg(f + 1)), to_f64(__ldg(f + 2))};
        for (int k = __ldg(p.site_ptr + s); k < __ldg(p.site_ptr + s + 1); ++k) {
          double eps[3];
          bead_noise<T>(p, t, __ldg(p.site_beads + k), eps);
          const double w = p.kbt_over_var * __ldg(p.site_w + k);
          v[0] = fma(w, eps[0], v[0]);
          v[1] = fma(w, eps[1], v[1]);
          v[2] = fma(w, eps[2], v[2]);
        }
        of[idx * 3] = static_cast<T>(v[0]);
        of[idx * 3 + 1] = static_cast<T>(v[1]);
        of[idx * 3 + 2] = static_cast<T>(v[2]);
      }
    } else {
      const int c = s - p.n_sites;
      double eps[3];
      bead_noise<T>(p, t, c, eps);
      if (oc) {
        double m[3] = {0.0, 0.0, 0.0};
        for (int k = __ldg(p.bead_ptr + c); k < __ldg(p.bead_ptr + c + 1); ++k) {
          const T* x = coords + in_frame + 3 * __ldg(p.bead_sites + k);
          const double w = __ldg(p.bead_w + k);
          m[0] = fma(w, to_f64(__ldg(x)), m[0]);
          m[1] = fma(w, to_f64(__ldg(x + 1)), m[1]);
          m[2] = fma(w, to_f64(__ldg(x + 2)), m[2]);
        }
        oc[idx * 3] = static_cast<T>(m[0] + eps[0]);
        oc[idx * 3 + 1] = static_cast<T>(m[1] + eps[1]);
        oc[idx * 3 + 2] = static_cast<T>(m[2] + eps[2]);
      }
      if (of) {
        of[idx * 3] = static_cast<T>(-p.kbt_over_var * eps[0]);
        of[idx * 3 + 1] = static_cast<T>(-p.kbt_over_var * eps[1]);
        of[idx * 3 + 2] = static_cast<T>(-p.kbt_over_var * eps[2]);
      }
    }
  }
}

// Row-wise variant (one CTA per frame, grid-stride): the site part of both arrays is a 16-byte vector copy into
// the wider output rows, the noise of a frame is drawn once into shared memory, and only the sites that carry a
// bead (the rows of A^T with an entry) are revisited for the force correction -- same float64 arithmetic and the
// same rounding as augment_kernel, at copy bandwidth instead of three 4-byte accesses per thread.
template <typename T>
__global__ void __launch_bounds__(256) augment_rows_kernel(const __grid_constant__ AugmentParams p) {
  extern __shared__ __align__(16) unsigned char ar_smem[];
  double* s_eps = reinterpret_cast<double*>(ar_smem);                        // [n_cg][3]
  int32_t* s_touched = reinterpret_cast<int32_t*>(s_eps + (size_t)p.n_cg * 3);  // sites with a row in A^T
  __shared__ int n_touched;
  if (threadIdx.x == 0) n_touched = 0;
  __syncthreads();
  for (int s = threadIdx.x; s < p.n_sites; s += blockDim.x)
    if (__ldg(p.site_ptr + s + 1) > __ldg(p.site_ptr + s)) s_touched[atomicAdd(&n_touched, 1)] = s;
  const int n_all = p.n_sites + p.n_cg;
  const T* coords = reinterpret_cast<const T*>(p.coords);
  const T* forces = reinterpret_cast<const T*>(p.forces);
  T* oc = reinterpret_cast<T*>(p.out_coords);
  T* of = reinterpret_cast<T*>(p.out_forces);
  const int n_vec = (int)((size_t)p.n_sites * 3 * sizeof(T) / 16);
  for (int64_t t = blockIdx.x; t < p.n_frames; t += gridDim.x) {
    __syncthreads();  // previous frame's readers of s_eps are done (and the touched list is complete)
    for (int c = threadIdx.x; c < p.n_cg; c += blockDim.x) {
      double eps[3];
      bead_noise<T>(p, t, c, eps);
      s_eps[3 * c] = eps[0];
      s_eps[3 * c + 1] = eps[1];
      s_eps[3 * c + 2] = eps[2];
    }
    const int64_t in_frame = t * (int64_t)p.n_sites * 3, out_frame = t * (int64_t)n_all * 3;
    if (oc) {
      const uint4* src = reinterpret_cast<const uint4*>(coords + in_frame);
      uint4* dst = reinterpret_cast<uint4*>(oc + out_frame);
      for (int i = threadIdx.x; i < n_vec; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (of) {
      const uint4* src = reinterpret_cast<const uint4*>(forces + in_frame);
      uint4* dst = reinterpret_cast<uint4*>(of + out_frame);
      for (int i = threadIdx.x; i < n_vec; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (of) {
      for (int i = threadIdx.x; i < n_touched; i += blockDim.x) {
        const int s = s_touched[i];
        const T* f = forces + in_frame + 3 * s;
        double v[3] = {to_f64(__ldg(f)), to_f64(__ldg(f + 1)), to_f64(__ldg(f + 2))};
        for (int k = __ldg(p.site_ptr + s); k < __ldg(p.site_ptr + s + 1); ++k) {
          const double* eps = s_eps + 3 * __ldg(p.site_beads + k);
          const double w = p.kbt_over_var * __ldg(p.site_w + k);
          v[0] = fma(w, eps[0], v[0]);
          v[1] = fma(w, eps[1], v[1]);
          v[2] = fma(w, eps[2], v[2]);
        }
        T* o = of + out_frame + 3 * s;
        o[0] = static_cast<T>(v[0]);
        o[1] = static_cast<T>(v[1]);
        o[2] = static_cast<T>(v[2]);
      }
    }
    for (int c = threadIdx.x; c < p.n_cg; c += blockDim.x) {
      const double* eps = s_eps + 3 * c;
      const int64_t o = out_frame + 3 * (int64_t)(p.n_sites + c);
      if (oc) {
        double m[3] = {0.0, 0.0, 0.0};
        for (int k = __ldg(p.bead_ptr + c); k < __ldg(p.bead_ptr + c + 1); ++k) {
          const T* x = coords + in_frame + 3 * __ldg(p.bead_sites + k);
          const double w = __ldg(p.bead_w + k);
          m[0] = fma(w, to_f64(__ldg(x)), m[0]);
          m[1] = fma(w, to_f64(__ldg(x + 1)), m[1]);
          m[2] = fma(w, to_f64(__ldg(x + 2)), m[2]);
        }
        oc[o] = static_cast<T>(m[0] + eps[0]);
        oc[o + 1] = static_cast<T>(m[1] + eps[1]);
        oc[o + 2] = static_cast<T>(m[2] + eps[2]);
      }
      if (of) {
        of[o] = static_cast<T>(-p.kbt_over_var * eps[0]);
        of[o + 1] = static_cast<T>(-p.kbt_over_var * eps[1]);
        of[o + 2] = static_cast<T>(-p.kbt_over_var * eps[2]);
      }
    }
  }
}

}  // namespace agf

extern "C" int agf_gauss_augment(const void* coords, const void* forces, int dtype, int64_t n_frames, int32_t n_sites,
                                 const int32_t* bead_ptr, const int32_t* bead_sites, const double* bead_w, int32_t n_cg,
                                 const int32_t* site_ptr, const int32_t* site_beads, const double* site_w, double var,
                                 double kbt, const void* noise, uint64_t seed, uint32_t draw, int64_t frame0,
                                 void* out_coords, void* out_forces, void* stream) {
  using namespace agf;
  AGF_REQUIRE(bead_ptr && bead_sites && bead_w && site_ptr && site_beads && site_w, "agf_gauss_augment: null map");
  AGF_REQUIRE(n_frames >= 0 && n_sites > 0 && n_cg > 0 && var > 0, "agf_gauss_augment: bad sizes / variance");
  AGF_REQUIRE(dtype == AGF_F32 || dtype == AGF_F64, "agf_gauss_augment: bad dtype");
  AGF_REQUIRE((out_coords == nullptr || coords != nullptr) && (out_forces == nullptr || forces != nullptr),
              "agf_gauss_augment: an output was requested without its input");
  if (n_frames == 0 || (!out_coords && !out_forces)) return AGF_OK;
  AugmentParams p;
  p.coords = coords;
  p.forces = forces;
  p.n_frames = n_frames;
  p.n_sites = n_sites;
  p.n_cg = n_cg;
  p.bead_ptr = bead_ptr;
  p.bead_sites = bead_sites;
  p.bead_w = bead_w;
  p.site_ptr = site_ptr;
  p.site_beads = site_beads;
  p.site_w = site_w;
  p.sd = sqrt(var);
  p.kbt_over_var = kbt / var;
  p.noise = noise;
  p.seed = seed;
  p.draw = draw;
  p.frame0 = frame0;
  p.out_coords = out_coords;
  p.out_forces = out_forces;
  const int64_t total = n_frames * (int64_t)(n_sites + n_cg);
  const int64_t want = (total + 255) / 256;
  const int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // rows of both the input and the output are whole 16-byte vectors: the row-wise kernel applies
  const size_t elem = dtype == AGF_F32 ? 4 : 8;
  const size_t rows_smem = (size_t)n_cg * 3 * sizeof(double) + (size_t)n_sites * sizeof(int32_t);
  auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % 16) == 0; };
  if (((size_t)n_sites * 3 * elem) % 16 == 0 && ((size_t)(n_sites + n_cg) * 3 * elem) % 16 == 0 && aligned(coords) &&
      aligned(forces) && aligned(out_coords) && aligned(out_forces) && rows_smem <= (size_t)200 * 1024 &&
      getenv("AGF_AUGMENT_ROWS_OFF") == nullptr) {
    const int rblocks = (int)(n_frames < (int64_t)sm_count() * 8 ? n_frames : (int64_t)sm_count() * 8);
    if (dtype == AGF_F32) {
      AGF_CUDA_TRY(cudaFuncSetAttribute(augment_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem));
      augment_rows_kernel<float><<<rblocks, 256, rows_smem, s>>>(p);
    } else {
      AGF_CUDA_TRY(cudaFuncSetAttribute(augment_rows_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem));
      augment_rows_kernel<double><<<rblocks, 256, rows_smem, s>>>(p);
    }
    AGF_CUDA_TRY(cudaGetLastError());
    return AGF_OK;
  }
  if (dtype == AGF_F32) augment_kernel<float><<<blocks, 256, 0, s>>>(p);
  else augment_kernel<double><<<blocks, 256, 0, s>>>(p);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}
