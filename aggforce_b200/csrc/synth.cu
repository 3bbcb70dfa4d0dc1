// Counter-based synthetic trajectory generator (benchmarks / tests).  Every value is a pure
// function of (seed, global frame index, site, stream), so any frame range produced by any
// rank or chunking is identical -- frames shard across GPUs without communicating.
//
// Model (adapted): heavy atoms jitter around a reference structure with
// N(0, pos_sigma^2); every hydrogen sits at EXACTLY bond_len from its parent in a direction that
// wobbles per frame (so X-H distances are rigid to f32 rounding while H-H and H-other
// distances fluctuate); forces are N(0, force_sigma^2) with F_H += -h_coupling * F_parent so
// the second-moment matrix is not diagonal.
#include "common.cuh"
#include "philox.cuh"

namespace agf {

__global__ void __launch_bounds__(256) synth_kernel(const float* __restrict__ ref_pos, const int32_t* __restrict__ parent,
                                                    const float* __restrict__ bond_len, int n_sites, int64_t frame0,
                                                    int64_t n_frames, uint64_t seed, float pos_sigma,
                                                    float force_sigma, float h_coupling, float* __restrict__ coords,
                                                    float* __restrict__ forces) {
  const int64_t total = n_frames * n_sites;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tl = idx / n_sites;
    const int a = (int)(idx - tl * n_sites);
    const uint64_t frame = (uint64_t)(frame0 + tl);
    const int par = parent[a];
    if (coords) {
      float x, y, zc;
      if (par < 0) {
        float z[3];
        normal3(seed, frame, (uint32_t)a, 0u, z);
        x = ref_pos[3 * a] + pos_sigma * z[0];
        y = ref_pos[3 * a + 1] + pos_sigma * z[1];
        zc = ref_pos[3 * a + 2] + pos_sigma * z[2];
      } else {
        float zp[3], zw[3];
        normal3(seed, frame, (uint32_t)par, 0u, zp);
        normal3(seed, frame, (uint32_t)a, 2u, zw);
        const float px = ref_pos[3 * par] + pos_sigma * zp[0];
        const float py = ref_pos[3 * par + 1] + pos_sigma * zp[1];
        const float pz = ref_pos[3 * par + 2] + pos_sigma * zp[2];
        float bx = ref_pos[3 * a] - ref_pos[3 * par], by = ref_pos[3 * a + 1] - ref_pos[3 * par + 1],
              bz = ref_pos[3 * a + 2] - ref_pos[3 * par + 2];
        const float bn = rsqrtf(bx * bx + by * by + bz * bz + 1e-20f);
        bx = bx * bn + 0.35f * zw[0];
        by = by * bn + 0.35f * zw[1];
        bz = bz * bn + 0.35f * zw[2];
        // normalise in double so the bond length is exact to f32 rounding of the sum below
        const double dn = (double)bond_len[a] / sqrt((double)bx * bx + (double)by * by + (double)bz * bz);
        x = (float)((double)px + dn * bx);
        y = (float)((double)py + dn * by);
        zc = (float)((double)pz + dn * bz);
      }
      float* o = coords + idx * 3;
      o[0] = x;
      o[1] = y;
      o[2] = zc;
    }
    if (forces) {
      float z[3];
      normal3(seed, frame, (uint32_t)a, 1u, z);
      float fx = force_sigma * z[0], fy = force_sigma * z[1], fz = force_sigma * z[2];
      if (par >= 0) {
        float zp[3];
        normal3(seed, frame, (uint32_t)par, 1u, zp);
        fx -= h_coupling * force_sigma * zp[0];
        fy -= h_coupling * force_sigma * zp[1];
        fz -= h_coupling * force_sigma * zp[2];
      }
      float* o = forces + idx * 3;
      o[0] = fx;
      o[1] = fy;
      o[2] = fz;
    }
  }
}

}  // namespace agf

extern "C" int agf_synth_frames(const float* ref_pos, const int32_t* parent, const float* bond_len, int32_t n_sites,
                                int64_t frame0, int64_t n_frames, uint64_t seed, float pos_sigma, float force_sigma,
                                float h_coupling, float* coords, float* forces, void* stream) {
  using namespace agf;
  AGF_REQUIRE(ref_pos && parent && bond_len, "agf_synth_frames: null topology pointer");
  AGF_REQUIRE(n_sites > 0 && n_frames >= 0, "agf_synth_frames: bad sizes");
  if (n_frames == 0 || (!coords && !forces)) return AGF_OK;
  int64_t total = n_frames * n_sites;
  int64_t want = (total + 255) / 256;
  int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
  synth_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ref_pos, parent, bond_len, n_sites, frame0, n_frames, seed, pos_sigma, force_sigma, h_coupling, coords, forces);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}
