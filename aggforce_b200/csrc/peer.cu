// One-shot all-reduce / all-gather of SMALL float64 payloads over NVLink peer memory.
//
// The frame-sharded fit exchanges a few small accumulators per call (screening matrix 245 KB, Gram
// 75 KB, moment records and residual sums a few hundred bytes): latency-bound messages for which a
// library collective costs a stream hop plus its own launch.  Here every rank owns one SYMMETRIC
// buffer that all peers have mapped (torch.distributed._symmetric_memory does the rendezvous; the
// pointers arrive as a device array), and ONE kernel on the caller's stream does the whole exchange:
//   1. copy the local contribution into this rank's buffer (slot = call parity),
//   2. signal every peer and wait for every peer's signal (release / acquire at system scope),
//   3. read all ranks' slots through the peer mappings, combine, write the result locally.
// Buffers are double-buffered by call parity: a rank that starts call s+2 has seen every peer's
// signal of call s+1, i.e. every peer has finished reading call s -- no second barrier.  The CTAs of
// the grid are independent: CTA c exchanges element range c of every rank and has its own signal
// row, so no grid-wide synchronisation is needed either.
// Layout of a rank's buffer:  [signals: kPeerMaxCtas x kPeerMaxRanks u32 | pad | slot 0 | slot 1].
#include "common.cuh"

namespace agf {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerMaxCtas = 32;
constexpr int kPeerThreads = 512;
constexpr size_t kPeerSignalBytes = 4096;  // >= kPeerMaxCtas * kPeerMaxRanks * 4, keeps the slots 16-byte aligned

struct PeerParams {
  const uint64_t* ptrs;  // device array [world]: base address of every rank's buffer in THIS process
  int32_t rank, world;
  uint32_t seq;          // call number, identical on all ranks, starts at 1
  int32_t op;            // 0 sum, 1 max, 2 gather (out[r * count + i] = rank r's in[i]), 3 sum for i < n_sum else max
  int64_t n_sum;
  const double* in;
  double* out;
  int64_t count;
  int64_t slot_doubles;  // capacity of one slot
  int32_t* error;        // set to 1 when a peer's signal did not arrive (spin limit)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_peer(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kPeerThreads) peer_exchange_kernel(const __grid_constant__ PeerParams p) {
  __shared__ uint64_t s_base[kPeerMaxRanks];
  __shared__ int s_fail;
  if (threadIdx.x < p.world) s_base[threadIdx.x] = p.ptrs[threadIdx.x];
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  const int64_t per = (p.count + gridDim.x - 1) / gridDim.x;
  const int64_t lo = min(p.count, (int64_t)blockIdx.x * per), hi = min(p.count, lo + per);
  const size_t slot_off = kPeerSignalBytes + (size_t)(p.seq & 1u) * p.slot_doubles * sizeof(double);
  double* mine = reinterpret_cast<double*>(s_base[p.rank] + slot_off);
  for (int64_t i = lo + threadIdx.x; i < hi; i += kPeerThreads) mine[i] = p.in[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    // my arrival, written into the peer's signal row of this CTA
    uint32_t* theirs = reinterpret_cast<uint32_t*>(s_base[peer]) + blockIdx.x * kPeerMaxRanks + p.rank;
    st_release_sys(theirs, p.seq);
    // the peer's arrival in my row; sequence numbers only grow (wrap-around safe compare)
    const uint32_t* own = reinterpret_cast<const uint32_t*>(s_base[p.rank]) + blockIdx.x * kPeerMaxRanks + peer;
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys(own) - p.seq) < 0) {
      if (++spins > (1ll << 27)) {  // ~ seconds: a peer never arrived; report instead of hanging the GPU
        s_fail = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_fail) {
    if (threadIdx.x == 0) *p.error = 1;
    return;
  }
  if (p.op == 2) {
    for (int r = 0; r < p.world; ++r) {
      const double* src = reinterpret_cast<const double*>(s_base[r] + slot_off);
      for (int64_t i = lo + threadIdx.x; i < hi; i += kPeerThreads) p.out[(int64_t)r * p.count + i] = ld_peer(src + i);
    }
    return;
  }
  for (int64_t i = lo + threadIdx.x; i < hi; i += kPeerThreads) {
    // fixed rank order: every rank computes bit-identical results
    double acc = ld_peer(reinterpret_cast<const double*>(s_base[0] + slot_off) + i);
    for (int r = 1; r < p.world; ++r) {
      const double v = ld_peer(reinterpret_cast<const double*>(s_base[r] + slot_off) + i);
      const bool add = p.op == 0 || (p.op == 3 && i < p.n_sum);
      acc = add ? acc + v : (v > acc || v != v ? v : acc);
    }
    p.out[i] = acc;
  }
}

}  // namespace agf

extern "C" size_t agf_peer_buffer_bytes(int64_t slot_doubles) {
  return agf::kPeerSignalBytes + (size_t)2 * (size_t)slot_doubles * sizeof(double);
}

extern "C" int agf_peer_exchange(const uint64_t* peer_ptrs, int32_t rank, int32_t world, uint32_t seq, int32_t op,
                                 const double* in, double* out, int64_t count, int64_t slot_doubles, int32_t* error,
                                 void* stream) {
  using namespace agf;
  AGF_REQUIRE(peer_ptrs && in && out && error, "agf_peer_exchange: null pointer");
  AGF_REQUIRE(world >= 1 && world <= kPeerMaxRanks && rank >= 0 && rank < world, "agf_peer_exchange: bad rank/world");
  const int64_t n_sum = op >> 8;  // op = 3 | (n_sum << 8): the first n_sum values are summed, the rest max-ed
  op &= 0xff;
  AGF_REQUIRE(op >= 0 && op <= 3 && seq != 0, "agf_peer_exchange: bad op / sequence number");
  AGF_REQUIRE(count >= 0 && count <= slot_doubles, "agf_peer_exchange: %lld values exceed the slot (%lld)",
              (long long)count, (long long)slot_doubles);
  PeerParams p;
  p.ptrs = peer_ptrs;
  p.rank = rank;
  p.world = world;
  p.seq = seq;
  p.op = op;
  p.n_sum = n_sum;
  p.in = in;
  p.out = out;
  p.count = count;
  p.slot_doubles = slot_doubles;
  p.error = error;
  // every rank must use the same grid (CTA c pairs with CTA c of the peers): a function of count only
  int64_t ctas = (count + 2047) / 2048;
  if (ctas < 1) ctas = 1;
  if (ctas > kPeerMaxCtas) ctas = kPeerMaxCtas;
  peer_exchange_kernel<<<(int)ctas, kPeerThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  AGF_CUDA_TRY(cudaGetLastError());
  return AGF_OK;
}
