// Row-shard exchange over NVLink peer memory, without a collective library call in the step:
// every rank stores its [B][k] (score | label) block straight into each peer's receive buffer
// (P2P stores through NVSwitch), publishes an epoch flag with system-scope release, and the merge
// kernel of each rank waits for the peers' flags with system-scope acquire before it reads.
// The buffers are symmetric allocations mapped into every rank (torch symmetric memory); this
// file only sees raw peer pointers. Included by aux_kernels.cuh.
#pragma once

namespace keds {

struct P2PPush {
  int n_ranks, my_rank;
  unsigned int epoch;
  long long n16;                       // 16-byte units to copy
  const uint4* src;                    // this rank's block (local memory)
  uint4* dst[P2P_MAX_RANKS];           // where that block lives in each rank's buffer (peer pointers)
  unsigned int* flag[P2P_MAX_RANKS];   // flag word "rank my_rank has delivered" in each rank's buffer
  unsigned int* ticket;                // local counter for "last CTA publishes"
};

__global__ void __launch_bounds__(256)
k_p2p_push(const P2PPush p) {
  griddep_wait();  // the local search results must be complete
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int r = 0; r < p.n_ranks; ++r) {
    if (r == p.my_rank) continue;
    uint4* d = p.dst[r];
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n16; i += stride)
      d[i] = p.src[i];
  }
  __threadfence_system();  // this thread's peer stores are ordered before the ticket below
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    // every CTA has fenced its stores: publish the epoch to the peers, reset the ticket
    if (threadIdx.x < p.n_ranks && static_cast<int>(threadIdx.x) != p.my_rank) {
      __threadfence_system();
      st_release_sys(p.flag[threadIdx.x], p.epoch);
    }
    if (threadIdx.x == 0) *p.ticket = 0u;
  }
}

// Spin (bounded) until every peer's flag has reached `epoch`. Called by one thread per block.
__device__ __forceinline__ bool p2p_wait_flags(const unsigned int* flags, int n_ranks, int my_rank,
                                               unsigned int epoch, unsigned int* err_word) {
  const long long t0 = clock64();
  for (int r = 0; r < n_ranks; ++r) {
    if (r == my_rank) continue;
    // epochs only grow; the signed difference also survives wrap-around
    while (static_cast<int>(ld_acquire_sys(flags + r) - epoch) < 0) {
      if (clock64() - t0 > (1ll << 32)) {  // ~2 s: a peer is gone; report instead of hanging
        atomicCAS(err_word, 0u, 0x500u + r);
        return false;
      }
      __nanosleep(200);
    }
  }
  return true;
}

}  // namespace keds
