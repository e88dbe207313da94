// Peer-memory plumbing of the sharded step: the cross-GPU barrier that orders the fused
// "kernel writes straight into the peer's buffer over NVLink" exchanges.
//
// The exchanges themselves are not here: they are the stores of route_scatter_kernel (ids +
// occurrence counts into the owners' inboxes), gather_kernel<SEG> (rows into the requesters'
// buffers) and scatter_rows_n_kernel (summed gradients into the owners' buffers).  What those
// need is a point after which every peer's stores are visible — this barrier — and which is
// graph-capturable (a kernel on the stream, no host involvement).
//
// Reference path being replaced: TF's send/recv between PS and worker around
// KvVariableGatherOrInsertV2 / the sparse apply; there is no reference kernel.
#include "common.cuh"
#include "table.h"

namespace kvhbm {

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// flags layout (every rank, symmetric): uint32 flags[world] — flags[p] is written by rank p.
// state (local): state[0] = epoch of the last completed barrier, state[1] = timeout count.
// Thread p publishes epoch e to peer p's flags[rank] and waits for peer p's e in flags[p].
// Epochs only grow and the wait is ">= e", so a peer that is already one barrier ahead is fine.
__global__ void peer_barrier_kernel(uint32_t* const* __restrict__ peer_flags, uint32_t* my_flags,
                                    uint32_t* state, int rank, int world,
                                    unsigned long long timeout_ns) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = state[0] + 1;
  __syncthreads();
  const uint32_t e = s_epoch;
  const int p = threadIdx.x;
  if (p < world && p != rank) {
    __threadfence_system();  // stores of the kernels before this one, to whoever acquires e
    st_release_sys(peer_flags[p] + rank, e);
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(ld_acquire_sys(my_flags + p) - e) < 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {  // a lost peer must not hang the GPU: report and go on
        atomicAdd(&state[1], 1u);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) state[0] = e;
}

}  // namespace

int do_peer_barrier(uint32_t* const* peer_flags, uint32_t* my_flags, uint32_t* state, int rank,
                    int world, int64_t timeout_ms, cudaStream_t st) {
  if (world < 1 || world > 256 || rank < 0 || rank >= world)
    return fail(1, "peer_barrier: bad rank / world");
  if (!peer_flags || !my_flags || !state) return fail(1, "peer_barrier: null buffer");
  if (world == 1) return 0;
  const int threads = (world + 31) / 32 * 32;
  peer_barrier_kernel<<<1, threads, 0, st>>>(peer_flags, my_flags, state, rank, world,
                                             (unsigned long long)timeout_ms * 1000000ull);
  KV_LAUNCHED();
  return 0;
}

}  // namespace kvhbm
