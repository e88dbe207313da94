// dedup.cu — the stock TensorFlow ops that sit on the KvVariable path, on the
// device and without sorting: Unique / UniqueWithCounts (first-occurrence
// order, int32 inverse index), UnsortedSegmentSum, and the id routing used by
// key-hash sharding.
#include "table.h"

namespace kvhbm {

struct Workspace {
  int device = 0;
  void* buf = nullptr;
  size_t bytes = 0;
  int grab(size_t need, cudaStream_t st) {
    if (need <= bytes) return 0;
    KV_CUDA(cudaStreamSynchronize(st));
    if (buf) cudaFree(buf);
    buf = nullptr;
    bytes = 0;
    size_t want = need + need / 4;
    KV_CUDA(cudaMalloc(&buf, want));
    bytes = want;
    return 0;
  }
  ~Workspace() { if (buf) cudaFree(buf); }
};

namespace {

constexpr int UB = 1024;  // ids per block in the scan kernels (256 threads x 4)

struct __align__(16) USlot {
  long long key;
  int first;  // smallest position holding this key
  int rank;   // index among the unique keys
};

__global__ void unique_init_kernel(USlot* tab, unsigned long long cap, int* counts, long long n) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  int4 e;
  e.x = 0; e.y = (int)0x80000000u; e.z = 0x7fffffff; e.w = 0;
  for (unsigned long long j = i; j < cap; j += stride) reinterpret_cast<int4*>(tab)[j] = e;
  if (counts)
    for (unsigned long long j = i; j < (unsigned long long)n; j += stride) counts[j] = 0;
}

__global__ void unique_insert_kernel(USlot* tab, unsigned long long mask, int shift,
                                     const long long* __restrict__ ids, long long n,
                                     int* __restrict__ slot_of) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const long long key = ids[i];
    // the sentinel itself is a legal id here: remap it onto a private slot key
    unsigned long long pos = mix64((unsigned long long)key) >> shift;
    for (;;) {
      long long cur = __ldcg(&tab[pos].key);
      if (cur == KEY_EMPTY) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&tab[pos].key),
                                           (unsigned long long)KEY_EMPTY, (unsigned long long)key);
        cur = old == (unsigned long long)KEY_EMPTY ? key : (long long)old;
      }
      if (cur == key) break;
      pos = (pos + 1) & mask;
    }
    atomicMin(&tab[pos].first, (int)i);
    slot_of[i] = (int)pos;
  }
}

__device__ __forceinline__ int block_sum(int v, int* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  __syncthreads();
  return t;
}

// pass 1: number of first occurrences per block of UB ids
__global__ void __launch_bounds__(256)
unique_count_kernel(const USlot* __restrict__ tab, const int* __restrict__ slot_of, long long n,
                    int* __restrict__ block_counts) {
  __shared__ int sh[8];
  const long long base = blockIdx.x * (long long)UB;
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < n) c += (tab[slot_of[i]].first == (int)i);
  }
  const int total = block_sum(c, sh);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// pass 2: exclusive scan of the block counts (one block)
__global__ void __launch_bounds__(1024)
unique_scan_kernel(int* __restrict__ block_counts, int nb, int* __restrict__ num_unique) {
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? block_counts[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const int incl = sh[threadIdx.x];
    const int c0 = carry;
    if (i < nb) block_counts[i] = c0 + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c0 + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_unique = carry;
}

// pass 3: rank of every first occurrence, in position order
__global__ void __launch_bounds__(256)
unique_assign_kernel(USlot* __restrict__ tab, const int* __restrict__ slot_of,
                     const long long* __restrict__ ids, long long n,
                     const int* __restrict__ block_offsets, long long* __restrict__ uniq) {
  __shared__ int warp_tot[8];
  const long long base = blockIdx.x * (long long)UB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread t owns 4 consecutive positions so that ranks follow position order
  const long long i0 = base + threadIdx.x * 4;
  int f[4], s[4];
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = i0 + k;
    s[k] = i < n ? slot_of[i] : 0;
    f[k] = i < n ? (tab[s[k]].first == (int)i) : 0;
    c += f[k];
  }
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += warp_tot[w];
  int r = block_offsets[blockIdx.x] + woff + incl - c;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (f[k]) {
      uniq[r] = ids[i0 + k];
      tab[s[k]].rank = r;
      ++r;
    }
  }
}

__global__ void unique_index_kernel(const USlot* __restrict__ tab, const int* __restrict__ slot_of,
                                    long long n, int* __restrict__ idx, int* __restrict__ counts) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int r = tab[slot_of[i]].rank;
    idx[i] = r;
    if (counts) atomicAdd(&counts[r], 1);
  }
}

// ---- UnsortedSegmentSum ----------------------------------------------------
__global__ void zero_rows_kernel(float* out, long long max_rows, const int* d_rows, int dim) {
  long long rows = max_rows;
  if (d_rows) { long long r = *d_rows; if (r < rows) rows = r; }
  const long long total = rows * dim;
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if ((dim & 3) == 0) {
    for (; e < total / 4; e += stride) reinterpret_cast<float4*>(out)[e] = make_float4(0, 0, 0, 0);
  } else {
    for (; e < total; e += stride) out[e] = 0.f;
  }
}

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// A block takes TILE consecutive rows of `data`.  Rows of the tile that share a
// segment are first summed in shared memory (hot Zipf ids repeat thousands of
// times per batch: without this every duplicate would be a same-address atomic
// in L2), then each distinct segment of the tile is flushed once with
// vectorised reductions.
template <int TILE>
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ data, const int* __restrict__ idx, long long n,
                   int dim, float* __restrict__ out) {
  extern __shared__ __align__(16) float acc[];  // [TILE][dim]
  __shared__ int seg[TILE];                     // segment id of local slot j
  __shared__ int local_of[TILE];                // local slot of row r
  __shared__ int htab[2 * TILE];                // open addressing: segment id -> local slot
  __shared__ int hval[2 * TILE];
  __shared__ int n_local;
  const long long base = blockIdx.x * (long long)TILE;
  const int rows = (int)((n - base) < TILE ? (n - base) : TILE);
  for (int j = threadIdx.x; j < 2 * TILE; j += blockDim.x) htab[j] = -1;
  if (threadIdx.x == 0) n_local = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const int sgm = idx[base + r];
    unsigned h = ((unsigned)sgm * 2654435761u) & (2 * TILE - 1);
    for (;;) {
      int cur = atomicCAS(&htab[h], -1, sgm);
      if (cur == -1) {
        const int j = atomicAdd(&n_local, 1);
        seg[j] = sgm;
        // publish the local slot; losers spin on hval below
        atomicExch(&hval[h], j + 1);
        local_of[r] = j;
        break;
      }
      if (cur == sgm) { local_of[r] = -(int)h - 1; break; }  // resolve after the barrier
      h = (h + 1) & (2 * TILE - 1);
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < rows; r += blockDim.x)
    if (local_of[r] < 0) local_of[r] = hval[-local_of[r] - 1] - 1;
  const int nl = n_local;
  const int d4 = dim >> 2;
  const bool vec = (dim & 3) == 0;
  for (int e = threadIdx.x; e < nl * dim; e += blockDim.x) acc[e] = 0.f;
  __syncthreads();
  if (vec) {
    for (int e = threadIdx.x; e < rows * d4; e += blockDim.x) {
      const int r = e / d4, c = e - r * d4;
      const float4 v = __ldcs(reinterpret_cast<const float4*>(data + (base + r) * dim) + c);
      float* a = acc + local_of[r] * dim + c * 4;
      atomicAdd(a + 0, v.x); atomicAdd(a + 1, v.y); atomicAdd(a + 2, v.z); atomicAdd(a + 3, v.w);
    }
  } else {
    for (int e = threadIdx.x; e < rows * dim; e += blockDim.x) {
      const int r = e / dim, c = e - r * dim;
      atomicAdd(acc + local_of[r] * dim + c, __ldcs(data + (base + r) * dim + c));
    }
  }
  __syncthreads();
  if (vec) {
    for (int e = threadIdx.x; e < nl * d4; e += blockDim.x) {
      const int j = e / d4, c = e - j * d4;
      const float4 v = *reinterpret_cast<const float4*>(acc + j * dim + c * 4);
      red_add_v4(out + (long long)seg[j] * dim + c * 4, v);
    }
  } else {
    for (int e = threadIdx.x; e < nl * dim; e += blockDim.x) {
      const int j = e / dim, c = e - j * dim;
      atomicAdd(out + (long long)seg[j] * dim + c, acc[e]);
    }
  }
}

// ---- id routing for key-hash sharding -----------------------------------------
constexpr int MAX_SHARDS = 256;

__device__ __forceinline__ int owner_of(long long id, int num_shards, int mode) {
  if (mode == 1) {  // google_floor_mod, kernels/utility.h:95-101
    long long m = id % num_shards;
    return (int)(m < 0 ? m + num_shards : m);
  }
  // low bits of the hash; the slot index uses the high bits (common.cuh)
  return (int)(mix64((unsigned long long)id ^ 0x5446534dULL) % (unsigned long long)num_shards);
}

__global__ void __launch_bounds__(256)
partition_count_kernel(const long long* __restrict__ ids, long long n, const int* d_n,
                       int num_shards, int mode, int* __restrict__ shard_counts) {
  __shared__ int hist[MAX_SHARDS];
  if (d_n) { long long dn = *d_n; if (dn < n) n = dn; }
  for (int g = threadIdx.x; g < num_shards; g += blockDim.x) hist[g] = 0;
  __syncthreads();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) atomicAdd(&hist[owner_of(ids[i], num_shards, mode)], 1);
  __syncthreads();
  for (int g = threadIdx.x; g < num_shards; g += blockDim.x)
    if (hist[g]) atomicAdd(&shard_counts[g], hist[g]);
}

__global__ void partition_offsets_kernel(const int* shard_counts, int num_shards, int* offsets,
                                         int* cursors) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int g = 0; g < num_shards; ++g) {
      offsets[g] = acc;
      cursors[g] = 0;
      acc += shard_counts[g];
    }
  }
}

__global__ void __launch_bounds__(256)
partition_scatter_kernel(const long long* __restrict__ ids, long long n, const int* d_n,
                         int num_shards, int mode, const int* __restrict__ offsets,
                         int* __restrict__ cursors, long long* __restrict__ sorted_ids,
                         int* __restrict__ perm) {
  __shared__ int hist[MAX_SHARDS];
  __shared__ int basepos[MAX_SHARDS];
  if (d_n) { long long dn = *d_n; if (dn < n) n = dn; }
  const long long per_block = 256 * 8;
  for (long long b0 = blockIdx.x * per_block; b0 < n; b0 += gridDim.x * per_block) {
    for (int g = threadIdx.x; g < num_shards; g += blockDim.x) hist[g] = 0;
    __syncthreads();
    int own[8], lr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long i = b0 + k * 256 + threadIdx.x;
      own[k] = -1;
      if (i < n) {
        own[k] = owner_of(ids[i], num_shards, mode);
        lr[k] = atomicAdd(&hist[own[k]], 1);
      }
    }
    __syncthreads();
    for (int g = threadIdx.x; g < num_shards; g += blockDim.x)
      basepos[g] = hist[g] ? offsets[g] + atomicAdd(&cursors[g], hist[g]) : 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long i = b0 + k * 256 + threadIdx.x;
      if (own[k] >= 0) {
        const int p = basepos[own[k]] + lr[k];
        sorted_ids[p] = ids[i];
        perm[i] = p;
      }
    }
    __syncthreads();
  }
}

size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

int do_unique(Workspace* ws, const int64_t* ids, int64_t n, int64_t* uniq, int32_t* idx,
              int32_t* counts, int32_t* num_unique, cudaStream_t st) {
  if (n < 0 || n > (1LL << 30)) return fail(1, "unique: n out of range");
  if (n == 0) {
    KV_CUDA(cudaMemsetAsync(num_unique, 0, sizeof(int32_t), st));
    return 0;
  }
  unsigned long long cap = 1024;
  while (cap < (unsigned long long)n * 2) cap <<= 1;
  int lg = 0;
  while ((1ULL << lg) < cap) ++lg;
  const int nb = (int)((n + UB - 1) / UB);
  const size_t b_tab = align_up(cap * sizeof(USlot));
  const size_t b_slot = align_up((size_t)n * sizeof(int));
  const size_t b_blk = align_up((size_t)nb * sizeof(int));
  KV_TRY(ws->grab(b_tab + b_slot + b_blk, st));
  char* p = static_cast<char*>(ws->buf);
  USlot* tab = reinterpret_cast<USlot*>(p);
  int* slot_of = reinterpret_cast<int*>(p + b_tab);
  int* blk = reinterpret_cast<int*>(p + b_tab + b_slot);
  const long long* k = reinterpret_cast<const long long*>(ids);
  const int dev = ws->device;
  unique_init_kernel<<<blocks_for(cap, 256, dev), 256, 0, st>>>(tab, cap, counts, n);
  KV_LAUNCHED();
  unique_insert_kernel<<<blocks_for(n, 256, dev), 256, 0, st>>>(tab, cap - 1, 64 - lg, k, n, slot_of);
  KV_LAUNCHED();
  unique_count_kernel<<<nb, 256, 0, st>>>(tab, slot_of, n, blk);
  KV_LAUNCHED();
  unique_scan_kernel<<<1, 1024, 0, st>>>(blk, nb, num_unique);
  KV_LAUNCHED();
  unique_assign_kernel<<<nb, 256, 0, st>>>(tab, slot_of, k, n, blk, reinterpret_cast<long long*>(uniq));
  KV_LAUNCHED();
  unique_index_kernel<<<blocks_for(n, 256, dev), 256, 0, st>>>(tab, slot_of, n, idx, counts);
  KV_LAUNCHED();
  return 0;
}

int do_segment_sum(Workspace* ws, const float* data, const int32_t* idx, int64_t n, int dim,
                   int64_t max_segments, const int32_t* d_num_segments, float* out,
                   cudaStream_t st) {
  if (dim <= 0) return fail(1, "segment_sum: dim must be positive");
  if (max_segments > 0) {
    const int64_t work = max_segments * (int64_t)((dim & 3) == 0 ? dim / 4 : dim);
    zero_rows_kernel<<<blocks_for(work, 256, ws->device, 16), 256, 0, st>>>(
        out, max_segments, d_num_segments, dim);
    KV_LAUNCHED();
  }
  if (n <= 0) return 0;
  // tile so that the shared accumulators stay <= 32 KB
  if (dim <= 64) {
    constexpr int TILE = 128;
    const size_t smem = (size_t)TILE * dim * sizeof(float);
    segment_sum_kernel<TILE><<<(unsigned)((n + TILE - 1) / TILE), 256, smem, st>>>(data, idx, n, dim, out);
  } else if (dim <= 256) {
    constexpr int TILE = 32;
    const size_t smem = (size_t)TILE * dim * sizeof(float);
    segment_sum_kernel<TILE><<<(unsigned)((n + TILE - 1) / TILE), 256, smem, st>>>(data, idx, n, dim, out);
  } else {
    constexpr int TILE = 8;
    const size_t smem = (size_t)TILE * dim * sizeof(float);
    if (smem > 48 * 1024) return fail(3, "segment_sum: dim too large");
    segment_sum_kernel<TILE><<<(unsigned)((n + TILE - 1) / TILE), 256, smem, st>>>(data, idx, n, dim, out);
  }
  KV_LAUNCHED();
  return 0;
}

int do_partition_ids(Workspace* ws, const int64_t* ids, int64_t n, const int32_t* d_n,
                     int num_shards, int mode, int64_t* sorted_ids, int32_t* perm,
                     int32_t* shard_counts, cudaStream_t st) {
  if (num_shards < 1 || num_shards > MAX_SHARDS)
    return fail(1, "partition_ids: num_shards must be in [1, 256]");
  KV_CUDA(cudaMemsetAsync(shard_counts, 0, sizeof(int32_t) * num_shards, st));
  if (n <= 0) return 0;
  KV_TRY(ws->grab(2 * MAX_SHARDS * sizeof(int), st));
  int* offsets = static_cast<int*>(ws->buf);
  int* cursors = offsets + MAX_SHARDS;
  const long long* k = reinterpret_cast<const long long*>(ids);
  const int dev = ws->device;
  partition_count_kernel<<<blocks_for(n, 256 * 8, dev), 256, 0, st>>>(k, n, d_n, num_shards, mode,
                                                                     shard_counts);
  KV_LAUNCHED();
  partition_offsets_kernel<<<1, 32, 0, st>>>(shard_counts, num_shards, offsets, cursors);
  KV_LAUNCHED();
  partition_scatter_kernel<<<blocks_for(n, 256 * 8, dev), 256, 0, st>>>(
      k, n, d_n, num_shards, mode, offsets, cursors, reinterpret_cast<long long*>(sorted_ids), perm);
  KV_LAUNCHED();
  return 0;
}

Workspace* workspace_new() {
  Workspace* w = new Workspace();
  cudaGetDevice(&w->device);
  return w;
}
void workspace_delete(Workspace* w) { delete w; }

}  // namespace kvhbm
