// dedup.cu — the stock TensorFlow ops that sit on the KvVariable path, on the
// device and without sorting: Unique / UniqueWithCounts (first-occurrence
// order, int32 inverse index), UnsortedSegmentSum, and the id routing used by
// key-hash sharding.
#include <cstdlib>

#include "plan.h"
#include "table.h"

namespace kvhbm {

struct Workspace {
  int device = 0;
  void* buf = nullptr;
  size_t bytes = 0;
  // unique(): scratch tables + selector, see UScratch
  void* ubuf = nullptr;
  unsigned long long ucap = 0;
  int unb = 0;
  void* wiped_buf = nullptr;  // ubuf for which unique_wipe_kernel has been enqueued
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
  ~Workspace() {
    if (buf) cudaFree(buf);
    if (ubuf) cudaFree(ubuf);
  }
};

int set_trace_unique(unsigned long long*) { return 0; }

namespace {

constexpr int UB = 1024;  // ids per block in the scan kernels (256 threads x 4)

constexpr int MAX_SHARDS = 256;

__device__ __forceinline__ int owner_of(long long id, int num_shards, int mode) {
  if (mode == 1) {  // google_floor_mod, kernels/utility.h:95-101
    long long m = id % num_shards;
    return (int)(m < 0 ? m + num_shards : m);
  }
  // low bits of the hash; the slot index uses the high bits (common.cuh)
  return (int)(mix64((unsigned long long)id ^ 0x5446534dULL) % (unsigned long long)num_shards);
}

// Where the rank kernel sends each distinct id when dedup and routing are fused
// (kv_unique_route_peer): shard g's row is dst_ids[g] / dst_occ[g], normally peer g's inbox.
struct RouteOut {
  long long* const* dst_ids;
  int* const* dst_occ;
  int* perm;          // [n]: padded position g * cap + r of unique id r, or -1 (overflow)
  int* shard_counts;  // [num_shards], zeroed by kv_route_fill_peer
  int* overflow;
  int num_shards, mode, cap;
};

// Extra outputs of the dedup when it builds a plan (kv_plan_build): the occurrences of every
// distinct id as a CSR over positions — seg_off[r] = first entry of unique id r (exclusive
// prefix of the counts in rank order), within[i] = entry of position i inside its segment (in
// arrival order; plan_sort_* puts the segments into increasing position afterwards) — and the
// list of "heavy" ids (more than heavy_t occurrences), which the fused apply sums on a
// dedicated pipeline.
struct PlanOut {
  int* first;
  uint2* hint;   // per rank, reset to "unknown" by every build
  int* seg_off;
  int* within;
  int* pos_u;
  int* heavy;
  int* heavy_n;
  int heavy_t;
  int heavy_cap;
};

struct __align__(16) USlot {
  long long key;
  int first;  // smallest position holding this key; once ranked, ~rank (negative)
  int count;  // occurrences of the key
};

// unique() keeps TWO scratch hash tables and alternates between them: the last kernel of a
// call wipes the table the next call will use, so no call pays for a clearing pass on its
// critical path.  Which table is current lives in device memory (sel[0]; sel[1] is the copy
// the last kernel reads), not in host state, so the scheme survives CUDA-graph replay:
// insert and rank read sel[0]; rank publishes sel[1] = sel[0]; index reads sel[1] and writes
// sel[0] = sel[1] ^ 1.  No kernel writes a word that its own blocks read.
struct UScratch {
  USlot* tab[2];
  unsigned long long* status[2];  // look-back words of the rank kernel, one per block
  int* sel;
  unsigned long long cap;  // slots per table (power of two)
  int shift;               // 64 - log2(cap)
  int nb_max;              // status words per table
};

__device__ __forceinline__ void wipe(USlot* tab, unsigned long long cap,
                                     unsigned long long* status, int nb,
                                     unsigned long long i, unsigned long long stride) {
  int4 e;
  e.x = 0; e.y = (int)0x80000000u; e.z = 0x7fffffff; e.w = 0;
  for (unsigned long long j = i; j < cap; j += stride) reinterpret_cast<int4*>(tab)[j] = e;
  for (unsigned long long j = i; j < (unsigned long long)nb; j += stride) status[j] = 0ULL;
}

__global__ void unique_wipe_kernel(UScratch s) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  wipe(s.tab[0], s.cap, s.status[0], s.nb_max, i, stride);
  wipe(s.tab[1], s.cap, s.status[1], s.nb_max, i, stride);
  if (i == 0) { s.sel[0] = 0; s.sel[1] = 0; }
}

// One probe per distinct id of a warp: the copies of a hot id inside a warp elect a leader
// (match.any), so a Zipf head id costs one 64-bit CAS per warp instead of one per occurrence,
// and the atomicMin is skipped once an earlier position is already recorded.  Occurrences are
// counted here too, one fire-and-forget add per (warp, distinct id): counting per occurrence
// in the index kernel serialised ~7 500 adds on the head id's counter (13 us of a 29 us call).
template <bool COUNT>
__global__ void __launch_bounds__(256)
unique_insert_kernel(UScratch s, const long long* __restrict__ ids, long long n,
                     int* __restrict__ slot_of, int* __restrict__ within, int* heavy_n) {
  if (heavy_n && blockIdx.x == 0 && threadIdx.x == 0) *heavy_n = 0;
  USlot* tab = s.tab[s.sel[0] & 1];
  const unsigned long long mask = s.cap - 1;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < n;
       i0 += stride) {
    const long long i = i0 + lane;
    const bool valid = i < n;
    const long long key = valid ? ids[i] : 0;
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) continue;
    const unsigned peers = __match_any_sync(active, key);
    const int leader = __ffs(peers) - 1;  // lowest lane = smallest position
    unsigned long long pos = 0;
    int wbase = 0;
    if (lane == leader) {
      pos = mix64((unsigned long long)key) >> s.shift;
      for (;;) {
        long long cur = __ldcg(&tab[pos].key);
        if (cur == KEY_EMPTY) {
          const unsigned long long old = atomicCAS(
              reinterpret_cast<unsigned long long*>(&tab[pos].key),
              (unsigned long long)KEY_EMPTY, (unsigned long long)key);
          cur = old == (unsigned long long)KEY_EMPTY ? key : (long long)old;
        }
        if (cur == key) break;
        pos = (pos + 1) & mask;
      }
      if (__ldcg(&tab[pos].first) > (int)i) atomicMin(&tab[pos].first, (int)i);
      if (COUNT) {
        // the value the add returns places this warp's occurrences inside the id's segment
        if (within) wbase = atomicAdd(&tab[pos].count, __popc(peers));
        else atomicAdd(&tab[pos].count, __popc(peers));
      }
    }
    pos = __shfl_sync(peers, pos, leader);
    slot_of[i] = (int)pos;
    if (within) {
      wbase = __shfl_sync(peers, wbase, leader);
      within[i] = wbase + __popc(peers & ((1u << lane) - 1u));
    }
  }
}

// Ranks of the first occurrences in position order, in one pass: every block of UB ids
// counts its first occurrences, publishes the count, and obtains the number of first
// occurrences before it by decoupled look-back over the earlier blocks' status words
// (flag 1 = block aggregate, flag 2 = inclusive prefix, in bits 62-63; the word carries TWO
// running sums: first occurrences in bits 0-30, and their occurrence counts in bits 31-61 —
// the second is the exclusive prefix that becomes seg_off when a plan is built; both are
// bounded by n <= 2^30).
//
// ROUTE: the same launch is the id exchange of the sharded step — a first occurrence is also
// given a position in its owner's row (shared histogram, one global add per (block, shard))
// and stored, with its occurrence count, straight into that row (peer memory).
__device__ __forceinline__ unsigned long long st_pack(unsigned long long flag, int ranks,
                                                      int occ) {
  return (flag << 62) | ((unsigned long long)(unsigned)occ << 31) | (unsigned long long)(unsigned)ranks;
}
template <bool ROUTE, bool PLAN>
__global__ void __launch_bounds__(256)
unique_rank_kernel(UScratch sc, const int* __restrict__ slot_of,
                   const long long* __restrict__ ids, long long n,
                   long long* __restrict__ uniq, int* __restrict__ counts,
                   int* __restrict__ num_unique, RouteOut ro, PlanOut po) {
  __shared__ int warp_tot[8], warp_occ[8];
  __shared__ int block_prefix, block_occ;
  const int which = sc.sel[0] & 1;
  USlot* __restrict__ tab = sc.tab[which];
  unsigned long long* status = sc.status[which];
  if (blockIdx.x == 0 && threadIdx.x == 0) sc.sel[1] = which;
  const long long base = blockIdx.x * (long long)UB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long i0 = base + threadIdx.x * 4;  // 4 consecutive positions per thread
  int f[4], s[4], cnt[4];
  int c = 0, oc = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = i0 + k;
    s[k] = i < n ? slot_of[i] : 0;
    f[k] = 0;
    cnt[k] = 0;
    if (i < n) {
      const int4 v = __ldcg(reinterpret_cast<const int4*>(&tab[s[k]]));  // {key, first, count}
      f[k] = v.z == (int)i;
      cnt[k] = v.w;
    }
    c += f[k];
    if (PLAN) oc += f[k] ? cnt[k] : 0;
  }
  int incl = c, oincl = oc;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    const int t2 = PLAN ? __shfl_up_sync(0xffffffffu, oincl, o) : 0;
    if (lane >= o) { incl += t; oincl += t2; }
  }
  if (lane == 31) { warp_tot[warp] = incl; warp_occ[warp] = oincl; }
  __syncthreads();
  int woff = 0, total = 0, wocc = 0, tocc = 0;
  for (int w = 0; w < 8; ++w) {
    if (w < warp) { woff += warp_tot[w]; wocc += warp_occ[w]; }
    total += warp_tot[w];
    tocc += warp_occ[w];
  }
  if (warp == 0) {
    const int b = blockIdx.x;
    volatile unsigned long long* st = status;
    if (lane == 0) st[b] = st_pack(b == 0 ? 2ULL : 1ULL, total, tocc);
    int prefix = 0, oprefix = 0;
    for (int j = b - 1; j >= 0; j -= 32) {
      const int idx = j - lane;
      unsigned long long v = 2ULL << 62;  // lanes past block 0 act as a zero prefix
      if (idx >= 0) {
        do { v = st[idx]; } while ((v >> 62) == 0);
      }
      const unsigned done = __ballot_sync(0xffffffffu, (v >> 62) == 2);
      const int stop = done ? __ffs(done) - 1 : 32;  // nearest block that has its prefix
      int add = lane <= stop ? (int)(v & 0x7fffffffULL) : 0;
      int oadd = lane <= stop ? (int)((v >> 31) & 0x7fffffffULL) : 0;
      for (int o = 16; o > 0; o >>= 1) {
        add += __shfl_xor_sync(0xffffffffu, add, o);
        oadd += __shfl_xor_sync(0xffffffffu, oadd, o);
      }
      prefix += add;
      oprefix += oadd;
      if (done) break;
    }
    if (lane == 0) {
      if (b > 0) st[b] = st_pack(2ULL, prefix + total, oprefix + tocc);
      block_prefix = prefix;
      block_occ = oprefix;
      if (b == (int)gridDim.x - 1) *num_unique = prefix + total;
    }
  }
  __syncthreads();
  int r = block_prefix + woff + incl - c;
  int off = block_occ + wocc + oincl - oc;
  long long key[4];
  int rk[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (f[k]) {
      // the other occurrences only test first == their own position: ~r is negative, so
      // overwriting first with the rank cannot turn any of them into a first occurrence
      key[k] = ids[i0 + k];
      uniq[r] = key[k];
      if (counts) counts[r] = cnt[k];
      tab[s[k]].first = ~r;
      rk[k] = r;
      if (PLAN) {
        po.first[r] = (int)(i0 + k);
        po.seg_off[r] = off;
        if (cnt[k] > po.heavy_t) {
          const int h = atomicAdd(po.heavy_n, 1);
          if (h < po.heavy_cap) po.heavy[h] = r;
        }
        off += cnt[k];
      }
      ++r;
    }
  }
  if (ROUTE) {
    __shared__ int hist[MAX_SHARDS];
    __shared__ int basepos[MAX_SHARDS];
    for (int g = threadIdx.x; g < ro.num_shards; g += blockDim.x) hist[g] = 0;
    __syncthreads();
    int own[4], lr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (f[k]) {
        own[k] = owner_of(key[k], ro.num_shards, ro.mode);
        lr[k] = atomicAdd(&hist[own[k]], 1);
      }
    }
    __syncthreads();
    for (int g = threadIdx.x; g < ro.num_shards; g += blockDim.x)
      basepos[g] = hist[g] ? atomicAdd(&ro.shard_counts[g], hist[g]) : 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (f[k]) {
        const int pos = basepos[own[k]] + lr[k];
        if (pos < ro.cap) {
          ro.dst_ids[own[k]][pos] = key[k];
          ro.dst_occ[own[k]][pos] = cnt[k];
          ro.perm[rk[k]] = own[k] * ro.cap + pos;
        } else {
          ro.perm[rk[k]] = -1;  // does not fit: reported, the caller falls back to the exact path
          atomicExch(ro.overflow, 1);
        }
      }
    }
  }
}

// idx[i] = rank of id i's slot; with a plan also pos_u[seg_off[rank] + within[i]] = i (the
// CSR of occurrences, segments still in arrival order); the same launch wipes the other table
// for the next call.
__global__ void unique_index_kernel(UScratch s, const int* __restrict__ slot_of, long long n,
                                    int* __restrict__ idx, PlanOut po) {
  const int which = s.sel[1] & 1;
  const USlot* __restrict__ tab = s.tab[which];
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = i; j < n; j += stride) {
    const int r = ~tab[slot_of[j]].first;
    idx[j] = r;
    if (po.pos_u) {
      po.pos_u[po.seg_off[r] + po.within[j]] = (int)j;
      po.hint[j] = make_uint2(0xffffffffu, 0u);  // (indexed by rank; ranks < n as well)
    }
  }
  // NB: the table just used stays dirty until the call after next wipes it
  wipe(s.tab[which ^ 1], s.cap, s.status[which ^ 1], s.nb_max, (unsigned long long)i,
       (unsigned long long)stride);
  if (i == 0) s.sel[0] = which ^ 1;
}

// ---- plan: segments into increasing position ---------------------------------------------
// Light segments (<= 32 occurrences): a warp loads the segment and every lane ranks its
// position among the others (positions are distinct, so the ranks are a permutation).
__device__ __forceinline__ void
plan_sort_light(const int* __restrict__ counts, const int* __restrict__ seg_off,
                const int* __restrict__ num_unique, const int* __restrict__ pos_u,
                int* __restrict__ pos, int heavy_t, long long block, long long nblocks) {
  const int lane = threadIdx.x & 31;
  const long long warp = (block * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (nblocks * blockDim.x) >> 5;
  const int U = *num_unique;
  for (long long r0 = warp * 32; r0 < U; r0 += nwarps * 32) {
    const long long r = r0 + lane;
    const int c = r < U ? counts[r] : 0;
    const int off = r < U ? seg_off[r] : 0;
    if (c == 1) pos[off] = pos_u[off];
    unsigned todo = __ballot_sync(0xffffffffu, c >= 2 && c <= heavy_t);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int cc = __shfl_sync(0xffffffffu, c, src);
      const int oo = __shfl_sync(0xffffffffu, off, src);
      for (int k0 = 0; k0 < cc; k0 += 32) {   // heavy_t may exceed 32: rank chunk by chunk
        const int mine = k0 + lane < cc ? pos_u[oo + k0 + lane] : 0x7fffffff;
        int rank = 0;
        for (int q0 = 0; q0 < cc; q0 += 32) {
          const int other = q0 == k0 ? mine : (q0 + lane < cc ? pos_u[oo + q0 + lane] : 0x7fffffff);
          const int lim = cc - q0 < 32 ? cc - q0 : 32;
          for (int j = 0; j < lim; ++j) rank += __shfl_sync(0xffffffffu, other, j) < mine;
        }
        if (k0 + lane < cc) pos[oo + rank] = mine;
      }
    }
  }
}

// Heavy segments: a block per segment marks the segment's positions in a shared-memory bitmap
// (one chunk of PLAN_CHUNK positions at a time) and writes the set bits back in order.
constexpr int PLAN_CHUNK = 1 << 18;              // positions per bitmap chunk (32 KB of bits)
constexpr int PLAN_WORDS = PLAN_CHUNK / 32;
__device__ __forceinline__ void
plan_sort_heavy(const int* __restrict__ counts, const int* __restrict__ seg_off,
                const int* __restrict__ heavy, const int* __restrict__ heavy_n,
                int heavy_cap, const int* __restrict__ pos_u, int* __restrict__ pos,
                unsigned char* __restrict__ eflag, long long n, int block, int nblocks) {
  __shared__ unsigned bm[PLAN_WORDS];
  __shared__ int wsum[16];
  __shared__ int run_base;
  int H = *heavy_n;
  if (H > heavy_cap) H = heavy_cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int h = block; h < H; h += nblocks) {
    const int r = heavy[h];
    const int c = counts[r], off = seg_off[r];
    if (threadIdx.x == 0) run_base = 0;
    for (long long c0 = 0; c0 < n; c0 += PLAN_CHUNK) {
      const int words = (int)(((n - c0 < PLAN_CHUNK ? n - c0 : PLAN_CHUNK) + 31) >> 5);
      for (int w = threadIdx.x; w < words; w += blockDim.x) bm[w] = 0u;
      __syncthreads();
      for (int e = threadIdx.x; e < c; e += blockDim.x) {
        const long long p = pos_u[off + e] - c0;
        if (p >= 0 && p < PLAN_CHUNK) atomicOr(&bm[p >> 5], 1u << (p & 31));
      }
      __syncthreads();
      // every thread owns a contiguous range of words
      const int per = (words + blockDim.x - 1) / blockDim.x;
      const int w0 = threadIdx.x * per;
      int mine = 0;
      for (int w = w0; w < w0 + per && w < words; ++w) mine += __popc(bm[w]);
      int incl = mine;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      int before = run_base + incl - mine;
      int total = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        if (w < warp) before += wsum[w];
        total += wsum[w];
      }
      for (int w = w0; w < w0 + per && w < words; ++w) {
        unsigned bits = bm[w];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          eflag[off + before] = 1;
          pos[off + before++] = (int)(c0 + (long long)w * 32 + b);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) run_base += total;
      __syncthreads();
    }
  }
}

// Both sorts in one launch: the first `hb` blocks take the heavy segments (the longest work,
// started first), the others the light ones.
__global__ void __launch_bounds__(512)
plan_sort_kernel(const int* __restrict__ counts, const int* __restrict__ seg_off,
                 const int* __restrict__ num_unique, const int* __restrict__ heavy,
                 const int* __restrict__ heavy_n, int heavy_cap, int heavy_t,
                 const int* __restrict__ pos_u, int* __restrict__ pos,
                 unsigned char* __restrict__ eflag, long long n, int hb) {
  if ((int)blockIdx.x < hb)
    plan_sort_heavy(counts, seg_off, heavy, heavy_n, heavy_cap, pos_u, pos, eflag, n, blockIdx.x, hb);
  else
    plan_sort_light(counts, seg_off, num_unique, pos_u, pos, heavy_t, (long long)blockIdx.x - hb,
                    (long long)gridDim.x - hb);
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

// A block takes TILE consecutive rows of `data` and issues their loads first.  While they
// are in flight the rows' segment ids are hashed in shared memory, which tells every row
// whether its segment occurs once in the tile (the common case: the row goes straight from
// registers to the output with one vectorised reduction per 16 bytes) or several times (hot
// Zipf ids repeat thousands of times per batch: those rows are staged in shared memory,
// summed per segment by walking a per-segment row list, and flushed once — otherwise every
// duplicate would be a same-address atomic in L2).
template <int TILE>
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ data, const int* __restrict__ idx, long long n,
                   int dim, float* __restrict__ out) {
  extern __shared__ __align__(16) float stage[];  // [TILE][dim], only duplicate rows land here
  __shared__ int seg[TILE];                       // segment id of local slot j
  __shared__ int local_of[TILE];                  // local slot of row r
  __shared__ int htab[2 * TILE];                  // open addressing: segment id -> local slot
  __shared__ int hval[2 * TILE];
  __shared__ int head[TILE], nxt[TILE], cnt[TILE];
  __shared__ int n_local;
  const long long base = blockIdx.x * (long long)TILE;
  const int rows = (int)((n - base) < TILE ? (n - base) : TILE);
  const int d4 = dim >> 2;
  const bool vec = (dim & 3) == 0;
  constexpr int PRE = 8;
  float4 pre[PRE];
  if (vec) {
#pragma unroll
    for (int k = 0; k < PRE; ++k) {
      const int e = threadIdx.x + k * 256;
      if (e < rows * d4) {
        const int r = e / d4, c = e - r * d4;
        pre[k] = __ldcs(reinterpret_cast<const float4*>(data + (base + r) * dim) + c);
      }
    }
  }
  for (int j = threadIdx.x; j < 2 * TILE; j += blockDim.x) htab[j] = -1;
  for (int j = threadIdx.x; j < TILE; j += blockDim.x) { head[j] = -1; cnt[j] = 0; }
  if (threadIdx.x == 0) n_local = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const int sgm = idx[base + r];
    unsigned h = ((unsigned)sgm * 2654435761u) & (2 * TILE - 1);
    for (;;) {
      const int cur = atomicCAS(&htab[h], -1, sgm);
      if (cur == -1) {
        const int j = atomicAdd(&n_local, 1);
        seg[j] = sgm;
        hval[h] = j;
        local_of[r] = j;
        break;
      }
      if (cur == sgm) { local_of[r] = -(int)h - 1; break; }  // resolved after the barrier
      h = (h + 1) & (2 * TILE - 1);
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    int j = local_of[r];
    if (j < 0) { j = hval[-j - 1]; local_of[r] = j; }
    nxt[r] = atomicExch(&head[j], r);
    atomicAdd(&cnt[j], 1);
  }
  __syncthreads();
  const int nl = n_local;
  if (vec) {
#pragma unroll
    for (int k = 0; k < PRE; ++k) {
      const int e = threadIdx.x + k * 256;
      if (e < rows * d4) {
        const int r = e / d4, c = e - r * d4;
        const int j = local_of[r];
        if (cnt[j] == 1) red_add_v4(out + (long long)seg[j] * dim + c * 4, pre[k]);
        else *reinterpret_cast<float4*>(stage + r * dim + c * 4) = pre[k];
      }
    }
    for (int e = threadIdx.x + PRE * 256; e < rows * d4; e += blockDim.x) {
      const int r = e / d4, c = e - r * d4;
      const int j = local_of[r];
      const float4 v = __ldcs(reinterpret_cast<const float4*>(data + (base + r) * dim) + c);
      if (cnt[j] == 1) red_add_v4(out + (long long)seg[j] * dim + c * 4, v);
      else *reinterpret_cast<float4*>(stage + r * dim + c * 4) = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nl * d4; e += blockDim.x) {
      const int j = e / d4, c = e - j * d4;
      if (cnt[j] < 2) continue;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = head[j]; r >= 0; r = nxt[r]) {
        const float4 v = *reinterpret_cast<const float4*>(stage + r * dim + c * 4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      red_add_v4(out + (long long)seg[j] * dim + c * 4, a);
    }
  } else {
    for (int e = threadIdx.x; e < rows * dim; e += blockDim.x) {
      const int r = e / dim, c = e - r * dim;
      const int j = local_of[r];
      const float v = __ldcs(data + (base + r) * dim + c);
      if (cnt[j] == 1) atomicAdd(out + (long long)seg[j] * dim + c, v);
      else stage[r * dim + c] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nl * dim; e += blockDim.x) {
      const int j = e / dim, c = e - j * dim;
      if (cnt[j] < 2) continue;
      float a = 0.f;
      for (int r = head[j]; r >= 0; r = nxt[r]) a += stage[r * dim + c];
      atomicAdd(out + (long long)seg[j] * dim + c, a);
    }
  }
}

// ---- id routing for key-hash sharding -----------------------------------------
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

// ---- fixed-capacity routing: the sync-free variant of partition_ids -------------------------
// send_ids / send_occ are [num_shards][cap]: the ids owned by shard g go to [g][0 .. count_g),
// the rest of each row is padding (KEY_PAD / 0).  Shapes do not depend on the data, so the
// exchange that follows needs no host synchronisation and can be captured in a CUDA graph.
// pairs != 0: send_ids holds interleaved {id, occurrence count} int64 pairs (one exchange
// carries both); otherwise ids and counts go to two separate arrays.
// dst_ids != null: shard g's row lives at dst_ids[g] / dst_occ[g] — peer g's inbox mapped over
// NVLink — instead of send_ids + g * cap, so routing and the id exchange are one kernel.
__global__ void route_fill_kernel(long long* send_ids, int* send_occ, long long total, int* counts,
                                  int num_shards, int pairs,
                                  long long* const* __restrict__ dst_ids,
                                  int* const* __restrict__ dst_occ, int cap) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = i; j < total; j += stride) {
    if (dst_ids) {
      const long long g = j / cap;
      dst_ids[g][j - g * cap] = KEY_PAD;
      dst_occ[g][j - g * cap] = 0;
    } else if (pairs) { send_ids[2 * j] = KEY_PAD; send_ids[2 * j + 1] = 0; }
    else {
      send_ids[j] = KEY_PAD;
      if (send_occ) send_occ[j] = 0;
    }
  }
  if (i < num_shards) counts[i] = 0;
}

__global__ void unzip_pairs_kernel(const long long* __restrict__ pairs, long long n,
                                   long long* __restrict__ ids, int* __restrict__ occ) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(pairs) + i);
    ids[i] = v.x;
    occ[i] = (int)v.y;
  }
}

constexpr int ROUTE_K = 2;

__global__ void __launch_bounds__(256)
route_scatter_kernel(const long long* __restrict__ ids, const int* __restrict__ occ, long long n,
                     const int* d_n, int num_shards, int mode, int cap,
                     long long* __restrict__ send_ids, int* __restrict__ send_occ,
                     int* __restrict__ perm, int* __restrict__ counts, int* __restrict__ overflow,
                     int pairs, long long* const* __restrict__ dst_ids,
                     int* const* __restrict__ dst_occ) {
  __shared__ int hist[MAX_SHARDS];
  __shared__ int basepos[MAX_SHARDS];
  if (d_n) { long long dn = *d_n; if (dn < n) n = dn; }
  // 2 ids per thread: the kernel is a chain of dependent phases (load, shared histogram,
  // one global add per shard, stores), so it wants many small blocks, not few long ones
  constexpr int K = ROUTE_K;
  const long long per_block = 256 * K;
  for (long long b0 = blockIdx.x * per_block; b0 < n; b0 += gridDim.x * per_block) {
    for (int g = threadIdx.x; g < num_shards; g += blockDim.x) hist[g] = 0;
    __syncthreads();
    int own[K], lr[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const long long i = b0 + k * 256 + threadIdx.x;
      own[k] = -1;
      if (i < n) {
        own[k] = owner_of(ids[i], num_shards, mode);
        lr[k] = atomicAdd(&hist[own[k]], 1);
      }
    }
    __syncthreads();
    for (int g = threadIdx.x; g < num_shards; g += blockDim.x)
      basepos[g] = hist[g] ? atomicAdd(&counts[g], hist[g]) : 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const long long i = b0 + k * 256 + threadIdx.x;
      if (own[k] >= 0) {
        const int r = basepos[own[k]] + lr[k];
        if (r < cap) {
          const long long p = (long long)own[k] * cap + r;
          if (dst_ids) {
            dst_ids[own[k]][r] = ids[i];
            dst_occ[own[k]][r] = occ ? occ[i] : 1;
          } else if (pairs) {
            reinterpret_cast<longlong2*>(send_ids)[p] = make_longlong2(ids[i], occ ? occ[i] : 1);
          } else {
            send_ids[p] = ids[i];
            if (send_occ) send_occ[p] = occ ? occ[i] : 1;
          }
          perm[i] = (int)p;
        } else {
          perm[i] = -1;  // does not fit: reported, the caller falls back to the exact path
          atomicExch(overflow, 1);
        }
      }
    }
    __syncthreads();
  }
}

size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

static bool ws_wiped(const Workspace* ws) { return ws->wiped_buf == ws->ubuf; }
static void ws_mark_wiped(Workspace* ws) { ws->wiped_buf = ws->ubuf; }

int do_unique_impl(Workspace* ws, const int64_t* ids, int64_t n, int64_t* uniq, int32_t* idx,
                   int32_t* counts, int32_t* num_unique, const RouteOut* route, cudaStream_t st,
                   const PlanOut* plan = nullptr) {
  if (n < 0 || n > (1LL << 30)) return fail(1, "unique: n out of range");
  if (n == 0) {
    KV_CUDA(cudaMemsetAsync(num_unique, 0, sizeof(int32_t), st));
    return 0;
  }
  unsigned long long cap = 1024;
  while (cap < (unsigned long long)n * 2) cap <<= 1;
  const int nb = (int)((n + UB - 1) / UB);
  const int dev = ws->device;
  if (cap > ws->ucap || nb > ws->unb) {  // (re)allocate and wipe both tables: rare, synchronises
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) cudaGetLastError();
    if (cs != cudaStreamCaptureStatusNone)
      return fail(2, "unique: scratch must be sized before CUDA-graph capture (run once eagerly)");
    KV_CUDA(cudaStreamSynchronize(st));
    if (ws->ubuf) cudaFree(ws->ubuf);
    ws->ubuf = nullptr;
    ws->wiped_buf = nullptr;
    ws->ucap = cap > ws->ucap ? cap : ws->ucap;
    ws->unb = nb > ws->unb ? nb : ws->unb;
    const size_t b_tab = align_up(ws->ucap * sizeof(USlot));
    const size_t b_st = align_up((size_t)ws->unb * sizeof(unsigned long long));
    KV_CUDA(cudaMalloc(&ws->ubuf, 2 * (b_tab + b_st) + 256));
  }
  UScratch sc;
  {
    const size_t b_tab = align_up(ws->ucap * sizeof(USlot));
    const size_t b_st = align_up((size_t)ws->unb * sizeof(unsigned long long));
    char* p = static_cast<char*>(ws->ubuf);
    sc.tab[0] = reinterpret_cast<USlot*>(p);
    sc.tab[1] = reinterpret_cast<USlot*>(p + b_tab);
    sc.status[0] = reinterpret_cast<unsigned long long*>(p + 2 * b_tab);
    sc.status[1] = reinterpret_cast<unsigned long long*>(p + 2 * b_tab + b_st);
    sc.sel = reinterpret_cast<int*>(p + 2 * (b_tab + b_st));
    sc.cap = ws->ucap;
    int lg = 0;
    while ((1ULL << lg) < sc.cap) ++lg;
    sc.shift = 64 - lg;
    sc.nb_max = ws->unb;
  }
  static_assert(sizeof(UScratch) <= 64, "passed by value");
  if (!ws->bytes || ws->bytes < align_up((size_t)n * sizeof(int))) KV_TRY(ws->grab(align_up((size_t)n * sizeof(int)), st));
  int* slot_of = static_cast<int*>(ws->buf);
  const long long* k = reinterpret_cast<const long long*>(ids);
  if (ws->ubuf && !ws_wiped(ws)) {
    unique_wipe_kernel<<<blocks_for(sc.cap, 256, dev), 256, 0, st>>>(sc);
    KV_LAUNCHED();
    ws_mark_wiped(ws);
  }
  const PlanOut po = plan ? *plan : PlanOut{};
  if (counts || route || plan)
    unique_insert_kernel<true><<<blocks_for(n, 256, dev), 256, 0, st>>>(sc, k, n, slot_of, po.within,
                                                                      po.heavy_n);
  else
    unique_insert_kernel<false><<<blocks_for(n, 256, dev), 256, 0, st>>>(sc, k, n, slot_of, nullptr,
                                                                       nullptr);
  KV_LAUNCHED();
  long long* uq = reinterpret_cast<long long*>(uniq);
  if (route)
    unique_rank_kernel<true, false><<<nb, 256, 0, st>>>(sc, slot_of, k, n, uq, counts, num_unique,
                                                        *route, po);
  else if (plan)
    unique_rank_kernel<false, true><<<nb, 256, 0, st>>>(sc, slot_of, k, n, uq, counts, num_unique,
                                                        RouteOut{}, po);
  else
    unique_rank_kernel<false, false><<<nb, 256, 0, st>>>(sc, slot_of, k, n, uq, counts, num_unique,
                                                         RouteOut{}, po);
  KV_LAUNCHED();
  const long long span = n > (long long)sc.cap ? n : (long long)sc.cap;
  unique_index_kernel<<<blocks_for(span, 256, dev), 256, 0, st>>>(sc, slot_of, n, idx, po);
  KV_LAUNCHED();
  return 0;
}

int do_unique(Workspace* ws, const int64_t* ids, int64_t n, int64_t* uniq, int32_t* idx,
              int32_t* counts, int32_t* num_unique, cudaStream_t st) {
  return do_unique_impl(ws, ids, n, uniq, idx, counts, num_unique, nullptr, st);
}

// kv_unique + kv_route_ids_peer in the same three launches (the rank kernel routes).
int do_unique_route(Workspace* ws, const int64_t* ids, int64_t n, int64_t* uniq, int32_t* idx,
                    int32_t* counts, int32_t* num_unique, int num_shards, int mode, int cap,
                    int64_t* const* dst_ids, int32_t* const* dst_occ, int32_t* perm,
                    int32_t* shard_counts, int32_t* overflow, cudaStream_t st) {
  if (num_shards < 1 || num_shards > MAX_SHARDS)
    return fail(1, "unique_route: num_shards must be in [1, 256]");
  if (cap < 1) return fail(1, "unique_route: capacity must be positive");
  RouteOut ro;
  ro.dst_ids = reinterpret_cast<long long* const*>(dst_ids);
  ro.dst_occ = dst_occ;
  ro.perm = perm;
  ro.shard_counts = shard_counts;
  ro.overflow = overflow;
  ro.num_shards = num_shards;
  ro.mode = mode;
  ro.cap = cap;
  return do_unique_impl(ws, ids, n, uniq, idx, counts, num_unique, &ro, st);
}

// Padding + zeroed shard counters for the next kv_unique_route_peer / kv_route_ids_peer.
int do_route_fill(int num_shards, int cap, int64_t* const* dst_ids, int32_t* const* dst_occ,
                  int32_t* shard_counts, cudaStream_t st) {
  if (num_shards < 1 || num_shards > MAX_SHARDS)
    return fail(1, "route_fill: num_shards must be in [1, 256]");
  if (cap < 1) return fail(1, "route_fill: capacity must be positive");
  int dev = 0;
  KV_CUDA(cudaGetDevice(&dev));
  const long long total = (long long)num_shards * cap;
  route_fill_kernel<<<blocks_for(total, 256, dev), 256, 0, st>>>(
      nullptr, nullptr, total, shard_counts, num_shards, 0,
      reinterpret_cast<long long* const*>(dst_ids), dst_occ, cap);
  KV_LAUNCHED();
  return 0;
}

int do_zero_rows(float* out, int64_t max_rows, const int32_t* d_rows, int dim, cudaStream_t st) {
  if (max_rows <= 0) return 0;
  int dev = 0;
  KV_CUDA(cudaGetDevice(&dev));
  const int64_t work = max_rows * (int64_t)((dim & 3) == 0 ? dim / 4 : dim);
  zero_rows_kernel<<<blocks_for(work, 256, dev, 16), 256, 0, st>>>(out, max_rows, d_rows, dim);
  KV_LAUNCHED();
  return 0;
}

int do_segment_sum(Workspace* ws, const float* data, const int32_t* idx, int64_t n, int dim,
                   int64_t max_segments, const int32_t* d_num_segments, float* out,
                   int accumulate, cudaStream_t st) {
  if (dim <= 0) return fail(1, "segment_sum: dim must be positive");
  if (max_segments > 0 && !accumulate) {
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

int do_route_ids(Workspace* ws, const int64_t* ids, const int32_t* occ, int64_t n,
                 const int32_t* d_n, int num_shards, int mode, int cap, int64_t* send_ids,
                 int32_t* send_occ, int32_t* perm, int32_t* counts, int32_t* overflow, int pairs,
                 int64_t* const* dst_ids, int32_t* const* dst_occ, cudaStream_t st) {
  if (num_shards < 1 || num_shards > MAX_SHARDS)
    return fail(1, "route_ids: num_shards must be in [1, 256]");
  if (cap < 1) return fail(1, "route_ids: capacity must be positive");
  if (!send_ids && !(dst_ids && dst_occ))
    return fail(1, "route_ids: needs a send buffer or destination pointers");
  const int dev = ws->device;
  const long long total = (long long)num_shards * cap;
  route_fill_kernel<<<blocks_for(total, 256, dev), 256, 0, st>>>(
      reinterpret_cast<long long*>(send_ids), send_occ, total, counts, num_shards, pairs,
      reinterpret_cast<long long* const*>(dst_ids), dst_occ, cap);
  KV_LAUNCHED();
  if (n <= 0) return 0;
  route_scatter_kernel<<<blocks_for(n, 256 * ROUTE_K, dev), 256, 0, st>>>(
      reinterpret_cast<const long long*>(ids), occ, n, d_n, num_shards, mode, cap,
      reinterpret_cast<long long*>(send_ids), send_occ, perm, counts, overflow, pairs,
      reinterpret_cast<long long* const*>(dst_ids), dst_occ);
  KV_LAUNCHED();
  return 0;
}

int do_unzip_pairs(const int64_t* pairs, int64_t n, int64_t* ids, int32_t* occ, cudaStream_t st) {
  if (n <= 0) return 0;
  int dev = 0;
  KV_CUDA(cudaGetDevice(&dev));
  unzip_pairs_kernel<<<blocks_for(n, 256, dev), 256, 0, st>>>(
      reinterpret_cast<const long long*>(pairs), n, reinterpret_cast<long long*>(ids), occ);
  KV_LAUNCHED();
  return 0;
}

// ---- dedup plan -------------------------------------------------------------------------------
// Everything the hot path derives from the ids of one batch alone — tf.unique_with_counts plus
// the CSR of occurrences — kept on the device for the lookup and the fused apply of that batch
// (kvhbm.h kv_plan_*).  Depends on nothing but the ids, so it can be built ahead of the step.
struct Plan {
  int device = 0;
  int64_t cap = 0;        // ids the buffers hold
  int64_t n = 0;          // ids of the last build
  int heavy_t = 32;
  int heavy_cap = 0;
  long long* uniq = nullptr;
  int *idx = nullptr, *counts = nullptr, *num = nullptr, *seg_off = nullptr, *pos = nullptr;
  int *pos_u = nullptr, *within = nullptr, *heavy = nullptr, *heavy_n = nullptr;
  int* first = nullptr;
  uint2* hint = nullptr;
  // scratch of the fused apply's heavy path: gradient sums [heavy_cap][sum_dim], arrival counters
  unsigned char* eflag = nullptr;
  float* heavy_sum = nullptr;
  float* staged = nullptr;
  int sum_dim = 0;
  unsigned* heavy_done = nullptr;
  unsigned* work = nullptr;
  void* block = nullptr;
  ~Plan() {
    if (block) cudaFree(block);
    if (heavy_sum) cudaFree(heavy_sum);
    if (staged) cudaFree(staged);
  }
};

__global__ void plan_clear_hints_kernel(uint2* hint, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) hint[i] = make_uint2(0xffffffffu, 0u);
}
// A new build invalidates what the previous batch's lookup left behind.
int plan_forget_hints(Plan* p, cudaStream_t st) {
  plan_clear_hints_kernel<<<blocks_for(p->cap, 256, p->device), 256, 0, st>>>(p->hint, p->cap);
  KV_LAUNCHED();
  return 0;
}

Plan* plan_new(int64_t max_ids, int heavy_t, int* rc) {
  *rc = 0;
  if (max_ids < 1 || max_ids > (1LL << 30)) { *rc = fail(1, "plan: max_ids must be in [1, 2^30]"); return nullptr; }
  Plan* p = new Plan();
  cudaGetDevice(&p->device);
  p->cap = max_ids;
  static const int ht_env = getenv("KVHBM_PLAN_HEAVY") ? atoi(getenv("KVHBM_PLAN_HEAVY")) : 0;
  p->heavy_t = heavy_t > 0 ? heavy_t : (ht_env > 0 ? ht_env : 32);
  p->heavy_cap = (int)(max_ids / (p->heavy_t + 1)) + 1;
  const size_t a_i = align_up((size_t)max_ids * sizeof(int));
  const size_t a_l = align_up((size_t)max_ids * sizeof(long long));
  const size_t a_h = align_up((size_t)p->heavy_cap * sizeof(int));
  const size_t a_f = align_up((size_t)max_ids);
  const size_t total = 2 * a_l + 7 * a_i + 2 * a_h + a_f + 512;
  cudaError_t e = cudaMalloc(&p->block, total);
  if (e != cudaSuccess) { *rc = cuda_fail(e, "cudaMalloc(plan)"); delete p; return nullptr; }
  cudaMemset(p->block, 0, total);
  char* c = static_cast<char*>(p->block);
  p->uniq = reinterpret_cast<long long*>(c); c += a_l;
  p->idx = reinterpret_cast<int*>(c); c += a_i;
  p->counts = reinterpret_cast<int*>(c); c += a_i;
  p->seg_off = reinterpret_cast<int*>(c); c += a_i;
  p->pos = reinterpret_cast<int*>(c); c += a_i;
  p->pos_u = reinterpret_cast<int*>(c); c += a_i;
  p->within = reinterpret_cast<int*>(c); c += a_i;
  p->first = reinterpret_cast<int*>(c); c += a_i;
  p->hint = reinterpret_cast<uint2*>(c); c += a_l;
  p->heavy = reinterpret_cast<int*>(c); c += a_h;
  p->heavy_done = reinterpret_cast<unsigned*>(c); c += a_h;
  p->eflag = reinterpret_cast<unsigned char*>(c); c += a_f;
  p->num = reinterpret_cast<int*>(c); c += 256;
  p->heavy_n = reinterpret_cast<int*>(c);
  p->work = reinterpret_cast<unsigned*>(c + 128);
  plan_forget_hints(p, 0);
  return p;
}
void plan_delete(Plan* p) { delete p; }
int64_t plan_capacity(const Plan* p) { return p->cap; }

PlanView plan_view(const Plan* p) {
  PlanView v;
  v.uniq = p->uniq; v.idx = p->idx; v.counts = p->counts; v.num = p->num;
  v.seg_off = p->seg_off; v.pos = p->pos; v.heavy = p->heavy; v.heavy_n = p->heavy_n;
  v.first = p->first; v.hint = p->hint; v.heavy_sum = p->heavy_sum; v.heavy_done = p->heavy_done;
  v.work = p->work; v.eflag = p->eflag; v.staged = p->staged;
  v.staged_units = (p->cap + 3) / 4 + 1;
  v.heavy_t = p->heavy_t; v.heavy_cap = p->heavy_cap; v.sum_dim = p->sum_dim; v.n = p->n;
  return v;
}

// The heavy path of the fused apply parks partial gradient sums here; sized on first use
// (not during CUDA-graph capture: run the step once eagerly first).
int plan_need_scratch(Plan* p, int dim, cudaStream_t st) {
  const int want = (dim + 31) / 32 * 32;
  if (p->heavy_sum && p->sum_dim >= want) return 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) cudaGetLastError();
  if (cs != cudaStreamCaptureStatusNone)
    return fail(2, "plan: scratch must be sized before CUDA-graph capture (run once eagerly)");
  KV_CUDA(cudaStreamSynchronize(st));
  if (p->heavy_sum) cudaFree(p->heavy_sum);
  if (p->staged) cudaFree(p->staged);
  p->heavy_sum = p->staged = nullptr;
  KV_CUDA(cudaMalloc(&p->heavy_sum, (size_t)p->heavy_cap * want * sizeof(float)));
  KV_CUDA(cudaMalloc(&p->staged, (size_t)((p->cap + 3) / 4 + 1) * 4 * want * sizeof(float)));
  p->sum_dim = want;
  return 0;
}

int do_plan_build(Plan* p, Workspace* ws, const int64_t* ids, int64_t n, cudaStream_t st) {
  if (n < 0 || n > p->cap) return fail(1, "plan_build: n exceeds the plan's capacity");
  p->n = n;
  PlanOut po;
  po.first = p->first; po.hint = p->hint; po.seg_off = p->seg_off; po.within = p->within; po.pos_u = p->pos_u;
  po.heavy = p->heavy; po.heavy_n = p->heavy_n; po.heavy_t = p->heavy_t; po.heavy_cap = p->heavy_cap;
  if (n == 0) {
    KV_CUDA(cudaMemsetAsync(p->num, 0, sizeof(int), st));
    KV_CUDA(cudaMemsetAsync(p->heavy_n, 0, sizeof(int), st));
    return 0;
  }
  KV_TRY(do_unique_impl(ws, ids, n, reinterpret_cast<int64_t*>(p->uniq), p->idx, p->counts, p->num,
                        nullptr, st, &po));
  const int dev = p->device;
  KV_CUDA(cudaMemsetAsync(p->eflag, 0, (size_t)((n + 3) & ~3LL), st));  // read four at a time
  // heavy segments: a block each; light segments: one warp per 32 distinct ids (U <= n)
  const int sms = sm_count(dev);
  int hb = p->heavy_cap < sms ? p->heavy_cap : sms;
  long long lb = ((n + 31) / 32 * 32 + 511) / 512;
  if (lb > 3LL * sms) lb = 3LL * sms;
  if (lb < 1) lb = 1;
  plan_sort_kernel<<<(unsigned)(hb + lb), 512, 0, st>>>(p->counts, p->seg_off, p->num, p->heavy,
                                                       p->heavy_n, p->heavy_cap, p->heavy_t, p->pos_u,
                                                       p->pos, p->eflag, n, hb);
  KV_LAUNCHED();
  return 0;
}

// ---- embedding_lookup_sparse combiner ---------------------------------------------------------
// python/ops/embedding_ops.py:403-441: after unique -> gather, the rows of the distinct ids are
// expanded through the inverse index, optionally weighted, and reduced per SparseTensor row:
// sparse_segment_sum / mean / sqrt_n (no weights) or segment_sum of weighted rows divided by
// segment_sum(w) / sqrt(segment_sum(w^2)).  One warp per output row: the row's entries are a
// contiguous run of `seg` (SparseTensor indices are row-major sorted), found by binary search,
// and added in increasing position - the order of TF's CPU kernels - so the result does not
// depend on scheduling.  Rows without entries are zero.
__global__ void __launch_bounds__(256)
sparse_combine_kernel(const float* __restrict__ emb, const int* __restrict__ idx,
                      const long long* __restrict__ seg, const float* __restrict__ w,
                      long long nnz, long long n_rows, int dim, int combiner,
                      float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (; r < n_rows; r += nw) {
    long long lo = 0, hi = nnz;          // first entry with seg >= r
    while (lo < hi) { const long long m = (lo + hi) >> 1; if (__ldg(seg + m) < r) lo = m + 1; else hi = m; }
    const long long first = lo;
    hi = nnz;                            // first entry with seg > r
    while (lo < hi) { const long long m = (lo + hi) >> 1; if (__ldg(seg + m) <= r) lo = m + 1; else hi = m; }
    const long long last = lo;
    float den = 0.f;
    for (long long i = first; i < last; ++i) {
      const float wi = w ? __ldg(w + i) : 1.0f;
      den += combiner == 2 ? wi * wi : wi;
    }
    if (combiner == 2) den = sqrtf(den);
    for (int c0 = 0; c0 < dim; c0 += 32 * 8) {
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      for (long long i = first; i < last; ++i) {
        const float* row = emb + (long long)__ldg(idx + i) * dim;
        const float wi = w ? __ldg(w + i) : 1.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = c0 + q * 32 + lane;
          if (c < dim) acc[q] += w ? row[c] * wi : row[c];
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = c0 + q * 32 + lane;
        if (c < dim) {
          float v = acc[q];
          if (combiner != 0) v = last > first ? v / den : 0.f;
          out[r * dim + c] = v;
        }
      }
    }
  }
}

int do_sparse_combine(const float* emb, const int32_t* idx, const int64_t* seg, const float* w,
                      int64_t nnz, int64_t n_rows, int dim, int combiner, float* out,
                      cudaStream_t st) {
  if (n_rows <= 0) return 0;
  if (combiner < 0 || combiner > 2) return fail(1, "combiner must be one of 'mean', 'sqrtn' or 'sum'");
  int dev = 0;
  KV_CUDA(cudaGetDevice(&dev));
  sparse_combine_kernel<<<blocks_for(n_rows * 32, 256, dev), 256, 0, st>>>(
      emb, idx, reinterpret_cast<const long long*>(seg), w, nnz, n_rows, dim, combiner, out);
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
