// common.cuh — device-side layout of the HBM-resident KvVariable table and the
// helpers every kernel shares (hash, probe, row-tile geometry, initializer).
//
// Layout (DESIGN.md "Data layout in HBM"):
//   slots : Slot[capacity]    16 B each, open addressing, linear probing,
//                             capacity a power of two, load factor <= 0.5
//   rows  : float[n_rows][row_stride]   densely packed row arena (a CUDA
//                             virtual-memory reservation that grows in place)
// A slot is {int64 key, u32 freq, u32 ctl}:
//   freq = count << 16 | day   (the reference packs lo16 = count, hi16 = day,
//                               embedding_value.h:229-234; swapped here so a
//                               plain atomicAdd on the word can only carry out
//                               of the top, never into the day)
//   ctl  = READY | BLACK | UNDER | row index (29 bits)
#ifndef KVHBM_COMMON_CUH_
#define KVHBM_COMMON_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace kvhbm {

constexpr long long KEY_EMPTY = (long long)0x8000000000000000ULL;
constexpr long long KEY_TOMB = (long long)0x8000000000000001ULL;
// Also the padding id of the fixed-capacity shard exchange: lookups of it return zeros, applies
// skip it, nothing probes the table for it (kv_route_ids pads its send buffers with it).
constexpr long long KEY_PAD = KEY_TOMB;
// The two sentinel values cannot be stored: every kernel treats them as padding (lookups
// return zeros, updates skip them) instead of letting INT64_MIN match an empty slot.
__host__ __device__ __forceinline__ bool key_reserved(long long k) {
  return k == KEY_EMPTY || k == KEY_TOMB;
}

constexpr uint32_t CTL_READY = 0x80000000u;  // row contents are published
constexpr uint32_t CTL_BLACK = 0x40000000u;  // EmbeddingValue::in_black_
constexpr uint32_t CTL_UNDER = 0x20000000u;  // EmbeddingValue::under_threshold_
constexpr uint32_t CTL_ROW_MASK = 0x1FFFFFFFu;

constexpr float DEFAULT_CUTOFF = 1.0e-20f;  // kv_variable_interface.h:55

struct __align__(16) Slot {
  long long key;
  uint32_t freq;
  uint32_t ctl;
};

struct Counters {
  unsigned long long used;       // slots ever claimed since the last rehash (live + tombstones)
  unsigned long long rows_bump;  // rows handed out by the bump allocator
  long long free_top;            // entries on the free-row stack
  unsigned long long tombstones;
  unsigned long long scratch[4];  // per-call reduction results (size, sum_freq, export counts)
  unsigned int apply_done;        // blocks of the running apply launch that have finished
  unsigned int overflow;          // sticky: a kernel ran out of mapped rows (bit 0) or of probe
                                  // room (bit 1) - only reachable when captured work is replayed
                                  // past what kv_reserve made room for; reported by the host
};

// Entry of a kernel that may have been launched with programmatic stream serialization: wait for
// the previous kernel of the stream (completed and flushed), then let the next one be scheduled.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// What a kernel needs to know about one table; passed by value.
struct TableView {
  Slot* slots;
  unsigned long long mask;  // capacity - 1
  int shift;                // 64 - log2(capacity / 2): hash -> home bucket
  float* rows;
  int dim;
  int row_stride;  // floats, multiple of 4
  const float* init;
  long long init_rows;
  unsigned long long seed;
  uint32_t enter_threshold;
  Counters* ctr;
  uint32_t* free_rows;
  unsigned long long rows_cap;  // rows the arena has mapped: the bump allocator's bound
};

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}
constexpr unsigned long long GOLDEN = 0x9e3779b97f4a7c15ULL;

// ---- memory access helpers -------------------------------------------------
__device__ __forceinline__ Slot load_slot(const Slot* p) {
  // L2-coherent 128-bit load: slots change under concurrent inserts.
  int4 v = __ldcg(reinterpret_cast<const int4*>(p));
  Slot s;
  s.key = (long long)(((unsigned long long)(uint32_t)v.y << 32) | (uint32_t)v.x);
  s.freq = (uint32_t)v.z;
  s.ctl = (uint32_t)v.w;
  return s;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- probing ------------------------------------------------------------------
// Bucketised linear probing: a bucket is two adjacent slots = one 32-byte DRAM
// sector, fetched with two 128-bit loads that the memory system merges.  A key
// lives in the first free slot (in bucket order) at or after its home bucket,
// so at load <= 0.5 almost every lookup is a single dependent memory round trip.
__device__ __forceinline__ unsigned long long home_bucket(const TableView& t, long long key) {
  return mix64((unsigned long long)key) >> t.shift;  // shift = 64 - log2(capacity / 2)
}

// A resumable probe: advances until it finds `key` or the first EMPTY slot on its path.
struct Probe {
  unsigned long long bucket;
  int sub;  // next slot inside the bucket to examine
};
__device__ __forceinline__ Probe probe_begin(const TableView& t, long long key) {
  Probe p;
  p.bucket = home_bucket(t, key);
  p.sub = 0;
  return p;
}
// Examines one bucket (s0, s1 = its two slots).  Returns 1 found (*pos, *out set), 0 reached
// an EMPTY slot (*pos = that slot), -1 keep going (p advanced to the next bucket).
__device__ __forceinline__ int probe_step(const TableView& t, long long key, Probe* p,
                                          const Slot& s0, const Slot& s1, long long* pos,
                                          Slot* out) {
  if (p->sub == 0) {
    if (s0.key == key) { *pos = (long long)(p->bucket * 2); *out = s0; return 1; }
    if (s0.key == KEY_EMPTY) { *pos = (long long)(p->bucket * 2); return 0; }
  }
  if (s1.key == key) { *pos = (long long)(p->bucket * 2 + 1); *out = s1; return 1; }
  if (s1.key == KEY_EMPTY) { *pos = (long long)(p->bucket * 2 + 1); p->sub = 1; return 0; }
  p->bucket = (p->bucket + 1) & (t.mask >> 1);
  p->sub = 0;
  return -1;
}
// After a failed claim of the EMPTY slot `pos` (another key took it): step past it.
__device__ __forceinline__ void probe_skip(const TableView& t, Probe* p, long long pos) {
  if (pos & 1) { p->bucket = (p->bucket + 1) & (t.mask >> 1); p->sub = 0; }
  else p->sub = 1;
}

// Read-only lookup.  Returns the slot index or -1.
__device__ __forceinline__ long long find_slot(const TableView& t, long long key, Slot* out) {
  Probe p = probe_begin(t, key);
  for (unsigned long long n = 0; n <= (t.mask >> 1); ++n) {
    const Slot s0 = load_slot(t.slots + p.bucket * 2);
    const Slot s1 = load_slot(t.slots + p.bucket * 2 + 1);
    long long pos;
    const int r = probe_step(t, key, &p, s0, s1, &pos, out);
    if (r == 1) return pos;
    if (r == 0) return -1;
  }
  return -1;
}

// Try to claim the empty slot `pos` for `key`.  1 = claimed, 2 = the same key got there first
// (found, possibly not yet published), 0 = another key took it: keep probing.
__device__ __forceinline__ int claim_slot(const TableView& t, long long key, long long pos) {
  const unsigned long long old = atomicCAS(
      reinterpret_cast<unsigned long long*>(&t.slots[pos].key), (unsigned long long)KEY_EMPTY,
      (unsigned long long)key);
  if (old == (unsigned long long)KEY_EMPTY) {
    atomicAdd(&t.ctr->used, 1ULL);
    return 1;
  }
  return old == (unsigned long long)key ? 2 : 0;
}

// Find-or-claim.  Returns the slot index; *claimed is true when this thread
// won the atomicCAS on an empty slot (it must then allocate a row, fill it and
// publish ctl).  A thread that loses the race to the same key gets
// claimed=false and may see ctl without CTL_READY.
__device__ __forceinline__ long long find_or_claim(const TableView& t, long long key,
                                                   Slot* out, bool* claimed) {
  Probe p = probe_begin(t, key);
  *claimed = false;
  for (unsigned long long n = 0; n <= t.mask + 2; ++n) {
    const Slot s0 = load_slot(t.slots + p.bucket * 2);
    const Slot s1 = load_slot(t.slots + p.bucket * 2 + 1);
    long long pos;
    const int r = probe_step(t, key, &p, s0, s1, &pos, out);
    if (r == 1) return pos;
    if (r == 0) {
      const int c = claim_slot(t, key, pos);
      if (c) {
        *claimed = c == 1;
        out->key = key; out->freq = 0; out->ctl = 0;  // a loser re-reads ctl
        return pos;
      }
      probe_skip(t, &p, pos);
    }
  }
  atomicOr(&t.ctr->overflow, 2u);
  return -1;  // table full: only reachable past the host's reservation (see Counters::overflow)
}

// Row allocator: pop the free stack (filled only by delete kernels, which never
// run concurrently with this), else bump.
__device__ __forceinline__ uint32_t alloc_row(const TableView& t) {
  if (t.free_rows != nullptr) {
    long long top = atomicAdd(reinterpret_cast<unsigned long long*>(&t.ctr->free_top),
                              (unsigned long long)(-1LL));
    if (top > 0) return t.free_rows[top - 1];
    atomicAdd(reinterpret_cast<unsigned long long*>(&t.ctr->free_top), 1ULL);
  }
  const unsigned long long r = atomicAdd(&t.ctr->rows_bump, 1ULL);
  if (r >= t.rows_cap) {
    // CUDA-graph replays inserted more keys than the host reserved for (the host books nothing
    // during replay).  Never touch unmapped memory: park the key on the last mapped row and
    // raise the sticky flag - the next host-side call on the table fails loudly.
    atomicOr(&t.ctr->overflow, 1u);
    return (uint32_t)(t.rows_cap - 1);
  }
  return (uint32_t)r;
}

__device__ __forceinline__ float* row_ptr(const TableView& t, uint32_t ctl) {
  return t.rows + (size_t)(ctl & CTL_ROW_MASK) * (size_t)t.row_stride;
}

// ---- frequency word --------------------------------------------------------------
// freq = count << 16 | day.  Saturating add of `cnt` (<= 65535) to the count and
// day := today, using only native atomics (no CAS loop, so thousands of
// duplicates of one hot key do not serialise on retries):
//  * atomicAdd(cnt << 16): a carry leaves the word at the top, the day is safe;
//  * whoever sees old_count + cnt > 65535 pins the count to 0xFFFF with an
//    atomicOr.  Every add after the last such OR would itself overflow and OR
//    again, so once any add overflowed the final count is 0xFFFF, and
//    otherwise it is the exact sum: min(65535, sum) as utility.h:65-70;
//  * the day changes once per key per day: only then AND it out and OR it in
//    (all writers of one launch write the same `today`).
// Second half of add_frequency, given the value the atomicAdd returned.
__device__ __forceinline__ void finish_frequency(uint32_t* freq, uint32_t old, uint32_t cnt,
                                                 uint32_t today) {
  if ((old >> 16) + cnt > 0xFFFFu) atomicOr(freq, 0xFFFF0000u);
  if ((old & 0xFFFFu) != today) {
    atomicAnd(freq, 0xFFFF0000u);
    atomicOr(freq, today);
  }
}
__device__ __forceinline__ void add_frequency(uint32_t* freq, uint32_t cnt, uint32_t today) {
  finish_frequency(freq, atomicAdd(freq, cnt << 16), cnt, today);
}
// SaturateMaxFrequency, utility.h:47-49 (including its wrap of negative counts).
__device__ __forceinline__ uint32_t saturate_count(int c) {
  return (uint32_t)(uint16_t)(c < 65535 ? c : 65535);
}
__host__ __device__ __forceinline__ uint32_t freq_count(uint32_t w) { return w >> 16; }
__host__ __device__ __forceinline__ uint32_t freq_day(uint32_t w) { return w & 0xFFFFu; }
// reference packing: lo16 = count, hi16 = day
__host__ __device__ __forceinline__ uint32_t freq_to_ref(uint32_t w) { return (w << 16) | (w >> 16); }

// ---- row tiles ----------------------------------------------------------------------
// A hint only: pull the line holding `p` into L2 (no register, no wait).
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// A row of `dim` floats is moved by a tile of `tpr` lanes (power of two <= 32),
// each lane owning `CPL` chunks of VEC floats: chunk c of lane l covers floats
// [(c * tpr + l) * VEC, +VEC).  VEC = 4 (128-bit) when dim % 4 == 0, else 1.
struct RowGeom {
  int vec;  // 4 or 1
  int tpr;  // lanes per row
  int cpl;  // chunks per lane
};
inline RowGeom row_geom(int dim) {
  RowGeom g;
  g.vec = (dim % 4 == 0) ? 4 : 1;
  int nvec = dim / g.vec;
  int tpr = 1;
  while (tpr < nvec && tpr < 32) tpr <<= 1;
  g.tpr = tpr;
  g.cpl = (nvec + tpr - 1) / tpr;
  return g;
}

template <int VEC> struct Chunk;
template <> struct Chunk<4> {
  float v[4];
  __device__ __forceinline__ void load_cg(const float* p) {
    float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void load_nc(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void load_plain(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  // weak load that does not allocate in L1: rows are never L1-resident, so a row written by
  // another SM earlier in the same launch can never be read stale
  __device__ __forceinline__ void load_na(const float* p) {
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
  }
  __device__ __forceinline__ void load_stream(const float* p) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ __forceinline__ void store_stream(float* p) const {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <> struct Chunk<1> {
  float v[1];
  __device__ __forceinline__ void load_cg(const float* p) { v[0] = __ldcg(p); }
  __device__ __forceinline__ void load_nc(const float* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void load_plain(const float* p) { v[0] = *p; }
  __device__ __forceinline__ void load_na(const float* p) {
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
  }
  __device__ __forceinline__ void load_stream(const float* p) { v[0] = __ldcs(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
  __device__ __forceinline__ void store_stream(float* p) const { __stcs(p, v[0]); }
};
template <int VEC>
__device__ __forceinline__ void chunk_zero(Chunk<VEC>& c) {
#pragma unroll
  for (int k = 0; k < VEC; ++k) c.v[k] = 0.f;
}

// Deterministic replacement of GenerateRandomInitialValue (kv_variable.h:889-898):
// r = mix64(key ^ seed [^ GOLDEN]) % R instead of std::rand() % R; the
// arithmetic (t[r1] + t[r2]) * 0.5f is the reference's.
__device__ __forceinline__ void init_rows_of(const TableView& t, long long key,
                                             long long* r1, long long* r2) {
  if (t.init_rows <= 0) { *r1 = -1; *r2 = -1; return; }
  unsigned long long k = (unsigned long long)key ^ t.seed;
  *r1 = (long long)(mix64(k) % (unsigned long long)t.init_rows);
  *r2 = (long long)(mix64(k ^ GOLDEN) % (unsigned long long)t.init_rows);
}
template <int VEC>
__device__ __forceinline__ void init_chunk(const TableView& t, long long r1, long long r2,
                                           int off, Chunk<VEC>& c) {
  if (r1 < 0) { chunk_zero(c); return; }
  Chunk<VEC> a, b;
  a.load_nc(t.init + r1 * t.dim + off);
  b.load_nc(t.init + r2 * t.dim + off);
#pragma unroll
  for (int k = 0; k < VEC; ++k) c.v[k] = (a.v[k] + b.v[k]) * 0.5f;
}
template <int VEC>
__device__ __forceinline__ bool chunk_over_cutoff(const Chunk<VEC>& c, float cutoff) {
  bool big = false;
#pragma unroll
  for (int k = 0; k < VEC; ++k) big |= (fabsf(c.v[k]) >= cutoff);
  return big;
}

__device__ __forceinline__ long long shfl_ll(long long v, int src) {
  int lo = __shfl_sync(0xffffffffu, (int)(v & 0xffffffffLL), src);
  int hi = __shfl_sync(0xffffffffu, (int)(v >> 32), src);
  return ((long long)hi << 32) | (unsigned int)lo;
}
template <typename T>
__device__ __forceinline__ T* shfl_ptr(T* p, int src) {
  return reinterpret_cast<T*>(shfl_ll((long long)p, src));
}

}  // namespace kvhbm
#endif  // KVHBM_COMMON_CUH_
