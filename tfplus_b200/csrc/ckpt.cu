// ckpt.cu — whole-table scans: size / frequency gauges, export, import,
// delete and timestamp eviction, all directly on the device table.
#include "table.h"

namespace kvhbm {

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ bool live(const Slot& s) {
  return s.key != KEY_EMPTY && s.key != KEY_TOMB && (s.ctl & CTL_READY);
}

__device__ __forceinline__ void warp_add(unsigned long long* dst, unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, v);
}

// KvVariable::size_unsafe / sum_freq_unsafe (kv_variable.h:144-175) and
// table_->size() in one scan: scratch[0] = size, [1] = sum_freq, [2] = map size.
__global__ void stats_kernel(TableView t) {
  unsigned long long sz = 0, fr = 0, all = 0;
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i <= t.mask; i += stride) {
    Slot s = load_slot(t.slots + i);
    if (!live(s)) continue;
    ++all;
    if (!(s.ctl & CTL_BLACK) && freq_count(s.freq) >= t.enter_threshold) {
      ++sz;
      fr += freq_count(s.freq);
    }
  }
  warp_add(&t.ctr->scratch[0], sz);
  warp_add(&t.ctr->scratch[1], fr);
  warp_add(&t.ctr->scratch[2], all);
}

// RefreshAllUnderThresholds, kv_variable.h:995-1012 with UpdateUnderThreshold
// (:837-861): one warp per slot row.
__global__ void refresh_under_kernel(TableView t, int enable_cutoff, float cutoff) {
  const int lane = threadIdx.x & 31;
  unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (; w <= t.mask; w += nw) {
    Slot s = load_slot(t.slots + w);
    if (!live(s)) continue;
    bool under;
    if (s.ctl & CTL_BLACK) under = true;
    else if (!enable_cutoff) under = false;
    else {
      const float* r = row_ptr(t, s.ctl);
      bool big = false;
      for (int j = lane; j < t.dim; j += 32) big |= fabsf(__ldcg(r + j)) >= cutoff;
      under = __ballot_sync(FULL, big) == 0;
    }
    if (lane == 0) {
      const uint32_t n = under ? (s.ctl | CTL_UNDER) : (s.ctl & ~CTL_UNDER);
      if (n != s.ctl) t.slots[w].ctl = n;
    }
  }
}

// Classification of ExportValues, dynamic_save.hpp:71-82,142-174.
__device__ __forceinline__ int export_class(const TableView& t, const Slot& s, int first_n) {
  if (s.ctl & CTL_BLACK) return 1;  // blacklist
  if ((first_n <= 3 || freq_count(s.freq) >= t.enter_threshold) && !(s.ctl & CTL_UNDER)) return 0;
  return 2;  // neither
}

// scratch[0] = keys/values rows, [1] = blacklist, [2] = all keys (freq table)
__global__ void export_count_kernel(TableView t, int first_n) {
  unsigned long long nk = 0, nb = 0, all = 0;
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i <= t.mask; i += stride) {
    Slot s = load_slot(t.slots + i);
    if (!live(s)) continue;
    ++all;
    const int c = export_class(t, s, first_n);
    nk += c == 0;
    nb += c == 1;
  }
  warp_add(&t.ctr->scratch[0], nk);
  warp_add(&t.ctr->scratch[1], nb);
  warp_add(&t.ctr->scratch[2], all);
}

// One warp per slot: lane 0 reserves the output positions, the warp copies the row.
__global__ void export_kernel(TableView t, int first_n, long long* keys, float* values,
                              long long* blacklist, long long* freq_keys, void* freq_values,
                              int freq_u32) {
  const int lane = threadIdx.x & 31;
  unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (; w <= t.mask; w += nw) {
    Slot s = load_slot(t.slots + w);
    if (!live(s)) continue;
    const int c = export_class(t, s, first_n);
    unsigned long long p = 0;
    if (lane == 0) {
      if (c == 0 && keys) { p = atomicAdd(&t.ctr->scratch[0], 1ULL); keys[p] = s.key; }
      if (c == 1 && blacklist) blacklist[atomicAdd(&t.ctr->scratch[1], 1ULL)] = s.key;
      if (freq_keys) {
        const unsigned long long q = atomicAdd(&t.ctr->scratch[2], 1ULL);
        freq_keys[q] = s.key;
        if (freq_u32) static_cast<uint32_t*>(freq_values)[q] = freq_to_ref(s.freq);
        else static_cast<uint16_t*>(freq_values)[q] = (uint16_t)freq_count(s.freq);
      }
    }
    if (c == 0 && values) {
      p = __shfl_sync(FULL, p, 0);
      const float* r = row_ptr(t, s.ctl);
      float* o = values + p * (unsigned long long)t.dim;
      for (int j = lane; j < t.dim; j += 32) o[j] = __ldcg(r + j);
    }
  }
}

// ImportValues stage 1, dynamic_restore.hpp:177-196: one warp per key.
__global__ void import_rows_kernel(TableView t, const long long* __restrict__ keys,
                                   const float* __restrict__ values, long long n) {
  const int lane = threadIdx.x & 31;
  long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (; w < n; w += nw) {
    long long pos = -1;
    uint32_t row = 0;
    bool claimed = false;
    if (lane == 0) {
      Slot s;
      pos = find_or_claim(t, keys[w], &s, &claimed);
      if (pos >= 0) row = claimed ? alloc_row(t) : (s.ctl & CTL_ROW_MASK);
    }
    pos = shfl_ll(pos, 0);
    row = __shfl_sync(FULL, row, 0);
    if (pos < 0) continue;
    float* r = t.rows + (size_t)row * t.row_stride;
    const float* v = values + w * (long long)t.dim;
    for (int j = lane; j < t.dim; j += 32) r[j] = v[j];
    __syncwarp();
    if (lane == 0) {
      t.slots[pos].freq = 1u << 16;  // ctor freq 1; no under-threshold refresh on import
      __threadfence();
      t.slots[pos].ctl = CTL_READY | row;
    }
  }
}

// stage 3, dynamic_restore.hpp:204-213 -> MarkBlacklistUnsafe(key, nullptr)
__global__ void import_blacklist_kernel(TableView t, const long long* __restrict__ keys, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    Slot s;
    bool claimed;
    const long long pos = find_or_claim(t, keys[i], &s, &claimed);
    if (pos < 0) continue;
    if (claimed) {
      const uint32_t row = alloc_row(t);
      t.slots[pos].freq = 1u << 16;
      __threadfence();
      t.slots[pos].ctl = CTL_READY | CTL_BLACK | row;  // under_threshold stays false
    } else if (!(s.ctl & CTL_BLACK)) {
      t.slots[pos].ctl = s.ctl | CTL_BLACK | CTL_UNDER;
    }
  }
}

// stage 4, dynamic_restore.hpp:217-245: only keys that exist
__global__ void import_freq_kernel(TableView t, const long long* __restrict__ keys,
                                   const void* __restrict__ vals, long long n, int freq_u32) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    Slot s;
    const long long pos = find_slot(t, keys[i], &s);
    if (pos < 0) continue;
    const uint32_t ref = freq_u32 ? static_cast<const uint32_t*>(vals)[i]
                                  : (uint32_t) static_cast<const uint16_t*>(vals)[i];
    t.slots[pos].freq = freq_to_ref(ref);  // the swap is its own inverse
  }
}

// TableManager::DeleteKey, table_manager.h:405-416
__global__ void delete_kernel(TableView t, const long long* __restrict__ keys, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    Slot s;
    const long long pos = find_slot(t, keys[i], &s);
    if (pos < 0) continue;
    // duplicates of one id in the request: only one thread wins the tombstone
    const unsigned long long old = atomicCAS(
        reinterpret_cast<unsigned long long*>(&t.slots[pos].key), (unsigned long long)keys[i],
        (unsigned long long)KEY_TOMB);
    if (old != (unsigned long long)keys[i]) continue;
    const long long top = (long long)atomicAdd(
        reinterpret_cast<unsigned long long*>(&t.ctr->free_top), 1ULL);
    t.free_rows[top] = s.ctl & CTL_ROW_MASK;
    atomicAdd(&t.ctr->tombstones, 1ULL);
  }
}

// KvVariable::DeleteWithTimestamp, kv_variable.h:756-789
__global__ void delete_older_kernel(TableView t, int threshold, int today, long long* out_keys,
                                    long long cap) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i <= t.mask; i += stride) {
    Slot s = load_slot(t.slots + i);
    if (!live(s)) continue;
    const int day = (int)freq_day(s.freq);
    if (!(day > 0 && today - day >= threshold)) continue;
    t.slots[i].key = KEY_TOMB;
    const long long top = (long long)atomicAdd(
        reinterpret_cast<unsigned long long*>(&t.ctr->free_top), 1ULL);
    t.free_rows[top] = s.ctl & CTL_ROW_MASK;
    atomicAdd(&t.ctr->tombstones, 1ULL);
    const unsigned long long q = atomicAdd(&t.ctr->scratch[0], 1ULL);
    if (out_keys && (long long)q < cap) out_keys[q] = s.key;
  }
}

int zero_scratch(Table* tb, cudaStream_t st) {
  KV_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(tb->d_ctr) + offsetof(Counters, scratch), 0,
                          sizeof(((Counters*)0)->scratch), st));
  return 0;
}

}  // namespace

int do_stats(Table* tb, cudaStream_t st, int64_t* size, int64_t* sum_freq, int64_t* map_size) {
  KV_TRY(zero_scratch(tb, st));
  stats_kernel<<<blocks_for(tb->capacity, 256, tb->device), 256, 0, st>>>(tb->view());
  KV_LAUNCHED();
  KV_TRY(tb->sync_counters(st));
  if (size) *size = (int64_t)tb->h_ctr->scratch[0];
  if (sum_freq) *sum_freq = (int64_t)tb->h_ctr->scratch[1];
  if (map_size) *map_size = (int64_t)tb->h_ctr->scratch[2];
  return 0;
}

int do_export_count(Table* tb, int first_n, int enable_cutoff, float cutoff, cudaStream_t st,
                    int64_t* n_keys, int64_t* n_black, int64_t* n_freq) {
  if ((enable_cutoff != 0) != true || cutoff != DEFAULT_CUTOFF) {
    refresh_under_kernel<<<blocks_for(tb->capacity * 32, 256, tb->device), 256, 0, st>>>(
        tb->view(), enable_cutoff, cutoff);
    KV_LAUNCHED();
  }
  KV_TRY(zero_scratch(tb, st));
  export_count_kernel<<<blocks_for(tb->capacity, 256, tb->device), 256, 0, st>>>(tb->view(), first_n);
  KV_LAUNCHED();
  KV_TRY(tb->sync_counters(st));
  *n_keys = (int64_t)tb->h_ctr->scratch[0];
  *n_black = first_n > 3 ? (int64_t)tb->h_ctr->scratch[1] : 0;
  *n_freq = first_n > 4 ? (int64_t)tb->h_ctr->scratch[2] : 0;
  return 0;
}

int do_export(Table* tb, int first_n, int64_t* keys, float* values, int64_t* blacklist,
              int64_t* freq_keys, void* freq_values, int freq_u32, cudaStream_t st) {
  KV_TRY(zero_scratch(tb, st));
  if (first_n <= 3) blacklist = nullptr;
  if (first_n <= 4) { freq_keys = nullptr; freq_values = nullptr; }
  if (freq_values == nullptr) freq_keys = nullptr;
  export_kernel<<<blocks_for(tb->capacity * 32, 256, tb->device), 256, 0, st>>>(
      tb->view(), first_n, reinterpret_cast<long long*>(keys), values,
      reinterpret_cast<long long*>(blacklist), reinterpret_cast<long long*>(freq_keys),
      freq_values, freq_u32);
  KV_LAUNCHED();
  return 0;
}

int do_set_init_table(Table* tb, const float* d_table, int64_t rows, cudaStream_t st, bool force);

int do_import(Table* tb, const int64_t* keys, const float* values, int64_t n,
              const float* init_table, int64_t init_rows, const int64_t* blacklist,
              int64_t n_black, const int64_t* freq_keys, const void* freq_values, int64_t n_freq,
              int freq_u32, cudaStream_t st) {
  KV_TRY(tb->clear(st));
  KV_TRY(tb->ensure(n + n_black, st));
  if (n > 0) {
    import_rows_kernel<<<blocks_for(n * 32, 256, tb->device), 256, 0, st>>>(
        tb->view(), reinterpret_cast<const long long*>(keys), values, n);
    KV_LAUNCHED();
  }
  if (init_table && init_rows > 0) KV_TRY(do_set_init_table(tb, init_table, init_rows, st, true));
  if (n_black > 0) {
    import_blacklist_kernel<<<blocks_for(n_black, 256, tb->device), 256, 0, st>>>(
        tb->view(), reinterpret_cast<const long long*>(blacklist), n_black);
    KV_LAUNCHED();
  }
  if (n_freq > 0) {
    import_freq_kernel<<<blocks_for(n_freq, 256, tb->device), 256, 0, st>>>(
        tb->view(), reinterpret_cast<const long long*>(freq_keys), freq_values, n_freq, freq_u32);
    KV_LAUNCHED();
  }
  tb->initialized = true;
  return 0;
}

int do_delete(Table* tb, const int64_t* ids, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  KV_TRY(tb->ensure_free_list((uint64_t)n, st));
  delete_kernel<<<blocks_for(n, 256, tb->device), 256, 0, st>>>(
      tb->view(), reinterpret_cast<const long long*>(ids), n);
  KV_LAUNCHED();
  return 0;
}

int do_delete_older(Table* tb, int threshold, uint16_t today, int64_t* out_keys, int64_t cap,
                    cudaStream_t st, int64_t* n_deleted) {
  // worst case every key goes
  KV_TRY(tb->sync_counters(st));
  KV_TRY(tb->ensure_free_list(tb->h_ctr->used, st));
  KV_TRY(zero_scratch(tb, st));
  delete_older_kernel<<<blocks_for(tb->capacity, 256, tb->device), 256, 0, st>>>(
      tb->view(), (int)(uint16_t)threshold, (int)today, reinterpret_cast<long long*>(out_keys),
      out_keys ? cap : 0);
  KV_LAUNCHED();
  KV_TRY(tb->sync_counters(st));
  if (n_deleted) *n_deleted = (int64_t)tb->h_ctr->scratch[0];
  return 0;
}

}  // namespace kvhbm
