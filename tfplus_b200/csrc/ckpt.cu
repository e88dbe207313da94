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
// cap_*: sizes of the caller's buffers (what kv_export_count returned).  The count and the export
// are two calls; should another thread have inserted in between, entries beyond a capacity are
// dropped, as the reference's `key_row < num_rows` guards do (dynamic_save.hpp:142-174).
__global__ void export_kernel(TableView t, int first_n, long long* keys, float* values,
                              long long* blacklist, long long* freq_keys, void* freq_values,
                              int freq_u32, unsigned long long cap_k, unsigned long long cap_b,
                              unsigned long long cap_f) {
  const int lane = threadIdx.x & 31;
  unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (; w <= t.mask; w += nw) {
    Slot s = load_slot(t.slots + w);
    if (!live(s)) continue;
    const int c = export_class(t, s, first_n);
    unsigned long long p = 0;
    if (lane == 0) {
      if (c == 0 && (keys || values)) {
        p = atomicAdd(&t.ctr->scratch[0], 1ULL);
        if (keys && p < cap_k) keys[p] = s.key;
      }
      if (c == 1 && blacklist) {
        const unsigned long long q = atomicAdd(&t.ctr->scratch[1], 1ULL);
        if (q < cap_b) blacklist[q] = s.key;
      }
      if (freq_keys) {
        const unsigned long long q = atomicAdd(&t.ctr->scratch[2], 1ULL);
        if (q < cap_f) {
          freq_keys[q] = s.key;
          if (freq_u32) static_cast<uint32_t*>(freq_values)[q] = freq_to_ref(s.freq);
          else static_cast<uint16_t*>(freq_values)[q] = (uint16_t)freq_count(s.freq);
        }
      }
    }
    p = __shfl_sync(FULL, p, 0);
    if (c == 0 && values && p < cap_k) {
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
                                    long long cap, TableView set, int has_set) {
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
    if (has_set) {   // kv_variable.h:772-774
      Slot d;
      bool claimed;
      const long long pos = find_or_claim(set, s.key, &d, &claimed);
      if (pos >= 0 && claimed) { set.slots[pos].freq = 1u << 16; set.slots[pos].ctl = CTL_READY; }
    }
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
              int64_t* freq_keys, void* freq_values, int freq_u32, cudaStream_t st,
              int64_t cap_k, int64_t cap_b, int64_t cap_f) {
  const unsigned long long NOCAP = ~0ULL;
  KV_TRY(zero_scratch(tb, st));
  if (first_n <= 3) blacklist = nullptr;
  if (first_n <= 4) { freq_keys = nullptr; freq_values = nullptr; }
  if (freq_values == nullptr) freq_keys = nullptr;
  export_kernel<<<blocks_for(tb->capacity * 32, 256, tb->device), 256, 0, st>>>(
      tb->view(), first_n, reinterpret_cast<long long*>(keys), values,
      reinterpret_cast<long long*>(blacklist), reinterpret_cast<long long*>(freq_keys),
      freq_values, freq_u32, cap_k < 0 ? NOCAP : (unsigned long long)cap_k,
      cap_b < 0 ? NOCAP : (unsigned long long)cap_b, cap_f < 0 ? NOCAP : (unsigned long long)cap_f);
  KV_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------
// Delta checkpoints: KvVariable::DeltaExport / DeltaImport (dynamic_save.hpp:197-449,
// dynamic_restore.hpp:28-153).  train_deltalist_ / prediction_deltalist_ (kv_variable.h:870-871)
// are two more device hash tables used as key SETS (dim 1, rows unused); every entry point
// that the reference marks inserts its keys there (capi.cu), the export walks the set and
// classifies each key against the main table.
// ---------------------------------------------------------------------------
// `if (NeedDeltaInfo()) train_deltalist_.insert(key)`.  filter != null: the apply ops only mark
// the ids they did not skip (MarkAsDeltaListElements of the non-filtered indices,
// training_ops.cc:7200-7201): a key that exists with a count under enter_threshold is left out.
__global__ void delta_mark_kernel(TableView set, const long long* __restrict__ ids, long long n_in,
                                  const int* __restrict__ d_n, TableView main, int filter) {
  long long n = n_in;
  if (d_n) { const long long dn = *d_n; if (dn < n) n = dn; }
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const long long key = ids[i];
    if (key_reserved(key)) continue;
    if (filter) {
      Slot m;
      if (find_slot(main, key, &m) >= 0 && freq_count(m.freq) < main.enter_threshold) continue;
    }
    Slot s;
    bool claimed;
    const long long pos = find_or_claim(set, key, &s, &claimed);
    if (pos >= 0 && claimed) {
      set.slots[pos].freq = 1u << 16;
      set.slots[pos].ctl = CTL_READY;   // a set member has no row
    }
  }
}

// src's keys into dst (training-mode export: train_deltalist_ -> prediction_deltalist_).
__global__ void delta_merge_kernel(TableView src, TableView dst) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i <= src.mask; i += stride) {
    Slot s = load_slot(src.slots + i);
    if (!live(s)) continue;
    Slot d;
    bool claimed;
    const long long pos = find_or_claim(dst, s.key, &d, &claimed);
    if (pos >= 0 && claimed) {
      dst.slots[pos].freq = 1u << 16;
      dst.slots[pos].ctl = CTL_READY;
    }
  }
}

// One warp per slot of the delta set.  Classes (dynamic_save.hpp:230-248,333-339): absent ->
// delete_keys; low frequency -> nothing; blacklisted -> blacklist (training mode) or delete_keys
// (first_n <= 3); else keys/values.  first_n > 4: every delta key goes to the frequency table
// with its full word (0 when absent).  write == 0 only counts.  Counters: main.ctr->scratch
// [0] rows [1] blacklist [2] freq [3] delete.  `skip`: keys also present there were handled by
// the pass over that set.
__global__ void delta_export_kernel(TableView set, TableView skip, int has_skip, TableView main,
                                    int first_n, int write, long long* keys, float* values,
                                    long long* blacklist, long long* freq_keys,
                                    uint32_t* freq_values, long long* delete_keys, long long cap_k,
                                    long long cap_b, long long cap_f, long long cap_d) {
  const int lane = threadIdx.x & 31;
  unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (; w <= set.mask; w += nw) {
    const Slot ds = load_slot(set.slots + w);
    if (!live(ds)) continue;
    const long long key = ds.key;
    int cls = -1;            // 0 rows, 1 blacklist, 3 delete, -1 nothing
    uint32_t fword = 0;
    uint32_t ctl = 0;
    if (lane == 0) {
      Slot sk;
      if (has_skip && find_slot(skip, key, &sk) >= 0) cls = -2;   // the other pass owns it
      else {
        Slot m;
        if (find_slot(main, key, &m) < 0) cls = 3;
        else {
          fword = freq_to_ref(m.freq);
          ctl = m.ctl;
          if (freq_count(m.freq) < main.enter_threshold) cls = -1;
          else if (m.ctl & CTL_BLACK) cls = first_n <= 3 ? 3 : 1;
          else cls = 0;
        }
      }
    }
    cls = __shfl_sync(FULL, cls, 0);
    if (cls == -2) continue;
    unsigned long long p = 0;
    if (lane == 0) {
      if (cls == 0) { p = atomicAdd(&main.ctr->scratch[0], 1ULL); if (write && (long long)p < cap_k) keys[p] = key; }
      if (cls == 1) { const unsigned long long q = atomicAdd(&main.ctr->scratch[1], 1ULL); if (write && (long long)q < cap_b) blacklist[q] = key; }
      if (cls == 3) { const unsigned long long q = atomicAdd(&main.ctr->scratch[3], 1ULL); if (write && (long long)q < cap_d) delete_keys[q] = key; }
      if (first_n > 4) {
        const unsigned long long q = atomicAdd(&main.ctr->scratch[2], 1ULL);
        if (write && (long long)q < cap_f) { freq_keys[q] = key; freq_values[q] = fword; }
      }
    }
    if (cls == 0 && write) {
      p = __shfl_sync(FULL, p, 0);
      ctl = __shfl_sync(FULL, ctl, 0);
      if ((long long)p < cap_k) {
        const float* r = row_ptr(main, ctl);
        float* o = values + p * (unsigned long long)main.dim;
        for (int j = lane; j < main.dim; j += 32) o[j] = __ldcg(r + j);
      }
    }
  }
}

// DeltaImport stage 1 (dynamic_restore.hpp:60-79): overwrite-or-insert the row, RemoveBlacklist,
// UpdateUnderThreshold; a new key starts at the constructor's frequency 1.  One warp per key.
__global__ void delta_import_rows_kernel(TableView t, const long long* __restrict__ keys,
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
    claimed = __shfl_sync(FULL, (int)claimed, 0) != 0;
    if (pos < 0) continue;
    float* r = t.rows + (size_t)row * t.row_stride;
    const float* v = values + w * (long long)t.dim;
    bool big = false;
    for (int j = lane; j < t.dim; j += 32) {
      const float x = v[j];
      r[j] = x;
      big |= fabsf(x) >= DEFAULT_CUTOFF;
    }
    big = __any_sync(FULL, big);
    if (lane == 0) {
      if (claimed) t.slots[pos].freq = 1u << 16;
      __threadfence();
      t.slots[pos].ctl = CTL_READY | (big ? 0u : CTL_UNDER) | row;
    }
  }
}

int do_delta_mark(Table* set, const int64_t* ids, int64_t n, const int32_t* d_n, Table* main,
                  bool filter, cudaStream_t st) {
  if (n <= 0) return 0;
  KV_TRY(set->ensure(n, st));
  delta_mark_kernel<<<blocks_for(n, 256, set->device), 256, 0, st>>>(
      set->view(), reinterpret_cast<const long long*>(ids), n, d_n, main->view(), filter ? 1 : 0);
  KV_LAUNCHED();
  return 0;
}

// pass == 0: count; pass == 1: write and update the delta lists as the reference does.
int do_delta_export(Table* tb, Table* train, Table* pred, bool support_pred, int first_n, int write,
                    int64_t* keys, float* values, int64_t* blacklist, int64_t* freq_keys,
                    uint32_t* freq_values, int64_t* delete_keys, int64_t cap_k, int64_t cap_b,
                    int64_t cap_f, int64_t cap_d, cudaStream_t st, int64_t* counts) {
  KV_TRY(zero_scratch(tb, st));
  auto pass = [&](Table* set, Table* skip) -> int {
    delta_export_kernel<<<blocks_for(set->capacity * 32, 256, tb->device), 256, 0, st>>>(
        set->view(), skip ? skip->view() : set->view(), skip != nullptr, tb->view(), first_n, write,
        reinterpret_cast<long long*>(keys), values, reinterpret_cast<long long*>(blacklist),
        reinterpret_cast<long long*>(freq_keys), freq_values,
        reinterpret_cast<long long*>(delete_keys), cap_k, cap_b, cap_f, cap_d);
    KV_LAUNCHED();
    return 0;
  };
  KV_TRY(pass(train, nullptr));
  if (first_n <= 3 && pred) KV_TRY(pass(pred, train));   // inference mode: train + prediction
  KV_TRY(tb->sync_counters(st));
  for (int i = 0; i < 4; ++i) counts[i] = (int64_t)tb->h_ctr->scratch[i];
  if (write) {
    if (first_n <= 3) {
      if (pred) KV_TRY(pred->clear(st));
    } else {
      if (support_pred && pred) {
        KV_TRY(train->sync_counters(st));
        KV_TRY(pred->ensure((int64_t)train->h_ctr->used, st));
        delta_merge_kernel<<<blocks_for(train->capacity, 256, tb->device), 256, 0, st>>>(
            train->view(), pred->view());
        KV_LAUNCHED();
      }
      KV_TRY(train->clear(st));
    }
  }
  return 0;
}

int do_delete(Table* tb, const int64_t* ids, int64_t n, cudaStream_t st);

int do_delta_import(Table* tb, int first_n, const int64_t* keys, const float* values, int64_t n,
                    const int64_t* blacklist, int64_t n_black, const int64_t* freq_keys,
                    const uint32_t* freq_values, int64_t n_freq, const int64_t* delete_keys,
                    int64_t n_delete, cudaStream_t st) {
  KV_TRY(tb->ensure(n + n_black, st));
  if (n > 0) {
    delta_import_rows_kernel<<<blocks_for(n * 32, 256, tb->device), 256, 0, st>>>(
        tb->view(), reinterpret_cast<const long long*>(keys), values, n);
    KV_LAUNCHED();
  }
  if (n_black > 0) {
    if (first_n > 3) {
      import_blacklist_kernel<<<blocks_for(n_black, 256, tb->device), 256, 0, st>>>(
          tb->view(), reinterpret_cast<const long long*>(blacklist), n_black);
      KV_LAUNCHED();
    } else {
      KV_TRY(do_delete(tb, blacklist, n_black, st));   // inference load: drop them (:104-110)
    }
  }
  if (n_freq > 0) {
    import_freq_kernel<<<blocks_for(n_freq, 256, tb->device), 256, 0, st>>>(
        tb->view(), reinterpret_cast<const long long*>(freq_keys), freq_values, n_freq, 1);
    KV_LAUNCHED();
  }
  if (n_delete > 0) KV_TRY(do_delete(tb, delete_keys, n_delete, st));
  tb->initialized = true;
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
                    cudaStream_t st, int64_t* n_deleted, Table* delta_set) {
  // worst case every key goes
  KV_TRY(tb->sync_counters(st));
  KV_TRY(tb->ensure_free_list(tb->h_ctr->used, st));
  if (delta_set) KV_TRY(delta_set->ensure((int64_t)tb->h_ctr->used, st));
  KV_TRY(zero_scratch(tb, st));
  delete_older_kernel<<<blocks_for(tb->capacity, 256, tb->device), 256, 0, st>>>(
      tb->view(), (int)(uint16_t)threshold, (int)today, reinterpret_cast<long long*>(out_keys),
      out_keys ? cap : 0, delta_set ? delta_set->view() : tb->view(), delta_set != nullptr);
  KV_LAUNCHED();
  KV_TRY(tb->sync_counters(st));
  if (n_deleted) *n_deleted = (int64_t)tb->h_ctr->scratch[0];
  return 0;
}

}  // namespace kvhbm
