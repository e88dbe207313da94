// table.cu — host-side table management: allocation, growth, rehash.
#include <cstdlib>

#include "table.h"

#include <atomic>
#include <cstdio>
#include <cstring>

namespace kvhbm {

// ---------------------------------------------------------------------------
// errors / bookkeeping
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }
int fail(int code, const std::string& msg) {
  set_error(msg);
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  std::string m = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
  set_error(m);
  // leave the sticky error state clean for the next call where possible
  cudaGetLastError();
  return e == cudaErrorMemoryAllocation ? 4 : 5;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  if (device >= 0 && device < 64) cached[device] = n;
  return n;
}

// ---------------------------------------------------------------------------
// GrowableArena
// ---------------------------------------------------------------------------
namespace {
struct DriverApi {
  CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*memAddressFree)(CUdeviceptr, size_t);
  CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                        unsigned long long);
  CUresult (*memRelease)(CUmemGenericAllocationHandle);
  CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle,
                     unsigned long long);
  CUresult (*memUnmap)(CUdeviceptr, size_t);
  CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*,
                                          CUmemAllocationGranularity_flags);
  bool ok = false;
};
template <typename F>
bool load_sym(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess ||
      st != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}
const DriverApi& driver() {
  static DriverApi api = [] {
    DriverApi a{};
    a.ok = load_sym("cuMemAddressReserve", &a.memAddressReserve) &&
           load_sym("cuMemAddressFree", &a.memAddressFree) &&
           load_sym("cuMemCreate", &a.memCreate) &&
           load_sym("cuMemRelease", &a.memRelease) &&
           load_sym("cuMemMap", &a.memMap) &&
           load_sym("cuMemUnmap", &a.memUnmap) &&
           load_sym("cuMemSetAccess", &a.memSetAccess) &&
           load_sym("cuMemGetAllocationGranularity", &a.memGetAllocationGranularity);
    return a;
  }();
  return api;
}
CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp p;
  std::memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}
}  // namespace

int GrowableArena::init(int device, size_t reserve_bytes) {
  device_ = device;
  const DriverApi& d = driver();
  vmm_ = false;
  if (d.ok && getenv("KVHBM_NO_VMM") == nullptr) {
    int supported = 0;
    cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, device);
    CUmemAllocationProp p = alloc_prop(device);
    size_t gran = 0;
    if (d.memGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) ==
            CUDA_SUCCESS && gran > 0) {
      gran_ = gran;
      reserved_ = (reserve_bytes + gran - 1) / gran * gran;
      if (d.memAddressReserve(&base_, reserved_, 0, 0, 0) == CUDA_SUCCESS) vmm_ = true;
    }
  }
  if (!vmm_) { base_ = 0; reserved_ = 0; gran_ = 2u << 20; }
  mapped_ = 0;
  return 0;
}

int GrowableArena::ensure(size_t bytes, cudaStream_t stream) {
  if (bytes <= mapped_) return 0;
  if (vmm_) {
    const DriverApi& d = driver();
    // grow by at least 25 % / 64 MiB so that mapping stays rare
    size_t want = bytes - mapped_;
    size_t min_step = mapped_ / 4 > (64u << 20) ? mapped_ / 4 : (64u << 20);
    if (want < min_step) want = min_step;
    want = (want + gran_ - 1) / gran_ * gran_;
    if (mapped_ + want > reserved_) {
      want = reserved_ - mapped_;
      if (mapped_ + want < bytes)
        return fail(4, "row arena: virtual reservation exhausted");
    }
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (want > free_b) {  // fall back to the exact need before giving up
      size_t exact = (bytes - mapped_ + gran_ - 1) / gran_ * gran_;
      if (exact > free_b) return fail(4, "row arena: out of device memory");
      want = exact;
    }
    CUmemAllocationProp p = alloc_prop(device_);
    CUmemGenericAllocationHandle h;
    CUresult r = d.memCreate(&h, want, &p, 0);
    if (r != CUDA_SUCCESS) return fail(4, "row arena: cuMemCreate failed (out of device memory?)");
    r = d.memMap(base_ + mapped_, want, 0, h, 0);
    if (r != CUDA_SUCCESS) { d.memRelease(h); return fail(5, "row arena: cuMemMap failed"); }
    CUmemAccessDesc acc;
    std::memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = d.memSetAccess(base_ + mapped_, want, &acc, 1);
    if (r != CUDA_SUCCESS) return fail(5, "row arena: cuMemSetAccess failed");
    handles_.push_back(h);
    handle_bytes_.push_back(want);
    mapped_ += want;
    return 0;
  }
  // Fallback: cudaMalloc + copy.
  size_t want = bytes > mapped_ * 2 ? bytes : mapped_ * 2;
  want = (want + gran_ - 1) / gran_ * gran_;
  void* n = nullptr;
  cudaError_t e = cudaMalloc(&n, want);
  if (e != cudaSuccess && want > bytes) {
    cudaGetLastError();
    want = (bytes + gran_ - 1) / gran_ * gran_;
    e = cudaMalloc(&n, want);
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(row arena)");
  if (base_ != 0) {
    KV_CUDA(cudaMemcpyAsync(n, reinterpret_cast<void*>(base_), mapped_,
                            cudaMemcpyDeviceToDevice, stream));
    KV_CUDA(cudaStreamSynchronize(stream));
    cudaFree(reinterpret_cast<void*>(base_));
  }
  base_ = reinterpret_cast<CUdeviceptr>(n);
  mapped_ = want;
  return 0;
}

GrowableArena::~GrowableArena() {
  if (vmm_) {
    const DriverApi& d = driver();
    size_t off = 0;
    for (size_t i = 0; i < handles_.size(); ++i) {
      d.memUnmap(base_ + off, handle_bytes_[i]);
      d.memRelease(handles_[i]);
      off += handle_bytes_[i];
    }
    if (base_) d.memAddressFree(base_, reserved_);
  } else if (base_) {
    cudaFree(reinterpret_cast<void*>(base_));
  }
}

// ---------------------------------------------------------------------------
// kernels used by table management
// ---------------------------------------------------------------------------
__global__ void fill_empty_kernel(Slot* slots, unsigned long long n) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  int4 e;
  e.x = 0; e.y = (int)0x80000000u; e.z = 0; e.w = 0;  // key = INT64_MIN, freq 0, ctl 0
  for (; i < n; i += stride) reinterpret_cast<int4*>(slots)[i] = e;
}

__global__ void rehash_kernel(const Slot* old_slots, unsigned long long old_cap, TableView nt) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i < old_cap; i += stride) {
    Slot s = load_slot(old_slots + i);
    if (s.key == KEY_EMPTY || s.key == KEY_TOMB) continue;
    unsigned long long pos = home_bucket(nt, s.key) * 2;
    for (;;) {
      unsigned long long old = atomicCAS(
          reinterpret_cast<unsigned long long*>(&nt.slots[pos].key),
          (unsigned long long)KEY_EMPTY, (unsigned long long)s.key);
      if (old == (unsigned long long)KEY_EMPTY) {
        nt.slots[pos].freq = s.freq;
        nt.slots[pos].ctl = s.ctl;
        break;
      }
      pos = (pos + 1) & nt.mask;
    }
  }
}

// ---------------------------------------------------------------------------
// Table
// ---------------------------------------------------------------------------
static uint64_t pow2_at_least(uint64_t x) {
  uint64_t p = 1024;
  while (p < x) p <<= 1;
  return p;
}

void plan_delete(Plan*);
void workspace_delete(Workspace*);

bool pdl_enabled() {
  static const bool on = !(getenv("KVHBM_PDL") && atoi(getenv("KVHBM_PDL")) == 0);
  return on;
}

Table::~Table() {
  cudaSetDevice(device);
  if (scatter_plan) plan_delete(scatter_plan);
  if (scatter_ws) workspace_delete(scatter_ws);
  if (d_slots) cudaFree(d_slots);
  if (d_init) cudaFree(d_init);
  if (d_ctr) cudaFree(d_ctr);
  if (h_ctr) cudaFreeHost(h_ctr);
  if (d_free) cudaFree(d_free);
}

int Table::create(int dim_, int thr, int64_t capacity_hint) {
  if (dim_ <= 0) return fail(1, "KvVariable: embedding dim must be positive");
  RowGeom g = row_geom(dim_);
  if (g.cpl > 8)
    return fail(3, "KvVariable: embedding dim " + std::to_string(dim_) +
                       " not supported (max 1024 when a multiple of 4, else 256)");
  KV_CUDA(cudaGetDevice(&device));
  dim = dim_;
  row_stride = (dim + 3) / 4 * 4;
  // SaturateMaxFrequency(enter_threshold), kv_variable.h:99
  enter_threshold = (uint32_t)(uint16_t)(thr < 65535 ? thr : 65535);
  KV_CUDA(cudaMalloc(&d_ctr, sizeof(Counters)));
  KV_CUDA(cudaMemset(d_ctr, 0, sizeof(Counters)));
  KV_CUDA(cudaMallocHost(&h_ctr, sizeof(Counters)));
  std::memset(h_ctr, 0, sizeof(Counters));
  size_t total = 0, free_b = 0;
  KV_CUDA(cudaMemGetInfo(&free_b, &total));
  // reserve address space for the whole GPU; physical memory is mapped lazily
  KV_TRY(arena.init(device, total));
  uint64_t keys = capacity_hint > 0 ? (uint64_t)capacity_hint : 16384;
  capacity = pow2_at_least(keys * 2);
  KV_CUDA(cudaMalloc(&d_slots, capacity * sizeof(Slot)));
  fill_empty_kernel<<<blocks_for(capacity, 256, device), 256>>>(d_slots, capacity);
  KV_LAUNCHED();
  KV_TRY(arena.ensure((size_t)keys * row_stride * sizeof(float), 0));
  rows_mapped = arena.mapped() / (row_stride * sizeof(float));
  KV_CUDA(cudaDeviceSynchronize());
  return 0;
}

TableView Table::view() const {
  TableView v;
  v.slots = d_slots;
  v.mask = capacity - 1;
  int lg = 0;
  while ((1ULL << lg) < capacity) ++lg;
  v.shift = 64 - (lg - 1);
  v.rows = static_cast<float*>(arena.base());
  v.dim = dim;
  v.row_stride = row_stride;
  v.init = d_init;
  v.init_rows = init_rows;
  v.seed = seed;
  v.enter_threshold = enter_threshold;
  v.ctr = d_ctr;
  v.free_rows = d_free;
  v.rows_cap = rows_mapped;
  return v;
}

int Table::sync_counters(cudaStream_t stream) {
  KV_CUDA(cudaMemcpyAsync(h_ctr, d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
  KV_CUDA(cudaStreamSynchronize(stream));
  if (h_ctr->overflow != 0)
    return fail(2, "KvVariable: the table overflowed while captured work was replayed (more new "
                   "keys than kv_reserve made room for); its contents are invalid");
  return 0;
}

int Table::rehash(uint64_t new_capacity, cudaStream_t stream) {
  Slot* n = nullptr;
  KV_CUDA(cudaMalloc(&n, new_capacity * sizeof(Slot)));
  fill_empty_kernel<<<blocks_for(new_capacity, 256, device), 256, 0, stream>>>(n, new_capacity);
  KV_LAUNCHED();
  Slot* old = d_slots;
  uint64_t old_cap = capacity;
  d_slots = n;
  capacity = new_capacity;
  rehash_kernel<<<blocks_for(old_cap, 256, device), 256, 0, stream>>>(old, old_cap, view());
  KV_LAUNCHED();
  // used := live keys, tombstones := 0
  KV_TRY(sync_counters(stream));
  uint64_t live = h_ctr->used - h_ctr->tombstones;
  h_ctr->used = live;
  h_ctr->tombstones = 0;
  KV_CUDA(cudaMemcpyAsync(d_ctr, h_ctr, 4 * sizeof(unsigned long long),
                          cudaMemcpyHostToDevice, stream));
  KV_CUDA(cudaStreamSynchronize(stream));
  cudaFree(old);
  used_ub = live;
  return 0;
}

// Sizing policy.  Eager calls keep host-side upper bounds of the claimed slots and rows
// (each call can insert at most n keys) and only read the device counters back - one stream
// synchronisation - when a bound reaches a limit.  Under CUDA-graph capture nothing may
// synchronise or allocate: the call only checks that the table already has room (kv_reserve
// makes it) and books nothing, because the captured work runs later, any number of times;
// after a capture the bounds are no longer bounds, so eager calls re-read the counters.
int Table::ensure(int64_t n, cudaStream_t stream, bool exact) {
  if (n < 0) n = 0;
  const uint64_t un = (uint64_t)n;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) cudaGetLastError();
  if (cap != cudaStreamCaptureStatusNone) {
    captured = true;
    if (used_ub + un > capacity / 2 || rows_ub + un > rows_mapped)
      return fail(2, "KvVariable would have to grow during CUDA-graph capture: call kv_reserve "
                     "with the number of keys the captured work may insert first");
    return 0;
  }
  if (exact || captured || used_ub + un > capacity / 2 || rows_ub + un > rows_mapped) {
    KV_TRY(sync_counters(stream));
    used_ub = h_ctr->used;
    rows_ub = h_ctr->rows_bump;
    if (used_ub + un > capacity / 2) {
      uint64_t live = h_ctr->used - h_ctr->tombstones;
      // next power of two that leaves the table at most a quarter full
      KV_TRY(rehash(pow2_at_least((live + un) * 4), stream));
    }
    if (rows_ub + un > rows_mapped) {
      KV_TRY(arena.ensure((size_t)(rows_ub + un) * row_stride * sizeof(float), stream));
      rows_mapped = arena.mapped() / (row_stride * sizeof(float));
    }
  }
  used_ub += un;
  rows_ub += un;
  return 0;
}

int Table::ensure_free_list(uint64_t n, cudaStream_t stream) {
  KV_TRY(sync_counters(stream));
  uint64_t top = h_ctr->free_top > 0 ? (uint64_t)h_ctr->free_top : 0;
  if (top + n <= free_cap) return 0;
  uint64_t want = pow2_at_least(top + n);
  uint32_t* nf = nullptr;
  KV_CUDA(cudaMalloc(&nf, want * sizeof(uint32_t)));
  if (d_free && top)
    KV_CUDA(cudaMemcpyAsync(nf, d_free, top * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
  KV_CUDA(cudaStreamSynchronize(stream));
  if (d_free) cudaFree(d_free);
  d_free = nf;
  free_cap = want;
  return 0;
}

int Table::clear(cudaStream_t stream) {
  fill_empty_kernel<<<blocks_for(capacity, 256, device), 256, 0, stream>>>(d_slots, capacity);
  KV_LAUNCHED();
  KV_CUDA(cudaMemsetAsync(d_ctr, 0, sizeof(Counters), stream));
  used_ub = 0;
  rows_ub = 0;
  return 0;
}

}  // namespace kvhbm
