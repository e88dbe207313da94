// table.h — host side of one HBM-resident KvVariable: owns the slot array, the
// growable row arena, the init table and the counters; decides when to grow.
#ifndef KVHBM_TABLE_H_
#define KVHBM_TABLE_H_

#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace kvhbm {

struct Plan;
struct Workspace;


// Error plumbing shared by every translation unit.
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
#define KV_CUDA(expr)                                             \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return ::kvhbm::cuda_fail(_e, #expr);  \
  } while (0)
#define KV_TRY(expr)          \
  do {                        \
    int _s = (expr);          \
    if (_s != 0) return _s;   \
  } while (0)
// after a kernel launch
#define KV_LAUNCHED() \
  do { ::kvhbm::count_launch(); KV_CUDA(cudaGetLastError()); } while (0)

// A device allocation that grows in place: a large virtual-address reservation
// (cuMemAddressReserve) backed by physical chunks mapped on demand (cuMemCreate
// + cuMemMap), so row pointers stay valid while the table grows towards the
// 180 GB of a B200 and nothing is ever copied.  Falls back to cudaMalloc +
// copy if the driver refuses virtual memory management.
class GrowableArena {
 public:
  ~GrowableArena();
  int init(int device, size_t reserve_bytes);
  // Makes at least `bytes` addressable.  `stream` is only used by the fallback.
  int ensure(size_t bytes, cudaStream_t stream);
  void* base() const { return reinterpret_cast<void*>(base_); }
  size_t mapped() const { return mapped_; }
  bool vmm() const { return vmm_; }

 private:
  int device_ = 0;
  bool vmm_ = false;
  CUdeviceptr base_ = 0;
  size_t reserved_ = 0, mapped_ = 0, gran_ = 0;
  std::vector<CUmemGenericAllocationHandle> handles_;
  std::vector<size_t> handle_bytes_;
};

struct Table {
  int device = 0;
  int dim = 0;
  int row_stride = 0;
  uint32_t enter_threshold = 0;
  uint64_t seed = 0;

  Slot* d_slots = nullptr;
  uint64_t capacity = 0;  // power of two
  GrowableArena arena;
  uint64_t rows_mapped = 0;

  float* d_init = nullptr;
  int64_t init_rows = 0;
  bool initialized = false;

  Counters* d_ctr = nullptr;
  Counters* h_ctr = nullptr;  // pinned mirror
  uint32_t* d_free = nullptr;
  uint64_t free_cap = 0;

  // host-side upper bounds, so that the steady state never synchronises
  uint64_t used_ub = 0;
  uint64_t rows_ub = 0;
  bool captured = false;  // some call on this table was recorded into a CUDA graph
  // dedup plan + scratch of the duplicate-safe scatter (lookup.cu do_scatter), sized on first use
  Plan* scatter_plan = nullptr;
  Workspace* scatter_ws = nullptr;

  std::mutex mu;

  ~Table();
  int create(int dim, int enter_threshold, int64_t capacity_hint);
  TableView view() const;
  // Guarantee room for `n` more keys (slots at load <= 0.5 and rows).
  int ensure(int64_t n, cudaStream_t stream, bool exact = false);
  // Copy the counters to the host (synchronises `stream`).
  int sync_counters(cudaStream_t stream);
  int rehash(uint64_t new_capacity, cudaStream_t stream);
  int ensure_free_list(uint64_t n, cudaStream_t stream);
  int clear(cudaStream_t stream);
};

// Programmatic dependent launch for the kernels of the step's serial chain (lookup probe ->
// expand -> staging -> fused apply -> next lookup): the next kernel's blocks are scheduled while
// the previous kernel drains and wait in `griddepcontrol.wait` (pdl_wait(), first thing in the
// kernel) until it has completed and flushed.  KVHBM_PDL=0 turns the attribute off.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// launch geometry helpers
int sm_count(int device);
inline int blocks_for(int64_t work_items, int per_block, int device, int max_per_sm = 8) {
  int64_t b = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)sm_count(device) * max_per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace kvhbm
#endif  // KVHBM_TABLE_H_
