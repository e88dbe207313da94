// capi.cu — the extern "C" surface declared in include/kvhbm.h.
#include "../../include/kvhbm.h"

#include <string>

#include "plan.h"
#include "table.h"

namespace kvhbm {
struct Workspace;
struct Plan;
Plan* plan_new(int64_t max_ids, int heavy_t, int* rc);
void plan_delete(Plan*);
PlanView plan_view(const Plan*);
int do_plan_build(Plan*, Workspace*, const int64_t*, int64_t, cudaStream_t);
int do_gather_plan(Table*, bool insert, const PlanView&, const int*, float*, uint16_t, cudaStream_t);
int do_apply(int kind, Table*, Table*, Table*, const int64_t*, const float*, int64_t,
             const int32_t*, const float* hp, const float* d_hp, int update_slots, uint16_t,
             cudaStream_t, float* d_adv);
int do_apply_plan(int kind, Table*, Table*, Table*, Plan*, const float*, const float* hp,
                  const float* d_hp, int update_slots, uint16_t, cudaStream_t, float* d_adv);
int do_segment_sum_plan(Plan*, const float*, int, float*, cudaStream_t);
const std::string& last_error();
long long launch_count();
Workspace* workspace_new();
void workspace_delete(Workspace*);

int do_gather(Table*, bool insert, const int64_t*, const int32_t*, int64_t, float*, uint16_t,
              cudaStream_t, const int32_t* d_n = nullptr);
int do_scatter(Table*, int op, const int64_t*, const float*, int64_t, cudaStream_t, bool unique_ids);
int do_insert(Table*, const int64_t*, const float*, int64_t, const uint8_t*, const uint8_t*,
              cudaStream_t);
int do_get_count(Table*, const int64_t*, int64_t, int32_t*, cudaStream_t);
int do_get_timestamp(Table*, const int64_t*, int64_t, uint32_t*, uint16_t, cudaStream_t);
int set_trace(unsigned long long* d_buf);
int set_trace_apply(unsigned long long* d_buf);
int set_trace_unique(unsigned long long* d_buf);
int do_permute_rows(bool scatter, const float*, const int32_t*, int64_t, int, float*, cudaStream_t);

int do_apply_adagrad(Table*, Table*, const int64_t*, const float*, int64_t, const int32_t*,
                     const float* hp, const float* d_hp, int, uint16_t, cudaStream_t);
int do_apply_group_adam_v4(Table*, Table*, const int64_t*, const float*, int64_t, const int32_t*,
                           const float* hp, const float* d_hp, uint16_t, cudaStream_t,
                           float* d_adv = nullptr);
int do_apply_sparse_group_ftrl(Table*, Table*, Table*, const int64_t*, const float*, int64_t,
                               const int32_t*, const float* hp, const float* d_hp, uint16_t,
                               cudaStream_t);
int do_apply_adam(Table*, Table*, const int64_t*, const float*, int64_t, const int32_t*,
                  const float* hp, const float* d_hp, uint16_t, cudaStream_t,
                  float* d_adv = nullptr);

int do_unique(Workspace*, const int64_t*, int64_t, int64_t*, int32_t*, int32_t*, int32_t*,
              cudaStream_t);
int do_segment_sum(Workspace*, const float*, const int32_t*, int64_t, int, int64_t, const int32_t*,
                   float*, int accumulate, cudaStream_t);
int do_zero_rows(float*, int64_t, const int32_t*, int, cudaStream_t);
int do_partition_ids(Workspace*, const int64_t*, int64_t, const int32_t*, int, int, int64_t*,
                     int32_t*, int32_t*, cudaStream_t);

int do_route_ids(Workspace*, const int64_t*, const int32_t*, int64_t, const int32_t*, int, int, int,
                 int64_t*, int32_t*, int32_t*, int32_t*, int32_t*, int pairs, int64_t* const*,
                 int32_t* const*, cudaStream_t);
int do_gather_segments(Table*, const int64_t*, const int32_t*, int64_t, float* const*, int64_t,
                       uint16_t, cudaStream_t);
int do_peer_barrier(uint32_t* const*, uint32_t*, uint32_t*, int, int, int64_t, cudaStream_t);
int do_unzip_pairs(const int64_t*, int64_t, int64_t*, int32_t*, cudaStream_t);
int do_unique_route(Workspace*, const int64_t*, int64_t, int64_t*, int32_t*, int32_t*, int32_t*,
                    int, int, int, int64_t* const*, int32_t* const*, int32_t*, int32_t*,
                    int32_t*, cudaStream_t);
int do_route_fill(int, int, int64_t* const*, int32_t* const*, int32_t*, cudaStream_t);
int do_expand_rows(const float*, const int32_t*, const int32_t*, int64_t, int, float*, cudaStream_t);
int do_scatter_rows_n(const float*, const int32_t*, int64_t, const int32_t*, int, float*,
                      float* const*, int64_t, cudaStream_t);
int do_stats(Table*, cudaStream_t, int64_t*, int64_t*, int64_t*);
int do_export_count(Table*, int, int, float, cudaStream_t, int64_t*, int64_t*, int64_t*);
int do_export(Table*, int, int64_t*, float*, int64_t*, int64_t*, void*, int, cudaStream_t,
              int64_t cap_k, int64_t cap_b, int64_t cap_f);
int do_import(Table*, const int64_t*, const float*, int64_t, const float*, int64_t, const int64_t*,
              int64_t, const int64_t*, const void*, int64_t, int, cudaStream_t);
int do_delete(Table*, const int64_t*, int64_t, cudaStream_t);
int do_sparse_combine(const float* emb, const int32_t* idx, const int64_t* seg, const float* w,
                      int64_t nnz, int64_t n_rows, int dim, int combiner, float* out,
                      cudaStream_t st);
int do_delete_older(Table*, int, uint16_t, int64_t*, int64_t, cudaStream_t, int64_t*, Table* delta_set);
int do_delta_mark(Table* set, const int64_t* ids, int64_t n, const int32_t* d_n, Table* main,
                  bool filter, cudaStream_t st);
int do_delta_export(Table* tb, Table* train, Table* pred, bool support_pred, int first_n, int write,
                    int64_t* keys, float* values, int64_t* blacklist, int64_t* freq_keys,
                    uint32_t* freq_values, int64_t* delete_keys, int64_t cap_k, int64_t cap_b,
                    int64_t cap_f, int64_t cap_d, cudaStream_t st, int64_t* counts);
int do_delta_import(Table* tb, int first_n, const int64_t* keys, const float* values, int64_t n,
                    const int64_t* blacklist, int64_t n_black, const int64_t* freq_keys,
                    const uint32_t* freq_values, int64_t n_freq, const int64_t* delete_keys,
                    int64_t n_delete, cudaStream_t st);

// InitRandomValues: only the first call takes effect unless `force` (import).
int do_set_init_table(Table* tb, const float* d_table, int64_t rows, cudaStream_t st, bool force) {
  if (!force && tb->initialized && tb->init_rows > 0) return 0;
  if (rows < 0 || (rows > 0 && d_table == nullptr))
    return fail(1, "init table must be [rows, dim] in device memory");
  float* n = nullptr;
  if (rows > 0) {
    KV_CUDA(cudaMalloc(&n, (size_t)rows * tb->dim * sizeof(float)));
    KV_CUDA(cudaMemcpyAsync(n, d_table, (size_t)rows * tb->dim * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
  }
  if (tb->d_init) {
    KV_CUDA(cudaStreamSynchronize(st));  // kernels in flight may still read the old table
    cudaFree(tb->d_init);
  }
  tb->d_init = n;
  tb->init_rows = rows;
  tb->initialized = true;
  return 0;
}
}  // namespace kvhbm

using namespace kvhbm;

struct kv_table {
  Table t;
  // SUPPORT_DELTA_EXPORT (kv_variable.h:101-111): key sets of what changed since the last delta
  // export; null until kv_enable_delta_export
  Table* delta_train = nullptr;
  Table* delta_pred = nullptr;
  bool support_pred_delta = false;
  ~kv_table() { delete delta_train; delete delta_pred; }
};
struct kv_workspace { Workspace* w; };
struct kv_plan { Plan* p; };

namespace {
inline cudaStream_t S(kv_stream s) { return static_cast<cudaStream_t>(s); }

// Every entry point runs on the table's device and under the table's mutex
// (TF may call one kernel object from several executor threads).
struct Guard {
  std::unique_lock<std::mutex> l;
  int rc = 0;
  int prev = -1;   // the caller's current device, put back on return
  explicit Guard(kv_table* t) {
    if (!t) { rc = fail(KV_INVALID_ARGUMENT, "null table handle"); return; }
    l = std::unique_lock<std::mutex>(t->t.mu);
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev == t->t.device) { prev = -1; return; }
    cudaError_t e = cudaSetDevice(t->t.device);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaSetDevice");
  }
  ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define KV_ENTER(t)      \
  Guard _g(t);           \
  if (_g.rc) return _g.rc
#define KV_NEED(cond, msg) \
  if (!(cond)) return fail(KV_INVALID_ARGUMENT, msg)

// train_deltalist_.insert(key) for the ids of one call (no-op unless delta export is enabled).
// filter_by: the apply ops leave out the ids they skip (low-frequency keys of the value table).
inline int delta_mark(kv_table* t, const int64_t* ids, int64_t n, const int32_t* d_n,
                      kv_table* filter_by, cudaStream_t st) {
  if (!t || !t->delta_train || n <= 0 || !ids) return 0;
  return do_delta_mark(t->delta_train, ids, n, d_n, filter_by ? &filter_by->t : &t->t,
                       filter_by != nullptr, st);
}
#define KV_MARK(t, ids, n, d_n, by) KV_TRY(delta_mark(t, ids, n, d_n, by, S(stream)))
// MarkAsDeltaListElements on the value table and on every slot table of an apply op, for the
// ids the op does not skip.  Must run BEFORE the apply (a key the apply inserts is not filtered).
inline int delta_mark_apply(kv_table* var, kv_table* a, kv_table* b, const int64_t* ids, int64_t n,
                            const int32_t* d_n, bool filters, cudaStream_t st) {
  kv_table* by = filters ? var : nullptr;
  KV_TRY(delta_mark(var, ids, n, d_n, by, st));
  if (a && a->delta_train) KV_TRY(do_delta_mark(a->delta_train, ids, n, d_n, &var->t, filters, st));
  if (b && b->delta_train) KV_TRY(do_delta_mark(b->delta_train, ids, n, d_n, &var->t, filters, st));
  return 0;
}
}  // namespace

extern "C" {

const char* kv_last_error(void) { return last_error().c_str(); }
int64_t kv_launch_count(void) { return launch_count(); }

int kv_create(int dim, int enter_threshold, int64_t capacity_hint, kv_table** out) {
  KV_NEED(out != nullptr, "kv_create: out is null");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(KV_INTERNAL, "kvhbm needs a CUDA device: there is no CPU fallback");
  }
  kv_table* t = new kv_table();
  int rc = t->t.create(dim, enter_threshold, capacity_hint);
  if (rc) { delete t; return rc; }
  *out = t;
  return KV_OK;
}
int kv_destroy(kv_table* t) {
  if (!t) return KV_OK;
  cudaSetDevice(t->t.device);
  cudaDeviceSynchronize();
  delete t;
  return KV_OK;
}
int kv_dim(const kv_table* t) { return t ? t->t.dim : -1; }
int kv_enter_threshold(const kv_table* t) { return t ? (int)t->t.enter_threshold : -1; }
int kv_set_seed(kv_table* t, uint64_t seed) {
  KV_ENTER(t);
  t->t.seed = seed;
  return KV_OK;
}
int kv_reserve(kv_table* t, int64_t n_keys, kv_stream stream) {
  KV_ENTER(t);
  Table& tb = t->t;
  KV_NEED(n_keys >= 0, "reserve: negative size");
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(S(stream), &cs) != cudaSuccess) cudaGetLastError();
  KV_NEED(cs == cudaStreamCaptureStatusNone, "reserve: not while the stream is being captured");
  KV_TRY(tb.ensure(n_keys, S(stream), /*exact=*/true));
  // ensure() booked the keys as used (it always does outside capture); a reservation does not
  // insert anything
  tb.used_ub -= (uint64_t)n_keys;
  tb.rows_ub -= (uint64_t)n_keys;
  return KV_OK;
}
int kv_check_overflow(kv_table* t, kv_stream stream) {
  KV_ENTER(t);
  return t->t.sync_counters(S(stream));
}
int kv_set_init_table(kv_table* t, const float* d_table, int64_t rows, kv_stream stream) {
  KV_ENTER(t);
  return do_set_init_table(&t->t, d_table, rows, S(stream), false);
}
int kv_is_initialized(const kv_table* t, int* out) {
  KV_NEED(t && out, "kv_is_initialized: null argument");
  *out = t->t.initialized ? 1 : 0;
  return KV_OK;
}
int kv_init_table_rows(const kv_table* t, int64_t* rows) {
  KV_NEED(t && rows, "kv_init_table_rows: null argument");
  *rows = t->t.init_rows;
  return KV_OK;
}
int kv_get_init_table(const kv_table* t, float* d_out, kv_stream stream) {
  KV_NEED(t != nullptr, "null table handle");
  if (t->t.init_rows == 0) return KV_OK;
  KV_CUDA(cudaMemcpyAsync(d_out, t->t.d_init, (size_t)t->t.init_rows * t->t.dim * sizeof(float),
                          cudaMemcpyDeviceToDevice, S(stream)));
  return KV_OK;
}
int kv_size(kv_table* t, kv_stream stream, int64_t* out) {
  KV_ENTER(t);
  return do_stats(&t->t, S(stream), out, nullptr, nullptr);
}
int kv_sum_freq(kv_table* t, kv_stream stream, int64_t* out) {
  KV_ENTER(t);
  return do_stats(&t->t, S(stream), nullptr, out, nullptr);
}
int kv_map_size(kv_table* t, kv_stream stream, int64_t* out) {
  KV_ENTER(t);
  return do_stats(&t->t, S(stream), nullptr, nullptr, out);
}

int kv_gather_or_insert(kv_table* t, const int64_t* d_ids, const int32_t* d_counts, int64_t n,
                        float* d_out, uint16_t today, kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(n >= 0 && (n == 0 || (d_ids && d_out)), "gather_or_insert: bad arguments");
  KV_MARK(t, d_ids, n, nullptr, nullptr);
  return do_gather(&t->t, true, d_ids, d_counts, n, d_out, today, S(stream));
}
int kv_gather_or_insert_n(kv_table* t, const int64_t* d_ids, const int32_t* d_counts, int64_t n,
                          const int32_t* d_n, float* d_out, uint16_t today, kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(n >= 0 && (n == 0 || (d_ids && d_out)), "gather_or_insert_n: bad arguments");
  KV_MARK(t, d_ids, n, d_n, nullptr);
  return do_gather(&t->t, true, d_ids, d_counts, n, d_out, today, S(stream), d_n);
}
int kv_gather_or_zeros(kv_table* t, const int64_t* d_ids, int64_t n, float* d_out,
                       kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(n >= 0 && (n == 0 || (d_ids && d_out)), "gather_or_zeros: bad arguments");
  return do_gather(&t->t, false, d_ids, nullptr, n, d_out, 0, S(stream));
}
int kv_insert_or_update(kv_table* t, const int64_t* d_ids, const float* d_values, int64_t n,
                        const uint8_t* d_filter_out, const uint8_t* d_blacklist,
                        kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(n >= 0 && (n == 0 || (d_ids && d_values)), "insert_or_update: bad arguments");
  KV_MARK(t, d_ids, n, nullptr, nullptr);
  return do_insert(&t->t, d_ids, d_values, n, d_filter_out, d_blacklist, S(stream));
}
int kv_scatter(kv_table* t, int op, const int64_t* d_ids, const float* d_updates, int64_t n,
               kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(n >= 0 && (n == 0 || (d_ids && d_updates)), "scatter: bad arguments");
  KV_MARK(t, d_ids, n, nullptr, nullptr);
  return do_scatter(&t->t, op, d_ids, d_updates, n, S(stream), false);
}
int kv_scatter_unique(kv_table* t, int op, const int64_t* d_ids, const float* d_updates, int64_t n,
                      kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(n >= 0 && (n == 0 || (d_ids && d_updates)), "scatter: bad arguments");
  KV_MARK(t, d_ids, n, nullptr, nullptr);
  return do_scatter(&t->t, op, d_ids, d_updates, n, S(stream), true);
}
int kv_get_count(kv_table* t, const int64_t* d_ids, int64_t n, int32_t* d_out, kv_stream stream) {
  KV_ENTER(t);
  return do_get_count(&t->t, d_ids, n, d_out, S(stream));
}
int kv_get_timestamp(kv_table* t, const int64_t* d_ids, int64_t n, uint32_t* d_out,
                     uint16_t today, kv_stream stream) {
  KV_ENTER(t);
  return do_get_timestamp(&t->t, d_ids, n, d_out, today, S(stream));
}

// The apply ops take the mutexes of all their tables in address order, as
// MaybeLockVariableInputMutexesInOrder does (training_ops.cc:131-184).
namespace {
struct MultiGuard {
  std::unique_lock<std::mutex> l[3];
  int rc = 0;
  int prev = -1;
  ~MultiGuard() { if (prev >= 0) cudaSetDevice(prev); }
  MultiGuard(kv_table* a, kv_table* b, kv_table* c) {
    kv_table* v[3] = {a, b, c};
    int k = c ? 3 : 2;
    for (int i = 0; i < k; ++i)
      if (!v[i]) { rc = fail(KV_INVALID_ARGUMENT, "null table handle"); return; }
    for (int i = 0; i < k; ++i)
      for (int j = i + 1; j < k; ++j) {
        if (v[i] == v[j]) { rc = fail(KV_INVALID_ARGUMENT, "apply: var and slot are the same table"); return; }
        if (v[j] < v[i]) { kv_table* x = v[i]; v[i] = v[j]; v[j] = x; }
      }
    for (int i = 0; i < k; ++i) l[i] = std::unique_lock<std::mutex>(v[i]->t.mu);
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev == a->t.device) { prev = -1; return; }
    cudaError_t e = cudaSetDevice(a->t.device);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaSetDevice");
  }
};
}  // namespace

int kv_apply_adagrad(kv_table* var, kv_table* accum, const int64_t* d_ids, const float* d_grad,
                     int64_t n, const int32_t* d_n, float lr, int update_slots, uint16_t today,
                     kv_stream stream) {
  MultiGuard g(var, accum, nullptr);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, accum, nullptr, d_ids, n, d_n, true, S(stream)));
  const float hp[1] = {lr};
  return do_apply_adagrad(&var->t, &accum->t, d_ids, d_grad, n, d_n, hp, nullptr, update_slots,
                          today, S(stream));
}
int kv_apply_group_adam_v4(kv_table* var, kv_table* mvl, const int64_t* d_ids,
                           const float* d_grad, int64_t n, const int32_t* d_n, float lr,
                           float beta1_power, float beta2_power, float beta1, float beta2,
                           float epsilon, float l1, float l2, float l21, uint16_t today,
                           kv_stream stream) {
  MultiGuard g(var, mvl, nullptr);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, mvl, nullptr, d_ids, n, d_n, true, S(stream)));
  const float hp[9] = {lr, beta1_power, beta2_power, beta1, beta2, epsilon, l1, l2, l21};
  return do_apply_group_adam_v4(&var->t, &mvl->t, d_ids, d_grad, n, d_n, hp, nullptr, today,
                                S(stream));
}
int kv_apply_sparse_group_ftrl(kv_table* var, kv_table* accum, kv_table* linear,
                               const int64_t* d_ids, const float* d_grad, int64_t n,
                               const int32_t* d_n, float lr, float l1, float l2, float l21,
                               float l2_shrinkage, float lr_power, uint16_t today,
                               kv_stream stream) {
  MultiGuard g(var, accum, linear);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, accum, linear, d_ids, n, d_n, true, S(stream)));
  const float hp[6] = {lr, l1, l2, l21, l2_shrinkage, lr_power};
  return do_apply_sparse_group_ftrl(&var->t, &accum->t, &linear->t, d_ids, d_grad, n, d_n, hp,
                                    nullptr, today, S(stream));
}
int kv_apply_adam(kv_table* var, kv_table* m_v, const int64_t* d_ids, const float* d_grad,
                  int64_t n, const int32_t* d_n, float lr, float beta1, float beta2, float epsilon,
                  float beta1_power, float beta2_power, uint16_t today, kv_stream stream) {
  MultiGuard g(var, m_v, nullptr);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, m_v, nullptr, d_ids, n, d_n, false, S(stream)));
  const float hp[6] = {lr, beta1, beta2, epsilon, beta1_power, beta2_power};
  return do_apply_adam(&var->t, &m_v->t, d_ids, d_grad, n, d_n, hp, nullptr, today, S(stream));
}

// Device-resident hyper-parameters (TF scalar inputs that were not pinned to HostMemory).
int kv_apply_adagrad_dev(kv_table* var, kv_table* accum, const int64_t* d_ids,
                         const float* d_grad, int64_t n, const int32_t* d_n, const float* d_hp,
                         int update_slots, uint16_t today, kv_stream stream) {
  MultiGuard g(var, accum, nullptr);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, accum, nullptr, d_ids, n, d_n, true, S(stream)));
  KV_NEED(d_hp != nullptr, "d_hp is null");
  return do_apply_adagrad(&var->t, &accum->t, d_ids, d_grad, n, d_n, nullptr, d_hp, update_slots,
                          today, S(stream));
}
int kv_apply_group_adam_v4_dev(kv_table* var, kv_table* mvl, const int64_t* d_ids,
                               const float* d_grad, int64_t n, const int32_t* d_n,
                               const float* d_hp, uint16_t today, kv_stream stream) {
  MultiGuard g(var, mvl, nullptr);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, mvl, nullptr, d_ids, n, d_n, true, S(stream)));
  KV_NEED(d_hp != nullptr, "d_hp is null");
  return do_apply_group_adam_v4(&var->t, &mvl->t, d_ids, d_grad, n, d_n, nullptr, d_hp, today,
                                S(stream));
}
int kv_apply_group_adam_v4_dev_advance(kv_table* var, kv_table* mvl, const int64_t* d_ids,
                                       const float* d_grad, int64_t n, const int32_t* d_n,
                                       float* d_hp, uint16_t today, kv_stream stream) {
  MultiGuard g(var, mvl, nullptr);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, mvl, nullptr, d_ids, n, d_n, true, S(stream)));
  KV_NEED(d_hp != nullptr, "d_hp is null");
  return do_apply_group_adam_v4(&var->t, &mvl->t, d_ids, d_grad, n, d_n, nullptr, d_hp, today,
                                S(stream), d_hp);
}
int kv_apply_sparse_group_ftrl_dev(kv_table* var, kv_table* accum, kv_table* linear,
                                   const int64_t* d_ids, const float* d_grad, int64_t n,
                                   const int32_t* d_n, const float* d_hp, uint16_t today,
                                   kv_stream stream) {
  MultiGuard g(var, accum, linear);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, accum, linear, d_ids, n, d_n, true, S(stream)));
  KV_NEED(d_hp != nullptr, "d_hp is null");
  return do_apply_sparse_group_ftrl(&var->t, &accum->t, &linear->t, d_ids, d_grad, n, d_n,
                                    nullptr, d_hp, today, S(stream));
}
int kv_apply_adam_dev(kv_table* var, kv_table* m_v, const int64_t* d_ids, const float* d_grad,
                      int64_t n, const int32_t* d_n, const float* d_hp, uint16_t today,
                      kv_stream stream) {
  MultiGuard g(var, m_v, nullptr);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, m_v, nullptr, d_ids, n, d_n, false, S(stream)));
  KV_NEED(d_hp != nullptr, "d_hp is null");
  return do_apply_adam(&var->t, &m_v->t, d_ids, d_grad, n, d_n, nullptr, d_hp, today, S(stream));
}

int kv_apply_adam_dev_advance(kv_table* var, kv_table* m_v, const int64_t* d_ids,
                              const float* d_grad, int64_t n, const int32_t* d_n, float* d_hp,
                              uint16_t today, kv_stream stream) {
  MultiGuard g(var, m_v, nullptr);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, m_v, nullptr, d_ids, n, d_n, false, S(stream)));
  KV_NEED(d_hp != nullptr, "d_hp is null");
  return do_apply_adam(&var->t, &m_v->t, d_ids, d_grad, n, d_n, nullptr, d_hp, today, S(stream),
                       d_hp);
}

// ---- optimizer variants sharing the apply template ----
int kv_apply_group_adam_v3(kv_table* var, kv_table* mvl, const int64_t* d_ids,
                           const float* d_grad, int64_t n, const int32_t* d_n, float lr,
                           float beta1_power, float beta2_power, float beta1, float beta2,
                           float epsilon, float l1, float l2, float l21, uint16_t today,
                           kv_stream stream) {
  MultiGuard g(var, mvl, nullptr);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, mvl, nullptr, d_ids, n, d_n, true, S(stream)));
  const float hp[9] = {lr, beta1_power, beta2_power, beta1, beta2, epsilon, l1, l2, l21};
  return do_apply(KV_OPT_GROUP_ADAM_V3, &var->t, &mvl->t, nullptr, d_ids, d_grad, n, d_n, hp,
                  nullptr, 1, today, S(stream), nullptr);
}
int kv_apply_sparse_ftrl_v2(kv_table* var, kv_table* accum, kv_table* linear,
                            const int64_t* d_ids, const float* d_grad, int64_t n,
                            const int32_t* d_n, float lr, float l1, float l2, float l2_shrinkage,
                            float lr_power, uint16_t today, kv_stream stream) {
  MultiGuard g(var, accum, linear);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, accum, linear, d_ids, n, d_n, true, S(stream)));
  const float hp[6] = {lr, l1, l2, 0.f, l2_shrinkage, lr_power};
  return do_apply(KV_OPT_SPARSE_FTRL_V2, &var->t, &accum->t, &linear->t, d_ids, d_grad, n, d_n, hp,
                  nullptr, 1, today, S(stream), nullptr);
}
int kv_apply_group_sparse_ftrl_v2(kv_table* var, kv_table* accum, kv_table* linear,
                                  const int64_t* d_ids, const float* d_grad, int64_t n,
                                  const int32_t* d_n, float lr, float l1, float l2,
                                  float l2_shrinkage, float lr_power, uint16_t today,
                                  kv_stream stream) {
  MultiGuard g(var, accum, linear);
  if (g.rc) return g.rc;
  KV_TRY(delta_mark_apply(var, accum, linear, d_ids, n, d_n, true, S(stream)));
  const float hp[6] = {lr, l1, l2, 0.f, l2_shrinkage, lr_power};
  return do_apply(KV_OPT_GROUP_SPARSE_FTRL_V2, &var->t, &accum->t, &linear->t, d_ids, d_grad, n,
                  d_n, hp, nullptr, 1, today, S(stream), nullptr);
}

// ---- dedup plan ----
int kv_plan_create(int64_t max_ids, kv_plan** out) {
  KV_NEED(out != nullptr, "kv_plan_create: out is null");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(KV_INTERNAL, "kvhbm needs a CUDA device: there is no CPU fallback");
  }
  int rc = 0;
  Plan* p = plan_new(max_ids, 0, &rc);
  if (!p) return rc;
  kv_plan* h = new kv_plan();
  h->p = p;
  *out = h;
  return KV_OK;
}
int kv_plan_destroy(kv_plan* plan) {
  if (!plan) return KV_OK;
  cudaDeviceSynchronize();
  plan_delete(plan->p);
  delete plan;
  return KV_OK;
}
int kv_plan_build(kv_plan* plan, kv_workspace* ws, const int64_t* d_ids, int64_t n,
                  kv_stream stream) {
  KV_NEED(plan && ws && (n == 0 || d_ids), "plan_build: bad arguments");
  return do_plan_build(plan->p, ws->w, d_ids, n, S(stream));
}
int kv_plan_arrays(const kv_plan* plan, const int64_t** d_uniq, const int32_t** d_idx,
                   const int32_t** d_counts, const int32_t** d_num_unique,
                   const int32_t** d_seg_off, const int32_t** d_pos) {
  KV_NEED(plan != nullptr, "plan_arrays: null plan");
  const PlanView v = plan_view(plan->p);
  if (d_uniq) *d_uniq = reinterpret_cast<const int64_t*>(v.uniq);
  if (d_idx) *d_idx = v.idx;
  if (d_counts) *d_counts = v.counts;
  if (d_num_unique) *d_num_unique = v.num;
  if (d_seg_off) *d_seg_off = v.seg_off;
  if (d_pos) *d_pos = v.pos;
  return KV_OK;
}
int kv_gather_or_insert_plan(kv_table* t, kv_plan* plan, float* d_out, uint16_t today,
                             kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(plan && d_out, "gather_or_insert_plan: bad arguments");
  const PlanView v = plan_view(plan->p);
  KV_MARK(t, reinterpret_cast<const int64_t*>(v.uniq), v.n, v.num, nullptr);
  return do_gather_plan(&t->t, true, v, v.first, d_out, today, S(stream));
}
int kv_gather_or_zeros_plan(kv_table* t, kv_plan* plan, float* d_out, kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(plan && d_out, "gather_or_zeros_plan: bad arguments");
  const PlanView v = plan_view(plan->p);
  return do_gather_plan(&t->t, false, v, v.first, d_out, 0, S(stream));
}
int kv_segment_sum_plan(kv_plan* plan, const float* d_data, int dim, float* d_out,
                        kv_stream stream) {
  KV_NEED(plan && d_data && d_out && dim > 0, "segment_sum_plan: bad arguments");
  return do_segment_sum_plan(plan->p, d_data, dim, d_out, S(stream));
}
static int n_hp_of(int kind) {
  switch (kind) {
    case KV_OPT_ADAGRAD: return 1;
    case KV_OPT_GROUP_ADAM_V4: case KV_OPT_GROUP_ADAM_V3: return 9;
    case KV_OPT_SPARSE_GROUP_FTRL: case KV_OPT_SPARSE_FTRL_V2: case KV_OPT_GROUP_SPARSE_FTRL_V2:
    case KV_OPT_ADAM: return 6;
  }
  return -1;
}
int kv_apply_plan(int kind, kv_table* var, kv_table* slot_a, kv_table* slot_b, kv_plan* plan,
                  const float* d_grad, const float* hp, int n_hp, int update_slots, uint16_t today,
                  kv_stream stream) {
  MultiGuard g(var, slot_a, slot_b);
  if (g.rc) return g.rc;
  KV_NEED(plan && hp, "apply_plan: bad arguments");
  KV_NEED(n_hp_of(kind) == n_hp, "apply_plan: wrong number of scalar inputs for this optimizer");
  {
    const PlanView v = plan_view(plan->p);
    KV_TRY(delta_mark_apply(var, slot_a, slot_b, reinterpret_cast<const int64_t*>(v.uniq), v.n, v.num,
                            kind != KV_OPT_ADAM, S(stream)));
  }
  return do_apply_plan(kind, &var->t, &slot_a->t, slot_b ? &slot_b->t : nullptr, plan->p, d_grad,
                       hp, nullptr, update_slots, today, S(stream), nullptr);
}
int kv_apply_plan_dev(int kind, kv_table* var, kv_table* slot_a, kv_table* slot_b, kv_plan* plan,
                      const float* d_grad, float* d_hp, int advance_powers, int update_slots,
                      uint16_t today, kv_stream stream) {
  MultiGuard g(var, slot_a, slot_b);
  if (g.rc) return g.rc;
  KV_NEED(plan && d_hp && n_hp_of(kind) > 0, "apply_plan_dev: bad arguments");
  {
    const PlanView v = plan_view(plan->p);
    KV_TRY(delta_mark_apply(var, slot_a, slot_b, reinterpret_cast<const int64_t*>(v.uniq), v.n, v.num,
                            kind != KV_OPT_ADAM, S(stream)));
  }
  return do_apply_plan(kind, &var->t, &slot_a->t, slot_b ? &slot_b->t : nullptr, plan->p, d_grad,
                       nullptr, d_hp, update_slots, today, S(stream),
                       advance_powers ? d_hp : nullptr);
}

int kv_workspace_create(kv_workspace** out) {
  KV_NEED(out != nullptr, "kv_workspace_create: out is null");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(KV_INTERNAL, "kvhbm needs a CUDA device: there is no CPU fallback");
  }
  kv_workspace* w = new kv_workspace();
  w->w = workspace_new();
  *out = w;
  return KV_OK;
}
int kv_workspace_destroy(kv_workspace* ws) {
  if (!ws) return KV_OK;
  workspace_delete(ws->w);
  delete ws;
  return KV_OK;
}
int kv_unique(kv_workspace* ws, const int64_t* d_ids, int64_t n, int64_t* d_uniq, int32_t* d_idx,
              int32_t* d_counts, int32_t* d_num_unique, kv_stream stream) {
  KV_NEED(ws && d_num_unique && (n == 0 || (d_ids && d_uniq && d_idx)), "unique: bad arguments");
  return do_unique(ws->w, d_ids, n, d_uniq, d_idx, d_counts, d_num_unique, S(stream));
}
int kv_segment_sum(kv_workspace* ws, const float* d_data, const int32_t* d_idx, int64_t n, int dim,
                   int64_t max_segments, const int32_t* d_num_segments, float* d_out,
                   int accumulate, kv_stream stream) {
  KV_NEED(ws && (n == 0 || (d_data && d_idx)) && (max_segments == 0 || d_out),
          "segment_sum: bad arguments");
  return do_segment_sum(ws->w, d_data, d_idx, n, dim, max_segments, d_num_segments, d_out,
                        accumulate, S(stream));
}
int kv_sparse_combine(const float* d_emb, const int32_t* d_idx, const int64_t* d_segment_ids,
                      const float* d_weights, int64_t nnz, int64_t n_rows, int dim, int combiner,
                      float* d_out, kv_stream stream) {
  KV_NEED(nnz >= 0 && n_rows >= 0 && dim > 0 && (n_rows == 0 || d_out) &&
              (nnz == 0 || (d_emb && d_idx && d_segment_ids)),
          "sparse_combine: bad arguments");
  return do_sparse_combine(d_emb, d_idx, d_segment_ids, d_weights, nnz, n_rows, dim, combiner, d_out,
                           S(stream));
}
int kv_zero_rows(float* d_out, int64_t max_rows, const int32_t* d_num_rows, int dim,
                 kv_stream stream) {
  KV_NEED(max_rows == 0 || d_out, "zero_rows: bad arguments");
  return do_zero_rows(d_out, max_rows, d_num_rows, dim, S(stream));
}

int kv_export_count(kv_table* t, int first_n, int enable_cutoff, float cutoff_value,
                    kv_stream stream, int64_t* n_keys, int64_t* n_blacklist, int64_t* n_freq) {
  KV_ENTER(t);
  KV_NEED(n_keys && n_blacklist && n_freq, "export_count: null output");
  return do_export_count(&t->t, first_n, enable_cutoff, cutoff_value, S(stream), n_keys,
                         n_blacklist, n_freq);
}
int kv_export(kv_table* t, int first_n, int64_t* d_keys, float* d_values, int64_t* d_blacklist,
              int64_t* d_freq_keys, void* d_freq_values, int freq_u32, kv_stream stream) {
  KV_ENTER(t);
  return do_export(&t->t, first_n, d_keys, d_values, d_blacklist, d_freq_keys, d_freq_values,
                   freq_u32, S(stream), -1, -1, -1);
}
int kv_export_bounded(kv_table* t, int first_n, int64_t* d_keys, float* d_values, int64_t cap_keys,
                      int64_t* d_blacklist, int64_t cap_blacklist, int64_t* d_freq_keys,
                      void* d_freq_values, int64_t cap_freq, int freq_u32, kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(cap_keys >= 0 && cap_blacklist >= 0 && cap_freq >= 0, "export: negative capacity");
  return do_export(&t->t, first_n, d_keys, d_values, d_blacklist, d_freq_keys, d_freq_values,
                   freq_u32, S(stream), cap_keys, cap_blacklist, cap_freq);
}
int kv_import(kv_table* t, const int64_t* d_keys, const float* d_values, int64_t n,
              const float* d_init_table, int64_t init_rows, const int64_t* d_blacklist,
              int64_t n_blacklist, const int64_t* d_freq_keys, const void* d_freq_values,
              int64_t n_freq, int freq_u32, kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(n >= 0 && n_blacklist >= 0 && n_freq >= 0, "import: negative size");
  return do_import(&t->t, d_keys, d_values, n, d_init_table, init_rows, d_blacklist, n_blacklist,
                   d_freq_keys, d_freq_values, n_freq, freq_u32, S(stream));
}
int kv_enable_delta_export(kv_table* t, int support_prediction_delta) {
  KV_ENTER(t);
  t->support_pred_delta = support_prediction_delta != 0;
  for (Table** set : {&t->delta_train, &t->delta_pred}) {
    if (*set) continue;
    *set = new Table();
    const int rc = (*set)->create(1, 0, 1024);
    if (rc) { delete *set; *set = nullptr; return rc; }
  }
  return KV_OK;
}
int kv_delta_export_count(kv_table* t, int first_n, kv_stream stream, int64_t* n_keys,
                          int64_t* n_blacklist, int64_t* n_freq, int64_t* n_delete) {
  KV_ENTER(t);
  KV_NEED(t->delta_train != nullptr, "delta export is not enabled on this table");
  int64_t c[4];
  KV_TRY(do_delta_export(&t->t, t->delta_train, t->delta_pred, t->support_pred_delta, first_n, 0,
                         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, S(stream), c));
  if (n_keys) *n_keys = c[0];
  if (n_blacklist) *n_blacklist = c[1];
  if (n_freq) *n_freq = c[2];
  if (n_delete) *n_delete = c[3];
  return KV_OK;
}
int kv_delta_export(kv_table* t, int first_n, int64_t* d_keys, float* d_values, int64_t cap_keys,
                    int64_t* d_blacklist, int64_t cap_blacklist, int64_t* d_freq_keys,
                    uint32_t* d_freq_values, int64_t cap_freq, int64_t* d_delete_keys,
                    int64_t cap_delete, kv_stream stream, int64_t* counts) {
  KV_ENTER(t);
  KV_NEED(t->delta_train != nullptr, "delta export is not enabled on this table");
  KV_NEED(cap_keys >= 0 && cap_blacklist >= 0 && cap_freq >= 0 && cap_delete >= 0 &&
              (cap_keys == 0 || (d_keys && d_values)) && (cap_blacklist == 0 || d_blacklist) &&
              (cap_freq == 0 || (d_freq_keys && d_freq_values)) && (cap_delete == 0 || d_delete_keys),
          "delta_export: bad arguments");
  int64_t c[4];
  KV_TRY(do_delta_export(&t->t, t->delta_train, t->delta_pred, t->support_pred_delta, first_n, 1,
                         d_keys, d_values, d_blacklist, d_freq_keys, d_freq_values, d_delete_keys,
                         cap_keys, cap_blacklist, cap_freq, cap_delete, S(stream), c));
  if (counts) for (int i = 0; i < 4; ++i) counts[i] = c[i];
  return KV_OK;
}
int kv_delta_import(kv_table* t, int first_n, const int64_t* d_keys, const float* d_values,
                    int64_t n, const int64_t* d_blacklist, int64_t n_blacklist,
                    const int64_t* d_freq_keys, const uint32_t* d_freq_values, int64_t n_freq,
                    const int64_t* d_delete_keys, int64_t n_delete, kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(n >= 0 && n_blacklist >= 0 && n_freq >= 0 && n_delete >= 0, "delta_import: negative size");
  KV_NEED((n == 0 || (d_keys && d_values)) && (n_blacklist == 0 || d_blacklist) &&
              (n_freq == 0 || (d_freq_keys && d_freq_values)) && (n_delete == 0 || d_delete_keys),
          "delta_import: bad arguments");
  return do_delta_import(&t->t, first_n, d_keys, d_values, n, d_blacklist, n_blacklist, d_freq_keys,
                         d_freq_values, n_freq, d_delete_keys, n_delete, S(stream));
}
int kv_delta_size(kv_table* t, kv_stream stream, int64_t* out) {
  KV_ENTER(t);
  KV_NEED(t->delta_train != nullptr && out, "delta export is not enabled on this table");
  KV_TRY(t->delta_train->sync_counters(S(stream)));
  *out = (int64_t)t->delta_train->h_ctr->used;
  return KV_OK;
}
int kv_delete(kv_table* t, const int64_t* d_ids, int64_t n, kv_stream stream) {
  KV_ENTER(t);
  KV_MARK(t, d_ids, n, nullptr, nullptr);
  return do_delete(&t->t, d_ids, n, S(stream));
}
int kv_delete_with_timestamp(kv_table* t, int threshold, uint16_t today, int64_t* d_out_keys,
                             int64_t cap, kv_stream stream, int64_t* n_deleted) {
  KV_ENTER(t);
  return do_delete_older(&t->t, threshold, today, d_out_keys, cap, S(stream), n_deleted,
                         t->delta_train);
}

int kv_partition_ids(kv_workspace* ws, const int64_t* d_ids, int64_t n, const int32_t* d_n,
                     int num_shards, int mode, int64_t* d_sorted_ids, int32_t* d_perm,
                     int32_t* d_shard_counts, kv_stream stream) {
  KV_NEED(ws && d_shard_counts && (n == 0 || (d_ids && d_sorted_ids && d_perm)),
          "partition_ids: bad arguments");
  return do_partition_ids(ws->w, d_ids, n, d_n, num_shards, mode, d_sorted_ids, d_perm,
                          d_shard_counts, S(stream));
}
int kv_debug_set_trace(void* d_buf) {
  set_trace_apply(static_cast<unsigned long long*>(d_buf));
  set_trace_unique(static_cast<unsigned long long*>(d_buf));
  return set_trace(static_cast<unsigned long long*>(d_buf));
}
int kv_route_ids(kv_workspace* ws, const int64_t* d_ids, const int32_t* d_occ, int64_t n,
                 const int32_t* d_n, int num_shards, int mode, int capacity, int64_t* d_send_ids,
                 int32_t* d_send_occ, int32_t* d_perm, int32_t* d_counts, int32_t* d_overflow,
                 kv_stream stream) {
  KV_NEED(ws && d_send_ids && d_counts && d_overflow && (n == 0 || (d_ids && d_perm)),
          "route_ids: bad arguments");
  return do_route_ids(ws->w, d_ids, d_occ, n, d_n, num_shards, mode, capacity, d_send_ids,
                      d_send_occ, d_perm, d_counts, d_overflow, /*pairs=*/0, nullptr, nullptr, S(stream));
}
int kv_route_id_pairs(kv_workspace* ws, const int64_t* d_ids, const int32_t* d_occ, int64_t n,
                      const int32_t* d_n, int num_shards, int mode, int capacity,
                      int64_t* d_send_pairs, int32_t* d_perm, int32_t* d_counts,
                      int32_t* d_overflow, kv_stream stream) {
  KV_NEED(ws && d_send_pairs && d_counts && d_overflow && (n == 0 || (d_ids && d_perm)),
          "route_id_pairs: bad arguments");
  return do_route_ids(ws->w, d_ids, d_occ, n, d_n, num_shards, mode, capacity, d_send_pairs,
                      nullptr, d_perm, d_counts, d_overflow, /*pairs=*/1, nullptr, nullptr, S(stream));
}
int kv_route_ids_peer(kv_workspace* ws, const int64_t* d_ids, const int32_t* d_occ, int64_t n,
                      const int32_t* d_n, int num_shards, int mode, int capacity,
                      int64_t* const* d_seg_ids, int32_t* const* d_seg_occ, int32_t* d_perm,
                      int32_t* d_counts, int32_t* d_overflow, kv_stream stream) {
  KV_NEED(ws && d_seg_ids && d_seg_occ && d_counts && d_overflow && (n == 0 || (d_ids && d_perm)),
          "route_ids_peer: bad arguments");
  return do_route_ids(ws->w, d_ids, d_occ, n, d_n, num_shards, mode, capacity, nullptr, nullptr,
                      d_perm, d_counts, d_overflow, /*pairs=*/0, d_seg_ids, d_seg_occ, S(stream));
}
int kv_unique_route_peer(kv_workspace* ws, const int64_t* d_ids, int64_t n, int64_t* d_uniq,
                         int32_t* d_idx, int32_t* d_counts, int32_t* d_num_unique,
                         int num_shards, int mode, int capacity, int64_t* const* d_seg_ids,
                         int32_t* const* d_seg_occ, int32_t* d_perm, int32_t* d_shard_counts,
                         int32_t* d_overflow, kv_stream stream) {
  KV_NEED(ws && d_num_unique && d_seg_ids && d_seg_occ && d_shard_counts && d_overflow &&
              (n == 0 || (d_ids && d_uniq && d_idx && d_perm)),
          "unique_route_peer: bad arguments");
  return do_unique_route(ws->w, d_ids, n, d_uniq, d_idx, d_counts, d_num_unique, num_shards, mode,
                         capacity, d_seg_ids, d_seg_occ, d_perm, d_shard_counts, d_overflow,
                         S(stream));
}
int kv_route_fill_peer(int num_shards, int capacity, int64_t* const* d_seg_ids,
                       int32_t* const* d_seg_occ, int32_t* d_shard_counts, kv_stream stream) {
  KV_NEED(d_seg_ids && d_seg_occ && d_shard_counts, "route_fill_peer: bad arguments");
  return do_route_fill(num_shards, capacity, d_seg_ids, d_seg_occ, d_shard_counts, S(stream));
}
int kv_gather_or_insert_peer(kv_table* t, const int64_t* d_ids, const int32_t* d_counts,
                             int64_t n, float* const* d_seg_rows, int64_t capacity,
                             uint16_t today, kv_stream stream) {
  KV_ENTER(t);
  KV_NEED(n >= 0 && (n == 0 || (d_ids && d_seg_rows)), "gather_or_insert_peer: bad arguments");
  KV_MARK(t, d_ids, n, nullptr, nullptr);
  return do_gather_segments(&t->t, d_ids, d_counts, n, d_seg_rows, capacity, today, S(stream));
}
int kv_scatter_rows_n_peer(const float* d_src, const int32_t* d_perm, int64_t n,
                           const int32_t* d_n, int dim, float* const* d_seg_rows,
                           int64_t capacity, kv_stream stream) {
  return do_scatter_rows_n(d_src, d_perm, n, d_n, dim, nullptr, d_seg_rows, capacity, S(stream));
}
int kv_peer_barrier(uint32_t* const* d_peer_flags, uint32_t* d_my_flags, uint32_t* d_state,
                    int rank, int world, int64_t timeout_ms, kv_stream stream) {
  return do_peer_barrier(d_peer_flags, d_my_flags, d_state, rank, world, timeout_ms, S(stream));
}
int kv_unzip_pairs(const int64_t* d_pairs, int64_t n, int64_t* d_ids, int32_t* d_occ,
                   kv_stream stream) {
  return do_unzip_pairs(d_pairs, n, d_ids, d_occ, S(stream));
}
int kv_expand_rows(const float* d_src, const int32_t* d_perm, const int32_t* d_idx, int64_t n,
                   int dim, float* d_out, kv_stream stream) {
  return do_expand_rows(d_src, d_perm, d_idx, n, dim, d_out, S(stream));
}
int kv_scatter_rows_n(const float* d_src, const int32_t* d_perm, int64_t n, const int32_t* d_n,
                      int dim, float* d_out, kv_stream stream) {
  return do_scatter_rows_n(d_src, d_perm, n, d_n, dim, d_out, nullptr, 0, S(stream));
}
int kv_permute_rows(const float* d_src, const int32_t* d_perm, int64_t n, int dim, float* d_out,
                    kv_stream stream) {
  return do_permute_rows(false, d_src, d_perm, n, dim, d_out, S(stream));
}
int kv_scatter_rows(const float* d_src, const int32_t* d_perm, int64_t n, int dim, float* d_out,
                    kv_stream stream) {
  return do_permute_rows(true, d_src, d_perm, n, dim, d_out, S(stream));
}

}  // extern "C"
