/* kvhbm.h — C ABI of the B200-native KvVariable hot path (libkvhbm.so).
 *
 * This is the drop-in boundary: each entry point is what a TensorFlow 2.13
 * DEVICE_GPU OpKernel for the TFPlus op of the same meaning calls on the op's
 * CUDA stream (glue in tf_ops/, binding shown in INTEGRATION.md).  Plain
 * pointers and sizes only; every `d_*` pointer is DEVICE memory on the table's
 * GPU, everything else is host.  All functions return a kv_status (0 = OK);
 * kv_last_error() gives the message of the calling thread's last failure.
 * Work is enqueued on `stream` (a cudaStream_t passed as void*); functions that
 * return a host scalar synchronise that stream and say so.
 *
 * Keys are int64, values fp32 (the reference also registers int32/uint64 keys
 * and half values, kv_variable_ops.cc:127-156; those return KV_UNIMPLEMENTED).
 * Two key values are reserved as slot sentinels: INT64_MIN and INT64_MIN+1.
 *
 * There is no CPU fallback: without a CUDA device every call fails with
 * KV_INTERNAL.
 *
 * Citations are file:line under /root/reference/tfplus/kv_variable/.
 */
#ifndef KVHBM_H_
#define KVHBM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kv_table kv_table;         /* one KvVariable<int64,float> resource */
typedef struct kv_workspace kv_workspace; /* scratch for dedup / routing kernels */
typedef void* kv_stream;                  /* cudaStream_t */

typedef enum {
  KV_OK = 0,
  KV_INVALID_ARGUMENT = 1,    /* errors::InvalidArgument */
  KV_FAILED_PRECONDITION = 2, /* errors::FailedPrecondition (uninitialized var) */
  KV_UNIMPLEMENTED = 3,
  KV_RESOURCE_EXHAUSTED = 4,  /* out of HBM */
  KV_INTERNAL = 5             /* CUDA error */
} kv_status;

/* kernels/kv_variable_interface.h ScatterUpdateOps */
typedef enum {
  KV_SCATTER_ASSIGN = 0, KV_SCATTER_ADD = 1, KV_SCATTER_SUB = 2,
  KV_SCATTER_MUL = 3, KV_SCATTER_DIV = 4, KV_SCATTER_MIN = 5, KV_SCATTER_MAX = 6
} kv_scatter_op;

const char* kv_last_error(void);
/* Number of kernel launches this library has issued in this process (bench
 * bookkeeping: the `gpu_launches` claim is read from here, not guessed). */
int64_t kv_launch_count(void);

/* ---- lifecycle ---------------------------------------------------------- */

/* CreateKvVariableOp::Compute, kernels/kv_variable_ops.cc:58-116 ->
 * KvVariable ctor kernels/kv_variable.h:92-114.  `capacity_hint` keys are
 * pre-sized (0 = default); the table grows on demand.  Uses the current CUDA
 * device. */
int kv_create(int dim, int enter_threshold, int64_t capacity_hint, kv_table** out);
/* DestroyKvVariableOp, kernels/kv_variable_ops.cc:295-323. */
int kv_destroy(kv_table* t);
int kv_dim(const kv_table* t);
int kv_enter_threshold(const kv_table* t);
/* Seed of the deterministic initializer that replaces std::rand()
 * (kernels/kv_variable.h:889-898; DESIGN.md "initializer"). */
int kv_set_seed(kv_table* t, uint64_t seed);
/* Make room for `n_keys` more keys now, so that later calls never have to
 * resize (needed before CUDA-graph capture). */
int kv_reserve(kv_table* t, int64_t n_keys, kv_stream stream);
/* Under CUDA-graph capture the host books nothing, so replays can insert more keys than
 * kv_reserve made room for.  The kernels then never touch unmapped memory but raise a sticky
 * flag; this call (and every call that reads the table's counters: size, export, growth)
 * returns an error once the flag is set.  Synchronises the stream. */
int kv_check_overflow(kv_table* t, kv_stream stream);
/* InitKvVariableOp -> KvVariable::InitRandomValues, kernels/kv_variable.h:184-206:
 * copies d_table[rows, dim]; only the first call takes effect. */
int kv_set_init_table(kv_table* t, const float* d_table, int64_t rows, kv_stream stream);
/* KvVariableIsInitializedOp, kernels/kv_variable_ops.cc:214-237. */
int kv_is_initialized(const kv_table* t, int* out);
int kv_init_table_rows(const kv_table* t, int64_t* rows);
int kv_get_init_table(const kv_table* t, float* d_out, kv_stream stream);
/* KvVariableSizeOp / KvVariableFrequencyOp / KvVariableShapeOp,
 * kernels/kv_variable_ops.cc:239-293,159-186 -> kv_variable.h:139-182.
 * size = keys not blacklisted with freq >= enter_threshold; map_size = all
 * keys (shape[0]).  Synchronise `stream`. */
int kv_size(kv_table* t, kv_stream stream, int64_t* out);
int kv_sum_freq(kv_table* t, kv_stream stream, int64_t* out);
int kv_map_size(kv_table* t, kv_stream stream, int64_t* out);

/* ---- lookups ------------------------------------------------------------ */

/* KvVariableGatherOrInsertOp / ...WithCountsOp, kernels/kv_variable_ops.cc:498-631
 * -> KvVariable::FindOrInsert kernels/kv_variable.h:263-380.
 * d_ids[n] int64, d_counts[n] int32 or NULL, d_out[n, dim].  `today` is
 * time()/86400 truncated to 16 bits (kernels/utility.cc:38-40), injected. */
int kv_gather_or_insert(kv_table* t, const int64_t* d_ids, const int32_t* d_counts,
                        int64_t n, float* d_out, uint16_t today, kv_stream stream);
/* kv_gather_or_insert over the first min(n, *d_n) ids: the count comes from a kernel earlier on
 * the stream (kv_unique's d_num_unique), as in embedding_lookup's unique -> gather
 * (python/ops/embedding_ops.py:365-372), without a host round trip.  Rows past the count are
 * left untouched. */
int kv_gather_or_insert_n(kv_table* t, const int64_t* d_ids, const int32_t* d_counts, int64_t n,
                          const int32_t* d_n, float* d_out, uint16_t today, kv_stream stream);
/* KvVariableGatherOrZerosOp, kernels/kv_variable_ops.cc:348-429 ->
 * KvVariable::FindOrZeros kernels/kv_variable.h:239-254. */
int kv_gather_or_zeros(kv_table* t, const int64_t* d_ids, int64_t n, float* d_out,
                       kv_stream stream);
/* KvVariableInsertOp, kernels/kv_variable_ops.cc:703-747 ->
 * KvVariable::InsertOrUpdate kernels/kv_variable.h:423-485.  d_filter_out /
 * d_blacklist are uint8[n] masks or NULL.  Ids must be unique. */
int kv_insert_or_update(kv_table* t, const int64_t* d_ids, const float* d_values,
                        int64_t n, const uint8_t* d_filter_out,
                        const uint8_t* d_blacklist, kv_stream stream);
/* KvVariableScatterUpdateOP<op>, kernels/kv_variable_ops.cc:1097-1163 ->
 * KvVariable::ScatterUpdate kernels/kv_variable.h:616-734.  The ids may repeat
 * (GradientDescentOptimizer._resource_apply_sparse_duplicate_indices calls scatter_add with
 * the raw indices): every occurrence is applied, those of one key one after the other in
 * index order.  The call dedups the ids internally (a dedup plan kept with the table, sized on
 * first use - not under CUDA-graph capture). */
int kv_scatter(kv_table* t, int op, const int64_t* d_ids, const float* d_updates,
               int64_t n, kv_stream stream);
/* The same for ids the caller KNOWS to be distinct (after TF's _deduplicate_indexed_slices:
 * the tfplus-Adam path, python/training/adam.py:131,157): no dedup pass. */
int kv_scatter_unique(kv_table* t, int op, const int64_t* d_ids, const float* d_updates,
                      int64_t n, kv_stream stream);
/* KvVariable::GetCount / GetTimeStamp, kernels/kv_variable.h:503-561 (ops
 * KvVariableGetCountV2 / KvVariableGetTimeStamp have no kernel in the OSS tree). */
int kv_get_count(kv_table* t, const int64_t* d_ids, int64_t n, int32_t* d_out,
                 kv_stream stream);
int kv_get_timestamp(kv_table* t, const int64_t* d_ids, int64_t n, uint32_t* d_out,
                     uint16_t today, kv_stream stream);

/* ---- fused sparse optimizer applies -------------------------------------
 * d_ids[n] must be unique (what TF's _deduplicate_indexed_slices hands the op).
 * If d_n is non-NULL the number of valid ids is read from *d_n on the device
 * (<= n), which lets unique -> segment_sum -> apply run without a host sync. */

/* KvVariableSparseApplyAdagradOp, kernels/training_ops.cc:1372-1520. */
int kv_apply_adagrad(kv_table* var, kv_table* accum, const int64_t* d_ids,
                     const float* d_grad, int64_t n, const int32_t* d_n, float lr,
                     int update_slots, uint16_t today, kv_stream stream);
/* KvVariableGroupSparseApplyAdamV4Op, kernels/training_ops.cc:6980-7235.
 * m_v_linear has dim 3*dim(var) = [m | v | linear]. */
int kv_apply_group_adam_v4(kv_table* var, kv_table* m_v_linear, const int64_t* d_ids,
                           const float* d_grad, int64_t n, const int32_t* d_n,
                           float lr, float beta1_power, float beta2_power,
                           float beta1, float beta2, float epsilon, float l1,
                           float l2, float l21, uint16_t today, kv_stream stream);
/* KvVariableSparseGroupSparseApplyFtrlOp<has_l2_shrinkage=true>,
 * kernels/training_ops.cc:532-801. */
int kv_apply_sparse_group_ftrl(kv_table* var, kv_table* accum, kv_table* linear,
                               const int64_t* d_ids, const float* d_grad, int64_t n,
                               const int32_t* d_n, float lr, float l1, float l2,
                               float l21, float l2_shrinkage, float lr_power,
                               uint16_t today, kv_stream stream);
/* tfplus AdamOptimizer._tfplus_apply_sparse_shared, python/training/adam.py:93-163,
 * concatenated slot m_v = [m | v] (dim 2*dim(var)): the reference runs
 * gather(m_v) + TF elementwise ops + scatter_update(m_v) + scatter_sub(var);
 * this is the same arithmetic, each op rounded separately, in one pass. */
int kv_apply_adam(kv_table* var, kv_table* m_v, const int64_t* d_ids,
                  const float* d_grad, int64_t n, const int32_t* d_n, float lr,
                  float beta1, float beta2, float epsilon, float beta1_power,
                  float beta2_power, uint16_t today, kv_stream stream);

/* The same four ops with their scalar inputs left in DEVICE memory (a TF GPU
 * kernel sees lr, beta1_power, ... as device tensors unless they are pinned
 * to HostMemory).  d_hp holds the op's scalar inputs in op order:
 *   adagrad [lr]; group_adam_v4 [lr, beta1_power, beta2_power, beta1, beta2,
 *   epsilon, l1, l2, l21]; sparse_group_ftrl [lr, l1, l2, l21, l2_shrinkage,
 *   lr_power]; adam [lr, beta1, beta2, epsilon, beta1_power, beta2_power].
 * Nothing is read back, so a whole step can be captured in a CUDA graph and
 * replayed while beta^t advances on the device; the sign checks of the host
 * variants are skipped. */
int kv_apply_adagrad_dev(kv_table* var, kv_table* accum, const int64_t* d_ids,
                         const float* d_grad, int64_t n, const int32_t* d_n,
                         const float* d_hp, int update_slots, uint16_t today,
                         kv_stream stream);
int kv_apply_group_adam_v4_dev(kv_table* var, kv_table* m_v_linear, const int64_t* d_ids,
                               const float* d_grad, int64_t n, const int32_t* d_n,
                               const float* d_hp, uint16_t today, kv_stream stream);
int kv_apply_sparse_group_ftrl_dev(kv_table* var, kv_table* accum, kv_table* linear,
                                   const int64_t* d_ids, const float* d_grad, int64_t n,
                                   const int32_t* d_n, const float* d_hp, uint16_t today,
                                   kv_stream stream);
int kv_apply_adam_dev(kv_table* var, kv_table* m_v, const int64_t* d_ids,
                      const float* d_grad, int64_t n, const int32_t* d_n,
                      const float* d_hp, uint16_t today, kv_stream stream);
/* The `_dev` apply plus AdamOptimizer._finish in the same launch: once every row is updated,
 * beta1_power *= beta1 and beta2_power *= beta2 inside d_hp (group_adam.py inherits _finish
 * from tf.train.AdamOptimizer; python/training/adam.py keeps the powers as non-slot variables),
 * so a training step needs no separate scalar-update op between two applies.  The election of
 * the last block uses a counter of the var table: launches on one var table must not overlap
 * (they never do on one stream; callers using several streams order them with events). */
int kv_apply_group_adam_v4_dev_advance(kv_table* var, kv_table* m_v_linear, const int64_t* d_ids,
                                       const float* d_grad, int64_t n, const int32_t* d_n,
                                       float* d_hp, uint16_t today, kv_stream stream);
int kv_apply_adam_dev_advance(kv_table* var, kv_table* m_v, const int64_t* d_ids,
                              const float* d_grad, int64_t n, const int32_t* d_n, float* d_hp,
                              uint16_t today, kv_stream stream);

/* ---- optimizer variants that share the fused apply ------------------------- */

/* KvVariableGroupSparseApplyAdamV3Op, kernels/training_ops.cc:5710-5965: as V4 but lr divides
 * the curvature terms instead of scaling alpha / l1 / l2 / l21 (:5846-5849, :5893-5925). */
int kv_apply_group_adam_v3(kv_table* var, kv_table* m_v_linear, const int64_t* d_ids,
                           const float* d_grad, int64_t n, const int32_t* d_n,
                           float lr, float beta1_power, float beta2_power,
                           float beta1, float beta2, float epsilon, float l1,
                           float l2, float l21, uint16_t today, kv_stream stream);
/* KvVariableSparseApplyFtrlV2 (KvVariableSparseApplyFtrlOp<has_l2_shrinkage=true>),
 * kernels/training_ops.cc:281-530: FTRL-proximal per row, no group lasso, no blacklist. */
int kv_apply_sparse_ftrl_v2(kv_table* var, kv_table* accum, kv_table* linear,
                            const int64_t* d_ids, const float* d_grad, int64_t n,
                            const int32_t* d_n, float lr, float l1, float l2,
                            float l2_shrinkage, float lr_power, uint16_t today,
                            kv_stream stream);
/* KvVariableGroupSparseApplyFtrlV2, kernels/training_ops.cc:805-1062: group lasso on the
 * norm of `linear` itself with threshold l1 (:976-1013). */
int kv_apply_group_sparse_ftrl_v2(kv_table* var, kv_table* accum, kv_table* linear,
                                  const int64_t* d_ids, const float* d_grad, int64_t n,
                                  const int32_t* d_n, float lr, float l1, float l2,
                                  float l2_shrinkage, float lr_power, uint16_t today,
                                  kv_stream stream);

/* ---- dedup plan: unique_with_counts + occurrence lists of one id batch -------
 * What the hot path derives from the ids ALONE, built once per batch and handed to the
 * lookup and to the fused segment-sum + apply of that batch:
 *   - tf.unique_with_counts(ids): uniq (first-occurrence order), idx (int32 inverse),
 *     counts, num_unique — the Unique that embedding_lookup_sparse places before the
 *     gather (python/ops/embedding_ops.py:365-372) and that TF's
 *     Optimizer._deduplicate_indexed_slices places before the apply
 *     (python/ops/variable_scope.py:1096-1106);
 *   - the occurrences of every distinct id in increasing position (seg_off / pos, a CSR),
 *     which is what lets UnsortedSegmentSum run in TF's order without atomics.
 * It depends on nothing but the ids, so an input pipeline can build the plan of batch t+1
 * while batch t trains.  A plan holds up to `max_ids` ids; buffers are allocated once. */
typedef struct kv_plan kv_plan;
int kv_plan_create(int64_t max_ids, kv_plan** out);
int kv_plan_destroy(kv_plan* plan);
int kv_plan_build(kv_plan* plan, kv_workspace* ws, const int64_t* d_ids, int64_t n,
                  kv_stream stream);
/* Device pointers of the plan's arrays (any may be NULL): uniq[n], idx[n], counts[n],
 * num_unique[1], seg_off[n], pos[n]; valid until the plan is destroyed, contents until the
 * next build. */
int kv_plan_arrays(const kv_plan* plan, const int64_t** d_uniq, const int32_t** d_idx,
                   const int32_t** d_counts, const int32_t** d_num_unique,
                   const int32_t** d_seg_off, const int32_t** d_pos);
/* KvVariableGatherOrInsertV2 over the batch the plan was built from: d_out[n, dim].  Every
 * distinct id is found-or-inserted ONCE with its occurrence count (the reference's own
 * unique -> KvVariableGatherOrInsertWithCounts -> gather chain, embedding_ops.py:365-441), then
 * rows are expanded to all positions; table state and output equal kv_gather_or_insert on the
 * raw ids.  Leaves the slot of every distinct id in the plan for kv_apply_plan. */
int kv_gather_or_insert_plan(kv_table* t, kv_plan* plan, float* d_out, uint16_t today,
                             kv_stream stream);
int kv_gather_or_zeros_plan(kv_table* t, kv_plan* plan, float* d_out, kv_stream stream);
/* tf.math.unsorted_segment_sum(d_data[n, dim], plan.idx, num_unique) into d_out[num_unique, dim],
 * every segment summed in increasing position from +0 (TF's CPU order): deterministic. */
int kv_segment_sum_plan(kv_plan* plan, const float* d_data, int dim, float* d_out,
                        kv_stream stream);

typedef enum {
  KV_OPT_ADAGRAD = 0,              /* hp = [lr] */
  KV_OPT_GROUP_ADAM_V4 = 1,        /* [lr, beta1_power, beta2_power, beta1, beta2, epsilon, l1, l2, l21] */
  KV_OPT_SPARSE_GROUP_FTRL = 2,    /* [lr, l1, l2, l21, l2_shrinkage, lr_power] */
  KV_OPT_ADAM = 3,                 /* [lr, beta1, beta2, epsilon, beta1_power, beta2_power] */
  KV_OPT_GROUP_ADAM_V3 = 4,        /* as V4 */
  KV_OPT_SPARSE_FTRL_V2 = 5,       /* [lr, l1, l2, 0, l2_shrinkage, lr_power] */
  KV_OPT_GROUP_SPARSE_FTRL_V2 = 6  /* [lr, l1, l2, 0, l2_shrinkage, lr_power] */
} kv_optimizer;
/* UnsortedSegmentSum + KvVariable*Apply* of optimizer `kind` in ONE launch:
 * d_grad[n, dim] holds one gradient row per id OCCURRENCE of the plan's batch (the
 * IndexedSlices values as the backward pass produces them); the duplicate rows of an id are
 * summed in TF's order inside the kernel that updates the id's rows.  Equivalent to
 * kv_unique + kv_segment_sum + kv_apply_<kind>; this is what an optimizer's
 * _apply_sparse_duplicate_indices override calls.  slot_b is NULL unless the optimizer has
 * two slot tables (the FTRL family: slot_a = accum, slot_b = linear).  hp = the op's scalar
 * inputs in op order (n_hp of them, see kv_optimizer). */
int kv_apply_plan(int kind, kv_table* var, kv_table* slot_a, kv_table* slot_b, kv_plan* plan,
                  const float* d_grad, const float* hp, int n_hp, int update_slots,
                  uint16_t today, kv_stream stream);
/* Same with the scalar inputs in device memory (CUDA-graph capturable); advance_powers != 0
 * also performs AdamOptimizer._finish (beta^t *= beta) inside d_hp once every row is updated. */
int kv_apply_plan_dev(int kind, kv_table* var, kv_table* slot_a, kv_table* slot_b,
                      kv_plan* plan, const float* d_grad, float* d_hp, int advance_powers,
                      int update_slots, uint16_t today, kv_stream stream);

/* ---- dedup (stock TF ops on the path; TF 2.13 Unique / UnsortedSegmentSum) */

int kv_workspace_create(kv_workspace** out);
int kv_workspace_destroy(kv_workspace* ws);
/* tf.unique / tf.unique_with_counts: d_uniq in first-occurrence order, d_idx
 * int32 inverse, d_counts int32 or NULL, *d_num_unique int32 on the device.
 * Call sites: python/ops/variable_scope.py:1096-1106 (TF
 * _deduplicate_indexed_slices), python/ops/embedding_ops.py:365-372. */
int kv_unique(kv_workspace* ws, const int64_t* d_ids, int64_t n, int64_t* d_uniq,
              int32_t* d_idx, int32_t* d_counts, int32_t* d_num_unique,
              kv_stream stream);
/* tf.math.unsorted_segment_sum(data[n, dim], idx, num_segments): d_out must
 * hold max_segments rows; rows [0, *d_num_segments) are written (all
 * max_segments rows when d_num_segments is NULL).  With accumulate != 0 the
 * sums are added to what d_out already holds (the caller zeroed it, e.g.
 * with kv_zero_rows on another stream while the ids were being deduplicated). */
int kv_segment_sum(kv_workspace* ws, const float* d_data, const int32_t* d_idx,
                   int64_t n, int dim, int64_t max_segments,
                   const int32_t* d_num_segments, float* d_out, int accumulate,
                   kv_stream stream);
/* The combiner of embedding_lookup_sparse (python/ops/embedding_ops.py:403-441) in one launch:
 * d_out[r, :] = combine over the entries i with d_segment_ids[i] == r (sorted, as SparseTensor
 * indices are) of d_emb[d_idx[i], :] * (d_weights ? d_weights[i] : 1); combiner 0 sum, 1 mean
 * (divide by the weight sum / the count), 2 sqrtn (divide by sqrt of the sum of squared weights /
 * sqrt(count)).  d_emb are the rows of the DISTINCT ids, d_idx the inverse index of kv_unique.
 * Entries are added in increasing position (TF's CPU order); rows without entries are zero. */
int kv_sparse_combine(const float* d_emb, const int32_t* d_idx, const int64_t* d_segment_ids,
                      const float* d_weights, int64_t nnz, int64_t n_rows, int dim, int combiner,
                      float* d_out, kv_stream stream);
/* d_out[0 .. min(max_rows, *d_num_rows)) [dim] = 0. */
int kv_zero_rows(float* d_out, int64_t max_rows, const int32_t* d_num_rows, int dim,
                 kv_stream stream);

/* ---- checkpoint ---------------------------------------------------------- */

/* KvVariable::ExportValues, kernels/dynamic_save.hpp:48-195, in two steps
 * because the caller owns the output buffers: kv_export_count applies the
 * under-threshold refresh (kernels/kv_variable.h:995-1012), counts, and
 * synchronises `stream`; kv_export fills buffers of exactly those sizes (row
 * order is unspecified, as in the reference).  d_freq_values is uint32 when
 * freq_u32 else uint16.  Pointers for parts not exported may be NULL. */
int kv_export_count(kv_table* t, int first_n, int enable_cutoff, float cutoff_value,
                    kv_stream stream, int64_t* n_keys, int64_t* n_blacklist,
                    int64_t* n_freq);
int kv_export(kv_table* t, int first_n, int64_t* d_keys, float* d_values,
              int64_t* d_blacklist, int64_t* d_freq_keys, void* d_freq_values,
              int freq_u32, kv_stream stream);
/* The same with the sizes of the caller's buffers: entries beyond a capacity are dropped instead
 * of written (the reference's `key_row < num_rows` guards, dynamic_save.hpp:142-174).  Use this
 * one when another thread may insert between kv_export_count and the export. */
int kv_export_bounded(kv_table* t, int first_n, int64_t* d_keys, float* d_values, int64_t cap_keys,
                      int64_t* d_blacklist, int64_t cap_blacklist, int64_t* d_freq_keys,
                      void* d_freq_values, int64_t cap_freq, int freq_u32, kv_stream stream);
/* KvVariable::ImportValues, kernels/dynamic_restore.hpp:156-262: clears the
 * table, copies the rows (the reference aliases the input tensor), replaces
 * the init table if init_rows > 0, marks the blacklist, overwrites frequency
 * words of keys that exist.  Keys must be unique. */
int kv_import(kv_table* t, const int64_t* d_keys, const float* d_values, int64_t n,
              const float* d_init_table, int64_t init_rows,
              const int64_t* d_blacklist, int64_t n_blacklist,
              const int64_t* d_freq_keys, const void* d_freq_values, int64_t n_freq,
              int freq_u32, kv_stream stream);

/* ---- eviction ------------------------------------------------------------ */

/* ---- delta checkpoints (online learning) -----------------------------------
 * KvVariableFullOrDeltaExport / KvVariableFullOrDeltaImport[V2] in delta mode
 * (ops/kv_variable_ops.cc:576-660; python/ops/kv_variable_ops.py:1435-1459,1609-1621) ->
 * KvVariable::DeltaExport kernels/dynamic_save.hpp:197-449 and KvVariable::DeltaImport
 * kernels/dynamic_restore.hpp:28-153.  The reference switches delta tracking on with the
 * environment variables SUPPORT_DELTA_EXPORT / SUPPORT_PREDICTION_DELTA_EXPORT at construction
 * (kernels/kv_variable.h:101-111); here it is a call.  From then on every entry point that the
 * reference marks - GatherOrInsert (:316), InsertOrUpdate (:451), ScatterUpdate (:685), Delete
 * (:747), DeleteWithTimestamp (:772) and the apply ops through MarkAsDeltaListElements
 * (:791-799, for the ids they do not skip) - records its keys in a device-side key set. */
int kv_enable_delta_export(kv_table* t, int support_prediction_delta);
/* Keys marked since the last training-mode delta export (train_deltalist_.size()). */
int kv_delta_size(kv_table* t, kv_stream stream, int64_t* out);
/* Sizes of the four variable-length outputs of a delta export with attr first_n (nothing is
 * consumed): update rows, blacklist, frequency table (first_n > 4), delete_keys. */
int kv_delta_export_count(kv_table* t, int first_n, kv_stream stream, int64_t* n_keys,
                          int64_t* n_blacklist, int64_t* n_freq, int64_t* n_delete);
/* The export proper into caller buffers of the given capacities (entries beyond a capacity
 * are dropped, never written); counts[4] = what the table held.  freq_values are full uint32
 * words (lo16 count, hi16 day).  first_n <= 3 is the inference-mode export (train + prediction
 * sets, blacklisted keys become delete_keys, prediction set cleared); otherwise the train set
 * is moved to the prediction set (if enabled) and cleared. */
int kv_delta_export(kv_table* t, int first_n, int64_t* d_keys, float* d_values, int64_t cap_keys,
                    int64_t* d_blacklist, int64_t cap_blacklist, int64_t* d_freq_keys,
                    uint32_t* d_freq_values, int64_t cap_freq, int64_t* d_delete_keys,
                    int64_t cap_delete, kv_stream stream, int64_t* counts);
/* DeltaImport on top of the current contents: upsert rows (un-blacklisting them), mark /
 * (first_n <= 3) drop the blacklist keys, overwrite the frequency words of keys that exist,
 * delete delete_keys. */
int kv_delta_import(kv_table* t, int first_n, const int64_t* d_keys, const float* d_values,
                    int64_t n, const int64_t* d_blacklist, int64_t n_blacklist,
                    const int64_t* d_freq_keys, const uint32_t* d_freq_values, int64_t n_freq,
                    const int64_t* d_delete_keys, int64_t n_delete, kv_stream stream);

/* KvVariable::Delete, kernels/kv_variable.h:737-753. */
int kv_delete(kv_table* t, const int64_t* d_ids, int64_t n, kv_stream stream);
/* KvVariable::DeleteWithTimestamp, kernels/kv_variable.h:756-789: deletes keys
 * with day > 0 and today - day >= threshold; writes up to `cap` deleted keys to
 * d_out_keys (may be NULL) and the total to *n_deleted.  Synchronises. */
int kv_delete_with_timestamp(kv_table* t, int threshold, uint16_t today,
                             int64_t* d_out_keys, int64_t cap, kv_stream stream,
                             int64_t* n_deleted);

/* ---- multi-GPU routing (key-hash sharding; python/ops/embedding_ops.py:121-204
 * does the same with floormod + dynamic_partition + dynamic_stitch) --------- */

/* Groups ids by owner shard: d_sorted_ids[n] holds the ids of shard 0, then
 * shard 1, ...; d_perm[n] (int32) maps each input position to its position in
 * d_sorted_ids; d_shard_counts[num_shards] (int32).  mode 0: owner =
 * mix64(id) % num_shards, mode 1: floormod(id, num_shards). */
int kv_partition_ids(kv_workspace* ws, const int64_t* d_ids, int64_t n,
                     const int32_t* d_n, int num_shards, int mode,
                     int64_t* d_sorted_ids, int32_t* d_perm, int32_t* d_shard_counts,
                     kv_stream stream);
/* Fixed-capacity variant for a sync-free exchange: d_send_ids / d_send_occ are
 * [num_shards][capacity]; the ids owned by shard g (and their occurrence
 * counts d_occ, or 1) fill [g][0 .. d_counts[g]), the rest is padding
 * (INT64_MIN+1, which lookups answer with zeros and applies skip).  d_perm[i]
 * is the padded position of input i, or -1 if its shard row was full, in which
 * case *d_overflow is set to 1 (it is never cleared here).  Nothing depends on
 * the data, so the all-to-all that follows can be captured in a CUDA graph. */
int kv_route_ids(kv_workspace* ws, const int64_t* d_ids, const int32_t* d_occ, int64_t n,
                 const int32_t* d_n, int num_shards, int mode, int capacity,
                 int64_t* d_send_ids, int32_t* d_send_occ, int32_t* d_perm,
                 int32_t* d_counts, int32_t* d_overflow, kv_stream stream);
/* Same, but ids and occurrence counts travel together as interleaved
 * {int64 id, int64 count} pairs in d_send_pairs[num_shards][capacity][2], so
 * one exchange carries both; kv_unzip_pairs splits what was received. */
int kv_route_id_pairs(kv_workspace* ws, const int64_t* d_ids, const int32_t* d_occ, int64_t n,
                      const int32_t* d_n, int num_shards, int mode, int capacity,
                      int64_t* d_send_pairs, int32_t* d_perm, int32_t* d_counts,
                      int32_t* d_overflow, kv_stream stream);
int kv_unzip_pairs(const int64_t* d_pairs, int64_t n, int64_t* d_ids, int32_t* d_occ,
                   kv_stream stream);
/* out[i, :] = src[perm[idx[i]], :]; perm and/or idx may be NULL (identity); a
 * negative perm entry yields zeros. */
int kv_expand_rows(const float* d_src, const int32_t* d_perm, const int32_t* d_idx, int64_t n,
                   int dim, float* d_out, kv_stream stream);
/* out[perm[i], :] = src[i, :] for i < min(n, *d_n); negative perm entries are skipped. */
int kv_scatter_rows_n(const float* d_src, const int32_t* d_perm, int64_t n, const int32_t* d_n,
                      int dim, float* d_out, kv_stream stream);
/* out[i, :] = src[perm[i], :] (gather rows back into request order). */
int kv_permute_rows(const float* d_src, const int32_t* d_perm, int64_t n, int dim,
                    float* d_out, kv_stream stream);
/* out[perm[i], :] = src[i, :]. */
int kv_scatter_rows(const float* d_src, const int32_t* d_perm, int64_t n, int dim,
                    float* d_out, kv_stream stream);

/* ---- Peer-memory (NVLink / NVSwitch) variants: compute fused with its exchange ----
 * The three exchanges of the sharded step as stores of the producing kernel into
 * buffers that live on the PEER GPUs (memory the host maps into this process:
 * symmetric memory / cudaIpc; the library only sees device pointers).  A
 * `d_seg` argument is a DEVICE array of num_shards device pointers; segment g
 * holds `capacity` entries and normally points at peer g's buffer + rank * capacity.
 * Replaces the PS<->worker send/recv around the ops of SURVEY.md §8e.          */
/* kv_route_ids whose ids / occurrence counts go straight to d_seg_ids[g][0..capacity) /
 * d_seg_occ[g][0..capacity) (padding included). */
int kv_route_ids_peer(kv_workspace* ws, const int64_t* d_ids, const int32_t* d_occ, int64_t n,
                      const int32_t* d_n, int num_shards, int mode, int capacity,
                      int64_t* const* d_seg_ids, int32_t* const* d_seg_occ, int32_t* d_perm,
                      int32_t* d_counts, int32_t* d_overflow, kv_stream stream);
/* kv_unique and kv_route_ids_peer in the same three launches: the kernel that ranks the first
 * occurrences also gives each distinct id a position in its owner's row and stores {id,
 * occurrence count} there.  d_perm[r] is the padded position of unique id r (or -1 and
 * *d_overflow = 1).  The rows must have been padded, and d_shard_counts zeroed, by
 * kv_route_fill_peer since the previous call (it can run any time after the owners have read
 * the previous step's ids, e.g. under the rest of the step). */
int kv_unique_route_peer(kv_workspace* ws, const int64_t* d_ids, int64_t n, int64_t* d_uniq,
                         int32_t* d_idx, int32_t* d_counts, int32_t* d_num_unique,
                         int num_shards, int mode, int capacity, int64_t* const* d_seg_ids,
                         int32_t* const* d_seg_occ, int32_t* d_perm, int32_t* d_shard_counts,
                         int32_t* d_overflow, kv_stream stream);
int kv_route_fill_peer(int num_shards, int capacity, int64_t* const* d_seg_ids,
                       int32_t* const* d_seg_occ, int32_t* d_shard_counts, kv_stream stream);
/* kv_gather_or_insert whose row r is written to d_seg_rows[r / capacity] + (r % capacity) * dim;
 * rows of padding ids are not written at all. */
int kv_gather_or_insert_peer(kv_table* t, const int64_t* d_ids, const int32_t* d_counts,
                             int64_t n, float* const* d_seg_rows, int64_t capacity,
                             uint16_t today, kv_stream stream);
/* kv_scatter_rows_n whose destination row p is d_seg_rows[p / capacity] + (p % capacity) * dim. */
int kv_scatter_rows_n_peer(const float* d_src, const int32_t* d_perm, int64_t n,
                           const int32_t* d_n, int dim, float* const* d_seg_rows,
                           int64_t capacity, kv_stream stream);
/* Barrier among the `world` GPUs, as one kernel on `stream` (CUDA-graph capturable):
 * everything the peers stored before their call is visible after it.  d_peer_flags is a
 * device array of world pointers to every rank's flag array (uint32[world], zero-initialised,
 * peer-mapped); d_my_flags is this rank's own; d_state is local uint32[2], zero-initialised:
 * [0] counts completed barriers, [1] counts waits given up after timeout_ms. */
int kv_peer_barrier(uint32_t* const* d_peer_flags, uint32_t* d_my_flags, uint32_t* d_state,
                    int rank, int world, int64_t timeout_ms, kv_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* KVHBM_H_ */
