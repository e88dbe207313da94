// kv_variable_gpu_ops.cc — TensorFlow 2.13 DEVICE_GPU OpKernels for the KvVariable ops,
// each a thin call into the C ABI of libkvhbm.so (include/kvhbm.h) on the op's CUDA stream.
//
// This file is the reference-side binding: build it INSTEAD of the reference's
// kernels/kv_variable_ops.cc + kernels/training_ops.cc, together with the reference's own
// REGISTER_OP files (ops/kv_variable_ops.cc, ops/training_ops.cc — op names, attrs and input
// order stay byte-for-byte the reference's, including the `beat1` input name), into
// python/ops/_kv_variable_ops.so (see INTEGRATION.md for the Bazel / CMake lines).  The Python
// layer (get_kv_variable, embedding_lookup, the optimizers) is used unchanged.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no TensorFlow headers.  It depends
// only on the public OpKernel API and on kvhbm.h.
#include <string>
#include <vector>

#include "kvhbm.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/resource_mgr.h"
#include "tensorflow/core/framework/resource_op_kernel.h"
#include "tensorflow/core/platform/env.h"
#include "tensorflow/core/platform/mutex.h"
#include "tensorflow/core/util/gpu_device_functions.h"

namespace tfplus_b200 {
using namespace tensorflow;  // NOLINT

// The resource behind a KvVariable handle: owns one device table.
class KvHbmVariable : public ResourceBase {
 public:
  KvHbmVariable(const std::string& name, const TensorShape& value_shape, kv_table* t)
      : name_(name), value_shape_(value_shape), table_(t) {}
  ~KvHbmVariable() override { kv_destroy(table_); }
  std::string DebugString() const override { return name_; }
  kv_table* table() const { return table_; }
  const TensorShape& value_shape() const { return value_shape_; }
  const std::string& name() const { return name_; }

 private:
  std::string name_;
  TensorShape value_shape_;
  kv_table* table_;
};

static Status FromKv(int code) {
  if (code == KV_OK) return OkStatus();
  const std::string msg = kv_last_error();
  switch (code) {
    case KV_INVALID_ARGUMENT: return errors::InvalidArgument(msg);
    case KV_FAILED_PRECONDITION: return errors::FailedPrecondition(msg);
    case KV_UNIMPLEMENTED: return errors::Unimplemented(msg);
    case KV_RESOURCE_EXHAUSTED: return errors::ResourceExhausted(msg);
    default: return errors::Internal(msg);
  }
}
static kv_stream StreamOf(OpKernelContext* ctx) {
  return reinterpret_cast<kv_stream>(ctx->eigen_gpu_device().stream());
}
static uint16_t Today() {  // kernels/utility.cc:38-40
  return static_cast<uint16_t>(Env::Default()->NowSeconds() / 86400);
}
static Status Lookup(OpKernelContext* ctx, int input, KvHbmVariable** v) {
  return LookupResource(ctx, HandleFromInput(ctx, input), v);
}

// ---- KvVariable / KvVariableV2..V4: kernels/kv_variable_ops.cc:31-147 --------------------
class CreateKvVariableOp : public OpKernel {
 public:
  explicit CreateKvVariableOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("container", &container_));
    OP_REQUIRES_OK(c, c->GetAttr("shared_name", &shared_name_));
    OP_REQUIRES_OK(c, c->GetAttr("use_node_name_sharing", &node_name_sharing_));
    OP_REQUIRES_OK(c, c->GetAttr("key_dtype", &key_dtype_));
    OP_REQUIRES_OK(c, c->GetAttr("value_dtype", &value_dtype_));
    OP_REQUIRES_OK(c, c->GetAttr("value_shape", &value_shape_));
    OP_REQUIRES_OK(c, c->GetAttr("enter_threshold", &enter_threshold_));
    OP_REQUIRES(c, key_dtype_ == DT_INT64 && value_dtype_ == DT_FLOAT,
                errors::Unimplemented("kvhbm: KvVariable is int64 -> float32 on the GPU"));
  }
  void Compute(OpKernelContext* ctx) override {
    mutex_lock l(mu_);
    if (!initialized_) {
      const std::string name = shared_name_.empty() || node_name_sharing_ ? def().name() : shared_name_;
      KvHbmVariable* v = nullptr;
      OP_REQUIRES_OK(ctx, ctx->resource_manager()->LookupOrCreate<KvHbmVariable>(
                              container_.empty() ? ctx->resource_manager()->default_container() : container_,
                              name, &v, [&](KvHbmVariable** out) {
                                kv_table* t = nullptr;
                                TF_RETURN_IF_ERROR(FromKv(kv_create(
                                    static_cast<int>(value_shape_.num_elements()), enter_threshold_, 0, &t)));
                                *out = new KvHbmVariable(name, value_shape_, t);
                                return OkStatus();
                              }));
      core::ScopedUnref unref(v);
      handle_ = MakeResourceHandle<KvHbmVariable>(ctx, container_, name);
      initialized_ = true;
    }
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({}), &out));
    out->scalar<ResourceHandle>()() = handle_;
  }

 private:
  mutex mu_;
  bool initialized_ = false;
  ResourceHandle handle_;
  std::string container_, shared_name_;
  bool node_name_sharing_ = false;
  DataType key_dtype_, value_dtype_;
  TensorShape value_shape_;
  int enter_threshold_ = 0;
};
#define REGISTER_CREATE(NAME) \
  REGISTER_KERNEL_BUILDER(Name(NAME).Device(DEVICE_GPU).HostMemory("table_handle"), CreateKvVariableOp)
REGISTER_CREATE("KvVariable");
REGISTER_CREATE("KvVariableV2");
REGISTER_CREATE("KvVariableV3");
REGISTER_CREATE("KvVariableV4");

// ---- InitKvVariableV2: kernels/kv_variable_ops.cc:188-212 --------------------------------------
class InitKvVariableOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &v));
    core::ScopedUnref unref(v);
    const Tensor& tbl = ctx->input(1);  // device memory
    OP_REQUIRES_OK(ctx, FromKv(kv_set_init_table(v->table(), tbl.flat<float>().data(), tbl.dim_size(0),
                                                 StreamOf(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("InitKvVariableV2").Device(DEVICE_GPU).HostMemory("table_handle"),
                        InitKvVariableOp);

// ---- scalar gauges: kernels/kv_variable_ops.cc:159-293 -----------------------------------------
template <int WHICH>
class KvGaugeOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v = nullptr;
    Status s = Lookup(ctx, 0, &v);
    if (WHICH == 0) {  // IsInitialized: a failed lookup is "false", :226-229
      Tensor* out;
      OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({}), &out));
      int init = 0;
      if (s.ok()) { core::ScopedUnref unref(v); kv_is_initialized(v->table(), &init); }
      out->scalar<bool>()() = init != 0;
      return;
    }
    OP_REQUIRES_OK(ctx, s);
    core::ScopedUnref unref(v);
    int64_t val = 0;
    if (WHICH == 1) OP_REQUIRES_OK(ctx, FromKv(kv_size(v->table(), StreamOf(ctx), &val)));
    if (WHICH == 2) OP_REQUIRES_OK(ctx, FromKv(kv_sum_freq(v->table(), StreamOf(ctx), &val)));
    if (WHICH == 3) {  // shape = [map size] + value_shape
      OP_REQUIRES_OK(ctx, FromKv(kv_map_size(v->table(), StreamOf(ctx), &val)));
      Tensor* out;
      OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({1 + v->value_shape().dims()}), &out));
      auto o = out->flat<int64_t>();
      o(0) = val;
      for (int d = 0; d < v->value_shape().dims(); ++d) o(d + 1) = v->value_shape().dim_size(d);
      return;
    }
    Tensor* out;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({}), &out));
    out->scalar<int64_t>()() = val;
  }
};
#define REGISTER_GAUGE(NAME, W, OUT) \
  REGISTER_KERNEL_BUILDER(Name(NAME).Device(DEVICE_GPU).HostMemory("table_handle").HostMemory(OUT), KvGaugeOp<W>)
REGISTER_GAUGE("KvVariableIsInitializedV2", 0, "is_initialized");
REGISTER_GAUGE("KvVariableSizeV2", 1, "output");
REGISTER_GAUGE("KvVariableFrequency", 2, "output");
REGISTER_GAUGE("KvVariableShapeV2", 3, "output");

class DestroyKvVariableOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext* ctx) override {
    OP_REQUIRES_OK(ctx, DeleteResource(ctx, HandleFromInput(ctx, 0)));
  }
};
REGISTER_KERNEL_BUILDER(Name("DestroyKvVariableOpV2").Device(DEVICE_GPU).HostMemory("table_handle"),
                        DestroyKvVariableOp);

// ---- gathers: kernels/kv_variable_ops.cc:348-631 --------------------------------------------------
template <int MODE>  // 0 zeros, 1 insert, 2 insert with counts
class KvGatherOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &v));
    core::ScopedUnref unref(v);
    const Tensor& ids = ctx->input(1);
    TensorShape shape = ids.shape();  // result = indices.shape + value_shape, :515-524
    shape.AppendShape(v->value_shape());
    Tensor* out;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, shape, &out));
    const int64_t n = ids.NumElements();
    if (n == 0) return;
    const int64_t* d_ids = reinterpret_cast<const int64_t*>(ids.flat<int64_t>().data());
    float* d_out = out->flat<float>().data();
    if (MODE == 0) {
      OP_REQUIRES_OK(ctx, FromKv(kv_gather_or_zeros(v->table(), d_ids, n, d_out, StreamOf(ctx))));
      return;
    }
    const int32_t* d_counts = nullptr;
    if (MODE == 2) {
      const Tensor& counts = ctx->input(2);
      OP_REQUIRES(ctx, counts.shape() == ids.shape(),
                  errors::InvalidArgument("KvVariable ", v->name(), ": increment count, indices shape ",
                                          ids.shape().DebugString(), " does not match with counts shape ",
                                          counts.shape().DebugString()));
      d_counts = counts.flat<int32>().data();
    }
    OP_REQUIRES_OK(ctx, FromKv(kv_gather_or_insert(v->table(), d_ids, d_counts, n, d_out, Today(),
                                                   StreamOf(ctx))));
  }
};
#define REGISTER_GATHER(NAME, M)                                                                   \
  REGISTER_KERNEL_BUILDER(Name(NAME).Device(DEVICE_GPU).HostMemory("table_handle")                 \
                              .TypeConstraint<float>("dtype").TypeConstraint<int64_t>("Tindices"), \
                          KvGatherOp<M>)
REGISTER_GATHER("KvVariableGatherOrZerosV2", 0);
REGISTER_GATHER("KvVariableGatherOrInsertV2", 1);
REGISTER_GATHER("KvVariableGatherOrInsertWithCounts", 2);

// ---- InsertV2 and the seven scatter ops: kernels/kv_variable_ops.cc:703-747,1097-1163 -------------
template <int OP>  // -1 = KvVariableInsertV2, else kv_scatter_op
class KvScatterOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &v));
    core::ScopedUnref unref(v);
    const Tensor& ids = ctx->input(1);
    const Tensor& upd = ctx->input(2);
    const int64_t n = ids.NumElements();
    if (n == 0) return;
    const int64_t* d_ids = reinterpret_cast<const int64_t*>(ids.flat<int64_t>().data());
    const int rc = OP < 0 ? kv_insert_or_update(v->table(), d_ids, upd.flat<float>().data(), n, nullptr,
                                                nullptr, StreamOf(ctx))
                          : kv_scatter(v->table(), OP, d_ids, upd.flat<float>().data(), n, StreamOf(ctx));
    OP_REQUIRES_OK(ctx, FromKv(rc));
  }
};
#define REGISTER_SCATTER(NAME, OP)                                                                 \
  REGISTER_KERNEL_BUILDER(Name(NAME).Device(DEVICE_GPU).HostMemory("table_handle")                 \
                              .TypeConstraint<float>("dtype").TypeConstraint<int64_t>("Tindices"), \
                          KvScatterOp<OP>)
REGISTER_SCATTER("KvVariableInsertV2", -1);
REGISTER_SCATTER("KvVariableScatterUpdateV2", KV_SCATTER_ASSIGN);
REGISTER_SCATTER("KvVariableScatterAddV2", KV_SCATTER_ADD);
REGISTER_SCATTER("KvVariableScatterSubV2", KV_SCATTER_SUB);
REGISTER_SCATTER("KvVariableScatterMulV2", KV_SCATTER_MUL);
REGISTER_SCATTER("KvVariableScatterDivV2", KV_SCATTER_DIV);
REGISTER_SCATTER("KvVariableScatterMinV2", KV_SCATTER_MIN);
REGISTER_SCATTER("KvVariableScatterMaxV2", KV_SCATTER_MAX);

// ---- checkpoint: kernels/kv_variable_ops.cc:779-851,990-1017,325-346 --------------------------------
class KvExportOp : public OpKernel {
 public:
  explicit KvExportOp(OpKernelConstruction* c) : OpKernel(c) {
    if (!c->GetAttr("first_n", &first_n_).ok()) first_n_ = 2;  // ReadKvVariableOpV2
    if (!c->GetAttr("enable_cutoff", &enable_cutoff_).ok()) enable_cutoff_ = false;
    if (!c->GetAttr("cutoff_value", &cutoff_value_).ok()) cutoff_value_ = 0.f;
  }
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &v));
    core::ScopedUnref unref(v);
    int64_t nk = 0, nb = 0, nf = 0, rows = 0;
    OP_REQUIRES_OK(ctx, FromKv(kv_export_count(v->table(), first_n_, enable_cutoff_, cutoff_value_,
                                               StreamOf(ctx), &nk, &nb, &nf)));
    TensorShape vs = v->value_shape();
    vs.InsertDim(0, nk);
    Tensor *keys, *values, *init = nullptr, *black = nullptr, *fkeys = nullptr, *fvals = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({nk}), &keys));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, vs, &values));
    if (first_n_ > 2) {
      kv_init_table_rows(v->table(), &rows);
      TensorShape is = v->value_shape();
      is.InsertDim(0, first_n_ > 3 ? rows : 0);
      OP_REQUIRES_OK(ctx, ctx->allocate_output(2, is, &init));
      OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({nb}), &black));
      OP_REQUIRES_OK(ctx, ctx->allocate_output(4, TensorShape({nf}), &fkeys));
      OP_REQUIRES_OK(ctx, ctx->allocate_output(5, TensorShape({nf}), &fvals));
      if (first_n_ > 3 && rows > 0)
        OP_REQUIRES_OK(ctx, FromKv(kv_get_init_table(v->table(), init->flat<float>().data(), StreamOf(ctx))));
    }
    const bool u32 = fvals && fvals->dtype() == DT_UINT32;  // dynamic_save.hpp:139-140
    OP_REQUIRES_OK(ctx, FromKv(kv_export(
        v->table(), first_n_, reinterpret_cast<int64_t*>(keys->flat<int64_t>().data()),
        values->flat<float>().data(),
        black ? reinterpret_cast<int64_t*>(black->flat<int64_t>().data()) : nullptr,
        fkeys ? reinterpret_cast<int64_t*>(fkeys->flat<int64_t>().data()) : nullptr,
        fvals ? const_cast<char*>(fvals->tensor_data().data()) : nullptr, u32, StreamOf(ctx))));
  }

 private:
  int first_n_;
  bool enable_cutoff_;
  float cutoff_value_;
};
REGISTER_KERNEL_BUILDER(Name("KvVariableExport").Device(DEVICE_GPU).HostMemory("table_handle"), KvExportOp);
REGISTER_KERNEL_BUILDER(Name("ReadKvVariableOpV2").Device(DEVICE_GPU).HostMemory("table_handle"), KvExportOp);

class KvImportOp : public OpKernel {
 public:
  explicit KvImportOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("first_n", &first_n_)); }
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &v));
    core::ScopedUnref unref(v);
    const Tensor &keys = ctx->input(1), &values = ctx->input(2), &init = ctx->input(3);
    const Tensor &black = ctx->input(4), &fk = ctx->input(5), &fv = ctx->input(6);
    const bool use_black = first_n_ > 3, use_freq = first_n_ > 4;  // :806-822
    OP_REQUIRES_OK(ctx, FromKv(kv_import(
        v->table(), reinterpret_cast<const int64_t*>(keys.flat<int64_t>().data()),
        values.flat<float>().data(), keys.NumElements(), init.flat<float>().data(),
        init.NumElements() ? init.dim_size(0) : 0,
        reinterpret_cast<const int64_t*>(black.flat<int64_t>().data()), use_black ? black.NumElements() : 0,
        reinterpret_cast<const int64_t*>(fk.flat<int64_t>().data()), fv.tensor_data().data(),
        use_freq ? fk.NumElements() : 0, fv.dtype() == DT_UINT32, StreamOf(ctx))));
  }

 private:
  int first_n_;
};
REGISTER_KERNEL_BUILDER(Name("KvVariableImport").Device(DEVICE_GPU).HostMemory("table_handle"), KvImportOp);

// ---- fused sparse applies: kernels/training_ops.cc:532-801,1372-1520,6980-7235 -----------------------
// The scalar hyper-parameter inputs are pinned to host memory so that they can be validated as
// the reference validates them; registering them in device memory and calling the *_dev entry
// points instead removes the host read (and lets TF capture the step in a CUDA graph).
class KvAdagradOp : public OpKernel {
 public:
  explicit KvAdagradOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("update_slots", &update_slots_));
  }
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable *var, *acc;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &var));
    core::ScopedUnref u0(var);
    OP_REQUIRES_OK(ctx, Lookup(ctx, 1, &acc));
    core::ScopedUnref u1(acc);
    const Tensor &lr = ctx->input(2), &grad = ctx->input(3), &ids = ctx->input(4);
    OP_REQUIRES(ctx, TensorShapeUtils::IsVector(ids.shape()), errors::InvalidArgument("indices must be one-dimensional"));
    OP_REQUIRES(ctx, grad.dim_size(0) == ids.dim_size(0),
                errors::InvalidArgument("grad must be the same size as indices in the first dimension."));
    OP_REQUIRES_OK(ctx, FromKv(kv_apply_adagrad(
        var->table(), acc->table(), reinterpret_cast<const int64_t*>(ids.flat<int64_t>().data()),
        grad.flat<float>().data(), ids.dim_size(0), nullptr, lr.scalar<float>()(), update_slots_, Today(),
        StreamOf(ctx))));
  }

 private:
  bool update_slots_;
};
REGISTER_KERNEL_BUILDER(Name("KvVariableSparseApplyAdagrad").Device(DEVICE_GPU).HostMemory("var")
                            .HostMemory("accum").HostMemory("lr").TypeConstraint<float>("T")
                            .TypeConstraint<int64_t>("Tindices"), KvAdagradOp);

class KvGroupAdamV4Op : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable *var, *mvl;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &var));
    core::ScopedUnref u0(var);
    OP_REQUIRES_OK(ctx, Lookup(ctx, 1, &mvl));
    core::ScopedUnref u1(mvl);
    const Tensor &grad = ctx->input(2), &ids = ctx->input(3);
    float s[9];  // lr, beta1_power, beta2_power, beat1, beta2, epsilon, l1, l2, l21
    for (int i = 0; i < 9; ++i) {
      OP_REQUIRES(ctx, TensorShapeUtils::IsScalar(ctx->input(4 + i).shape()),
                  errors::InvalidArgument("hyper-parameter ", i, " is not a scalar"));
      s[i] = ctx->input(4 + i).scalar<float>()();
    }
    OP_REQUIRES(ctx, TensorShapeUtils::IsVector(ids.shape()), errors::InvalidArgument("indices must be one-dimensional"));
    OP_REQUIRES(ctx, grad.dim_size(0) == ids.dim_size(0),
                errors::InvalidArgument("grad must be the same size as indices in the first dimension."));
    OP_REQUIRES_OK(ctx, FromKv(kv_apply_group_adam_v4(
        var->table(), mvl->table(), reinterpret_cast<const int64_t*>(ids.flat<int64_t>().data()),
        grad.flat<float>().data(), ids.dim_size(0), nullptr, s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7],
        s[8], Today(), StreamOf(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("KvVariableGroupSparseApplyAdamV4").Device(DEVICE_GPU).HostMemory("var")
                            .HostMemory("m_v_linear").HostMemory("lr").HostMemory("beta1_power")
                            .HostMemory("beta2_power").HostMemory("beat1").HostMemory("beta2")
                            .HostMemory("epsilon").HostMemory("l1").HostMemory("l2").HostMemory("l21")
                            .TypeConstraint<float>("T").TypeConstraint<int64_t>("Tindices"), KvGroupAdamV4Op);

class KvSparseGroupFtrlOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable *var, *acc, *lin;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &var));
    core::ScopedUnref u0(var);
    OP_REQUIRES_OK(ctx, Lookup(ctx, 1, &acc));
    core::ScopedUnref u1(acc);
    OP_REQUIRES_OK(ctx, Lookup(ctx, 2, &lin));
    core::ScopedUnref u2(lin);
    const Tensor &grad = ctx->input(3), &ids = ctx->input(4);
    float s[6];  // lr, l1, l2, l21, l2_shrinkage, lr_power
    for (int i = 0; i < 6; ++i) s[i] = ctx->input(5 + i).scalar<float>()();
    OP_REQUIRES_OK(ctx, FromKv(kv_apply_sparse_group_ftrl(
        var->table(), acc->table(), lin->table(), reinterpret_cast<const int64_t*>(ids.flat<int64_t>().data()),
        grad.flat<float>().data(), ids.dim_size(0), nullptr, s[0], s[1], s[2], s[3], s[4], s[5], Today(),
        StreamOf(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("KvVariableSparseGroupSparseApplyFtrlV2").Device(DEVICE_GPU).HostMemory("var")
                            .HostMemory("accum").HostMemory("linear").HostMemory("lr").HostMemory("l1")
                            .HostMemory("l2").HostMemory("l21").HostMemory("l2_shrinkage").HostMemory("lr_power")
                            .TypeConstraint<float>("T").TypeConstraint<int64_t>("Tindices"), KvSparseGroupFtrlOp);

// ---- eviction: ops registered without kernels in the OSS tree (ops/kv_variable_ops.cc:349,681) --
class KvDeleteOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &v));
    core::ScopedUnref unref(v);
    const Tensor& ids = ctx->input(1);
    OP_REQUIRES_OK(ctx, FromKv(kv_delete(v->table(), reinterpret_cast<const int64_t*>(ids.flat<int64_t>().data()),
                                         ids.NumElements(), StreamOf(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("KvVariableDelete").Device(DEVICE_GPU).HostMemory("table_handle"), KvDeleteOp);

// KvVariableGetCountV2 / KvVariableGetTimeStamp (ops/kv_variable_ops.cc:349-358,687-696):
// KvVariable::GetCount / GetTimeStamp, kernels/kv_variable.h:503-561.
template <bool TIMESTAMP>
class KvGetMetaOp : public OpKernel {
 public:
  using OpKernel::OpKernel;
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &v));
    core::ScopedUnref unref(v);
    const Tensor& ids = ctx->input(1);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, ids.shape(), &out));
    const int64_t* d_ids = reinterpret_cast<const int64_t*>(ids.flat<int64_t>().data());
    if (TIMESTAMP)
      OP_REQUIRES_OK(ctx, FromKv(kv_get_timestamp(v->table(), d_ids, ids.NumElements(),
                                                  out->flat<uint32_t>().data(), Today(), StreamOf(ctx))));
    else
      OP_REQUIRES_OK(ctx, FromKv(kv_get_count(v->table(), d_ids, ids.NumElements(),
                                              out->flat<int32_t>().data(), StreamOf(ctx))));
  }
};
REGISTER_KERNEL_BUILDER(Name("KvVariableGetCountV2").Device(DEVICE_GPU).HostMemory("table_handle")
                            .TypeConstraint<int64_t>("Tindices"), KvGetMetaOp<false>);
REGISTER_KERNEL_BUILDER(Name("KvVariableGetTimeStamp").Device(DEVICE_GPU).HostMemory("table_handle")
                            .TypeConstraint<int64_t>("Tindices"), KvGetMetaOp<true>);

// KvVariableDeleteWithTimestamp (ops/kv_variable_ops.cc:698-706): KvVariable::DeleteWithTimestamp,
// kernels/kv_variable.h:756-789.  The output size is data dependent: a buffer as large as the
// table is filled on the device, the count read back, the prefix copied out.
class KvDeleteWithTimestampOp : public OpKernel {
 public:
  explicit KvDeleteWithTimestampOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("threshold", &threshold_));
  }
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &v));
    core::ScopedUnref unref(v);
    int64_t cap = 0, n = 0;
    OP_REQUIRES_OK(ctx, FromKv(kv_map_size(v->table(), StreamOf(ctx), &cap)));
    Tensor tmp;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_INT64, TensorShape({cap > 0 ? cap : 1}), &tmp));
    OP_REQUIRES_OK(ctx, FromKv(kv_delete_with_timestamp(
        v->table(), threshold_, Today(), reinterpret_cast<int64_t*>(tmp.flat<int64_t>().data()), cap,
        StreamOf(ctx), &n)));
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({n}), &out));
    if (n > 0)
      cudaMemcpyAsync(out->flat<int64_t>().data(), tmp.flat<int64_t>().data(), n * sizeof(int64_t),
                      cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(StreamOf(ctx)));
  }
 private:
  int threshold_;
};
REGISTER_KERNEL_BUILDER(Name("KvVariableDeleteWithTimestamp").Device(DEVICE_GPU).HostMemory("table_handle")
                            .TypeConstraint<int64_t>("Tkeys"), KvDeleteWithTimestampOp);

// KvVariableFullOrDeltaExport (ops/kv_variable_ops.cc:633-660), delta mode: KvVariable::DeltaExport,
// kernels/dynamic_save.hpp:197-449.  Eight outputs; sizes from kv_delta_export_count, buffers
// bounded by them (entries a concurrent insert adds in between are dropped, not written).
class KvDeltaExportOp : public OpKernel {
 public:
  explicit KvDeltaExportOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("first_n", &first_n_));
  }
  void Compute(OpKernelContext* ctx) override {
    KvHbmVariable* v;
    OP_REQUIRES_OK(ctx, Lookup(ctx, 0, &v));
    core::ScopedUnref unref(v);
    int64_t nk, nb, nf, nd, counts[4];
    OP_REQUIRES_OK(ctx, FromKv(kv_delta_export_count(v->table(), first_n_, StreamOf(ctx), &nk, &nb, &nf, &nd)));
    Tensor *keys, *values, *init, *black, *fk, *fv, *need_full, *del;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({nk}), &keys));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({nk, kv_dim(v->table())}), &values));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({0, kv_dim(v->table())}), &init));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({nb}), &black));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(4, TensorShape({nf}), &fk));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(5, TensorShape({nf}), &fv));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(6, TensorShape({1}), &need_full));   // HostMemory
    OP_REQUIRES_OK(ctx, ctx->allocate_output(7, TensorShape({nd}), &del));
    need_full->flat<bool>()(0) = false;
    OP_REQUIRES_OK(ctx, FromKv(kv_delta_export(
        v->table(), first_n_, reinterpret_cast<int64_t*>(keys->flat<int64_t>().data()),
        values->flat<float>().data(), nk, reinterpret_cast<int64_t*>(black->flat<int64_t>().data()), nb,
        reinterpret_cast<int64_t*>(fk->flat<int64_t>().data()), fv->flat<uint32_t>().data(), nf,
        reinterpret_cast<int64_t*>(del->flat<int64_t>().data()), nd, StreamOf(ctx), counts)));
  }
 private:
  int first_n_;
};
REGISTER_KERNEL_BUILDER(Name("KvVariableFullOrDeltaExport").Device(DEVICE_GPU).HostMemory("table_handle")
                            .HostMemory("need_full_import"), KvDeltaExportOp);

}  // namespace tfplus_b200
