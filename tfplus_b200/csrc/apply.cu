// apply.cu — fused per-row sparse optimizer updates on the device table.
//
// One kernel per optimizer step: for every (unique) id it finds-or-inserts the
// value row and the slot row(s), reads gradient + value + slots once, updates
// them in registers and writes them back once — the reference walks each row
// three times through lazy Eigen expressions and two or three hash maps
// (training_ops.cc:7166-7195, :713-751, :1473-1482).
//
// Arithmetic is the reference's, operation by operation, in fp32 without FMA
// contraction (this file is compiled with -fmad=false; sqrtf and `/` are the
// IEEE-rounded versions), the L2-norm reduction included (apply_math.cuh).
#include <cstdlib>

#include "apply_math.cuh"

namespace kvhbm {

namespace {

// The fused apply over deduplicated ids with their gradients already summed (the op surface of
// the reference: one gradient row per unique id).  A warp takes `kpw` ids at a time
// (apply_group, apply_math.cuh); kpw is chosen on the host so that one wave of warps covers
// the whole launch.
template <int VEC, int CPL, int KIND, int UNR>
__global__ void __launch_bounds__(128, CPL == 1 ? 5 : 1)
apply_kernel(TableView var, TableView sa, TableView sb, const long long* __restrict__ ids,
             const float* __restrict__ grad, long long n, const int* __restrict__ d_n,
             ApplyParams p, const float* __restrict__ d_hp, uint32_t today, int tpr, int kpw,
             float* d_adv) {
  if (d_hp) p = derive_params<KIND>(d_hp, var.dim, p.update_slots);
  __shared__ ApplySmem<4, VEC, CPL> sm;
  const int wib = threadIdx.x >> 5;
  const long long wpb = blockDim.x >> 5;
  if (d_n) { long long dn = *d_n; if (dn < n) n = dn; }
  GradSrc gs;
  gs.grad = grad; gs.row0 = 0; gs.counts = nullptr; gs.seg_off = nullptr; gs.pos = nullptr;
  gs.heavy_t = 0; gs.hint = nullptr; gs.cg = false;
  for (long long base = (blockIdx.x * wpb + wib) * kpw; base < n;
       base += (long long)gridDim.x * wpb * kpw)
    apply_group<4, VEC, CPL, KIND, UNR, 1>(sm, wib, var, sa, sb, ids, gs, base, n, p, today, tpr,
                                           kpw, false);
  // AdamOptimizer._finish (beta1_power *= beta1, beta2_power *= beta2; inherited by
  // python/training/group_adam.py) folded into this launch: every block read the powers when
  // it started, so the block that finishes last may advance them for the next step.
  if ((KindTraits<KIND>::ADAMISH || KIND == K_ADAM) && d_adv != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned done = atomicAdd(&var.ctr->apply_done, 1u);
      if (done == gridDim.x - 1) {
        constexpr int P = KIND == K_ADAM ? 4 : 1;   // index of beta1_power in hp
        constexpr int Bt = KIND == K_ADAM ? 1 : 3;  // index of beta1 in hp
        d_adv[P] = d_adv[P] * d_adv[Bt];
        d_adv[P + 1] = d_adv[P + 1] * d_adv[Bt + 1];
        var.ctr->apply_done = 0;
      }
    }
  }
}

__global__ void advance_powers_kernel(float* hp, int p, int b) {
  hp[p] = hp[p] * hp[b];
  hp[p + 1] = hp[p + 1] * hp[b + 1];
}

template <int VEC, int CPL, int KIND>
int launch_apply(Table* var, Table* sa, Table* sb, const int64_t* ids, const float* grad,
                 int64_t n, const int32_t* d_n, const ApplyParams& p, const float* d_hp,
                 uint16_t today, cudaStream_t st, int tpr, float* d_adv) {
  static const int kpw_env = getenv("KVHBM_APPLY_KPW") ? atoi(getenv("KVHBM_APPLY_KPW")) : 0;
  // ids per warp: the smallest power of two that keeps the launch within ~1.5 waves of warps
  // (20 resident per SM at this register footprint); measured best on B200 (profiles/).  With
  // d_n the launch is sized for the upper bound n; the usual caller (unique -> apply) has
  // ~n/3 valid ids, hence the /2.
  const int kpi = 32 / tpr;
  int kpw = kpi;
  const long long max_warps = (long long)sm_count(var->device) * 32;
  const long long n_est = d_n ? (n + 1) / 2 : n;
  while (kpw < 32 && (n_est + kpw - 1) / kpw > max_warps) kpw <<= 1;
  if (kpw_env >= kpi && kpw_env <= 32) kpw = kpw_env;
  const int64_t warps = (n + kpw - 1) / kpw;
  const int blocks = blocks_for(warps, 4, var->device, 32);
  TableView vb = sb ? sb->view() : sa->view();
  apply_kernel<VEC, CPL, KIND, 1><<<blocks, 128, 0, st>>>(
      var->view(), sa->view(), vb, reinterpret_cast<const long long*>(ids), grad, n, d_n, p, d_hp,
      today, tpr, kpw, d_adv);
  KV_LAUNCHED();
  return 0;
}

template <int KIND>
int dispatch_apply(Table* var, Table* sa, Table* sb, const int64_t* ids, const float* grad,
                   int64_t n, const int32_t* d_n, const ApplyParams& p, const float* d_hp,
                   uint16_t today, cudaStream_t st, float* d_adv) {
  if (n <= 0) {
    if (d_adv) {  // nothing to update, but the step still counts
      advance_powers_kernel<<<1, 1, 0, st>>>(d_adv, KIND == K_ADAM ? 4 : 1,
                                             KIND == K_ADAM ? 1 : 3);
      KV_LAUNCHED();
    }
    return 0;
  }
  KV_TRY(var->ensure(n, st));
  KV_TRY(sa->ensure(n, st));
  if (sb) KV_TRY(sb->ensure(n, st));
  RowGeom g = row_geom(var->dim);
  if (g.cpl > 4)
    return fail(3, "fused apply: embedding dim " + std::to_string(var->dim) +
                       " not supported (max 512 when a multiple of 4, else 128)");
  const int cpl = g.cpl == 3 ? 4 : g.cpl;
#define CALL(V, C) launch_apply<V, C, KIND>(var, sa, sb, ids, grad, n, d_n, p, d_hp, today, st, g.tpr, d_adv)
  if (g.vec == 4) {
    if (cpl == 1) return CALL(4, 1);
    if (cpl == 2) return CALL(4, 2);
    return CALL(4, 4);
  }
  if (cpl == 1) return CALL(1, 1);
  if (cpl == 2) return CALL(1, 2);
  return CALL(1, 4);
#undef CALL
}

int check_initialized(const Table* t, const char* what) {
  if (!t->initialized)
    return fail(2, std::string("Attempting to use uninitialized variables: ") + what);
  return 0;
}

}  // namespace

// Argument checks of the reference's Compute() bodies (training_ops.cc:7001-7103 GroupAdam,
// :560-659 / :309-405 / :833-930 the FTRL family, :1399-1440 Adagrad).  `hp` holds the op's
// scalar inputs on the host or is null when they live in device memory (then they are not
// validated: reading them back would need a stream synchronisation).
int apply_validate(int kind, Table* var, Table* sa, Table* sb, const float* hp) {
  const bool adamish = kind == K_GROUP_ADAM || kind == K_GROUP_ADAM_V3;
  const bool ftrlish = kind == K_FTRL || kind == K_FTRL_V2 || kind == K_GROUP_FTRL_V2;
  KV_TRY(check_initialized(var, "var"));
  if (adamish) {
    KV_TRY(check_initialized(sa, "m_v_linear"));
    if (hp) {
      if (!(hp[0] > 0.f)) return fail(1, "lr is not a positive scalar");
      if (!(hp[6] >= 0.f)) return fail(1, "l1 regularization strength is not a non-negative scalar");
      if (!(hp[7] >= 0.f)) return fail(1, "l2 regularization strength is not a non-negative scalar");
      if (!(hp[8] >= 0.f)) return fail(1, "l21 regularization strength is not a non-negative scalar");
    }
    if (sa->dim != 3 * var->dim) return fail(1, "kv_variable and linear do not have the same shape");
  } else if (ftrlish) {
    KV_TRY(check_initialized(sa, "accum"));
    KV_TRY(check_initialized(sb, "linear"));
    if (hp) {
      if (!(hp[0] > 0.f)) return fail(1, "lr is not a positive scalar");
      if (!(hp[1] >= 0.f)) return fail(1, "l1 regularization strength is not a non-negative scalar");
      if (!(hp[2] >= 0.f)) return fail(1, "l2 regularization strength is not a non-negative scalar");
      if (!(hp[3] >= 0.f)) return fail(1, "l21 regularization strength is not a non-negative scalar");
      if (!(hp[4] >= 0.f))
        return fail(1, "l2 shrinkage regularization strength is not a non-negative scalar");
      if (!(hp[5] <= 0.f)) return fail(1, "lr_power is not a non-positive scalar");
    }
    if (sa->dim != var->dim) return fail(1, "kv_varaible and accum do not have the same shape");
    if (sb->dim != var->dim) return fail(1, "kv_variable and linear do not have the same shape");
  } else if (kind == K_ADAGRAD) {
    KV_TRY(check_initialized(sa, "accum"));
    if (sa->dim != var->dim) return fail(1, "var and accum do not have the same shape");
  } else if (kind == K_ADAM) {
    KV_TRY(check_initialized(sa, "m_v"));
    if (sa->dim != 2 * var->dim) return fail(1, "m_v slot must have dim 2 * dim(var)");
  } else {
    return fail(1, "apply: unknown optimizer kind");
  }
  return 0;
}

// One entry for every optimizer kind (apply_math.cuh K_*).  `hp` = the op's scalar inputs on
// the host, or `d_hp` = the same in device memory; `d_adv` (Adam family) = where to advance
// beta^t once every row is updated.
int do_apply(int kind, Table* var, Table* sa, Table* sb, const int64_t* ids, const float* grad,
             int64_t n, const int32_t* d_n, const float* hp, const float* d_hp, int update_slots,
             uint16_t today, cudaStream_t st, float* d_adv) {
  KV_TRY(apply_validate(kind, var, sa, sb, d_hp ? nullptr : hp));
#define KV_KIND(K)                                                                          \
  case K: {                                                                                 \
    ApplyParams p{};                                                                        \
    if (d_hp == nullptr) p = derive_params<K>(hp, var->dim, update_slots);                  \
    p.update_slots = update_slots;                                                          \
    return dispatch_apply<K>(var, sa, Kind<K>::TWO ? sb : nullptr, ids, grad, n, d_n, p,    \
                             d_hp, today, st, d_adv);                                       \
  }
  switch (kind) {
    KV_KIND(K_ADAGRAD)
    KV_KIND(K_GROUP_ADAM)
    KV_KIND(K_FTRL)
    KV_KIND(K_ADAM)
    KV_KIND(K_GROUP_ADAM_V3)
    KV_KIND(K_FTRL_V2)
    KV_KIND(K_GROUP_FTRL_V2)
  }
#undef KV_KIND
  return fail(1, "apply: unknown optimizer kind");
}

int do_apply_adagrad(Table* var, Table* accum, const int64_t* ids, const float* grad, int64_t n,
                     const int32_t* d_n, const float* hp, const float* d_hp, int update_slots,
                     uint16_t today, cudaStream_t st) {
  return do_apply(K_ADAGRAD, var, accum, nullptr, ids, grad, n, d_n, hp, d_hp, update_slots,
                  today, st, nullptr);
}
int do_apply_group_adam_v4(Table* var, Table* mvl, const int64_t* ids, const float* grad,
                           int64_t n, const int32_t* d_n, const float* hp, const float* d_hp,
                           uint16_t today, cudaStream_t st, float* d_adv) {
  return do_apply(K_GROUP_ADAM, var, mvl, nullptr, ids, grad, n, d_n, hp, d_hp, 1, today, st,
                  d_adv);
}
int do_apply_sparse_group_ftrl(Table* var, Table* accum, Table* linear, const int64_t* ids,
                               const float* grad, int64_t n, const int32_t* d_n, const float* hp,
                               const float* d_hp, uint16_t today, cudaStream_t st) {
  return do_apply(K_FTRL, var, accum, linear, ids, grad, n, d_n, hp, d_hp, 1, today, st, nullptr);
}
int do_apply_adam(Table* var, Table* mv, const int64_t* ids, const float* grad, int64_t n,
                  const int32_t* d_n, const float* hp, const float* d_hp, uint16_t today,
                  cudaStream_t st, float* d_adv) {
  return do_apply(K_ADAM, var, mv, nullptr, ids, grad, n, d_n, hp, d_hp, 1, today, st, d_adv);
}

}  // namespace kvhbm
