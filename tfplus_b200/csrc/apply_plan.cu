// apply_plan.cu — UnsortedSegmentSum + fused sparse apply in ONE kernel, driven by a dedup plan.
//
// What TF's optimizer does with the IndexedSlices gradient of a KvVariable — Unique,
// UnsortedSegmentSum, then the KvVariable*Apply* op on the unique ids
// (python/ops/variable_scope.py:1096-1106 -> Optimizer._deduplicate_indexed_slices) — with the
// Unique part taken from the plan (dedup.cu; it depends on the ids alone) and the other two
// fused: the gradient of a distinct id is summed straight into the registers that update its
// rows, never written to memory.  The sum runs over the id's occurrences in increasing
// position starting from +0, which is exactly the order of TF's CPU UnsortedSegmentSum
// (out[idx[j]] += data[j], j = 0, 1, ...), so the result does not depend on scheduling and
// equals a sequential CPU sum bit for bit.
//
// That order is a serial chain per distinct id.  Zipf batches have a few very hot ids (the head
// id of the microbench occurs ~7 600 times in 65 536): at one dependent FADD (4 cycles) per
// occurrence its chain alone is ~16 us, so the kernel is organised around it:
//  * "heavy" ids (more than plan.heavy_t occurrences, listed by the plan) are work items
//    (id, 32-column part).  Warp 0 of a block walks the chain of its item out of a
//    shared-memory ring that warp 1 keeps full with bulk asynchronous copies (TMA,
//    cp.async.bulk + mbarrier: one copy per occurrence row part, hundreds in flight, no
//    registers), one float per lane, then parks the partial sum; whoever completes an id
//    applies it.  Heavy items take the lowest block indices, i.e. they start first.
//  * every other warp of every block processes light ids in groups (apply_group,
//    apply_math.cuh): a tile sums the <= heavy_t rows of its id, unrolled loads first.
#include <cstdlib>

#include "apply_math.cuh"
#include "async_copy.cuh"
#include "plan.h"

namespace kvhbm {

PlanView plan_view(const Plan* p);
int plan_need_scratch(Plan* p, int dim, cudaStream_t st);
int apply_validate(int kind, Table* var, Table* sa, Table* sb, const float* hp);

namespace {

constexpr int AP_THREADS = 320;  // warp 0: heavy consumer, warp 1: heavy producer, 2..9: light ids
constexpr int AP_NW = AP_THREADS / 32;
constexpr int RING_ROWS = 64;    // occurrence rows per ring stage
constexpr int RING_STAGES = 10;
constexpr int RING_PITCH = 128;  // bytes per row part (32 columns)
constexpr int RING_BYTES = RING_STAGES * RING_ROWS * RING_PITCH;  // 80 KB

struct Ring {
  unsigned char* base;
  unsigned long long* full;
  unsigned long long* empty;
  int stage;
  unsigned par;
  __device__ __forceinline__ void advance() {
    if (++stage == RING_STAGES) { stage = 0; par ^= 1u; }
  }
};

// Consumer side of one heavy work item: the sum of column `col` over the item's `c`
// occurrence rows in list order, out of the ring (ringed) or straight from memory.
__device__ __forceinline__ float heavy_consume(Ring& rg, bool ringed, const float* __restrict__ grad,
                                               const int* __restrict__ list, int c, int dim,
                                               int col, bool act, int lane) {
  float acc = 0.f;
  if (ringed) {
    for (int k0 = 0; k0 < c; k0 += RING_ROWS) {
      mbar_wait(&rg.full[rg.stage], rg.par);
      const float* st = reinterpret_cast<const float*>(rg.base + (size_t)rg.stage * RING_ROWS * RING_PITCH);
      const int rows = c - k0 < RING_ROWS ? c - k0 : RING_ROWS;
      if (rows == RING_ROWS) {
#pragma unroll 16
        for (int j = 0; j < RING_ROWS; ++j) acc += st[j * 32 + lane];
      } else {
        for (int j = 0; j < rows; ++j) acc += st[j * 32 + lane];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rg.empty[rg.stage]);
      rg.advance();
    }
  } else {
    for (int k0 = 0; k0 < c; k0 += 8) {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        t[j] = (act && k0 + j < c) ? __ldcs(grad + (long long)__ldg(list + k0 + j) * dim + col) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (k0 + j < c) acc += t[j];
    }
  }
  return acc;
}

// Producer side: one bulk copy per occurrence row part into the ring, a stage at a time.
__device__ __forceinline__ void heavy_produce(Ring& rg, const float* __restrict__ g0,
                                              const int* __restrict__ list, int c, int dim,
                                              unsigned pb, int lane) {
  for (int k0 = 0; k0 < c; k0 += RING_ROWS) {
    mbar_wait(&rg.empty[rg.stage], rg.par ^ 1u);
    unsigned char* st = rg.base + (size_t)rg.stage * RING_ROWS * RING_PITCH;
    int pz[RING_ROWS / 32];
    unsigned bytes = 0;
#pragma unroll
    for (int q = 0; q < RING_ROWS / 32; ++q) {
      const int k = k0 + q * 32 + lane;
      pz[q] = k < c ? __ldg(list + k) : -1;
      if (pz[q] >= 0) bytes += pb;
    }
    if (bytes) mbar_arrive_expect_tx(&rg.full[rg.stage], bytes);
    else mbar_arrive(&rg.full[rg.stage]);
#pragma unroll
    for (int q = 0; q < RING_ROWS / 32; ++q)
      if (pz[q] >= 0)
        bulk_load(st + (size_t)(q * 32 + lane) * RING_PITCH, g0 + (long long)pz[q] * dim, pb,
                  &rg.full[rg.stage]);
    rg.advance();
  }
}

// The apply of one heavy id by the warp that completed its sum.  Not inlined: it runs once
// per heavy id, and keeping a second copy of the group routine out of the kernel body leaves
// the registers to the light path.
template <int VEC, int CPL, int KIND>
__device__ __noinline__ void apply_heavy_id(ApplySmem<AP_NW, VEC, CPL>* sm, int wib,
                                            const TableView* var, const TableView* sa,
                                            const TableView* sb, const PlanView* pl, int h, int r,
                                            const ApplyParams* p, uint32_t today, int tpr) {
  GradSrc gs;
  gs.grad = pl->heavy_sum + (size_t)h * pl->sum_dim; gs.row0 = r; gs.counts = nullptr;
  gs.seg_off = nullptr; gs.pos = nullptr; gs.heavy_t = 0; gs.hint = pl->hint; gs.cg = true;
  apply_group<AP_NW, VEC, CPL, KIND, 1, 1>(*sm, wib, *var, *sa, *sb, pl->uniq, gs, r, (long long)r + 1,
                                           *p, today, tpr, 32 / tpr, true);
}

template <int VEC, int CPL, int KIND>
__global__ void __launch_bounds__(AP_THREADS, 2)
apply_plan_kernel(const __grid_constant__ TableView var, const __grid_constant__ TableView sa,
                  const __grid_constant__ TableView sb, const __grid_constant__ PlanView pl,
                  const float* __restrict__ grad, const __grid_constant__ ApplyParams p_in,
                  const float* __restrict__ d_hp, uint32_t today, int tpr, int kpw, float* d_adv,
                  int use_ring) {
  ApplyParams p = p_in;
  if (d_hp) p = derive_params<KIND>(d_hp, var.dim, p_in.update_slots);
  __shared__ ApplySmem<AP_NW, VEC, CPL> sm;
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ unsigned long long full_bar[RING_STAGES], empty_bar[RING_STAGES];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dim = var.dim;
  const long long U = *pl.num;
  const long long* ids = pl.uniq;
  const bool ringed = use_ring && VEC == 4;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RING_STAGES; ++s) { mbar_init(&full_bar[s], 32); mbar_init(&empty_bar[s], 1); }
    fence_async_smem();
  }
  __syncthreads();

  int H = *pl.heavy_n;
  if (H > pl.heavy_cap) H = pl.heavy_cap;
  const int parts = (dim + 31) / 32;
  const long long items = (long long)H * parts;

  Ring rg;
  rg.base = ring; rg.full = full_bar; rg.empty = empty_bar; rg.stage = 0; rg.par = 0;
  if (wib == 0) {
    // ---- heavy consumer: the serial chain of one (id, 32-column part) at a time ----
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
      const int h = (int)(it / parts), part = (int)(it - (long long)h * parts);
      const int r = pl.heavy[h];
      const int c = pl.counts[r], off = pl.seg_off[r];
      const int col = part * 32 + lane;
      const bool act = col < dim;
      const float acc = heavy_consume(rg, ringed, grad, pl.pos + off, c, dim, col, act, lane);
      if (act) __stcg(pl.heavy_sum + (size_t)h * pl.sum_dim + col, acc);
      __threadfence();
      unsigned last = 0;
      if (lane == 0) last = atomicAdd(&pl.heavy_done[h], 1u) == (unsigned)(parts - 1);
      last = __shfl_sync(APPLY_FULL, last, 0);
      if (last) {
        // every part of this id has been parked: apply it (the counter goes back to zero for
        // the next launch)
        if (lane == 0) pl.heavy_done[h] = 0u;
        __threadfence();
        apply_heavy_id<VEC, CPL, KIND>(&sm, wib, &var, &sa, &sb, &pl, h, r, &p, today, tpr);
      }
    }
  } else if (wib == 1) {
    // ---- heavy producer: keeps the ring full ----
    if (ringed) {
      for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int h = (int)(it / parts), part = (int)(it - (long long)h * parts);
        const int r = pl.heavy[h];
        const int width = dim - part * 32 < 32 ? dim - part * 32 : 32;
        heavy_produce(rg, grad + part * 32, pl.pos + pl.seg_off[r], pl.counts[r], dim,
                      (unsigned)width * 4u, lane);
      }
    }
  } else {
    // ---- light ids ----
    const long long lw = (long long)blockIdx.x * (AP_NW - 2) + (wib - 2);
    const long long nlw = (long long)gridDim.x * (AP_NW - 2);
    GradSrc gs;
    gs.grad = grad; gs.row0 = 0; gs.counts = pl.counts; gs.seg_off = pl.seg_off; gs.pos = pl.pos;
    gs.heavy_t = pl.heavy_t; gs.hint = pl.hint; gs.cg = false;
    for (long long base = lw * kpw; base < U; base += nlw * kpw)
      apply_group<AP_NW, VEC, CPL, KIND, 1, 4>(sm, wib, var, sa, sb, ids, gs, base, U, p, today, tpr,
                                               kpw, false);
  }

  // AdamOptimizer._finish folded into this launch (see apply.cu)
  if ((KindTraits<KIND>::ADAMISH || KIND == K_ADAM) && d_adv != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned done = atomicAdd(&var.ctr->apply_done, 1u);
      if (done == gridDim.x - 1) {
        constexpr int P = KIND == K_ADAM ? 4 : 1;
        constexpr int Bt = KIND == K_ADAM ? 1 : 3;
        d_adv[P] = d_adv[P] * d_adv[Bt];
        d_adv[P + 1] = d_adv[P + 1] * d_adv[Bt + 1];
        var.ctr->apply_done = 0;
      }
    }
  }
}

// tf.math.unsorted_segment_sum through the plan, on its own: out[r, :] = sum of the rows
// data[pos[seg_off[r] + k], :], k = 0 .. counts[r]-1, in that order from +0.  Same roles as the
// fused kernel: warps 0/1 run the heavy ids' chains through the ring, the others take light
// segments, a tile per segment.
__global__ void __launch_bounds__(AP_THREADS, 2)
segsum_plan_kernel(const __grid_constant__ PlanView pl, const float* __restrict__ data, int dim,
                   float* __restrict__ out, int tpr, int vec, int use_ring) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ unsigned long long full_bar[RING_STAGES], empty_bar[RING_STAGES];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long U = *pl.num;
  const bool ringed = use_ring && vec == 4;
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING_STAGES; ++s) { mbar_init(&full_bar[s], 32); mbar_init(&empty_bar[s], 1); }
    fence_async_smem();
  }
  __syncthreads();
  int H = *pl.heavy_n;
  if (H > pl.heavy_cap) H = pl.heavy_cap;
  const int parts = (dim + 31) / 32;
  const long long items = (long long)H * parts;
  Ring rg;
  rg.base = ring; rg.full = full_bar; rg.empty = empty_bar; rg.stage = 0; rg.par = 0;
  if (wib == 0) {
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
      const int h = (int)(it / parts), part = (int)(it - (long long)h * parts);
      const int r = pl.heavy[h];
      const int col = part * 32 + lane;
      const bool act = col < dim;
      const float acc = heavy_consume(rg, ringed, data, pl.pos + pl.seg_off[r], pl.counts[r], dim,
                                      col, act, lane);
      if (act) out[(long long)r * dim + col] = acc;
    }
  } else if (wib == 1) {
    if (ringed) {
      for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int h = (int)(it / parts), part = (int)(it - (long long)h * parts);
        const int r = pl.heavy[h];
        const int width = dim - part * 32 < 32 ? dim - part * 32 : 32;
        heavy_produce(rg, data + part * 32, pl.pos + pl.seg_off[r], pl.counts[r], dim,
                      (unsigned)width * 4u, lane);
      }
    }
  } else {
    // a tile of `tpr` lanes per segment, elements strided by the tile (any dim)
    const int tl = lane & (tpr - 1);
    const long long tile = ((long long)blockIdx.x * (AP_NW - 2) + (wib - 2)) * (32 / tpr) + lane / tpr;
    const long long ntiles = (long long)gridDim.x * (AP_NW - 2) * (32 / tpr);
    const int per = (dim / vec + tpr - 1) / tpr;  // chunks per lane (<= 8)
    for (long long r = tile; r < U; r += ntiles) {
      const int c = pl.counts[r];
      if (c > pl.heavy_t) continue;
      const int* list = pl.pos + pl.seg_off[r];
      float acc[8][4];
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[q][e] = 0.f;
      for (int k = 0; k < c; ++k) {
        const float* row = data + (long long)__ldg(list + k) * dim;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int off = (q * tpr + tl) * vec;
          if (q < per && off < dim) {
            if (vec == 4) {
              const float4 v = __ldcs(reinterpret_cast<const float4*>(row + off));
              acc[q][0] += v.x; acc[q][1] += v.y; acc[q][2] += v.z; acc[q][3] += v.w;
            } else {
              acc[q][0] += __ldcs(row + off);
            }
          }
        }
      }
      float* o = out + r * dim;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int off = (q * tpr + tl) * vec;
        if (q < per && off < dim) {
          if (vec == 4) *reinterpret_cast<float4*>(o + off) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
          else o[off] = acc[q][0];
        }
      }
    }
  }
}

__global__ void advance_powers_plan_kernel(float* hp, int p, int b) {
  hp[p] = hp[p] * hp[b];
  hp[p + 1] = hp[p + 1] * hp[b + 1];
}

template <int VEC, int CPL, int KIND>
int launch_apply_plan(Table* var, Table* sa, Table* sb, const PlanView& pv, const float* grad,
                      const ApplyParams& p, const float* d_hp, uint16_t today, cudaStream_t st,
                      int tpr, float* d_adv) {
  static const int kpw_env = getenv("KVHBM_APPLYP_KPW") ? atoi(getenv("KVHBM_APPLYP_KPW")) : 0;
  static const int ring_env = getenv("KVHBM_APPLYP_RING") ? atoi(getenv("KVHBM_APPLYP_RING")) : 1;
  static const int bps_env = getenv("KVHBM_APPLYP_BPS") ? atoi(getenv("KVHBM_APPLYP_BPS")) : 2;
  const int sms = sm_count(var->device);
  const int kpi = 32 / tpr;
  // light warps of a full grid; a Zipf batch has ~n/3 distinct ids
  const long long lwarps = (long long)sms * bps_env * (AP_NW - 2);
  const long long n_est = (pv.n + 2) / 3;
  int kpw = kpi;
  while (kpw < 32 && (n_est + kpw - 1) / kpw > lwarps) kpw <<= 1;
  if (kpw_env >= kpi && kpw_env <= 32) kpw = kpw_env;
  long long blocks = ((pv.n + kpw - 1) / kpw + (AP_NW - 3)) / (AP_NW - 2);
  if (blocks > (long long)sms * bps_env) blocks = (long long)sms * bps_env;
  if (blocks < 1) blocks = 1;
  auto kern = apply_plan_kernel<VEC, CPL, KIND>;
  static bool attr = false;  // per instantiation
  if (!attr) {
    KV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RING_BYTES));
    attr = true;
  }
  TableView vb = sb ? sb->view() : sa->view();
  kern<<<(unsigned)blocks, AP_THREADS, RING_BYTES, st>>>(var->view(), sa->view(), vb, pv, grad, p, d_hp,
                                                        today, tpr, kpw, d_adv, ring_env);
  KV_LAUNCHED();
  return 0;
}

template <int KIND>
int dispatch_apply_plan(Table* var, Table* sa, Table* sb, Plan* plan, const float* grad,
                        const ApplyParams& p, const float* d_hp, uint16_t today, cudaStream_t st,
                        float* d_adv) {
  KV_TRY(plan_need_scratch(plan, var->dim, st));
  const PlanView pv = plan_view(plan);
  if (pv.n <= 0) {
    if (d_adv) {  // nothing to update, but the step still counts
      advance_powers_plan_kernel<<<1, 1, 0, st>>>(d_adv, KIND == K_ADAM ? 4 : 1, KIND == K_ADAM ? 1 : 3);
      KV_LAUNCHED();
    }
    return 0;
  }
  KV_TRY(var->ensure(pv.n, st));
  KV_TRY(sa->ensure(pv.n, st));
  if (sb) KV_TRY(sb->ensure(pv.n, st));
  RowGeom g = row_geom(var->dim);
  if (g.cpl > 4)
    return fail(3, "fused apply: embedding dim " + std::to_string(var->dim) +
                       " not supported (max 512 when a multiple of 4, else 128)");
  const int cpl = g.cpl == 3 ? 4 : g.cpl;
#define CALL(V, C) launch_apply_plan<V, C, KIND>(var, sa, sb, pv, grad, p, d_hp, today, st, g.tpr, d_adv)
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

}  // namespace

int do_segment_sum_plan(Plan* plan, const float* data, int dim, float* out, cudaStream_t st) {
  const PlanView pv = plan_view(plan);
  if (pv.n <= 0) return 0;
  RowGeom g = row_geom(dim);
  if (g.cpl > 8) return fail(3, "segment_sum_plan: dim too large");
  static const int ring_env = getenv("KVHBM_APPLYP_RING") ? atoi(getenv("KVHBM_APPLYP_RING")) : 1;
  static bool attr = false;
  if (!attr) {
    KV_CUDA(cudaFuncSetAttribute(segsum_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 RING_BYTES));
    attr = true;
  }
  int dev = 0;
  KV_CUDA(cudaGetDevice(&dev));
  const int sms = sm_count(dev);
  const long long tiles_per_block = (long long)(AP_NW - 2) * (32 / g.tpr);
  long long blocks = (pv.n + tiles_per_block - 1) / tiles_per_block;
  if (blocks > 2LL * sms) blocks = 2LL * sms;
  if (blocks < 1) blocks = 1;
  segsum_plan_kernel<<<(unsigned)blocks, AP_THREADS, RING_BYTES, st>>>(pv, data, dim, out, g.tpr, g.vec,
                                                                      ring_env);
  KV_LAUNCHED();
  return 0;
}

// UnsortedSegmentSum(grad[n, dim], plan.idx) + the apply op `kind` on plan.uniq, in one launch.
int do_apply_plan(int kind, Table* var, Table* sa, Table* sb, Plan* plan, const float* grad,
                  const float* hp, const float* d_hp, int update_slots, uint16_t today,
                  cudaStream_t st, float* d_adv) {
  KV_TRY(apply_validate(kind, var, sa, sb, d_hp ? nullptr : hp));
#define KV_KIND(K)                                                                            \
  case K: {                                                                                   \
    ApplyParams p{};                                                                          \
    if (d_hp == nullptr) p = derive_params<K>(hp, var->dim, update_slots);                    \
    p.update_slots = update_slots;                                                            \
    return dispatch_apply_plan<K>(var, sa, Kind<K>::TWO ? sb : nullptr, plan, grad, p, d_hp,  \
                                  today, st, d_adv);                                          \
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

}  // namespace kvhbm
