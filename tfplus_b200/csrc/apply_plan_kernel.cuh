// apply_plan_kernel.cuh — the fused UnsortedSegmentSum + sparse apply kernel and its pieces, shared
// by apply_plan.cu (staging pass, plain segment sum, dispatch) and the per-optimizer translation
// units apply_plan_k*.cu (one optimizer each, so that nvcc compiles them in parallel).
#ifndef KVHBM_APPLY_PLAN_KERNEL_CUH_
#define KVHBM_APPLY_PLAN_KERNEL_CUH_

#include <cstdlib>

#include "apply_math.cuh"
#include "async_copy.cuh"
#include "plan.h"

namespace kvhbm {

PlanView plan_view(const Plan* p);
int plan_need_scratch(Plan* p, int dim, cudaStream_t st);
int launch_stage_heavy(const PlanView& pv, const float* data, int dim, int device, cudaStream_t st);

namespace {

// Optional per-warp timeline of the fused kernel (scripts/trace_apply_plan.py; compiled in by
// KVHBM_TRACE=1): {start, heavy items done, light groups done, groups taken} in ns.  One copy per
// translation unit; set_trace_apply (apply_plan.cu) sets them all.
__device__ unsigned long long* g_trace_plan = nullptr;
inline int set_trace_plan_local(unsigned long long* d_buf) {
  KV_CUDA(cudaMemcpyToSymbol(g_trace_plan, &d_buf, sizeof(d_buf)));
  return 0;
}


#ifdef KVHBM_TRACE
__device__ __forceinline__ unsigned long long gtime_plan() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#endif

// One block per SM, 20 warps.  A dependent FADD issues every 4 cycles only if its warp has the
// SM sub-partition (warps w with the same w % 4) to itself: with four light warps sharing the
// scheduler the chain ran at 10 cycles per add (measured, scripts/trace_apply_plan.py; alone:
// 4.15, scripts/ub/faddchain.cu).  So while warp 0 walks the block's chains, the other warps of
// its sub-partition (4, 8, 12, 16) sleep at a named barrier; the remaining 15 warps take light
// ids from the start.
constexpr int AP_THREADS = 640;
constexpr int AP_NW = AP_THREADS / 32;
constexpr int AP_MATES = AP_NW / 4;   // warps of sub-partition 0, the consumer included
constexpr int RING_ROWS = 128;    // occurrence rows per ring stage (one bulk copy, 16 KB)
constexpr int RING_STAGES = 5;
constexpr int RING_PITCH = 128;   // bytes per row part (32 columns)
constexpr int RING_BYTES = RING_STAGES * RING_ROWS * RING_PITCH;  // 80 KB

// The block's ring of stages.  `no` counts the stages the block has been through since the
// launch started (the same number in the consumer and the producer): stage `no` lives in ring
// entry no % STAGES and is use number no / STAGES of that entry.
struct Ring {
  unsigned char* base;
  unsigned long long* full;
  unsigned long long* empty;
  unsigned no;
};

// One heavy work item = the sum of a 32-column part of one id's occurrence rows — entries
// [e0, e0 + c) of pos — in list order from +0.  Warp 0 walks the chain out of the ring, one
// column per lane, four units (16 rows) pulled into registers ahead of the adds so that the
// chain itself is nothing but dependent FADDs; lane 0 of warp 1 keeps the ring full with one
// bulk copy per stage (32 units, 16 KB) out of the staged part `part0` (unit 0 of the part).
// The first and the last stage may hold rows of neighbouring ids: they are added under a
// predicate, the stages in between unconditionally.  Returns the sum (warp 0).
__device__ __forceinline__ float chain_item(Ring& rg, int wib, int lane, const float* __restrict__ part0,
                                            int e0, int c, int dbg = 0) {
  constexpr int SU = RING_ROWS / 4;     // units per stage
  const int e1 = e0 + c;
  const int u0 = e0 >> 2, u1 = (e1 + 3) >> 2;
  const int nst = (u1 - u0 + SU - 1) / SU;
  const unsigned no0 = rg.no;
  rg.no += (unsigned)nst;
  float acc = 0.f;
  if (wib == 0) {
    // running ring entry / parity of the stage being read (no divisions in the loop)
    unsigned e = no0 % RING_STAGES, par = (no0 / RING_STAGES) & 1u;
    auto stage_ptr = [&](unsigned ent) {
      return reinterpret_cast<const float4*>(rg.base + (size_t)ent * RING_ROWS * RING_PITCH) + lane;
    };
    auto next_entry = [&]() { if (++e == RING_STAGES) { e = 0; par ^= 1u; } };
    // a boundary stage: rows outside [e0, e1) belong to neighbouring ids.  Four units (16 rows)
    // at a time, loads first; only a batch that straddles an end of the run pays for predicates
    auto boundary = [&](int s) {
      mbar_wait(&rg.full[e], par);
      const float4* st = stage_ptr(e);
      const int ub = u0 + s * SU;
      const int nun = u1 - ub < SU ? u1 - ub : SU;
      for (int j0 = 0; j0 < nun; j0 += 4) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = j0 + j < nun ? st[(j0 + j) * 32] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int row0 = (ub + j0) * 4;
        if (row0 >= e0 && row0 + 16 <= e1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { acc += v[j].x; acc += v[j].y; acc += v[j].z; acc += v[j].w; }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int row = row0 + j * 4;
            if (row >= e0 && row < e1) acc += v[j].x;
            if (row + 1 >= e0 && row + 1 < e1) acc += v[j].y;
            if (row + 2 >= e0 && row + 2 < e1) acc += v[j].z;
            if (row + 3 >= e0 && row + 3 < e1) acc += v[j].w;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rg.empty[e]);
      next_entry();
    };
    boundary(0);
    if (nst > 2) {
      // interior stages: a tight loop of nothing but 128-bit loads issued a chunk (16 rows)
      // ahead, dependent adds, and one barrier round per stage
      static_assert(RING_ROWS == 128, "eight chunks per stage");
      float4 a[4], b[4];
#define KV_LOAD(dst, c0) _Pragma("unroll") for (int j = 0; j < 4; ++j) dst[j] = st[((c0) * 4 + j) * 32]
#define KV_ADD(v) _Pragma("unroll") for (int j = 0; j < 4; ++j) { acc += v[j].x; acc += v[j].y; acc += v[j].z; acc += v[j].w; }
      mbar_wait(&rg.full[e], par);
      const float4* st = stage_ptr(e);
      KV_LOAD(a, 0);
#pragma unroll 1
      for (int s = 1; s < nst - 1; ++s) {
        KV_LOAD(b, 1); KV_ADD(a);
        KV_LOAD(a, 2); KV_ADD(b);
        KV_LOAD(b, 3); KV_ADD(a);
        KV_LOAD(a, 4); KV_ADD(b);
        KV_LOAD(b, 5); KV_ADD(a);
        KV_LOAD(a, 6); KV_ADD(b);
        KV_LOAD(b, 7); KV_ADD(a);
        __syncwarp();
        if (lane == 0) mbar_arrive(&rg.empty[e]);   // the stage is in registers: hand it back
        next_entry();
        if (s + 1 < nst - 1) {
          unsigned ok = mbar_test(&rg.full[e], par);
          KV_ADD(b);
          while (!ok) ok = mbar_test(&rg.full[e], par);
          st = stage_ptr(e);
          KV_LOAD(a, 0);
        } else {
          KV_ADD(b);
        }
      }
#undef KV_LOAD
#undef KV_ADD
    }
    if (nst > 1) boundary(nst - 1);
  } else if (wib == 1 && lane == 0) {
    for (int s = 0; s < nst; ++s) {
      const unsigned no = no0 + (unsigned)s;
      const unsigned e = no % RING_STAGES, use = no / RING_STAGES;
      if (use > 0) mbar_wait(&rg.empty[e], (use - 1u) & 1u);
      const int ub = u0 + s * SU;
      const int nun = u1 - ub < SU ? u1 - ub : SU;
      const unsigned bytes = (unsigned)nun * 512u;
      mbar_arrive_expect_tx(&rg.full[e], bytes);
      bulk_load(rg.base + (size_t)e * RING_ROWS * RING_PITCH, part0 + (size_t)ub * 128, bytes, &rg.full[e]);
    }
  }
  return acc;
}

// Heavy items: a block's first item is static (item blockIdx.x), the following ones come from
// a counter, so a block that is still walking a long chain takes nothing more.  The consumer
// (warp 0) draws the item and posts it for the block's producer thread.
struct ItemMail {
  volatile long long item;
  volatile unsigned seq;
};
__device__ __forceinline__ long long next_item_consumer(ItemMail* mail, unsigned* work, unsigned& seq, int lane) {
  long long nxt = 0;
  if (lane == 0) {
    nxt = (long long)gridDim.x + atomicAdd(&work[2], 1u);
    mail->item = nxt;
    __threadfence_block();
    mail->seq = ++seq;
  }
  return __shfl_sync(0xffffffffu, nxt, 0);
}
__device__ __forceinline__ long long next_item_producer(ItemMail* mail, unsigned& seq) {
  ++seq;
  while (mail->seq < seq) __nanosleep(64);
  __threadfence_block();
  return mail->item;
}

// Light groups are handed out through a counter in the plan (blocks that spend their time on
// long chains take fewer); the last block to finish leaves both words at zero.
__device__ __forceinline__ long long next_light_group(unsigned* work, int kpw, int lane) {
  unsigned b = 0;
  if (lane == 0) b = atomicAdd(&work[0], (unsigned)kpw);
  return (long long)__shfl_sync(0xffffffffu, b, 0);
}
__device__ __forceinline__ void work_epilogue(unsigned* work) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&work[1], 1u) == gridDim.x - 1) {
      work[0] = 0u;
      work[1] = 0u;
      work[2] = 0u;
      work[3] = 0u;
    }
  }
}

template <int VEC, int CPL, int KIND>
__global__ void __launch_bounds__(AP_THREADS, 1)
apply_plan_kernel(const __grid_constant__ TableView var, const __grid_constant__ TableView sa,
                  const __grid_constant__ TableView sb, const __grid_constant__ PlanView pl,
                  const float* __restrict__ grad, const __grid_constant__ ApplyParams p_in,
                  const float* __restrict__ d_hp, uint32_t today, int tpr, int kpw, float* d_adv,
                  int guided) {
  pdl_wait();
  ApplyParams p = p_in;
  if (d_hp) p = derive_params<KIND>(d_hp, var.dim, p_in.update_slots);
  extern __shared__ __align__(128) unsigned char ring[];   // ring, then the light path's staging
  ApplySmem<AP_NW, VEC, CPL>& sm = *reinterpret_cast<ApplySmem<AP_NW, VEC, CPL>*>(ring + RING_BYTES);
  __shared__ unsigned long long full_bar[RING_STAGES], empty_bar[RING_STAGES];
  __shared__ ItemMail mail;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dim = var.dim;
  const long long U = *pl.num;
  const long long* ids = pl.uniq;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RING_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mail.seq = 0u;
    fence_async_smem();
  }
  __syncthreads();

  int H = *pl.heavy_n;
  if (H > pl.heavy_cap) H = pl.heavy_cap;
  const int parts = (dim + 31) / 32;
  const long long items = (long long)H * parts;

#ifdef KVHBM_TRACE
  unsigned long long* trace = g_trace_plan;
  unsigned long long t_start = 0, t_heavy = 0, n_groups = 0;
  if (trace) t_start = gtime_plan();
#endif
  // ---- heavy ids first: their chains are the longest thing in the launch ----
  Ring rg;
  rg.base = ring; rg.full = full_bar; rg.empty = empty_bar; rg.no = 0;
  if (wib < 2 && (wib == 0 || lane == 0)) {
    unsigned seq = 0;
    long long it = blockIdx.x;
    while (it < items) {
      const int h = (int)(it / parts), part = (int)(it - (long long)h * parts);
      const int r = pl.heavy[h];
      const int c = pl.counts[r], off = pl.seg_off[r];
      const int width = dim - part * 32 < 32 ? dim - part * 32 : 32;
      const float acc = chain_item(rg, wib, lane, pl.staged + (size_t)part * pl.staged_units * 128, off, c);
      if (wib == 0) {
        if (lane < width) __stcg(pl.heavy_sum + (size_t)h * pl.sum_dim + part * 32 + lane, acc);
        __threadfence();
        if (lane == 0) atomicAdd(&pl.heavy_done[h], 1u);   // the id is applied by whoever took it below
        it = next_item_consumer(&mail, pl.work, seq, lane);
      } else {
        it = next_item_producer(&mail, seq);
      }
    }
  }
  __syncwarp();

  // the consumer's sub-partition mates start their light work when the block's chains are done
  if ((wib & 3) == 0 && blockIdx.x < items)
    asm volatile("bar.sync 1, %0;" ::"n"(AP_MATES * 32) : "memory");
#ifdef KVHBM_TRACE
  if (trace) t_heavy = gtime_plan();
#endif
  // ---- the applies of the heavy ids: taken first, so that an id's probes and row loads are
  // long done when its chain delivers the sum (apply_group waits right before it needs it) ----
  {
    GradSrc gs;
    gs.grad = pl.heavy_sum; gs.row0 = 0; gs.counts = nullptr; gs.seg_off = nullptr; gs.pos = nullptr;
    gs.heavy_t = 0; gs.hint = pl.hint; gs.cg = true;
    gs.remap = pl.heavy; gs.ready = pl.heavy_done; gs.ready_n = (unsigned)parts; gs.gstride = pl.sum_dim;
    const int kpi = 32 / tpr;
    for (;;) {
      unsigned b = 0;
      if (lane == 0) b = atomicAdd(&pl.work[3], (unsigned)kpi);
      const long long base = (long long)__shfl_sync(0xffffffffu, b, 0);
      if (base >= H) break;
      apply_group<AP_NW, VEC, CPL, KIND, 1, 1>(sm, wib, var, sa, sb, ids, gs, base, H, p, today, tpr, kpi, true);
    }
  }
  // ---- light ids ----
  {
    GradSrc gs;
    gs.grad = grad; gs.row0 = 0; gs.counts = pl.counts; gs.seg_off = pl.seg_off; gs.pos = pl.pos;
    gs.heavy_t = pl.heavy_t; gs.hint = pl.hint; gs.cg = false;
    gs.prefetch = (guided & 2) != 0;
    guided &= 1;
    // Guided sizes: the first three quarters of the ranks go out in groups of `kpw` ids, the
    // last quarter in groups of one round (32 / tpr ids) - what is still unclaimed when the pool
    // runs dry costs the launch one group-time, so the last groups are the short ones.  Both
    // ranges are handed out strided (group g takes ranks g, g + G, g + 2G, ...).
    const int kpi_l = 32 / tpr;
    const long long UA = (guided && kpw > kpi_l) ? U - U / 4 : U;
    const long long GA = (UA + kpw - 1) / kpw;
    const long long GB = (U - UA + kpi_l - 1) / kpi_l;
    for (;;) {
      const long long g = next_light_group(pl.work, 1, lane);
      if (g >= GA + GB) break;
#ifdef KVHBM_TRACE
      if (trace && n_groups == 0) gs.trace = trace + 131072 + ((size_t)blockIdx.x * AP_NW + wib) * 16;
      else gs.trace = nullptr;
#endif
      const bool first = g < GA;
      gs.stride = first ? GA : GB;
      apply_group<AP_NW, VEC, CPL, KIND, 1, 4>(sm, wib, var, sa, sb, ids, gs, first ? g : UA + (g - GA),
                                               first ? UA : U, p, today, tpr, first ? kpw : kpi_l, false);
#ifdef KVHBM_TRACE
      ++n_groups;
#endif
    }
  }
#ifdef KVHBM_TRACE
  if (trace && lane == 0) {
    unsigned long long* r = trace + ((size_t)blockIdx.x * AP_NW + wib) * 4;
    r[0] = t_start; r[1] = t_heavy; r[2] = gtime_plan();
    r[3] = n_groups;
  }
#endif
  work_epilogue(pl.work);

  // AdamOptimizer._finish folded into this launch (see apply.cu)
  if ((KindTraits<KIND>::ADAMISH || KIND == K_ADAM) && d_adv != nullptr) {
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

__global__ void advance_powers_plan_kernel(float* hp, int p, int b) {
  hp[p] = hp[p] * hp[b];
  hp[p + 1] = hp[p + 1] * hp[b + 1];
}

template <int VEC, int CPL, int KIND>
int launch_apply_plan(Table* var, Table* sa, Table* sb, const PlanView& pv, const float* grad,
                      const ApplyParams& p, const float* d_hp, uint16_t today, cudaStream_t st,
                      int tpr, float* d_adv) {
  static const int kpw_env = getenv("KVHBM_APPLYP_KPW") ? atoi(getenv("KVHBM_APPLYP_KPW")) : 0;
  static const int guided_env = (getenv("KVHBM_APPLYP_GUIDED") ? atoi(getenv("KVHBM_APPLYP_GUIDED")) : 1) |
                                ((getenv("KVHBM_APPLYP_PREFETCH") ? atoi(getenv("KVHBM_APPLYP_PREFETCH")) : 1) << 1);
  constexpr int bps_env = 1;
  // One persistent block fills an SM (96 registers x 640 threads), so a full grid starves
  // whatever runs beside the step - the input pipeline's dedup plan of the next batch - until the
  // launch drains, and the chain then waits for that plan: 81.9 us per step.  The launch is
  // bound by the head id's chain and the light path's latency, not by SM count (47.6 us on 148
  // SMs, 48.6 on 132), so one SM in nine is left free: 67.6 us per step with the plan beside
  // it (64.1 with prebuilt plans).  KVHBM_APPLYP_SMS overrides.
  static const int sms_env = getenv("KVHBM_APPLYP_SMS") ? atoi(getenv("KVHBM_APPLYP_SMS")) : 0;
  int sms = sm_count(var->device);
  if (sms >= 36) sms -= sms / 9;
  if (sms_env > 0 && sms_env <= sm_count(var->device)) sms = sms_env;
  const int kpi = 32 / tpr;
  // light warps of a full grid; a Zipf batch has ~n/3 distinct ids
  // Light groups are latency chains: ~3 us of probes, then per round of 32 / tpr ids ~1.8 us for
  // the gradient rows and ~2 us of row math (measured, scripts/trace_apply_plan.py).  Small
  // groups balance best over the dynamic counter (measured at the microbench: 2 rounds per
  // group 50 us, 4 rounds 54-70, 6 rounds 83, 8 rounds 95); they only grow when there would be
  // more than four groups per warp.  A Zipf batch has ~n/3 distinct ids.
  const long long lwarps = (long long)sms * bps_env * AP_NW;
  const long long n_est = (pv.n + 2) / 3;
  int kpw = 2 * kpi <= 32 ? 2 * kpi : kpi;
  while (kpw + kpi <= 32 && (n_est + kpw - 1) / kpw > 4 * lwarps) kpw += kpi;
  if (kpw_env >= kpi && kpw_env <= 32) kpw = kpw_env;
  long long blocks = ((pv.n + kpw - 1) / kpw + AP_NW - 1) / AP_NW;
  if (blocks > (long long)sms * bps_env) blocks = (long long)sms * bps_env;
  if (blocks < 1) blocks = 1;
  auto kern = apply_plan_kernel<VEC, CPL, KIND>;
  const size_t smem = RING_BYTES + sizeof(ApplySmem<AP_NW, VEC, CPL>);
  static bool attr = false;  // per instantiation
  if (!attr) {
    KV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  TableView vb = sb ? sb->view() : sa->view();
  KV_TRY(launch_stage_heavy(pv, grad, var->dim, var->device, st));
  KV_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(AP_THREADS), smem, st, var->view(), sa->view(), vb,
                     pv, grad, p, d_hp, (uint32_t)today, tpr, kpw, d_adv, guided_env));
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
}  // namespace kvhbm
#endif  // KVHBM_APPLY_PLAN_KERNEL_CUH_
