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
#include "apply_plan_kernel.cuh"

namespace kvhbm {

int apply_validate(int kind, Table* var, Table* sa, Table* sb, const float* hp);
// one per optimizer, apply_plan_k*.cu
#define KV_DECL_KIND(N)                                                                             \
  int apply_plan_kind_##N(Table* var, Table* sa, Table* sb, Plan* plan, const float* grad,         \
                          const float* hp, const float* d_hp, int update_slots, uint16_t today,    \
                          cudaStream_t st, float* d_adv);                                          \
  int set_trace_plan_kind_##N(unsigned long long* d_buf);
KV_DECL_KIND(0) KV_DECL_KIND(1) KV_DECL_KIND(2) KV_DECL_KIND(3) KV_DECL_KIND(4) KV_DECL_KIND(5) KV_DECL_KIND(6)
#undef KV_DECL_KIND

int set_trace_apply(unsigned long long* d_buf) {
  KV_TRY(set_trace_plan_local(d_buf));
  KV_TRY(set_trace_plan_kind_0(d_buf)); KV_TRY(set_trace_plan_kind_1(d_buf));
  KV_TRY(set_trace_plan_kind_2(d_buf)); KV_TRY(set_trace_plan_kind_3(d_buf));
  KV_TRY(set_trace_plan_kind_4(d_buf)); KV_TRY(set_trace_plan_kind_5(d_buf));
  return set_trace_plan_kind_6(d_buf);
}

namespace {

// Why the heavy rows are staged first (measured on B200, scripts/ub/rowgather*.cu): one SM
// pulls randomly placed 128-byte lines out of HBM at ~22 GB/s with 9 warps x 16 loads in flight
// (a warp gets ~30 lines per microsecond, whatever the load flavour), while a chain that adds
// one row every 4 cycles consumes 63 GB/s.  Bulk copies of single rows (cp.async.bulk, 128 B
// each) and 16-byte cp.async fed the chain at 40 / 9 ns per row instead of the 2 ns it needs.
// So the random part of the access is spread over the whole chip: stage_heavy_kernel copies
// the occurrence rows of heavy ids into list order (one contiguous run per id and 32-column
// part, 8.8 MB at the microbench), and the chain's feed becomes 8 KB contiguous bulk copies
// out of L2, which a single thread keeps in flight.
//
// Staged layout, per 32-column part: "units" of four consecutive entries of pos (4 rows x 32
// columns, 512 bytes), stored column by column — float 4*c + j of unit u is column c of entry
// 4*u + j — so that a chain lane reads four consecutive rows of its column with one 128-bit
// shared-memory load, conflict-free across the warp.  Part p starts at unit p * staged_units.
template <int VEC>
__global__ void __launch_bounds__(256)
stage_heavy_kernel(const __grid_constant__ PlanView pl, const float* __restrict__ grad, int dim) {
  pdl_wait();
  const long long n = pl.n;
  const long long units = (n + 3) >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (VEC == 4) {
    // a thread owns (unit, part, 4 columns): four 16-byte row loads, a 4x4 transpose in
    // registers, 64 contiguous bytes out
    const int cpr = dim >> 2;                          // 16-byte chunks per row
    const long long total = units * cpr;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += stride) {
      const long long u = t / cpr;
      const int ch = (int)(t - u * cpr), part = ch >> 3, cc = ch & 7;
      const unsigned fl = *reinterpret_cast<const unsigned*>(pl.eflag + 4 * u);   // 4 entry flags
      if (fl == 0u) continue;
      const int4 pz = *reinterpret_cast<const int4*>(pl.pos + 4 * u);
      const int pzs[4] = {pz.x, pz.y, pz.z, pz.w};
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((fl >> (8 * j)) & 0xffu)
          v[j] = __ldcs(reinterpret_cast<const float4*>(grad + (long long)pzs[j] * dim + ch * 4));
      }
      float4* o = reinterpret_cast<float4*>(pl.staged + ((((long long)part * pl.staged_units + u) << 5) + cc * 4) * 4);
      o[0] = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
      o[1] = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
      o[2] = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
      o[3] = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
    }
  } else {
    const long long total = n * dim;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += stride) {
      const long long e = t / dim;
      const int col = (int)(t - e * dim);
      if (!pl.eflag[e]) continue;
      const float v = __ldcs(grad + (long long)__ldg(pl.pos + e) * dim + col);
      pl.staged[(((((long long)(col >> 5) * pl.staged_units + (e >> 2)) << 5) + (col & 31)) << 2) + (e & 3)] = v;
    }
  }
}

// tf.math.unsorted_segment_sum through the plan, on its own: out[r, :] = sum of the rows
// data[pos[seg_off[r] + k], :], k = 0 .. counts[r]-1, in that order from +0.  Same roles as the
// fused kernel: blocks walk the heavy ids' chains through the ring first, then every warp
// takes light segments, a tile per segment.
template <int VEC>
__global__ void __launch_bounds__(AP_THREADS, 1)
segsum_plan_kernel(const __grid_constant__ PlanView pl, const float* __restrict__ data, int dim,
                   float* __restrict__ out, int tpr) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ unsigned long long full_bar[RING_STAGES], empty_bar[RING_STAGES];
  __shared__ ItemMail mail;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long U = *pl.num;
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
  Ring rg;
  rg.base = ring; rg.full = full_bar; rg.empty = empty_bar; rg.no = 0;
  if (wib < 2 && (wib == 0 || lane == 0)) {
    unsigned seq = 0;
    long long it = blockIdx.x;
    while (it < items) {
      const int h = (int)(it / parts), part = (int)(it - (long long)h * parts);
      const int r = pl.heavy[h];
      const int width = dim - part * 32 < 32 ? dim - part * 32 : 32;
      const float acc = chain_item(rg, wib, lane, pl.staged + (size_t)part * pl.staged_units * 128,
                                   pl.seg_off[r], pl.counts[r]);
      if (wib == 0) {
        if (lane < width) out[(long long)r * dim + part * 32 + lane] = acc;
        it = next_item_consumer(&mail, pl.work, seq, lane);
      } else {
        it = next_item_producer(&mail, seq);
      }
    }
  }
  __syncwarp();
  if ((wib & 3) == 0 && blockIdx.x < items)
    asm volatile("bar.sync 1, %0;" ::"n"(AP_MATES * 32) : "memory");
  // a tile of `tpr` lanes per light segment, elements strided by the tile (any dim), the rows
  // of a segment fetched four at a time
  const int tl = lane & (tpr - 1);
  const int kpi = 32 / tpr;
  const int per = (dim / VEC + tpr - 1) / tpr;  // chunks per lane (<= 8)
  for (;;) {
    const long long r = next_light_group(pl.work, kpi, lane) + lane / tpr;
    if (r - lane / tpr >= U) break;
    const int c = r < U ? pl.counts[r] : 0;
    if (c == 0 || c > pl.heavy_t) continue;
    const int* list = pl.pos + pl.seg_off[r];
    float acc[8][VEC];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[q][e] = 0.f;
    if (per <= 2) {
      for (int k0 = 0; k0 < c; k0 += 4) {
        Chunk<VEC> t[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (k0 + j < c) {
            const float* row = data + (long long)__ldg(list + k0 + j) * dim;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int off = (q * tpr + tl) * VEC;
              if (q < per && off < dim) t[j][q].load_stream(row + off); else chunk_zero(t[j][q]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (k0 + j < c)
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
              for (int e = 0; e < VEC; ++e) acc[q][e] += t[j][q].v[e];
      }
    } else {
      for (int k = 0; k < c; ++k) {
        const float* row = data + (long long)__ldg(list + k) * dim;
        Chunk<VEC> t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int off = (q * tpr + tl) * VEC;
          if (q < per && off < dim) t[q].load_stream(row + off); else chunk_zero(t[q]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[q][e] += t[q].v[e];
      }
    }
    float* o = out + r * dim;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int off = (q * tpr + tl) * VEC;
      if (q < per && off < dim) {
        Chunk<VEC> w;
#pragma unroll
        for (int e = 0; e < VEC; ++e) w.v[e] = acc[q][e];
        w.store(o + off);
      }
    }
  }
  work_epilogue(pl.work);
}

// The staging pass that precedes either chain kernel.
int launch_stage_heavy_impl(const PlanView& pv, const float* data, int dim, int device, cudaStream_t st) {
  const int vec = (dim & 3) == 0 ? 4 : 1;
  const long long total = vec == 4 ? ((pv.n + 3) / 4) * (dim / 4) : pv.n * dim;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count(device) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (vec == 4) KV_CUDA(launch_pdl(stage_heavy_kernel<4>, dim3((unsigned)blocks), dim3(256), 0, st, pv, data, dim));
  else KV_CUDA(launch_pdl(stage_heavy_kernel<1>, dim3((unsigned)blocks), dim3(256), 0, st, pv, data, dim));
  KV_LAUNCHED();
  return 0;
}

}  // namespace

int launch_stage_heavy(const PlanView& pv, const float* data, int dim, int device, cudaStream_t st) {
  return launch_stage_heavy_impl(pv, data, dim, device, st);
}

int do_segment_sum_plan(Plan* plan, const float* data, int dim, float* out, cudaStream_t st) {
  if (plan_view(plan).n <= 0) return 0;
  KV_TRY(plan_need_scratch(plan, dim, st));
  const PlanView pv = plan_view(plan);
  RowGeom g = row_geom(dim);
  if (g.cpl > 8) return fail(3, "segment_sum_plan: dim too large");
  static bool attr = false;
  if (!attr) {
    KV_CUDA(cudaFuncSetAttribute(segsum_plan_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 RING_BYTES));
    KV_CUDA(cudaFuncSetAttribute(segsum_plan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 RING_BYTES));
    attr = true;
  }
  int dev = 0;
  KV_CUDA(cudaGetDevice(&dev));
  const int sms = sm_count(dev);
  const long long tiles_per_block = (long long)AP_NW * (32 / g.tpr);
  long long blocks = (pv.n + tiles_per_block - 1) / tiles_per_block;
  if (blocks > (long long)sms) blocks = sms;
  if (blocks < 1) blocks = 1;
  KV_TRY(launch_stage_heavy(pv, data, dim, dev, st));
  if (g.vec == 4)
    segsum_plan_kernel<4><<<(unsigned)blocks, AP_THREADS, RING_BYTES, st>>>(pv, data, dim, out, g.tpr);
  else
    segsum_plan_kernel<1><<<(unsigned)blocks, AP_THREADS, RING_BYTES, st>>>(pv, data, dim, out, g.tpr);
  KV_LAUNCHED();
  return 0;
}

// UnsortedSegmentSum(grad[n, dim], plan.idx) + the apply op `kind` on plan.uniq, in one launch.
int do_apply_plan(int kind, Table* var, Table* sa, Table* sb, Plan* plan, const float* grad,
                  const float* hp, const float* d_hp, int update_slots, uint16_t today,
                  cudaStream_t st, float* d_adv) {
  KV_TRY(apply_validate(kind, var, sa, sb, d_hp ? nullptr : hp));
  switch (kind) {
    case K_ADAGRAD: return apply_plan_kind_0(var, sa, sb, plan, grad, hp, d_hp, update_slots, today, st, d_adv);
    case K_GROUP_ADAM: return apply_plan_kind_1(var, sa, sb, plan, grad, hp, d_hp, update_slots, today, st, d_adv);
    case K_FTRL: return apply_plan_kind_2(var, sa, sb, plan, grad, hp, d_hp, update_slots, today, st, d_adv);
    case K_ADAM: return apply_plan_kind_3(var, sa, sb, plan, grad, hp, d_hp, update_slots, today, st, d_adv);
    case K_GROUP_ADAM_V3: return apply_plan_kind_4(var, sa, sb, plan, grad, hp, d_hp, update_slots, today, st, d_adv);
    case K_FTRL_V2: return apply_plan_kind_5(var, sa, sb, plan, grad, hp, d_hp, update_slots, today, st, d_adv);
    case K_GROUP_FTRL_V2: return apply_plan_kind_6(var, sa, sb, plan, grad, hp, d_hp, update_slots, today, st, d_adv);
  }
  return fail(1, "apply: unknown optimizer kind");
}

}  // namespace kvhbm
