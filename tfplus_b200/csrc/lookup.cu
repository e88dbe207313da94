// lookup.cu — find-or-insert / gather, scatter-update, insert-or-update.
//
// Kernel shape shared by everything here: a warp takes 32 ids.  Phase 1, one
// lane per id: hash, probe the 16-byte slots (all 32 probes of the warp are in
// flight together), claim an empty slot with atomicCAS when inserting.  Phase
// 2, the warp moves the 32 rows cooperatively: a row is split over a tile of
// `tpr` lanes doing 128-bit accesses, 32/tpr rows per step, four steps of loads
// issued before the first store.
#include <cstdlib>

#include "async_copy.cuh"
#include "plan.h"
#include "table.h"

namespace kvhbm {

// dedup.cu: the plan machinery the duplicate-safe scatter borrows
Plan* plan_new(int64_t max_ids, int heavy_t, int* rc);
void plan_delete(Plan* p);
int64_t plan_capacity(const Plan* p);
PlanView plan_view(const Plan* p);
int do_plan_build(Plan* p, Workspace* ws, const int64_t* ids, int64_t n, cudaStream_t st);
Workspace* workspace_new();

// Optional per-warp timeline for kernel tuning (scripts/trace_gather.py): when set, every
// gather warp records {start ns, end ns, SM id, unused}.
__device__ unsigned long long* g_trace = nullptr;
int set_trace(unsigned long long* d_buf) {
  KV_CUDA(cudaMemcpyToSymbol(g_trace, &d_buf, sizeof(d_buf)));
  return 0;
}

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %smid;" : "=r"(r));
  return r;
}

// modes of a lane's id after phase 1
constexpr int M_SKIP = -1;   // no output (padding lane / filtered id)
constexpr int M_ZERO = 0;    // output zeros (absent on predict, blacklisted)
constexpr int M_COPY = 1;    // row exists
constexpr int M_FRESH = 2;   // being inserted by another thread of this launch
constexpr int M_CLAIM = 3;   // this lane inserted the key and must fill the row

template <bool SEG>
__device__ __forceinline__ float* out_row(float* out, float* const* seg_out, int seg_len,
                                          long long r, int dim) {
  if (!SEG) return out + r * (long long)dim;
  const long long g = r / seg_len;
  return seg_out[g] + (r - g * seg_len) * (long long)dim;
}

// ---------------------------------------------------------------------------
// KvVariable::FindOrInsert (kv_variable.h:263-380) when INSERT, else
// KvVariable::FindOrZeros (kv_variable.h:239-254).
// ---------------------------------------------------------------------------
//
// SEG: the output is `seg_len`-row segments living at seg_out[0], seg_out[1], … instead of one
// array — the owner side of the shard exchange, where segment g is peer g's receive buffer
// mapped over NVLink, so the lookup and the "rows back" transfer are one kernel.  Padding ids
// are skipped outright there (their rows are never read by the requester).
template <int VEC, int CPL, bool INSERT, int UQ, bool SEG>
__global__ void __launch_bounds__(256)
gather_kernel(TableView t, const long long* __restrict__ ids, const int* __restrict__ counts,
              long long n, float* __restrict__ out, uint32_t today, int tpr, int kpw, int flags,
              float* const* __restrict__ seg_out, int seg_len, const int* __restrict__ d_n,
              uint2* __restrict__ hint) {
  pdl_wait();
  if (d_n) { const long long dn = *d_n; if (dn < n) n = dn; }  // count produced on the device
  // flags bit 1: frequencies untouched; bit 2: no row output at all — the caller only wants
  // the keys resolved (inserted, counted) and `hint[i]` = {slot, ctl} of id i, and moves the
  // rows itself (expand_plan_kernel)
  const bool no_out = (flags & 4) != 0;
  // A warp takes `kpw` ids (lanes < kpw probe): small kpw = more warps, so the machine is
  // full even for a 64 K-id batch and instruction latency hides behind other warps.
  constexpr int UNR = UQ / CPL > 0 ? UQ / CPL : 1;  // rows in flight per lane
  const int lane = threadIdx.x & 31;
  const long long wpb = blockDim.x >> 5;
  const long long warp0 = blockIdx.x * wpb + (threadIdx.x >> 5);
  const long long nwarps = gridDim.x * wpb;
  const int kpi = 32 / tpr;  // rows per step
  const int tl = lane & (tpr - 1);
  const int tq = lane / tpr;
  const unsigned tmask = tpr == 32 ? FULL : ((1u << tpr) - 1u);
  const int dim = t.dim;

  const int steps = kpw / kpi;  // row-movement steps per warp (kpw >= kpi)
  __shared__ const float* s_src[8][32];
  __shared__ signed char s_mode[8][32];
  __shared__ unsigned char s_big[8][32];
  const int wib = threadIdx.x >> 5;
#ifdef KVHBM_TRACE
  unsigned long long* trace = g_trace;
  unsigned long long t_start = 0, t_probe = 0;
  if (trace) t_start = gtime();
#endif
  for (long long base = warp0 * kpw; base < n; base += nwarps * kpw) {
    const long long i = base + lane;
    const bool valid = lane < kpw && i < n;
    const long long key = valid ? ids[i] : 0;
    int mode = M_SKIP;
    const float* src = nullptr;
    float* claim_row = nullptr;
    long long pos = -1;
    uint32_t ctl = 0;

    if (valid && key_reserved(key)) {
      mode = SEG ? M_SKIP : M_ZERO;  // padding id of the shard exchange: no table access
    } else if (valid) {
      Slot s;
      mode = M_ZERO;
      if (INSERT) {
        bool claimed;
        pos = find_or_claim(t, key, &s, &claimed);
        if (pos >= 0) {
          if (claimed) {
            ctl = alloc_row(t);
            claim_row = row_ptr(t, ctl);
            mode = M_CLAIM;
          } else {
            ctl = s.ctl;
            if (!(ctl & CTL_READY)) ctl = ld_acquire_u32(&t.slots[pos].ctl);
            if (!(ctl & CTL_READY)) mode = M_FRESH;
            else if (ctl & CTL_BLACK) mode = M_ZERO;
            else { mode = M_COPY; src = row_ptr(t, ctl); }
          }
        }
      } else {
        pos = find_slot(t, key, &s);
        if (pos >= 0) ctl = s.ctl;
        if (pos >= 0 && (s.ctl & CTL_READY) && !(s.ctl & CTL_BLACK)) {
          mode = M_COPY;
          src = row_ptr(t, s.ctl);
        }
      }
    }

#ifdef KVHBM_TRACE
    if (trace) t_probe = gtime();
#endif
    // find_func / insert_func frequency bookkeeping (kv_variable.h:323-350), aggregated over
    // the duplicates inside this warp.  The atomic is issued now and its result is consumed
    // after the rows have been moved, so its round trip overlaps the row traffic.
    bool f_lead = false;
    uint32_t f_cnt = 0, f_old = 0;
    if (INSERT && !(flags & 1)) {
      const bool has = valid && pos >= 0;
      const unsigned active = __ballot_sync(FULL, has);
      if (has) {
        uint32_t cnt = counts ? saturate_count(counts[i]) : 1u;
        const unsigned peers = __match_any_sync(active, pos);
        uint32_t sum = 0;
        for (unsigned p = peers; p; p &= p - 1)
          sum += __shfl_sync(peers, cnt, __ffs(p) - 1);
        if (lane == __ffs(peers) - 1) {
          f_lead = true;
          f_cnt = sum < 0xFFFFu ? sum : 0xFFFFu;
          f_old = atomicAdd(&t.slots[pos].freq, f_cnt << 16);
        }
      }
    }

    // ---- cooperative row movement ----
    // Row pointers and modes go through shared memory (one broadcast LDS per row) instead of
    // warp shuffles: with only a few warps per scheduler the dependent shuffle chains, not
    // memory, were what a warp spent its time on.
    s_src[wib][lane] = src;
    s_mode[wib][lane] = (signed char)mode;
    __syncwarp();
    if (no_out && mode == M_COPY) {
      // start the row on its way from HBM to L2: the expansion reads it next
      for (int b = 0; b < dim * 4; b += 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(src) + b));
    }
    for (int it = 0; it < (no_out ? 0 : steps); it += UNR) {
      Chunk<VEC> c[UNR][CPL];
      int m[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        m[u] = M_SKIP;
        if (it + u < steps) {
          const int kl = (it + u) * kpi + tq;
          m[u] = s_mode[wib][kl];
          const float* sp = s_src[wib][kl];
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            const int off = (q * tpr + tl) * VEC;
            if (m[u] == M_COPY && off < dim) c[u][q].load_cg(sp + off);
            else chunk_zero(c[u][q]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (it + u < steps) {
          const int kl = (it + u) * kpi + tq;
          bool big = false;
          float* op = out_row<SEG>(out, seg_out, seg_len, base + kl, dim);
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            const int off = (q * tpr + tl) * VEC;
            if (m[u] >= M_ZERO && m[u] <= M_COPY && off < dim) c[u][q].store(op + off);
            big |= chunk_over_cutoff(c[u][q], DEFAULT_CUTOFF);
          }
          if (INSERT) {
            // UpdateUnderThreshold on a hit (kv_variable.h:329): the tile's verdict
            const unsigned bal = __ballot_sync(FULL, big);
            if (tl == 0) s_big[wib][kl] = ((bal >> (tq * tpr)) & tmask) != 0;
          }
        }
      }
    }
    __syncwarp();
    const bool my_under = (INSERT && !no_out) ? !s_big[wib][lane] : false;
    __syncwarp();
    if (INSERT && !no_out && mode == M_COPY && my_under != ((ctl & CTL_UNDER) != 0)) {
      if (my_under) atomicOr(&t.slots[pos].ctl, CTL_UNDER);
      else atomicAnd(&t.slots[pos].ctl, ~CTL_UNDER);
    }
    if (INSERT && f_lead) finish_frequency(&t.slots[pos].freq, f_old, f_cnt, today);

    if (INSERT) {
      // New keys (rare after warm-up): the whole warp builds one row at a time.
      // A lane that lost the claim race to the same key does not wait for the
      // winner: the initial row is a pure function of the key.
      unsigned need = __ballot_sync(FULL, mode >= M_FRESH);
      const int nvec = dim / VEC;
      while (need) {
        const int kl = __ffs(need) - 1;
        need &= need - 1;
        const long long k = shfl_ll(key, kl);
        const int m = __shfl_sync(FULL, mode, kl);
        float* dr = shfl_ptr(claim_row, kl);
        long long r1, r2;
        init_rows_of(t, k, &r1, &r2);
        bool big = false;
        float* op = no_out ? nullptr : out_row<SEG>(out, seg_out, seg_len, base + kl, dim);
        for (int j = lane; j < nvec; j += 32) {
          Chunk<VEC> c;
          init_chunk<VEC>(t, r1, r2, j * VEC, c);
          if (!no_out) c.store(op + j * VEC);
          if (m == M_CLAIM) c.store(dr + j * VEC);
          big |= chunk_over_cutoff(c, DEFAULT_CUTOFF);
        }
        const unsigned bal = __ballot_sync(FULL, big);
        __syncwarp();
        if (m == M_CLAIM && lane == kl) {
          ctl = CTL_READY | (bal ? 0u : CTL_UNDER) | ctl;
          __threadfence();
          st_release_u32(&t.slots[pos].ctl, ctl);
        }
      }
    }
    if (hint != nullptr && valid) {
      // M_FRESH (a duplicate of a key another lane is inserting) cannot happen for the
      // deduplicated ids a plan hands in; it reports "no slot" and the consumer probes
      uint2 h;
      h.x = (pos >= 0 && mode != M_FRESH) ? (uint32_t)pos : 0xffffffffu;
      h.y = (pos >= 0 && mode != M_FRESH) ? ctl : 0u;
      hint[i] = h;
    }
  }
#ifdef KVHBM_TRACE
  if (trace && lane == 0) {
    unsigned long long* r = trace + warp0 * 4;
    r[0] = t_start; r[1] = gtime(); r[2] = smid(); r[3] = t_probe;
  }
#endif
}

// ---------------------------------------------------------------------------
// Bulk-copy (TMA) variant of the gather for 16-byte-multiple rows (KVHBM_GATHER_BULK=1; off
// by default: measured 17.5 us vs 15.8 us for the register path at B = 65 536 — the gather is
// bound by dependent round trips in phase 1, not by bytes in flight).  The same phase 1, but the
// rows never pass through registers.  Every lane that found its key issues ONE
// cp.async.bulk (global -> shared, completion counted on the warp's mbarrier), so all of a
// warp's rows are in flight at once with no register cost; the warp's rows sit contiguously
// in shared memory in id order and leave with ONE bulk store (shared -> global) of up to
// 32 x row bytes.  SASS: UBLKCP.
// ---------------------------------------------------------------------------
template <bool INSERT>
__global__ void __launch_bounds__(128)
gather_bulk_kernel(TableView t, const long long* __restrict__ ids,
                   const int* __restrict__ counts, long long n, float* __restrict__ out,
                   uint32_t today, int tpr, int kpw) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ unsigned long long s_bar[4];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long wpb = blockDim.x >> 5;
  const long long warp0 = blockIdx.x * wpb + wib;
  const long long nwarps = gridDim.x * wpb;
  const int dim = t.dim;
  const unsigned row_bytes = (unsigned)dim * 4u;
  float* stage = reinterpret_cast<float*>(dyn_smem + (size_t)wib * 32 * row_bytes);  // [32][dim]
  unsigned long long* bar = &s_bar[wib];
  const int tl = lane & (tpr - 1);
  const int tq = lane / tpr;
  const int kpi = 32 / tpr;
  const unsigned tmask = tpr == 32 ? FULL : ((1u << tpr) - 1u);
  const int d4 = dim >> 2;
  if (lane == 0) mbar_init(bar, 32);
  fence_async_smem();
  __syncwarp();
  unsigned parity = 0;

  for (long long base = warp0 * kpw; base < n; base += nwarps * kpw) {
    const long long i = base + lane;
    const bool valid = lane < kpw && i < n;
    const long long key = valid ? ids[i] : 0;
    int mode = M_SKIP;
    const float* src = nullptr;
    float* claim_row = nullptr;
    long long pos = -1;
    uint32_t ctl = 0;
    if (valid && key_reserved(key)) {
      mode = M_ZERO;  // padding id of the shard exchange: zeros, no table access
    } else if (valid) {
      Slot s;
      mode = M_ZERO;
      if (INSERT) {
        bool claimed;
        pos = find_or_claim(t, key, &s, &claimed);
        if (pos >= 0) {
          if (claimed) {
            ctl = alloc_row(t);
            claim_row = row_ptr(t, ctl);
            mode = M_CLAIM;
          } else {
            ctl = s.ctl;
            if (!(ctl & CTL_READY)) ctl = ld_acquire_u32(&t.slots[pos].ctl);
            if (!(ctl & CTL_READY)) mode = M_FRESH;
            else if (ctl & CTL_BLACK) mode = M_ZERO;
            else { mode = M_COPY; src = row_ptr(t, ctl); }
          }
        }
      } else {
        pos = find_slot(t, key, &s);
        if (pos >= 0 && (s.ctl & CTL_READY) && !(s.ctl & CTL_BLACK)) {
          mode = M_COPY;
          src = row_ptr(t, s.ctl);
        }
      }
    }
    // rows that exist: one bulk copy each, all in flight together
    float* my_stage = stage + (size_t)lane * dim;
    if (mode == M_COPY) {
      mbar_arrive_expect_tx(bar, row_bytes);
      bulk_load(my_stage, src, row_bytes, bar);
    } else {
      mbar_arrive(bar);
    }

    // frequency bookkeeping while the rows travel (kv_variable.h:323-350)
    bool f_lead = false;
    uint32_t f_cnt = 0, f_old = 0;
    if (INSERT) {
      const bool has = valid && pos >= 0;
      const unsigned active = __ballot_sync(FULL, has);
      if (has) {
        uint32_t cnt = counts ? saturate_count(counts[i]) : 1u;
        const unsigned peers = __match_any_sync(active, pos);
        uint32_t sum = 0;
        for (unsigned p = peers; p; p &= p - 1) sum += __shfl_sync(peers, cnt, __ffs(p) - 1);
        if (lane == __ffs(peers) - 1) {
          f_lead = true;
          f_cnt = sum < 0xFFFFu ? sum : 0xFFFFu;
          f_old = atomicAdd(&t.slots[pos].freq, f_cnt << 16);
        }
      }
    }

    // rows that are zeros / being inserted are produced by the warp (rare)
    unsigned special = __ballot_sync(FULL, valid && mode != M_COPY);
    bool fresh_under = false;
    while (special) {
      const int kl = __ffs(special) - 1;
      special &= special - 1;
      const int m = __shfl_sync(FULL, mode, kl);
      const long long k = shfl_ll(key, kl);
      float* dr = shfl_ptr(claim_row, kl);
      float* sp = stage + (size_t)kl * dim;
      long long r1 = -1, r2 = -1;
      if (m >= M_FRESH) init_rows_of(t, k, &r1, &r2);
      bool big = false;
      for (int j = lane; j < d4; j += 32) {
        Chunk<4> c;
        if (m >= M_FRESH) init_chunk<4>(t, r1, r2, j * 4, c); else chunk_zero(c);
        c.store(sp + j * 4);
        if (m == M_CLAIM) c.store(dr + j * 4);
        big |= chunk_over_cutoff(c, DEFAULT_CUTOFF);
      }
      const unsigned bal = __ballot_sync(FULL, big);
      if (lane == kl) fresh_under = bal == 0;
    }

    mbar_wait(bar, parity);
    parity ^= 1;

    // UpdateUnderThreshold on hits (kv_variable.h:329): tiles scan the staged rows
    bool my_under = (ctl & CTL_UNDER) != 0;
    if (INSERT) {
      for (int it = 0; it * kpi < kpw; ++it) {
        const int kl = it * kpi + tq;
        bool big = false;
        for (int c = tl; c < d4; c += tpr) {
          const float4 v = *reinterpret_cast<const float4*>(stage + (size_t)kl * dim + c * 4);
          big |= fabsf(v.x) >= DEFAULT_CUTOFF || fabsf(v.y) >= DEFAULT_CUTOFF ||
                 fabsf(v.z) >= DEFAULT_CUTOFF || fabsf(v.w) >= DEFAULT_CUTOFF;
        }
        const unsigned bal = __ballot_sync(FULL, big);
        if (lane / kpi == it) my_under = ((bal >> ((lane % kpi) * tpr)) & tmask) == 0;
      }
    }

    // everything the warp owns leaves with one bulk store
    fence_async_smem();
    __syncwarp();
    const long long rows_here = (n - base) < kpw ? (n - base) : kpw;
    if (lane == 0) bulk_store(out + base * (long long)dim, stage, (unsigned)rows_here * row_bytes);

    if (INSERT) {
      if (mode == M_COPY && my_under != ((ctl & CTL_UNDER) != 0)) {
        if (my_under) atomicOr(&t.slots[pos].ctl, CTL_UNDER);
        else atomicAnd(&t.slots[pos].ctl, ~CTL_UNDER);
      }
      if (mode == M_CLAIM) {
        __threadfence();
        st_release_u32(&t.slots[pos].ctl, CTL_READY | (fresh_under ? CTL_UNDER : 0u) | ctl);
      }
      if (f_lead) finish_frequency(&t.slots[pos].freq, f_old, f_cnt, today);
    }
    if (lane == 0) bulk_store_wait_read();  // the staging rows are reused by the next pass
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// KvVariable::ScatterUpdate, kv_variable.h:616-734 (unique ids).
// ---------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ float cwise(float l, float r) {
  // kv_variable_cwise_op.h:19-63; min/max are Eigen's numext::mini/maxi
  switch (OP) {
    case 0: return r;
    case 1: return l + r;
    case 2: return l - r;
    case 3: return l * r;
    case 4: return l / r;
    case 5: return r < l ? r : l;
    default: return l < r ? r : l;
  }
}

// PLANNED: `ids` are the DISTINCT ids of a dedup plan and every id applies ALL its occurrences
// upd[pos[seg_off[i] + k]], k = 0 .. counts[i]-1, one after the other in increasing position -
// ScatterUpdate over raw indices with duplicates (what GradientDescentOptimizer.
// _resource_apply_sparse_duplicate_indices feeds scatter_add), in the order a sequential walk
// of the indices applies them.  Otherwise the ids must be distinct (one update row each).
template <int VEC, int CPL, int OP, bool PLANNED>
__global__ void __launch_bounds__(256)
scatter_kernel(TableView t, const long long* __restrict__ ids, const float* __restrict__ upd,
               long long n_in, int tpr, const int* __restrict__ d_n,
               const int* __restrict__ counts, const int* __restrict__ seg_off,
               const int* __restrict__ plist) {
  long long n = n_in;
  if (PLANNED) { const long long dn = *d_n; if (dn < n) n = dn; }
  const int lane = threadIdx.x & 31;
  const long long wpb = blockDim.x >> 5;
  const long long warp0 = blockIdx.x * wpb + (threadIdx.x >> 5);
  const long long nwarps = gridDim.x * wpb;
  const int kpi = 32 / tpr;
  const int tl = lane & (tpr - 1);
  const int tq = lane / tpr;
  const unsigned tmask = tpr == 32 ? FULL : ((1u << tpr) - 1u);
  const int dim = t.dim;

  for (long long base = warp0 * 32; base < n; base += nwarps * 32) {
    const long long i = base + lane;
    const bool valid = i < n;
    const long long key = valid ? ids[i] : 0;
    int mode = M_SKIP;
    float* row = nullptr;
    long long pos = -1;
    uint32_t ctl = 0;
    int cnt = 1, off0 = 0;
    if (valid && PLANNED) { cnt = counts[i]; off0 = seg_off[i]; }
    if (valid && !key_reserved(key)) {
      Slot s;
      bool claimed;
      pos = find_or_claim(t, key, &s, &claimed);
      if (pos >= 0) {
        if (claimed) {
          ctl = alloc_row(t);
          mode = M_CLAIM;
          row = row_ptr(t, ctl);
        } else {
          ctl = s.ctl;
          // a row pointer is only ever derived from a PUBLISHED ctl: with distinct ids a
          // found key is always published; should a caller break that contract, wait for
          // the claimer of this launch instead of touching row 0
          for (int spin = 0; !(ctl & CTL_READY) && spin < (1 << 14); ++spin)
            ctl = ld_acquire_u32(&t.slots[pos].ctl);
          if (!(ctl & CTL_READY) || (ctl & CTL_BLACK)) mode = M_SKIP;  // :690 blacklisted keys are skipped
          else { mode = M_COPY; row = row_ptr(t, ctl); }
        }
      }
    }
    bool my_under = false;
    for (int it = 0; it < tpr; ++it) {
      const int kl = it * kpi + tq;
      const int m = __shfl_sync(FULL, mode, kl);
      float* rp = shfl_ptr(row, kl);
      const long long k = shfl_ll(key, kl);
      bool big = false;
      const int c_k = PLANNED ? __shfl_sync(FULL, cnt, kl) : 1;
      const int o_k = PLANNED ? __shfl_sync(FULL, off0, kl) : 0;
      if (m >= M_COPY) {
        long long r1 = -1, r2 = -1;
        if (m == M_CLAIM) init_rows_of(t, k, &r1, &r2);
        Chunk<VEC> cur[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          const int off = (q * tpr + tl) * VEC;
          if (off < dim) {
            if (m == M_CLAIM) init_chunk<VEC>(t, r1, r2, off, cur[q]);
            else cur[q].load_cg(rp + off);
          }
        }
        for (int occ = 0; occ < c_k; ++occ) {
          const long long urow = PLANNED ? (long long)__ldg(plist + o_k + occ) : base + kl;
          const float* up = upd + urow * (long long)dim;
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            const int off = (q * tpr + tl) * VEC;
            if (off < dim) {
              Chunk<VEC> u;
              u.load_stream(up + off);
#pragma unroll
              for (int e = 0; e < VEC; ++e) cur[q].v[e] = cwise<OP>(cur[q].v[e], u.v[e]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          const int off = (q * tpr + tl) * VEC;
          if (off < dim) {
            cur[q].store(rp + off);
            big |= chunk_over_cutoff(cur[q], DEFAULT_CUTOFF);
          }
        }
      }
      const unsigned bal = __ballot_sync(FULL, big);
      if (lane / kpi == it) my_under = ((bal >> ((lane % kpi) * tpr)) & tmask) == 0;
    }
    if (mode == M_CLAIM) {
      t.slots[pos].freq = 1u << 16;  // EmbeddingValue ctor: freq_val_ = 1
      __threadfence();
      t.slots[pos].ctl = CTL_READY | (my_under ? CTL_UNDER : 0u) | ctl;
    } else if (mode == M_COPY && my_under != ((ctl & CTL_UNDER) != 0)) {
      t.slots[pos].ctl = my_under ? (ctl | CTL_UNDER) : (ctl & ~CTL_UNDER);
    }
  }
}

// ---------------------------------------------------------------------------
// KvVariable::InsertOrUpdate, kv_variable.h:423-485 (unique ids).
// ---------------------------------------------------------------------------
template <int VEC, int CPL>
__global__ void __launch_bounds__(256)
insert_kernel(TableView t, const long long* __restrict__ ids, const float* __restrict__ values,
              long long n, const uint8_t* __restrict__ filter_out,
              const uint8_t* __restrict__ blacklist, int tpr) {
  const int lane = threadIdx.x & 31;
  const long long wpb = blockDim.x >> 5;
  const long long warp0 = blockIdx.x * wpb + (threadIdx.x >> 5);
  const long long nwarps = gridDim.x * wpb;
  const int kpi = 32 / tpr;
  const int tl = lane & (tpr - 1);
  const int tq = lane / tpr;
  const unsigned tmask = tpr == 32 ? FULL : ((1u << tpr) - 1u);
  const int dim = t.dim;

  for (long long base = warp0 * 32; base < n; base += nwarps * 32) {
    const long long i = base + lane;
    const bool valid = i < n && !(filter_out && filter_out[i]);
    const long long key = valid ? ids[i] : 0;
    int mode = M_SKIP;
    float* row = nullptr;
    long long pos = -1;
    uint32_t ctl = 0;
    bool claimed = false;
    if (valid) {
      Slot s;
      pos = find_or_claim(t, key, &s, &claimed);
      if (pos >= 0) {
        ctl = claimed ? alloc_row(t) : s.ctl;
        // ids must be distinct; should a duplicate of this launch have claimed the key a moment
        // ago, wait for it to publish - a row pointer is only ever derived from a READY ctl
        for (int spin = 0; !claimed && !(ctl & CTL_READY) && spin < (1 << 14); ++spin)
          ctl = ld_acquire_u32(&t.slots[pos].ctl);
        if (!claimed && !(ctl & CTL_READY)) {
          pos = -1;   // never published: skip the key rather than touch row 0
        } else if (blacklist && blacklist[i]) {
          // TableManager::MarkBlacklistUnsafe, table_manager.h:335-357
          if (claimed) {
            t.slots[pos].freq = 1u << 16;
            __threadfence();
            t.slots[pos].ctl = CTL_READY | CTL_BLACK | ctl;
          } else if (!(ctl & CTL_BLACK)) {
            t.slots[pos].ctl = ctl | CTL_BLACK | CTL_UNDER;
          }
        } else {
          mode = M_COPY;
          row = row_ptr(t, ctl);
        }
      }
    }
    bool my_under = false;
    for (int it = 0; it < tpr; ++it) {
      const int kl = it * kpi + tq;
      const int m = __shfl_sync(FULL, mode, kl);
      float* rp = shfl_ptr(row, kl);
      bool big = false;
      if (m == M_COPY) {
        const float* vp = values + (base + kl) * (long long)dim;
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          const int off = (q * tpr + tl) * VEC;
          if (off < dim) {
            Chunk<VEC> c;
            c.load_stream(vp + off);
            c.store(rp + off);
            big |= chunk_over_cutoff(c, DEFAULT_CUTOFF);
          }
        }
      }
      const unsigned bal = __ballot_sync(FULL, big);
      if (lane / kpi == it) my_under = ((bal >> ((lane % kpi) * tpr)) & tmask) == 0;
    }
    if (mode == M_COPY) {
      if (claimed) {
        t.slots[pos].freq = 1u << 16;
        __threadfence();
        t.slots[pos].ctl = CTL_READY | (my_under ? CTL_UNDER : 0u) | ctl;
      } else {
        // a blacklisted key stays blacklisted and under threshold (:463, :840)
        const bool under = (ctl & CTL_BLACK) ? true : my_under;
        const uint32_t nctl = under ? (ctl | CTL_UNDER) : (ctl & ~CTL_UNDER);
        if (nctl != ctl) t.slots[pos].ctl = nctl;
      }
    }
  }
}

// KvVariable::GetCount / GetTimeStamp, kv_variable.h:503-561.
__global__ void get_count_kernel(TableView t, const long long* ids, long long n, int* out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    Slot s;
    out[i] = find_slot(t, ids[i], &s) >= 0 ? (int)freq_count(s.freq) : 0;
  }
}
__global__ void get_timestamp_kernel(TableView t, const long long* ids, long long n,
                                     uint32_t* out, uint32_t today) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    Slot s;
    out[i] = find_slot(t, ids[i], &s) >= 0 ? freq_day(s.freq) : today;
  }
}

// out[i,:] = src[perm[i],:]  /  out[perm[i],:] = src[i,:]
template <bool SCATTER>
__global__ void permute_rows_kernel(const float* __restrict__ src, const int* __restrict__ perm,
                                    long long n, int dim, float* __restrict__ out) {
  const long long total = n * (long long)dim;
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if ((dim & 3) == 0) {
    const int d4 = dim >> 2;
    const long long total4 = n * (long long)d4;
    for (; e < total4; e += stride) {
      const long long r = e / d4;
      const int c = (int)(e - r * d4);
      const long long p = perm[r];
      const long long from = SCATTER ? r : p, to = SCATTER ? p : r;
      reinterpret_cast<float4*>(out)[to * d4 + c] =
          __ldg(reinterpret_cast<const float4*>(src) + from * d4 + c);
    }
  } else {
    for (; e < total; e += stride) {
      const long long r = e / dim;
      const int c = (int)(e - r * dim);
      const long long p = perm[r];
      const long long from = SCATTER ? r : p, to = SCATTER ? p : r;
      out[to * dim + c] = src[from * dim + c];
    }
  }
}

// out[i,:] = src[perm[idx[i]],:] (perm and/or idx may be null = identity): the requester side
// of the shard exchange, un-permuting and expanding the deduplicated rows in one pass.
__global__ void expand_rows_kernel(const float* __restrict__ src, const int* __restrict__ perm,
                                   const int* __restrict__ idx, long long n, int dim,
                                   float* __restrict__ out) {
  const int d4 = dim >> 2;
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if ((dim & 3) == 0) {
    for (; e < n * d4; e += stride) {
      const long long r = e / d4;
      const int c = (int)(e - r * d4);
      long long p = idx ? idx[r] : r;
      if (perm) p = perm[p];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p >= 0) v = __ldg(reinterpret_cast<const float4*>(src) + p * d4 + c);
      reinterpret_cast<float4*>(out)[r * d4 + c] = v;
    }
  } else {
    for (; e < n * dim; e += stride) {
      const long long r = e / dim;
      const int c = (int)(e - r * dim);
      long long p = idx ? idx[r] : r;
      if (perm) p = perm[p];
      out[r * dim + c] = p >= 0 ? src[p * dim + c] : 0.f;
    }
  }
}

// out[perm[i],:] = src[i,:] for i < min(n, *d_n); rows with perm < 0 are dropped
// seg_out != null: destination row p lives at seg_out[p / seg_len] + (p % seg_len) * dim
// (the peers' receive buffers: the gradient exchange is this kernel's stores).
__global__ void scatter_rows_n_kernel(const float* __restrict__ src, const int* __restrict__ perm,
                                      long long n, const int* __restrict__ d_n, int dim,
                                      float* __restrict__ out, float* const* __restrict__ seg_out,
                                      int seg_len) {
  if (d_n) { long long dn = *d_n; if (dn < n) n = dn; }
  const int d4 = dim >> 2;
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if ((dim & 3) == 0) {
    for (; e < n * d4; e += stride) {
      const long long r = e / d4;
      const int c = (int)(e - r * d4);
      const long long p = perm[r];
      if (p < 0) continue;
      float* o = seg_out ? out_row<true>(nullptr, seg_out, seg_len, p, dim) : out + p * dim;
      reinterpret_cast<float4*>(o)[c] = __ldg(reinterpret_cast<const float4*>(src) + r * d4 + c);
    }
  } else {
    for (; e < n * dim; e += stride) {
      const long long r = e / dim;
      const int c = (int)(e - r * dim);
      const long long p = perm[r];
      if (p < 0) continue;
      float* o = seg_out ? out_row<true>(nullptr, seg_out, seg_len, p, dim) : out + p * dim;
      o[c] = src[r * dim + c];
    }
  }
}

template <int VEC, int CPL>
int launch_gather(Table* tb, bool insert, const int64_t* ids, const int32_t* counts, int64_t n,
                  float* out, uint16_t today, cudaStream_t st, int tpr,
                  float* const* seg_out = nullptr, int seg_len = 0, const int32_t* d_n = nullptr,
                  uint2* hint = nullptr) {
  static const int use_bulk = getenv("KVHBM_GATHER_BULK") ? atoi(getenv("KVHBM_GATHER_BULK")) : 0;
  if (use_bulk && !seg_out && !d_n && !hint && out && VEC == 4 && tb->dim * 4 <= 1024) {
    // 4 warps x 32 rows of staging per block
    const int kpi = 32 / tpr;
    int kpw = 32;
    const long long max_warps = (long long)sm_count(tb->device) * 8;
    while (kpw > kpi && (n + kpw - 1) / kpw < max_warps / 2) kpw >>= 1;
    static const int kpw_env2 = getenv("KVHBM_GATHER_KPW") ? atoi(getenv("KVHBM_GATHER_KPW")) : 0;
    if (kpw_env2 >= kpi && kpw_env2 <= 32) kpw = kpw_env2;
    const size_t smem = (size_t)4 * 32 * tb->dim * 4;
    const long long warps = (n + kpw - 1) / kpw;
    const int blocks = blocks_for(warps, 4, tb->device, 8);
    const long long* kk = reinterpret_cast<const long long*>(ids);
    if (insert) {
      static bool attr1 = false;
      if (!attr1) { cudaFuncSetAttribute(gather_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); attr1 = true; }
      gather_bulk_kernel<true><<<blocks, 128, smem, st>>>(tb->view(), kk, counts, n, out, today, tpr, kpw);
    } else {
      static bool attr0 = false;
      if (!attr0) { cudaFuncSetAttribute(gather_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); attr0 = true; }
      gather_bulk_kernel<false><<<blocks, 128, smem, st>>>(tb->view(), kk, counts, n, out, today, tpr, kpw);
    }
    KV_LAUNCHED();
    return 0;
  }
  static const int bs = getenv("KVHBM_GATHER_BS") ? atoi(getenv("KVHBM_GATHER_BS")) : 128;
  const int flags = out == nullptr && seg_out == nullptr ? 4 : 0;
  static const int kpw_env = getenv("KVHBM_GATHER_KPW") ? atoi(getenv("KVHBM_GATHER_KPW")) : 0;
  // ids per warp: measured on B200, a 64 K-id batch is fastest with 32 ids per warp (one
  // wave of ~14 warps per SM, 8 rows in flight per lane); smaller batches use fewer ids per
  // warp so that they still spread over the whole chip
  const int kpi = 32 / tpr;
  int kpw = kpi;
  const long long max_warps = (long long)sm_count(tb->device) * 8;
  while (kpw < 32 && (n + kpw - 1) / kpw > max_warps) kpw <<= 1;
  if (kpw_env >= kpi && kpw_env <= 32) kpw = kpw_env;
  if (flags & 4) {  // probe only: nothing to move, so a lane per id as soon as the chip is full
    static const int kpw_p = getenv("KVHBM_PROBE_KPW") ? atoi(getenv("KVHBM_PROBE_KPW")) : 0;
    kpw = kpi;
    while (kpw < 32 && (n + kpw - 1) / kpw > (long long)sm_count(tb->device) * 16) kpw <<= 1;
    if (kpw_p >= kpi && kpw_p <= 32) kpw = kpw_p;
  }
  const long long warps = (n + kpw - 1) / kpw;
  const int blocks = blocks_for(warps * 32, bs, tb->device, 2048 / bs);
  const long long* k = reinterpret_cast<const long long*>(ids);
#define KV_G(INS) KV_CUDA(launch_pdl(gather_kernel<VEC, CPL, INS, 8, false>, dim3(blocks), dim3(bs), 0, st, tb->view(), k, counts, (long long)n, out, (uint32_t)today, tpr, kpw, flags, (float* const*)nullptr, 0, d_n, hint))
  if (seg_out) {
    KV_CUDA(launch_pdl(gather_kernel<VEC, CPL, true, 8, true>, dim3(blocks), dim3(bs), 0, st, tb->view(), k,
                       counts, (long long)n, (float*)nullptr, (uint32_t)today, tpr, kpw, flags, seg_out,
                       (int)seg_len, d_n, hint));
  } else if (insert) KV_G(true);
  else KV_G(false);
#undef KV_G
  KV_LAUNCHED();
  return 0;
}

template <int VEC, int CPL>
int launch_scatter(Table* tb, int op, const int64_t* ids, const float* upd, int64_t n,
                   cudaStream_t st, int tpr, const PlanView* pv) {
  const int blocks = blocks_for(n, 256, tb->device);
  const long long* k = reinterpret_cast<const long long*>(pv ? pv->uniq : reinterpret_cast<const long long*>(ids));
  TableView v = tb->view();
#define KV_S(OP)                                                                                   \
  case OP:                                                                                         \
    if (pv) scatter_kernel<VEC, CPL, OP, true><<<blocks, 256, 0, st>>>(v, k, upd, n, tpr, pv->num,     \
                                                                    pv->counts, pv->seg_off, pv->pos); \
    else scatter_kernel<VEC, CPL, OP, false><<<blocks, 256, 0, st>>>(v, k, upd, n, tpr, nullptr,        \
                                                                    nullptr, nullptr, nullptr);     \
    break;
  switch (op) {
    KV_S(0) KV_S(1) KV_S(2) KV_S(3) KV_S(4) KV_S(5) KV_S(6)
    default: return fail(1, "KvVariable: unsupported scatter update operation");
  }
#undef KV_S
  KV_LAUNCHED();
  return 0;
}

template <int VEC, int CPL>
int launch_insert(Table* tb, const int64_t* ids, const float* values, int64_t n,
                  const uint8_t* filter_out, const uint8_t* blacklist, cudaStream_t st, int tpr) {
  const int blocks = blocks_for(n, 256, tb->device);
  insert_kernel<VEC, CPL><<<blocks, 256, 0, st>>>(
      tb->view(), reinterpret_cast<const long long*>(ids), values, n, filter_out, blacklist, tpr);
  KV_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------
// Second half of the planned lookup: out[i, :] = row of uniq[idx[i]], for every position of the
// batch.  The keys were resolved (inserted, counted) once per DISTINCT id by gather_kernel in
// its no-output mode, which left {slot, ctl} per distinct id in plan.hint; here nothing probes:
// a warp takes `kpw` positions, one lane per position reads idx and the hint (L2), then tiles
// move the rows, UQ in flight per lane.  The tile that handles an id's FIRST occurrence also
// performs FindOrInsert's UpdateUnderThreshold for it (kv_variable.h:329).
// ---------------------------------------------------------------------------
template <int VEC, int CPL, bool INSERT>
__global__ void __launch_bounds__(256)
expand_plan_kernel(TableView t, const int* __restrict__ idx, const uint2* __restrict__ hint,
                   const int* __restrict__ first, long long n, float* __restrict__ out, int tpr,
                   int kpw) {
  pdl_wait();
  constexpr int UNR = 8 / CPL > 0 ? 8 / CPL : 1;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long wpb = blockDim.x >> 5;
  const long long warp0 = blockIdx.x * wpb + wib;
  const long long nwarps = gridDim.x * wpb;
  const int kpi = 32 / tpr;
  const int tl = lane & (tpr - 1);
  const int tq = lane / tpr;
  const unsigned tmask = tpr == 32 ? FULL : ((1u << tpr) - 1u);
  const int dim = t.dim;
  const int steps = kpw / kpi;
  __shared__ const float* s_src[8][32];
  __shared__ unsigned char s_big[8][32];
  for (long long base = warp0 * kpw; base < n; base += nwarps * kpw) {
    const long long i = base + lane;
    const bool valid = lane < kpw && i < n;
    const float* src = nullptr;
    uint2 h = make_uint2(0xffffffffu, 0u);
    bool lead = false;
    if (valid) {
      const int r = idx[i];
      h = __ldg(hint + r);
      if (INSERT) lead = __ldg(first + r) == (int)i;
      if (h.x != 0xffffffffu && (h.y & CTL_READY) && !(h.y & CTL_BLACK)) src = row_ptr(t, h.y);
    }
    s_src[wib][lane] = src;
    __syncwarp();
    for (int it = 0; it < steps; it += UNR) {
      Chunk<VEC> c[UNR][CPL];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (it + u < steps) {
          const float* sp = s_src[wib][(it + u) * kpi + tq];
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            const int off = (q * tpr + tl) * VEC;
            if (sp != nullptr && off < dim) c[u][q].load_cg(sp + off);
            else chunk_zero(c[u][q]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (it + u < steps) {
          const int kl = (it + u) * kpi + tq;
          bool big = false;
          if (base + kl < n) {
            float* op = out + (base + kl) * (long long)dim;
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              const int off = (q * tpr + tl) * VEC;
              if (off < dim) c[u][q].store_stream(op + off);
              big |= chunk_over_cutoff(c[u][q], DEFAULT_CUTOFF);
            }
          }
          if (INSERT) {
            const unsigned bal = __ballot_sync(FULL, big);
            if (tl == 0) s_big[wib][kl] = ((bal >> (tq * tpr)) & tmask) != 0;
          }
        }
      }
    }
    __syncwarp();
    if (INSERT && lead && src != nullptr) {
      const bool under = !s_big[wib][lane];
      if (under != ((h.y & CTL_UNDER) != 0)) {
        if (under) atomicOr(&t.slots[h.x].ctl, CTL_UNDER);
        else atomicAnd(&t.slots[h.x].ctl, ~CTL_UNDER);
      }
    }
    __syncwarp();
  }
}

template <int VEC, int CPL>
int launch_expand_plan(Table* tb, bool insert, const PlanView& pv, const int* first, float* out,
                       cudaStream_t st, int tpr) {
  static const int kpw_env = getenv("KVHBM_EXPAND_KPW") ? atoi(getenv("KVHBM_EXPAND_KPW")) : 0;
  static const int bs = getenv("KVHBM_EXPAND_BS") ? atoi(getenv("KVHBM_EXPAND_BS")) : 256;
  const long long n = pv.n;
  const int kpi = 32 / tpr;
  int kpw = kpi;
  const long long max_warps = (long long)sm_count(tb->device) * 16;
  while (kpw < 32 && (n + kpw - 1) / kpw > max_warps) kpw <<= 1;
  if (kpw_env >= kpi && kpw_env <= 32) kpw = kpw_env;
  const long long warps = (n + kpw - 1) / kpw;
  const int blocks = blocks_for(warps * 32, bs, tb->device, 2048 / bs);
  const uint2* hint = reinterpret_cast<const uint2*>(pv.hint);
  if (insert)
    KV_CUDA(launch_pdl(expand_plan_kernel<VEC, CPL, true>, dim3(blocks), dim3(bs), 0, st, tb->view(), pv.idx,
                       hint, first, n, out, tpr, kpw));
  else
    KV_CUDA(launch_pdl(expand_plan_kernel<VEC, CPL, false>, dim3(blocks), dim3(bs), 0, st, tb->view(), pv.idx,
                       hint, first, n, out, tpr, kpw));
  KV_LAUNCHED();
  return 0;
}

}  // namespace

// Dispatch on the row geometry.  (VEC, CPL) is one of (4|1) x (1|2|4).
#define KV_DISPATCH_GEOM(g, CALL)                                   \
  do {                                                              \
    if (g.vec == 4) {                                               \
      if (g.cpl == 1) return CALL(4, 1);                            \
      if (g.cpl == 2) return CALL(4, 2);                            \
      if (g.cpl <= 4) return CALL(4, 4);                            \
      return CALL(4, 8);                                            \
    }                                                               \
    if (g.cpl == 1) return CALL(1, 1);                              \
    if (g.cpl == 2) return CALL(1, 2);                              \
    if (g.cpl <= 4) return CALL(1, 4);                              \
    return CALL(1, 8);                                              \
  } while (0)

int do_gather(Table* tb, bool insert, const int64_t* ids, const int32_t* counts, int64_t n,
              float* out, uint16_t today, cudaStream_t st, const int32_t* d_n) {
  if (n <= 0) return 0;
  if (insert) KV_TRY(tb->ensure(n, st));
  RowGeom g = row_geom(tb->dim);
#define CALL(V, C) launch_gather<V, C>(tb, insert, ids, counts, n, out, today, st, g.tpr, nullptr, 0, d_n)
  KV_DISPATCH_GEOM(g, CALL);
#undef CALL
}

// The lookup of a whole batch through its dedup plan: every distinct id is resolved once
// (FindOrInsert with the id's occurrence count, or FindOrZeros), then the rows are expanded to
// all positions.  Same table state and same output as do_gather over the raw ids.
int do_gather_plan(Table* tb, bool insert, const PlanView& pv, const int* first, float* out,
                   uint16_t today, cudaStream_t st) {
  if (pv.n <= 0) return 0;
  if (insert) KV_TRY(tb->ensure(pv.n, st));
  RowGeom g = row_geom(tb->dim);
  uint2* hint = reinterpret_cast<uint2*>(pv.hint);
  const int64_t* ids = reinterpret_cast<const int64_t*>(pv.uniq);
#define CALL(V, C) [&]() { \
    int rc = launch_gather<V, C>(tb, insert, ids, insert ? pv.counts : nullptr, pv.n, nullptr, today, st, g.tpr, nullptr, 0, pv.num, hint); \
    if (rc) return rc; \
    return launch_expand_plan<V, C>(tb, insert, pv, first, out, st, g.tpr); }()
  KV_DISPATCH_GEOM(g, CALL);
#undef CALL
}

// FindOrInsert whose output rows land in per-segment buffers (peer memory): row r goes to
// seg_out[r / seg_len] + (r % seg_len) * dim.  seg_out is a device array of device pointers.
int do_gather_segments(Table* tb, const int64_t* ids, const int32_t* counts, int64_t n,
                       float* const* seg_out, int64_t seg_len, uint16_t today, cudaStream_t st) {
  if (n <= 0) return 0;
  if (!seg_out || seg_len <= 0 || seg_len > 0x7fffffff)
    return fail(1, "gather_segments: needs segment pointers and a positive segment length");
  KV_TRY(tb->ensure(n, st));
  RowGeom g = row_geom(tb->dim);
#define CALL(V, C) launch_gather<V, C>(tb, true, ids, counts, n, nullptr, today, st, g.tpr, seg_out, (int)seg_len)
  KV_DISPATCH_GEOM(g, CALL);
#undef CALL
}

// unique_ids: the caller guarantees distinct ids (TF's _deduplicate_indexed_slices ran before
// the op: the tfplus-Adam path).  Otherwise duplicates are applied one after the other in
// index order, through a dedup plan of the ids kept with the table.
int do_scatter(Table* tb, int op, const int64_t* ids, const float* upd, int64_t n,
               cudaStream_t st, bool unique_ids) {
  if (n <= 0) return 0;
  KV_TRY(tb->ensure(n, st));
  RowGeom g = row_geom(tb->dim);
  PlanView pv;
  const PlanView* pvp = nullptr;
  if (!unique_ids && n > 1) {
    if (tb->scatter_plan == nullptr || plan_capacity(tb->scatter_plan) < n) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) cudaGetLastError();
      if (cs != cudaStreamCaptureStatusNone)
        return fail(2, "scatter: the duplicate-safe path sizes its dedup plan on first use; run "
                       "the call once outside CUDA-graph capture");
      KV_CUDA(cudaStreamSynchronize(st));
      if (tb->scatter_plan) plan_delete(tb->scatter_plan);
      int rc = 0;
      int64_t cap = 1024;
      while (cap < n) cap <<= 1;
      tb->scatter_plan = plan_new(cap, 0, &rc);
      if (rc) return rc;
      if (tb->scatter_ws == nullptr) tb->scatter_ws = workspace_new();
    }
    KV_TRY(do_plan_build(tb->scatter_plan, tb->scatter_ws, ids, n, st));
    pv = plan_view(tb->scatter_plan);
    pvp = &pv;
  }
#define CALL(V, C) launch_scatter<V, C>(tb, op, ids, upd, n, st, g.tpr, pvp)
  KV_DISPATCH_GEOM(g, CALL);
#undef CALL
}

int do_insert(Table* tb, const int64_t* ids, const float* values, int64_t n,
              const uint8_t* filter_out, const uint8_t* blacklist, cudaStream_t st) {
  if (n <= 0) return 0;
  KV_TRY(tb->ensure(n, st));
  RowGeom g = row_geom(tb->dim);
#define CALL(V, C) launch_insert<V, C>(tb, ids, values, n, filter_out, blacklist, st, g.tpr)
  KV_DISPATCH_GEOM(g, CALL);
#undef CALL
}

int do_get_count(Table* tb, const int64_t* ids, int64_t n, int32_t* out, cudaStream_t st) {
  if (n <= 0) return 0;
  get_count_kernel<<<blocks_for(n, 256, tb->device), 256, 0, st>>>(
      tb->view(), reinterpret_cast<const long long*>(ids), n, out);
  KV_LAUNCHED();
  return 0;
}
int do_get_timestamp(Table* tb, const int64_t* ids, int64_t n, uint32_t* out, uint16_t today,
                     cudaStream_t st) {
  if (n <= 0) return 0;
  get_timestamp_kernel<<<blocks_for(n, 256, tb->device), 256, 0, st>>>(
      tb->view(), reinterpret_cast<const long long*>(ids), n, out, today);
  KV_LAUNCHED();
  return 0;
}

int do_expand_rows(const float* src, const int32_t* perm, const int32_t* idx, int64_t n, int dim,
                   float* out, cudaStream_t st) {
  if (n <= 0) return 0;
  int dev = 0;
  KV_CUDA(cudaGetDevice(&dev));
  const int64_t work = n * (int64_t)((dim & 3) == 0 ? dim / 4 : dim);
  expand_rows_kernel<<<blocks_for(work, 256, dev, 16), 256, 0, st>>>(src, perm, idx, n, dim, out);
  KV_LAUNCHED();
  return 0;
}

int do_scatter_rows_n(const float* src, const int32_t* perm, int64_t n, const int32_t* d_n,
                      int dim, float* out, float* const* seg_out, int64_t seg_len,
                      cudaStream_t st) {
  if (n <= 0) return 0;
  if (!out && (!seg_out || seg_len <= 0))
    return fail(1, "scatter_rows_n: needs an output array or segment pointers");
  int dev = 0;
  KV_CUDA(cudaGetDevice(&dev));
  const int64_t work = n * (int64_t)((dim & 3) == 0 ? dim / 4 : dim);
  scatter_rows_n_kernel<<<blocks_for(work, 256, dev, 16), 256, 0, st>>>(
      src, perm, n, d_n, dim, out, out ? nullptr : seg_out, (int)seg_len);
  KV_LAUNCHED();
  return 0;
}

int do_permute_rows(bool scatter, const float* src, const int32_t* perm, int64_t n, int dim,
                    float* out, cudaStream_t st) {
  if (n <= 0) return 0;
  int dev = 0;
  KV_CUDA(cudaGetDevice(&dev));
  const int64_t work = n * (int64_t)((dim & 3) == 0 ? dim / 4 : dim);
  const int blocks = blocks_for(work, 256, dev, 16);
  if (scatter) permute_rows_kernel<true><<<blocks, 256, 0, st>>>(src, perm, n, dim, out);
  else permute_rows_kernel<false><<<blocks, 256, 0, st>>>(src, perm, n, dim, out);
  KV_LAUNCHED();
  return 0;
}

}  // namespace kvhbm
