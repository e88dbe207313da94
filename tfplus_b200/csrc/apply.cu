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
// IEEE-rounded versions).  Only the L2-norm reduction order differs (a tile
// butterfly instead of Eigen's packet reduction).
#include <cstdlib>

#include "table.h"

namespace kvhbm {

enum { K_ADAGRAD = 0, K_GROUP_ADAM = 1, K_FTRL = 2, K_ADAM = 3 };

struct ApplyParams {
  float lr;
  float beta1, beta2, one_minus_beta1, one_minus_beta2, epsilon;
  float alpha;      // GroupAdam: lr*sqrt(1-b2^t)/(1-b1^t); Adam: lr_t
  float l1, l2x2;   // l1 (scaled), 2*l2 (scaled)
  float l21_norm;   // l21 * sqrt(D)
  float shrink2;    // 2 * l2_shrinkage
  float neg_lr_power;
  int later_step;   // beta1 > beta1_power
  int update_slots;
  int fast_sqrt;    // lr_power == -0.5
};

// Scalars every row needs, derived from the op's hyper-parameter inputs exactly as the
// reference derives them.  One __host__ __device__ body so that the host path (scalars
// passed by value, as TF HostMemory inputs) and the device path (scalars read from HBM, so
// that a step can be captured in a CUDA graph and replayed while beta^t advances) round
// identically.  `hp` layout per optimizer = the op's scalar inputs in op order:
//   adagrad     [lr]
//   group adam  [lr, beta1_power, beta2_power, beta1, beta2, epsilon, l1, l2, l21]
//   ftrl        [lr, l1, l2, l21, l2_shrinkage, lr_power]
//   adam        [lr, beta1, beta2, epsilon, beta1_power, beta2_power]
template <int KIND>
__host__ __device__ __forceinline__ ApplyParams derive_params(const float* hp, int dim,
                                                              int update_slots) {
  ApplyParams p;
  p.lr = hp[0];
  p.beta1 = p.beta2 = p.one_minus_beta1 = p.one_minus_beta2 = p.epsilon = 0.f;
  p.alpha = p.l1 = p.l2x2 = p.l21_norm = p.shrink2 = p.neg_lr_power = 0.f;
  p.later_step = 0;
  p.update_slots = update_slots;
  p.fast_sqrt = 0;
  if (KIND == K_GROUP_ADAM) {
    const float lr = hp[0], b1p = hp[1], b2p = hp[2];
    p.beta1 = hp[3];
    p.beta2 = hp[4];
    p.epsilon = hp[5];
    p.one_minus_beta1 = 1.0f - p.beta1;
    p.one_minus_beta2 = 1.0f - p.beta2;
    const float l1s = hp[6] * lr, l2s = hp[7] * lr, l21s = hp[8] * lr;  // :7111-7113
    p.l1 = l1s;
    p.l2x2 = 2.0f * l2s;
    p.alpha = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);          // :7117-7119
    p.l21_norm = l21s * sqrtf(static_cast<float>(dim));       // :7120
    p.later_step = p.beta1 > b1p;                             // :7171
  } else if (KIND == K_FTRL) {
    p.l1 = hp[1];
    p.l2x2 = 2.0f * hp[2];
    p.l21_norm = hp[3] * sqrtf(static_cast<float>(dim));      // :728
    p.shrink2 = 2.0f * hp[4];
    p.neg_lr_power = -hp[5];
    p.fast_sqrt = hp[5] == -0.5f;                             // :715
  } else if (KIND == K_ADAM) {
    p.beta1 = hp[1];
    p.beta2 = hp[2];
    p.epsilon = hp[3];
    p.one_minus_beta1 = 1.0f - p.beta1;
    p.one_minus_beta2 = 1.0f - p.beta2;
    p.alpha = (hp[0] * sqrtf(1.0f - hp[5])) / (1.0f - hp[4]);  // adam.py:147-148
  }
  return p;
}

// Optional per-warp timeline for kernel tuning (scripts/trace_apply.py).
__device__ unsigned long long* g_trace_apply = nullptr;
int set_trace_apply(unsigned long long* d_buf) {
  KV_CUDA(cudaMemcpyToSymbol(g_trace_apply, &d_buf, sizeof(d_buf)));
  return 0;
}

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned long long gtime_a() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

constexpr int V_SKIP = -1;   // low-frequency key or padding: nothing happens
constexpr int V_ZERO = 0;    // blacklisted key revived at zeros (table_manager.h:359-372)
constexpr int V_COPY = 1;
constexpr int V_CLAIM = 3;   // key inserted by this lane: row starts at the initializer
constexpr int V_KEEP = 4;    // Adam path only: blacklisted var is left alone (kv_variable.h:690)

template <int KIND> struct Kind;
template <> struct Kind<K_ADAGRAD> { static constexpr int PARTS = 1; static constexpr bool TWO = false; };
template <> struct Kind<K_GROUP_ADAM> { static constexpr int PARTS = 3; static constexpr bool TWO = false; };
template <> struct Kind<K_FTRL> { static constexpr int PARTS = 2; static constexpr bool TWO = true; };
template <> struct Kind<K_ADAM> { static constexpr int PARTS = 2; static constexpr bool TWO = false; };

__device__ __forceinline__ float powp(const ApplyParams& p, float x) {
  return p.fast_sqrt ? sqrtf(x) : powf(x, p.neg_lr_power);
}
__device__ __forceinline__ float clip_l1(float lin, float l1) {
  // linear.cwiseMin(l1).cwiseMax(-l1) with Eigen's mini/maxi
  float a = l1 < lin ? l1 : lin;
  return a < -l1 ? -l1 : a;
}

// One slot-variable table of one key, resolved by the tile leader:
// FindOrInsertUnsafe(key, ctx, nullptr), kv_variable.h:382-416 (found: freq += 1, day = today;
// absent: EmbeddingValue ctor freq 1).  `r`/`pos`/`s` come from the read-only probe.  The
// frequency atomic is issued here; *f_old is consumed by finish_frequency after the row math.
template <int KIND>
__device__ __forceinline__ int resolve_slot_table(const TableView& t, long long key, int r,
                                                  long long* pos, Slot s, uint32_t today,
                                                  uint32_t* ctl, bool* f_lead, uint32_t* f_old) {
  bool claimed = false;
  if (r != 1) {
    const int c = claim_slot(t, key, *pos);
    if (c == 1) claimed = true;
    else if (c == 0) *pos = find_or_claim(t, key, &s, &claimed);  // rare: lost the slot to another key
    if (*pos < 0) return V_SKIP;
    if (!claimed && c != 1) {  // a duplicate id of this launch inserted it meanwhile
      s.ctl = ld_acquire_u32(&t.slots[*pos].ctl);
    }
  }
  if (claimed) {
    *ctl = alloc_row(t);
    // Adam reaches its slot through GatherOrInsert: insert_func writes {1, today}
    t.slots[*pos].freq = KIND == K_ADAM ? ((1u << 16) | today) : (1u << 16);
    return V_CLAIM;
  }
  *ctl = s.ctl;
  *f_old = atomicAdd(&t.slots[*pos].freq, 1u << 16);
  *f_lead = true;
  return V_COPY;
}

// Row math of one id, shared by every optimizer.  g/w/s are the tile's register copies of the
// gradient, the value row and the slot parts; on return they hold the updated rows.  Returns
// the blacklist verdict (group lasso only) and the tile-local "some |x| >= cutoff" flags.
template <int VEC, int CPL, int KIND>
__device__ __forceinline__ void row_update(const ApplyParams& p, int dim, int tpr, int vm,
                                           Chunk<VEC> (&g)[CPL], Chunk<VEC> (&w)[CPL],
                                           Chunk<VEC> (&s)[Kind<KIND>::PARTS][CPL], bool* vbig,
                                           bool* abig, bool* bbig, bool* black) {
  *vbig = *abig = *bbig = *black = false;
  if (KIND == K_ADAGRAD) {
    // training_ops.cc:1473-1482.  Under-threshold flags are those of the insert
    // (kv_variable.h:398), Adagrad never refreshes them.
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      *vbig |= chunk_over_cutoff(w[q], DEFAULT_CUTOFF);
      *abig |= chunk_over_cutoff(s[0][q], DEFAULT_CUTOFF);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float gg = g[q].v[e];
        float a = s[0][q].v[e];
        if (p.update_slots) a += gg * gg;
        s[0][q].v[e] = a;
        if (dim > 1) w[q].v[e] -= (p.lr * gg) * (1.0f / sqrtf(a));
        else w[q].v[e] -= (p.lr * gg) / sqrtf(a);
      }
    }
  } else if (KIND == K_ADAM) {
    // python/training/adam.py:116-156, every TF op rounded on its own
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float gg = g[q].v[e];
        const float m_t = (s[0][q].v[e] * p.beta1) + (gg * p.one_minus_beta1);
        const float v_t = (s[1][q].v[e] * p.beta2) + ((gg * gg) * p.one_minus_beta2);
        s[0][q].v[e] = m_t;
        s[1][q].v[e] = v_t;
        if (vm != V_KEEP) w[q].v[e] -= (p.alpha * m_t) / (sqrtf(v_t) + p.epsilon);
      }
      *vbig |= chunk_over_cutoff(w[q], DEFAULT_CUTOFF);
      *abig |= chunk_over_cutoff(s[0][q], DEFAULT_CUTOFF) |
               chunk_over_cutoff(s[1][q], DEFAULT_CUTOFF);
    }
  } else {
    // GroupAdam v4 (training_ops.cc:7166-7195) / SparseGroupFtrl (:713-751)
    Chunk<VEC> z[CPL], den[CPL], gs[CPL];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float gg = g[q].v[e];
        const float wv = w[q].v[e];
        float lin;
        if (KIND == K_GROUP_ADAM) {
          const float m = p.beta1 * s[0][q].v[e] + p.one_minus_beta1 * gg;
          const float vo = s[1][q].v[e];
          const float nv = p.beta2 * vo + p.one_minus_beta2 * (gg * gg);
          const float sq = sqrtf(nv);
          lin = s[2][q].v[e];
          if (p.later_step) lin += p.alpha * m - (sq - sqrtf(vo)) * wv;
          else lin += p.alpha * m - (sq + p.epsilon) * wv;
          s[0][q].v[e] = m;
          s[1][q].v[e] = nv;
          s[2][q].v[e] = lin;
          den[q].v[e] = sq + p.epsilon + p.l2x2;
          gs[q].v[e] = 0.f;
        } else {
          const float a = s[0][q].v[e];
          const float gsh = gg + p.shrink2 * wv;
          const float na = a + gsh * gsh;
          const float pna = powp(p, na);
          lin = s[1][q].v[e];
          lin += gsh - (pna - powp(p, a)) / p.lr * wv;
          s[1][q].v[e] = lin;
          gs[q].v[e] = gsh;
          den[q].v[e] = pna / p.lr + p.l2x2;
        }
        const float zz = clip_l1(lin, p.l1) - lin;
        z[q].v[e] = zz;
        ss += zz * zz;
      }
    }
    for (int o = tpr >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
    const float nrm = sqrtf(ss);
    *black = !(nrm > p.l21_norm);
    const float c = 1.0f - p.l21_norm / nrm;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (!*black) w[q].v[e] = z[q].v[e] * c / den[q].v[e];
        if (KIND == K_FTRL) {
          // accum += grad_to_use.square(), re-evaluated with the new var (old var after a
          // blacklist); see oracle/kv_oracle.cc
          const float g2 = *black ? gs[q].v[e] : g[q].v[e] + p.shrink2 * w[q].v[e];
          s[0][q].v[e] += g2 * g2;
        }
      }
      *vbig |= chunk_over_cutoff(w[q], DEFAULT_CUTOFF);
      if (KIND == K_GROUP_ADAM) {
        *abig |= chunk_over_cutoff(s[0][q], DEFAULT_CUTOFF) |
                 chunk_over_cutoff(s[1][q], DEFAULT_CUTOFF) |
                 chunk_over_cutoff(s[2][q], DEFAULT_CUTOFF);
      } else {
        *abig |= chunk_over_cutoff(s[0][q], DEFAULT_CUTOFF);
        *bbig |= chunk_over_cutoff(s[1][q], DEFAULT_CUTOFF);
      }
    }
  }
}

// The fused apply.  A warp takes `kpw` ids:
//  phase 1, lanes < kpw, one id each: probe the value table and the slot table(s) at the same
//           time (independent loads in flight together), claim missing keys, issue the slot
//           frequency atomics, and leave row pointers + modes in shared memory;
//  phase 2, tiles of `tpr` lanes: read gradient + value + slots of one id with 128-bit
//           accesses (UNR ids in flight per tile), update in registers, write back once;
//  phase 3, lanes < kpw: publish flags (under-threshold, blacklist) and finish the atomics.
// kpw is chosen on the host so that one wave of warps covers the whole launch.
template <int VEC, int CPL, int KIND, int UNR>
__global__ void __launch_bounds__(128, CPL == 1 ? 5 : 1)
apply_kernel(TableView var, TableView sa, TableView sb, const long long* __restrict__ ids,
             const float* __restrict__ grad, long long n, const int* __restrict__ d_n,
             ApplyParams p, const float* __restrict__ d_hp, uint32_t today, int tpr, int kpw,
             float* d_adv) {
  if (d_hp) p = derive_params<KIND>(d_hp, var.dim, p.update_slots);
  constexpr int PARTS = Kind<KIND>::PARTS;
  constexpr bool TWO = Kind<KIND>::TWO;
  __shared__ float* s_vp[4][32];
  __shared__ float* s_ap[4][32];
  __shared__ float* s_bp[4][32];
  __shared__ long long s_key[4][32];
  __shared__ int s_modes[4][32];          // vm | am << 8 | bm << 16 (each + 1, so SKIP = 0)
  __shared__ unsigned char s_res[4][32];  // bit0 v_under, bit1 a_under, bit2 b_under, bit3 black
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long wpb = blockDim.x >> 5;
  const int kpi = 32 / tpr;  // ids per tile round
  const int tl = lane & (tpr - 1);
  const int tq = lane / tpr;
  const unsigned tmask = tpr == 32 ? FULL : ((1u << tpr) - 1u);
  const int steps = kpw / kpi;
  const int dim = var.dim;
  if (d_n) { long long dn = *d_n; if (dn < n) n = dn; }
#ifdef KVHBM_TRACE
  unsigned long long* trace = g_trace_apply;
  unsigned long long t0 = 0, t1 = 0, t2 = 0;
  if (trace) t0 = gtime_a();
#endif

  for (long long base = (blockIdx.x * wpb + wib) * kpw; base < n;
       base += (long long)gridDim.x * wpb * kpw) {
    const long long i = base + lane;
    const bool valid = lane < kpw && i < n;

    // ---------------- phase 1 ----------------
    long long key = 0;
    int vmode = V_SKIP, amode = V_SKIP, bmode = V_SKIP;
    long long vpos = -1, apos = -1, bpos = -1;
    uint32_t vctl = 0, actl = 0, bctl = 0;
    bool a_lead = false, b_lead = false;
    uint32_t a_old = 0, b_old = 0;
    if (valid) key = ids[i];
    if (valid && key != KEY_PAD) {  // padding ids of the shard exchange are skipped
      Probe pv = probe_begin(var, key), pa = probe_begin(sa, key), pb = probe_begin(sb, key);
      int rv = -1, ra = -1, rb = TWO ? -1 : 0;
      Slot sv, ssa, ssb, x0, x1, y0, y1, z0, z1;
      for (unsigned long long guard = 0; guard <= var.mask + sa.mask + sb.mask; ++guard) {
        if (rv < 0) { x0 = load_slot(var.slots + pv.bucket * 2); x1 = load_slot(var.slots + pv.bucket * 2 + 1); }
        if (ra < 0) { y0 = load_slot(sa.slots + pa.bucket * 2); y1 = load_slot(sa.slots + pa.bucket * 2 + 1); }
        if (TWO && rb < 0) { z0 = load_slot(sb.slots + pb.bucket * 2); z1 = load_slot(sb.slots + pb.bucket * 2 + 1); }
        if (rv < 0) rv = probe_step(var, key, &pv, x0, x1, &vpos, &sv);
        if (ra < 0) ra = probe_step(sa, key, &pa, y0, y1, &apos, &ssa);
        if (TWO && rb < 0) rb = probe_step(sb, key, &pb, z0, z1, &bpos, &ssb);
        if (rv >= 0 && ra >= 0 && rb >= 0) break;
      }
      // value table: FindOrInsertUnsafe(key, ctx, &should_filter), kv_variable.h:382-408
      bool claimed = false;
      if (rv != 1) {
        const int c = claim_slot(var, key, vpos);
        if (c == 1) claimed = true;
        else if (c == 0) vpos = find_or_claim(var, key, &sv, &claimed);
        if (!claimed && vpos >= 0) {  // inserted meanwhile by a duplicate id of this launch
          sv.ctl = ld_acquire_u32(&var.slots[vpos].ctl);
          sv.freq = 1u << 16;
        }
      }
      if (vpos >= 0) {
        if (claimed) {
          vctl = alloc_row(var);
          var.slots[vpos].freq = 1u << 16;
          vmode = V_CLAIM;
        } else {
          vctl = sv.ctl;
          if (KIND != K_ADAM && freq_count(sv.freq) < var.enter_threshold) vmode = V_SKIP;
          else if (vctl & CTL_BLACK) vmode = KIND == K_ADAM ? V_KEEP : V_ZERO;
          else vmode = V_COPY;
        }
      }
      if (vmode != V_SKIP) {
        amode = resolve_slot_table<KIND>(sa, key, ra, &apos, ssa, today, &actl, &a_lead, &a_old);
        if (TWO)
          bmode = resolve_slot_table<KIND>(sb, key, rb, &bpos, ssb, today, &bctl, &b_lead, &b_old);
        if (amode == V_SKIP || (TWO && bmode == V_SKIP)) vmode = V_SKIP;
      }
    }
    s_vp[wib][lane] = row_ptr(var, vctl);
    s_ap[wib][lane] = row_ptr(sa, actl);
    if (TWO) s_bp[wib][lane] = row_ptr(sb, bctl);
    s_key[wib][lane] = key;
    s_modes[wib][lane] = (vmode + 1) | ((amode + 1) << 8) | ((bmode + 1) << 16);
    __syncwarp();
#ifdef KVHBM_TRACE
    if (trace) t1 = gtime_a();
#endif

    // ---------------- phase 2 ----------------
    for (int it = 0; it < steps; it += UNR) {
      Chunk<VEC> g[UNR][CPL], w[UNR][CPL], s[UNR][PARTS][CPL];
      int vm[UNR];
      float *vp[UNR], *ap[UNR], *bp[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        vm[u] = V_SKIP;
        vp[u] = ap[u] = bp[u] = nullptr;
        if (it + u < steps) {
          const int kl = (it + u) * kpi + tq;
          const int modes = s_modes[wib][kl];
          vm[u] = (modes & 0xff) - 1;
          const int am = ((modes >> 8) & 0xff) - 1;
          const int bm = ((modes >> 16) & 0xff) - 1;
          vp[u] = s_vp[wib][kl];
          ap[u] = s_ap[wib][kl];
          if (TWO) bp[u] = s_bp[wib][kl];
          const bool on = vm[u] != V_SKIP;
          long long v1 = -1, v2 = -1, a1 = -1, a2 = -1, b1 = -1, b2 = -1;
          if (vm[u] == V_CLAIM || am == V_CLAIM || bm == V_CLAIM) {
            const long long k = s_key[wib][kl];
            if (vm[u] == V_CLAIM) init_rows_of(var, k, &v1, &v2);
            if (am == V_CLAIM) init_rows_of(sa, k, &a1, &a2);
            if (TWO && bm == V_CLAIM) init_rows_of(sb, k, &b1, &b2);
          }
          const float* gp = grad + (base + kl) * (long long)dim;
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            const int off = (q * tpr + tl) * VEC;
            const bool in = on && off < dim;
            if (in) g[u][q].load_stream(gp + off); else chunk_zero(g[u][q]);
            if (in && (vm[u] == V_COPY || vm[u] == V_KEEP)) w[u][q].load_cg(vp[u] + off);
            else if (in && vm[u] == V_CLAIM) init_chunk<VEC>(var, v1, v2, off, w[u][q]);
            else chunk_zero(w[u][q]);
#pragma unroll
            for (int r = 0; r < PARTS; ++r) {
              const bool second = TWO && r == 1;
              const int md = second ? bm : am;
              float* rp = second ? bp[u] : ap[u];
              const int soff = (TWO ? 0 : r * dim) + off;
              if (in && md == V_COPY) s[u][r][q].load_cg(rp + soff);
              else if (in && md == V_CLAIM) {
                if (second) init_chunk<VEC>(sb, b1, b2, soff, s[u][r][q]);
                else init_chunk<VEC>(sa, a1, a2, soff, s[u][r][q]);
              } else chunk_zero(s[u][r][q]);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (it + u >= steps) continue;  // uniform across the warp
        const int kl = (it + u) * kpi + tq;
        const bool on = vm[u] != V_SKIP;
        bool vbig, abig, bbig, black;
        row_update<VEC, CPL, KIND>(p, dim, tpr, vm[u], g[u], w[u], s[u], &vbig, &abig, &bbig, &black);
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          const int off = (q * tpr + tl) * VEC;
          if (on && off < dim) {
            if (vm[u] != V_KEEP) w[u][q].store(vp[u] + off);
#pragma unroll
            for (int r = 0; r < PARTS; ++r) {
              const bool second = TWO && r == 1;
              float* rp = second ? bp[u] : ap[u];
              s[u][r][q].store(rp + (TWO ? 0 : r * dim) + off);
            }
          }
        }
        const int sh = tq * tpr;
        const unsigned vb = __ballot_sync(FULL, vbig);
        const unsigned ab = __ballot_sync(FULL, abig);
        const unsigned bb = TWO ? __ballot_sync(FULL, bbig) : 0u;
        if (tl == 0)
          s_res[wib][kl] = (((vb >> sh) & tmask) == 0 ? 1 : 0) | (((ab >> sh) & tmask) == 0 ? 2 : 0) |
                           (((bb >> sh) & tmask) == 0 ? 4 : 0) | (black ? 8 : 0);
      }
    }
    __syncwarp();
#ifdef KVHBM_TRACE
    if (trace) t2 = gtime_a();
#endif

    // ---------------- phase 3 ----------------
    if (vmode != V_SKIP) {
      const int res = s_res[wib][lane];
      const bool v_under = res & 1, a_under = res & 2, b_under = res & 4, black = res & 8;
      uint32_t nv = CTL_READY | (vctl & CTL_ROW_MASK);
      if (KIND == K_ADAGRAD) {
        if (vmode == V_CLAIM) nv |= v_under ? CTL_UNDER : 0u;       // insert-time flag
        else if (vmode == V_ZERO) nv |= CTL_UNDER;                   // RemoveBlacklistUnsafe
        else nv |= vctl & CTL_UNDER;                                 // untouched
      } else if (KIND == K_ADAM) {
        if (vmode == V_KEEP) nv = vctl;
        else nv |= v_under ? CTL_UNDER : 0u;                         // ScatterUpdate refresh
      } else {
        if (black) nv |= CTL_BLACK | CTL_UNDER;                      // MarkBlacklistUnsafe
        else nv |= v_under ? CTL_UNDER : 0u;                         // CoverUpdateUnsafe
      }
      if (vmode == V_CLAIM || amode == V_CLAIM || bmode == V_CLAIM) __threadfence();
      if (nv != vctl || vmode == V_CLAIM) var.slots[vpos].ctl = nv;

      uint32_t na = CTL_READY | (actl & CTL_ROW_MASK);
      if (KIND == K_ADAGRAD) na |= amode == V_CLAIM ? (a_under ? CTL_UNDER : 0u) : (actl & CTL_UNDER);
      else na |= a_under ? CTL_UNDER : 0u;
      if (na != actl || amode == V_CLAIM) sa.slots[apos].ctl = na;
      if (a_lead) finish_frequency(&sa.slots[apos].freq, a_old, 1u, today);

      if (TWO) {
        const uint32_t nb = CTL_READY | (bctl & CTL_ROW_MASK) | (b_under ? CTL_UNDER : 0u);
        if (nb != bctl || bmode == V_CLAIM) sb.slots[bpos].ctl = nb;
        if (b_lead) finish_frequency(&sb.slots[bpos].freq, b_old, 1u, today);
      }
    }
    __syncwarp();
  }
#ifdef KVHBM_TRACE
  if (trace && lane == 0 && t1) {
    unsigned long long* r = trace + (blockIdx.x * wpb + wib) * 4;
    r[0] = t0; r[1] = gtime_a(); r[2] = t2; r[3] = t1;
  }
#endif
  // AdamOptimizer._finish (beta1_power *= beta1, beta2_power *= beta2; inherited by
  // python/training/group_adam.py) folded into this launch: every block read the powers when
  // it started, so the block that finishes last may advance them for the next step.
  if ((KIND == K_GROUP_ADAM || KIND == K_ADAM) && d_adv != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned done = atomicAdd(&var.ctr->apply_done, 1u);
      if (done == gridDim.x - 1) {
        constexpr int P = KIND == K_GROUP_ADAM ? 1 : 4;   // index of beta1_power in hp
        constexpr int Bt = KIND == K_GROUP_ADAM ? 3 : 1;  // index of beta1 in hp
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
  static const int unr_env = getenv("KVHBM_APPLY_UNR") ? atoi(getenv("KVHBM_APPLY_UNR")) : 0;
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
  const bool two = (CPL == 1) && unr_env == 2;
#define KV_A(U) apply_kernel<VEC, CPL, KIND, U><<<blocks, 128, 0, st>>>(                          \
      var->view(), sa->view(), vb, reinterpret_cast<const long long*>(ids), grad, n, d_n, p,     \
      d_hp, today, tpr, kpw, d_adv)
  if (two) KV_A(2); else KV_A(1);
#undef KV_A
  KV_LAUNCHED();
  return 0;
}

template <int KIND>
int dispatch_apply(Table* var, Table* sa, Table* sb, const int64_t* ids, const float* grad,
                   int64_t n, const int32_t* d_n, const ApplyParams& p, const float* d_hp,
                   uint16_t today, cudaStream_t st, float* d_adv) {
  if (n <= 0) {
    if (d_adv) {  // nothing to update, but the step still counts
      advance_powers_kernel<<<1, 1, 0, st>>>(d_adv, KIND == K_GROUP_ADAM ? 1 : 4,
                                             KIND == K_GROUP_ADAM ? 3 : 1);
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

// Host-side entry points.  `hp` holds the op's scalar inputs on the host (validated as the
// reference validates them) or, when `d_hp` is given, in device memory (not validated: reading
// them back would need a stream synchronisation).
template <int KIND>
static int apply_common(Table* var, Table* sa, Table* sb, const int64_t* ids, const float* grad,
                        int64_t n, const int32_t* d_n, const float* hp, const float* d_hp,
                        int update_slots, uint16_t today, cudaStream_t st,
                        float* d_adv = nullptr) {
  ApplyParams p{};
  if (d_hp == nullptr) p = derive_params<KIND>(hp, var->dim, update_slots);
  p.update_slots = update_slots;
  return dispatch_apply<KIND>(var, sa, sb, ids, grad, n, d_n, p, d_hp, today, st, d_adv);
}

int do_apply_adagrad(Table* var, Table* accum, const int64_t* ids, const float* grad, int64_t n,
                     const int32_t* d_n, const float* hp, const float* d_hp, int update_slots,
                     uint16_t today, cudaStream_t st) {
  KV_TRY(check_initialized(var, "var"));
  KV_TRY(check_initialized(accum, "accum"));
  if (accum->dim != var->dim) return fail(1, "var and accum do not have the same shape");
  return apply_common<K_ADAGRAD>(var, accum, nullptr, ids, grad, n, d_n, hp, d_hp, update_slots,
                                 today, st);
}

int do_apply_group_adam_v4(Table* var, Table* mvl, const int64_t* ids, const float* grad,
                           int64_t n, const int32_t* d_n, const float* hp, const float* d_hp,
                           uint16_t today, cudaStream_t st, float* d_adv) {
  // argument checks of training_ops.cc:7001-7103
  KV_TRY(check_initialized(var, "var"));
  KV_TRY(check_initialized(mvl, "m_v_linear"));
  if (d_hp == nullptr) {
    if (!(hp[0] > 0.f)) return fail(1, "lr is not a positive scalar");
    if (!(hp[6] >= 0.f)) return fail(1, "l1 regularization strength is not a non-negative scalar");
    if (!(hp[7] >= 0.f)) return fail(1, "l2 regularization strength is not a non-negative scalar");
    if (!(hp[8] >= 0.f)) return fail(1, "l21 regularization strength is not a non-negative scalar");
  }
  if (mvl->dim != 3 * var->dim)
    return fail(1, "kv_variable and linear do not have the same shape");
  return apply_common<K_GROUP_ADAM>(var, mvl, nullptr, ids, grad, n, d_n, hp, d_hp, 1, today, st,
                                    d_adv);
}

int do_apply_sparse_group_ftrl(Table* var, Table* accum, Table* linear, const int64_t* ids,
                               const float* grad, int64_t n, const int32_t* d_n, const float* hp,
                               const float* d_hp, uint16_t today, cudaStream_t st) {
  // argument checks of training_ops.cc:560-659
  KV_TRY(check_initialized(var, "var"));
  KV_TRY(check_initialized(accum, "accum"));
  KV_TRY(check_initialized(linear, "linear"));
  if (d_hp == nullptr) {
    if (!(hp[0] > 0.f)) return fail(1, "lr is not a positive scalar");
    if (!(hp[1] >= 0.f)) return fail(1, "l1 regularization strength is not a non-negative scalar");
    if (!(hp[2] >= 0.f)) return fail(1, "l2 regularization strength is not a non-negative scalar");
    if (!(hp[3] >= 0.f)) return fail(1, "l21 regularization strength is not a non-negative scalar");
    if (!(hp[4] >= 0.f))
      return fail(1, "l2 shrinkage regularization strength is not a non-negative scalar");
    if (!(hp[5] <= 0.f)) return fail(1, "lr_power is not a non-positive scalar");
  }
  if (accum->dim != var->dim) return fail(1, "kv_varaible and accum do not have the same shape");
  if (linear->dim != var->dim) return fail(1, "kv_variable and linear do not have the same shape");
  return apply_common<K_FTRL>(var, accum, linear, ids, grad, n, d_n, hp, d_hp, 1, today, st);
}

int do_apply_adam(Table* var, Table* mv, const int64_t* ids, const float* grad, int64_t n,
                  const int32_t* d_n, const float* hp, const float* d_hp, uint16_t today,
                  cudaStream_t st, float* d_adv) {
  KV_TRY(check_initialized(var, "var"));
  KV_TRY(check_initialized(mv, "m_v"));
  if (mv->dim != 2 * var->dim) return fail(1, "m_v slot must have dim 2 * dim(var)");
  return apply_common<K_ADAM>(var, mv, nullptr, ids, grad, n, d_n, hp, d_hp, 1, today, st, d_adv);
}

}  // namespace kvhbm
