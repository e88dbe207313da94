// apply_math.cuh — what the two fused-apply kernels share (apply.cu: gradient rows already
// summed per unique id; apply_plan.cu: gradient rows per occurrence, summed in the kernel in
// TensorFlow's order): hyper-parameter derivation, the per-row optimizer arithmetic, and the
// warp routine that resolves a group of ids, updates their rows and publishes their flags.
//
// Arithmetic is the reference's, operation by operation, in fp32 without FMA contraction
// (-fmad=false; sqrtf and `/` are IEEE-rounded), including the ORDER of the one reduction on
// the path, ||z||^2 of the group-lasso branch (eigen_sum_tile below).
#ifndef KVHBM_APPLY_MATH_CUH_
#define KVHBM_APPLY_MATH_CUH_

#include "table.h"

namespace kvhbm {

// V1 kinds first; the variants share their row math (see row_update).
enum {
  K_ADAGRAD = 0, K_GROUP_ADAM = 1, K_FTRL = 2, K_ADAM = 3,
  K_GROUP_ADAM_V3 = 4,   // KvVariableGroupSparseApplyAdamV3, training_ops.cc:5710-5965
  K_FTRL_V2 = 5,         // KvVariableSparseApplyFtrlV2, training_ops.cc:281-530
  K_GROUP_FTRL_V2 = 6,   // KvVariableGroupSparseApplyFtrlV2, training_ops.cc:805-1062
  K_NUM_KINDS = 7
};

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

template <int KIND> struct Kind;
template <> struct Kind<K_ADAGRAD> { static constexpr int PARTS = 1; static constexpr bool TWO = false; };
template <> struct Kind<K_GROUP_ADAM> { static constexpr int PARTS = 3; static constexpr bool TWO = false; };
template <> struct Kind<K_FTRL> { static constexpr int PARTS = 2; static constexpr bool TWO = true; };
template <> struct Kind<K_ADAM> { static constexpr int PARTS = 2; static constexpr bool TWO = false; };
template <> struct Kind<K_GROUP_ADAM_V3> { static constexpr int PARTS = 3; static constexpr bool TWO = false; };
template <> struct Kind<K_FTRL_V2> { static constexpr int PARTS = 2; static constexpr bool TWO = true; };
template <> struct Kind<K_GROUP_FTRL_V2> { static constexpr int PARTS = 2; static constexpr bool TWO = true; };

template <int KIND> struct KindTraits {
  static constexpr bool ADAMISH = KIND == K_GROUP_ADAM || KIND == K_GROUP_ADAM_V3;
  static constexpr bool FTRLISH = KIND == K_FTRL || KIND == K_FTRL_V2 || KIND == K_GROUP_FTRL_V2;
  // ops that consult the low-frequency filter / blacklist of the value table
  static constexpr bool FILTERS = KIND != K_ADAM;
  // ops whose value row may be blacklisted by the update (group lasso)
  static constexpr bool LASSO = KIND == K_GROUP_ADAM || KIND == K_GROUP_ADAM_V3 || KIND == K_FTRL ||
                                KIND == K_GROUP_FTRL_V2;
};

// Scalars every row needs, derived from the op's hyper-parameter inputs exactly as the
// reference derives them.  One __host__ __device__ body so that the host path (scalars
// passed by value, as TF HostMemory inputs) and the device path (scalars read from HBM, so
// that a step can be captured in a CUDA graph and replayed while beta^t advances) round
// identically.  `hp` layout per optimizer = the op's scalar inputs in op order:
//   adagrad        [lr]
//   group adam v4  [lr, beta1_power, beta2_power, beta1, beta2, epsilon, l1, l2, l21]
//   group adam v3  same as v4
//   ftrl (all)     [lr, l1, l2, l21, l2_shrinkage, lr_power]  (FtrlV2: l21 unused = 0)
//   adam           [lr, beta1, beta2, epsilon, beta1_power, beta2_power]
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
  if (KindTraits<KIND>::ADAMISH) {
    const float lr = hp[0], b1p = hp[1], b2p = hp[2];
    p.beta1 = hp[3];
    p.beta2 = hp[4];
    p.epsilon = hp[5];
    p.one_minus_beta1 = 1.0f - p.beta1;
    p.one_minus_beta2 = 1.0f - p.beta2;
    if (KIND == K_GROUP_ADAM) {
      const float l1s = hp[6] * lr, l2s = hp[7] * lr, l21s = hp[8] * lr;  // :7111-7113
      p.l1 = l1s;
      p.l2x2 = 2.0f * l2s;
      p.alpha = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);          // :7117-7119
      p.l21_norm = l21s * sqrtf(static_cast<float>(dim));       // :7120
    } else {  // v3, :5846-5849: nothing is pre-scaled by lr
      p.l1 = hp[6];
      p.l2x2 = 2.0f * hp[7];
      p.alpha = sqrtf(1.0f - b2p) / (1.0f - b1p);
      p.l21_norm = hp[8] * sqrtf(static_cast<float>(dim));
    }
    p.later_step = p.beta1 > b1p;                             // :7171, :5899
  } else if (KindTraits<KIND>::FTRLISH) {
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

constexpr unsigned APPLY_FULL = 0xffffffffu;

constexpr int V_SKIP = -1;   // low-frequency key or padding: nothing happens
constexpr int V_ZERO = 0;    // blacklisted key revived at zeros (table_manager.h:359-372)
constexpr int V_COPY = 1;
constexpr int V_CLAIM = 3;   // key inserted by this lane: row starts at the initializer
constexpr int V_KEEP = 4;    // Adam path only: blacklisted var is left alone (kv_variable.h:690)

__device__ __forceinline__ float powp(const ApplyParams& p, float x) {
  return p.fast_sqrt ? sqrtf(x) : powf(x, p.neg_lr_power);
}
__device__ __forceinline__ float clip_l1(float lin, float l1) {
  // linear.cwiseMin(l1).cwiseMax(-l1) with Eigen's mini/maxi
  float a = l1 < lin ? l1 : lin;
  return a < -l1 ? -l1 : a;
}

// `expr.square().sum()` over one row, in the order Eigen's vectorised full reduction adds it
// on the reference's build (SSE2 Packet4f, TensorReduction.h InnerMostDimReducer<.., true,
// false>::reduce; the reference is compiled without -march, kv_variable/BUILD:71-76): four packet accumulators over the
// first (n/16)*16 coefficients, folded ((p0+p1)+p2)+p3, remaining packets into p0, scalar tail
// from 0, result = tail + ((p0[0]+p0[2]) + (p0[1]+p0[3])).  `z2` are the tile's squared
// coefficients; they go through `zs` (shared memory, one region of `dim` floats per tile) so
// that four lanes of the tile can each walk one SSE lane.  Every lane returns the sum.
template <int VEC, int CPL>
__device__ __forceinline__ float eigen_sum_tile(const Chunk<VEC> (&z2)[CPL], int dim, int tpr,
                                                int tl, int tile_lane0, float* zs) {
  __syncwarp();
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const int off = (q * tpr + tl) * VEC;
    if (off < dim) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) zs[off + e] = z2[q].v[e];
    }
  }
  __syncwarp();
  const int npk = dim >> 2;
  const int n4 = npk & ~3;
  auto sse_lane = [&](int m) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int P = 0; P < n4; P += 4) {
      a0 += zs[P * 4 + m];
      a1 += zs[P * 4 + 4 + m];
      a2 += zs[P * 4 + 8 + m];
      a3 += zs[P * 4 + 12 + m];
    }
    if (n4 > 0) { a0 = a0 + a1; a0 = a0 + a2; a0 = a0 + a3; }
    for (int P = n4; P < npk; ++P) a0 += zs[P * 4 + m];
    return a0;
  };
  float x0, x1, x2, x3;
  if (tpr >= 4) {
    const float mine = sse_lane(tl & 3);
    x0 = __shfl_sync(APPLY_FULL, mine, tile_lane0);
    x1 = __shfl_sync(APPLY_FULL, mine, tile_lane0 + 1);
    x2 = __shfl_sync(APPLY_FULL, mine, tile_lane0 + 2);
    x3 = __shfl_sync(APPLY_FULL, mine, tile_lane0 + 3);
  } else {
    x0 = sse_lane(0); x1 = sse_lane(1); x2 = sse_lane(2); x3 = sse_lane(3);
  }
  float tail = 0.f;
  for (int e = npk * 4; e < dim; ++e) tail += zs[e];
  const float pr = (x0 + x2) + (x1 + x3);
  return tail + pr;
}

// Row math of one id, shared by every optimizer.  g/w/s are the tile's register copies of the
// gradient, the value row and the slot parts; on return they hold the updated rows.  Returns
// the blacklist verdict (group lasso only) and the tile-local "some |x| >= cutoff" flags.
template <int VEC, int CPL, int KIND>
__device__ __forceinline__ void row_update(const ApplyParams& p, int dim, int tpr, int tl,
                                           int tile_lane0, float* zs, int vm,
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
  } else if (KIND == K_FTRL_V2) {
    // KvVariableSparseApplyFtrlV2 (has_l2_shrinkage), training_ops.cc:457-478: plain FTRL, no
    // group lasso, no blacklist, no under-threshold refresh.  Lazy Eigen expressions as in
    // SparseGroupFtrl: grad_to_use = grad + 2*l2s*var is re-evaluated by the final
    // `accum += grad_to_use.square()` with the NEW var.
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      *vbig |= chunk_over_cutoff(w[q], DEFAULT_CUTOFF);      // flags stay those of the insert
      *abig |= chunk_over_cutoff(s[0][q], DEFAULT_CUTOFF);
      *bbig |= chunk_over_cutoff(s[1][q], DEFAULT_CUTOFF);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float gg = g[q].v[e];
        const float wv = w[q].v[e];
        const float a = s[0][q].v[e];
        const float gsh = gg + p.shrink2 * wv;
        const float na = a + gsh * gsh;
        const float pna = powp(p, na);
        float lin = s[1][q].v[e];
        lin += gsh - (pna - powp(p, a)) / p.lr * wv;
        const float x = clip_l1(lin, p.l1) - lin;
        const float nw = x / (pna / p.lr + p.l2x2);
        const float g2 = gg + p.shrink2 * nw;
        w[q].v[e] = nw;
        s[0][q].v[e] = a + g2 * g2;
        s[1][q].v[e] = lin;
      }
    }
  } else {
    // GroupAdam v4 (training_ops.cc:7166-7195) / v3 (:5893-5925) / SparseGroupFtrl (:713-751) /
    // GroupSparseApplyFtrlV2 (:976-1013)
    constexpr bool GFV2 = KIND == K_GROUP_FTRL_V2;
    Chunk<VEC> z[CPL], den[CPL], gs[CPL], z2[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float gg = g[q].v[e];
        const float wv = w[q].v[e];
        float lin;
        if (KindTraits<KIND>::ADAMISH) {
          const float m = p.beta1 * s[0][q].v[e] + p.one_minus_beta1 * gg;
          const float vo = s[1][q].v[e];
          const float nv = p.beta2 * vo + p.one_minus_beta2 * (gg * gg);
          const float sq = sqrtf(nv);
          lin = s[2][q].v[e];
          if (KIND == K_GROUP_ADAM) {
            if (p.later_step) lin += p.alpha * m - (sq - sqrtf(vo)) * wv;
            else lin += p.alpha * m - (sq + p.epsilon) * wv;
            den[q].v[e] = sq + p.epsilon + p.l2x2;
          } else {
            // v3: lr divides the curvature terms instead of scaling alpha / l1 / l2 / l21
            if (p.later_step) lin += p.alpha * m - (sq - sqrtf(vo)) / p.lr * wv;
            else lin += p.alpha * m - (sq - sqrtf(vo) + p.epsilon) / p.lr * wv;
            den[q].v[e] = (sq + p.epsilon) / p.lr + p.l2x2;
          }
          s[0][q].v[e] = m;
          s[1][q].v[e] = nv;
          s[2][q].v[e] = lin;
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
        // the vector whose L2 norm decides: l1-clipped linear, or linear itself (GroupFtrlV2)
        const float zz = GFV2 ? lin : clip_l1(lin, p.l1) - lin;
        z[q].v[e] = zz;
        z2[q].v[e] = zz * zz;
      }
    }
    const float ss = eigen_sum_tile<VEC, CPL>(z2, dim, tpr, tl, tile_lane0, zs);
    const float nrm = sqrtf(ss);
    const float thr = GFV2 ? p.l1 : p.l21_norm;
    *black = !(nrm > thr);
    const float c = GFV2 ? p.l1 - nrm : 1.0f - p.l21_norm / nrm;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (!*black) {
          // GroupFtrlV2: coef = (l1 - norm) / ((eta_rec + 2*l2) * norm); var = coef * linear
          if (GFV2) w[q].v[e] = (c / (den[q].v[e] * nrm)) * z[q].v[e];
          else w[q].v[e] = z[q].v[e] * c / den[q].v[e];
        }
        if (KindTraits<KIND>::FTRLISH) {
          // accum += grad_to_use.square(), re-evaluated with the new var (old var after a
          // blacklist: the reference reads the freed row there); GroupFtrlV2 has the
          // statement twice (:1007-1008), restated as written
          const float g2 = *black ? gs[q].v[e] : g[q].v[e] + p.shrink2 * w[q].v[e];
          s[0][q].v[e] += g2 * g2;
          if (GFV2) s[0][q].v[e] += g2 * g2;
        }
      }
      *vbig |= chunk_over_cutoff(w[q], DEFAULT_CUTOFF);
      if (KindTraits<KIND>::ADAMISH) {
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

// One slot-variable table of one key, resolved by the lane that owns the id:
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

// Where the gradient of id i comes from.
//  dense   (counts == nullptr): row (i - row0) of `grad` — the ids were deduplicated and their
//          gradients summed before the op (TF's _deduplicate_indexed_slices);
//  planned (counts != nullptr): `grad` holds one row per id OCCURRENCE of the batch and the
//          dedup plan (dedup.cu) lists, for unique id i, its occurrences pos[seg_off[i] ..
//          + counts[i]) in increasing position: the tile adds those rows one by one, starting
//          from +0, which is the order of TF's UnsortedSegmentSum on the CPU
//          (out[idx[j]] += data[j] for j = 0, 1, ...).  Ids with more than `heavy_t`
//          occurrences belong to the heavy path of apply_plan.cu and are skipped here.
struct GradSrc {
  const float* grad;
  long long row0;
  const int* counts;
  const int* seg_off;
  const int* pos;
  int heavy_t;
  const uint2* hint;     // {slot, ctl} of id i in the value table as this batch's lookup left it, or null
  bool cg;               // dense rows were written by other SMs during this launch: read at L2
  // dense rows produced DURING this launch (the heavy ids of the planned kernel): id i of the
  // group is ids[remap[i]], its gradient row `i` (gstride floats apart) may be read once
  // ready[i] has reached ready_n; the reader puts the counter back to zero
  const int* remap = nullptr;
  unsigned* ready = nullptr;
  unsigned ready_n = 0;
  int gstride = 0;
  // planned groups only: lane l of the group takes id base + l * stride instead of base + l.
  // Ids are in first-occurrence order, so a Zipf batch has its hot ids at the lowest ranks; a
  // group of neighbours there would sum eight 30-row segments one after the other (measured:
  // 66 us for such a group, 17 us for a typical one), a strided group gets one of them at most
  long long stride = 1;
  // planned groups: pull the rows of the group's later rounds (and the occurrences beyond the
  // fourth) into L2 while the first round's loads are in flight - a later round then waits for
  // an L2 hit instead of a DRAM round trip, and it costs no registers
  bool prefetch = false;
#ifdef KVHBM_TRACE
  unsigned long long* trace = nullptr;   // [16] timestamps of one group (tuning aid)
#endif
};

template <int NW, int VEC, int CPL>
struct ApplySmem {
  float* vp[NW][32];
  float* ap[NW][32];
  float* bp[NW][32];
  long long key[NW][32];
  int modes[NW][32];          // vm | am << 8 | bm << 16 (each + 1, so SKIP = 0)
  int goff[NW][32];
  int gcnt[NW][32];
  int gpos[NW][32][4];        // the id's first four occurrence positions (planned gradients)
  unsigned char res[NW][32];  // bit0 v_under, bit1 a_under, bit2 b_under, bit3 black
  float zs[NW][32 * CPL * VEC];
};

// The fused apply of ids [base, base + kpw) ∩ [0, n) by one warp:
//  phase 1, lanes < kpw, one id each: find the value slot (directly at the lookup's hint when
//           there is one, validated by the key) and probe the slot table(s) at the same time
//           (independent loads in flight together), claim missing keys, issue the slot
//           frequency atomics, leave row pointers + modes in shared memory;
//  phase 2, tiles of `tpr` lanes: read gradient + value + slots of one id with 128-bit
//           accesses, update in registers, write back once;
//  phase 3, lanes < kpw: publish flags (under-threshold, blacklist) and finish the atomics.
template <int NW, int VEC, int CPL, int KIND, int UNR, int GU>
__device__ __forceinline__ void apply_group(ApplySmem<NW, VEC, CPL>& sm, int wib,
                                            const TableView& var, const TableView& sa,
                                            const TableView& sb,
                                            const long long* __restrict__ ids, const GradSrc& gs,
                                            long long base, long long n, const ApplyParams& p,
                                            uint32_t today, int tpr, int kpw, bool take_heavy) {
  constexpr int PARTS = Kind<KIND>::PARTS;
  constexpr bool TWO = Kind<KIND>::TWO;
  const int lane = threadIdx.x & 31;
  const int kpi = 32 / tpr;  // ids per tile round
  const int tl = lane & (tpr - 1);
  const int tq = lane / tpr;
  const unsigned tmask = tpr == 32 ? APPLY_FULL : ((1u << tpr) - 1u);
  const int steps = kpw / kpi;
  const int dim = var.dim;
  const bool planned = gs.counts != nullptr;
  const long long i = base + lane * gs.stride;
  const bool valid = lane < kpw && i < n;
#ifdef KVHBM_TRACE
#define KV_STAMP(k) do { if (gs.trace && lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); gs.trace[k] = t_; } } while (0)
#else
#define KV_STAMP(k) do { } while (0)
#endif
  KV_STAMP(0);

  // ---------------- phase 1 ----------------
  long long key = 0;
  int vmode = V_SKIP, amode = V_SKIP, bmode = V_SKIP;
  long long vpos = -1, apos = -1, bpos = -1;
  uint32_t vctl = 0, actl = 0, bctl = 0;
  bool a_lead = false, b_lead = false;
  uint32_t a_old = 0, b_old = 0;
  int gcnt = 1, goff = 0;
  bool mine = valid;
  const long long ii = (valid && gs.remap) ? (long long)gs.remap[i] : i;   // index into ids / hint
  if (valid) key = ids[ii];
  // the lookup's hint is indexed by rank, not by key: fetch it with the key, not after it
  uint32_t hint_slot = 0xffffffffu;
  if (valid && gs.hint) hint_slot = gs.hint[ii].x;
  int gp4[4] = {0, 0, 0, 0};
  if (valid && planned) {
    gcnt = gs.counts[i];
    goff = gs.seg_off[i];
    if (!take_heavy && gcnt > gs.heavy_t) mine = false;
    if (mine) {   // the first occurrences' positions now, so that phase 2 starts with the row loads
#pragma unroll
      for (int j = 0; j < 4; ++j) gp4[j] = j < gcnt ? __ldg(gs.pos + goff + j) : 0;
    }
  }
  if (mine && !key_reserved(key)) {  // padding ids of the shard exchange are skipped
    Probe pv = probe_begin(var, key), pa = probe_begin(sa, key), pb = probe_begin(sb, key);
    int rv = -1, ra = -1, rb = TWO ? -1 : 0;
    Slot sv, ssa, ssb, x0, x1, y0, y1, z0, z1;
    long long hpos = -1;
    if (hint_slot != 0xffffffffu && (unsigned long long)hint_slot <= var.mask) hpos = (long long)hint_slot;
    if (hpos >= 0) x0 = load_slot(var.slots + hpos);
    for (unsigned long long guard = 0; guard <= var.mask + sa.mask + sb.mask; ++guard) {
      if (rv < 0 && hpos < 0) { x0 = load_slot(var.slots + pv.bucket * 2); x1 = load_slot(var.slots + pv.bucket * 2 + 1); }
      if (ra < 0) { y0 = load_slot(sa.slots + pa.bucket * 2); y1 = load_slot(sa.slots + pa.bucket * 2 + 1); }
      if (TWO && rb < 0) { z0 = load_slot(sb.slots + pb.bucket * 2); z1 = load_slot(sb.slots + pb.bucket * 2 + 1); }
      if (rv < 0) {
        if (hpos >= 0) {
          if (x0.key == key) { rv = 1; vpos = hpos; sv = x0; }
          hpos = -1;  // a stale hint: probe from the home bucket on the next round
        } else {
          rv = probe_step(var, key, &pv, x0, x1, &vpos, &sv);
        }
      }
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
        if (KindTraits<KIND>::FILTERS && freq_count(sv.freq) < var.enter_threshold) vmode = V_SKIP;
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
  sm.vp[wib][lane] = row_ptr(var, vctl);
  sm.ap[wib][lane] = row_ptr(sa, actl);
  if (TWO) sm.bp[wib][lane] = row_ptr(sb, bctl);
  sm.key[wib][lane] = key;
  sm.modes[wib][lane] = (vmode + 1) | ((amode + 1) << 8) | ((bmode + 1) << 16);
  sm.goff[wib][lane] = goff;
  sm.gcnt[wib][lane] = gcnt;
  if (planned) {
#pragma unroll
    for (int j = 0; j < 4; ++j) sm.gpos[wib][lane][j] = gp4[j];
  }
  __syncwarp();
  KV_STAMP(1);

  // ---------------- phase 2 ----------------
  float* zs = &sm.zs[wib][tq * tpr * CPL * VEC];
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
        const int modes = sm.modes[wib][kl];
        vm[u] = (modes & 0xff) - 1;
        const int am = ((modes >> 8) & 0xff) - 1;
        const int bm = ((modes >> 16) & 0xff) - 1;
        vp[u] = sm.vp[wib][kl];
        ap[u] = sm.ap[wib][kl];
        if (TWO) bp[u] = sm.bp[wib][kl];
        const bool on = vm[u] != V_SKIP;
        long long v1 = -1, v2 = -1, a1 = -1, a2 = -1, b1 = -1, b2 = -1;
        if (vm[u] == V_CLAIM || am == V_CLAIM || bm == V_CLAIM) {
          const long long k = sm.key[wib][kl];
          if (vm[u] == V_CLAIM) init_rows_of(var, k, &v1, &v2);
          if (am == V_CLAIM) init_rows_of(sa, k, &a1, &a2);
          if (TWO && bm == V_CLAIM) init_rows_of(sb, k, &b1, &b2);
        }
        // value + slot rows first: their loads fly while the gradient rows are summed
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          const int off = (q * tpr + tl) * VEC;
          const bool in = on && off < dim;
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
        if (planned && gs.prefetch && it == 0 && u == 0) {
          // later rounds of this tile: value + slot rows and the first four gradient rows (all
          // addresses are in shared memory since phase 1)
          for (int r2 = UNR; r2 < steps; ++r2) {
            const int k2 = r2 * kpi + tq;
            const int m2 = sm.modes[wib][k2];
            const int v2m = (m2 & 0xff) - 1, a2m = ((m2 >> 8) & 0xff) - 1, b2m = ((m2 >> 16) & 0xff) - 1;
            if (v2m == V_SKIP) continue;
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              const int off = (q * tpr + tl) * VEC;
              if (off >= dim) continue;
              if (v2m == V_COPY || v2m == V_KEEP) prefetch_l2(sm.vp[wib][k2] + off);
#pragma unroll
              for (int r = 0; r < PARTS; ++r) {
                const bool second = TWO && r == 1;
                if ((second ? b2m : a2m) == V_COPY)
                  prefetch_l2((second ? sm.bp[wib][k2] : sm.ap[wib][k2]) + (TWO ? 0 : r * dim) + off);
              }
            }
            const int c2 = sm.gcnt[wib][k2] < 4 ? sm.gcnt[wib][k2] : 4;
            for (int j = 0; j < c2; ++j) {
              const float* gp = gs.grad + (long long)sm.gpos[wib][k2][j] * dim;
#pragma unroll
              for (int q = 0; q < CPL; ++q) {
                const int off = (q * tpr + tl) * VEC;
                if (off < dim) prefetch_l2(gp + off);
              }
            }
          }
        }
        if (!planned) {
          const long long gi = base + kl - gs.row0;
          if (gs.ready && gi < n - gs.row0) {
            // the sum is being produced by a chain of this launch: the value and slot rows are
            // already in flight, wait here
            while (*reinterpret_cast<volatile unsigned*>(gs.ready + gi) < gs.ready_n) __nanosleep(100);
            __threadfence();
          }
          const float* gp = gs.grad + gi * (long long)(gs.gstride ? gs.gstride : dim);
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            const int off = (q * tpr + tl) * VEC;
            if (on && off < dim) {
              if (gs.cg) g[u][q].load_cg(gp + off); else g[u][q].load_stream(gp + off);
            } else chunk_zero(g[u][q]);
          }
        } else {
          // UnsortedSegmentSum of this id's occurrences, in increasing position, from +0
#pragma unroll
          for (int q = 0; q < CPL; ++q) chunk_zero(g[u][q]);
          const int cnt = on ? sm.gcnt[wib][kl] : 0;
          const int* pl = gs.pos + sm.goff[wib][kl];
          for (int k0 = 0; k0 < cnt; k0 += GU) {
            Chunk<VEC> t[GU][CPL];
#pragma unroll
            for (int j = 0; j < GU; ++j) {
              if (k0 + j < cnt) {
                const int pz = (GU == 4 && k0 == 0) ? sm.gpos[wib][kl][j] : __ldg(pl + k0 + j);
                const float* gp = gs.grad + (long long)pz * dim;
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                  const int off = (q * tpr + tl) * VEC;
                  if (off < dim) t[j][q].load_stream(gp + off); else chunk_zero(t[j][q]);
                }
              }
            }
            if (gs.prefetch && k0 == 0 && cnt > GU) {
              // an id with more occurrences: its remaining rows, one occurrence per lane of the tile
              for (int k = GU + tl; k < cnt; k += tpr) {
                const char* gp = reinterpret_cast<const char*>(gs.grad + (long long)__ldg(pl + k) * dim);
                for (int b = 0; b < dim * 4; b += 128) prefetch_l2(gp + b);
              }
            }
#pragma unroll
            for (int j = 0; j < GU; ++j) {
              if (k0 + j < cnt) {
#pragma unroll
                for (int q = 0; q < CPL; ++q)
#pragma unroll
                  for (int e = 0; e < VEC; ++e) g[u][q].v[e] += t[j][q].v[e];
              }
            }
          }
        }
      }
    }
#ifdef KVHBM_TRACE
    if (it == 0 && gs.trace) {   // loads issued / loads landed (forced) / math done
      KV_STAMP(8);
      float sink = g[0][0].v[0] + w[0][0].v[0] + s[0][0][0].v[0] + s[0][PARTS - 1][0].v[0];
      if (sink == 1.2345e-30f) gs.trace[14] = 1;
      KV_STAMP(9);
    }
#endif
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (it + u >= steps) continue;  // uniform across the warp
      const int kl = (it + u) * kpi + tq;
      const bool on = vm[u] != V_SKIP;
      bool vbig, abig, bbig, black;
      row_update<VEC, CPL, KIND>(p, dim, tpr, tl, tq * tpr, zs, vm[u], g[u], w[u], s[u], &vbig,
                                 &abig, &bbig, &black);
#ifdef KVHBM_TRACE
      if (it == 0 && gs.trace) { if (vbig && black && w[u][0].v[0] == 1.2345e-30f) gs.trace[14] = 2; KV_STAMP(10); }
#endif
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
      const unsigned vb = __ballot_sync(APPLY_FULL, vbig);
      const unsigned ab = __ballot_sync(APPLY_FULL, abig);
      const unsigned bb = TWO ? __ballot_sync(APPLY_FULL, bbig) : 0u;
      if (tl == 0)
        sm.res[wib][kl] = (((vb >> sh) & tmask) == 0 ? 1 : 0) | (((ab >> sh) & tmask) == 0 ? 2 : 0) |
                          (((bb >> sh) & tmask) == 0 ? 4 : 0) | (black ? 8 : 0);
    }
    KV_STAMP(2 + (it < 12 ? it : 12));
  }
  __syncwarp();

  // ---------------- phase 3 ----------------
  if (vmode != V_SKIP) {
    const int res = sm.res[wib][lane];
    const bool v_under = res & 1, a_under = res & 2, b_under = res & 4, black = res & 8;
    // ops that never refresh under_threshold_ (no CoverUpdateUnsafe in their loop): the flag is
    // the one the insert computed (kv_variable.h:398), a revived key's is true (:404-406)
    constexpr bool KEEPS_FLAGS = KIND == K_ADAGRAD || KIND == K_FTRL_V2;
    uint32_t nv = CTL_READY | (vctl & CTL_ROW_MASK);
    if (KEEPS_FLAGS) {
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
    if (KEEPS_FLAGS) na |= amode == V_CLAIM ? (a_under ? CTL_UNDER : 0u) : (actl & CTL_UNDER);
    else na |= a_under ? CTL_UNDER : 0u;
    if (na != actl || amode == V_CLAIM) sa.slots[apos].ctl = na;
    if (a_lead) finish_frequency(&sa.slots[apos].freq, a_old, 1u, today);

    if (TWO) {
      uint32_t nb = CTL_READY | (bctl & CTL_ROW_MASK);
      if (KEEPS_FLAGS) nb |= bmode == V_CLAIM ? (b_under ? CTL_UNDER : 0u) : (bctl & CTL_UNDER);
      else nb |= b_under ? CTL_UNDER : 0u;
      if (nb != bctl || bmode == V_CLAIM) sb.slots[bpos].ctl = nb;
      if (b_lead) finish_frequency(&sb.slots[bpos].freq, b_old, 1u, today);
    }
  }
  __syncwarp();
  if (gs.ready && valid) gs.ready[i - gs.row0] = 0u;   // every tile is past its wait
  KV_STAMP(15);
#undef KV_STAMP
}

}  // namespace kvhbm
#endif  // KVHBM_APPLY_MATH_CUH_
