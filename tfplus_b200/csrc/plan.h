// plan.h — device view of a dedup plan (dedup.cu builds it; lookup.cu and apply_plan.cu use it).
#ifndef KVHBM_PLAN_H_
#define KVHBM_PLAN_H_

#include <cuda_runtime.h>
#include <stdint.h>

namespace kvhbm {

// tf.unique_with_counts of one id batch plus the CSR of its occurrences:
//   uniq[r], counts[r]      distinct ids in first-occurrence order and how often each occurs
//   idx[i]                  rank of position i's id
//   pos[seg_off[r] + k]     k-th position holding uniq[r], increasing in k
//   heavy[0 .. *heavy_n)    ranks with more than heavy_t occurrences (eflag marks their entries of pos)
//   first[r]                position of the first occurrence of uniq[r] (= pos[seg_off[r]])
//   hint[r]                 {slot, ctl} of uniq[r] in the value table as the lookup of this
//                           batch left them ({0xffffffff, 0} = unknown: consumers probe)
struct PlanView {
  const long long* uniq;
  const int* idx;
  const int* counts;
  const int* num;
  const int* seg_off;
  const int* pos;
  const int* heavy;
  const int* heavy_n;
  const int* first;
  uint2* hint;
  const unsigned char* eflag;  // [n] 1 where pos[e] belongs to a heavy id
  float* staged;          // [sum_dim / 32][staged_units][32][4]: occurrence rows of the heavy ids
                          // in list order (written by the apply's staging pass, apply_plan.cu)
  long long staged_units; // 512-byte units per 32-column part
  float* heavy_sum;       // [heavy_cap][sum_dim] gradient sums of the heavy ids
  unsigned* heavy_done;   // [heavy_cap] arrival counters, left at zero by every launch
  unsigned* work;         // {next light group, blocks finished}, left at zero by every launch
  int heavy_t;
  int heavy_cap;
  int sum_dim;
  long long n;            // ids in the batch
};

struct Plan;
struct Workspace;

}  // namespace kvhbm
#endif  // KVHBM_PLAN_H_
