// apply_plan_k0.cu — the fused segment-sum + apply kernel for optimizer kind 0 (apply_math.cuh
// ApplyKind), every row geometry; see apply_plan_kernel.cuh / apply_plan.cu.
#include "apply_plan_kernel.cuh"

namespace kvhbm {

int apply_plan_kind_0(Table* var, Table* sa, Table* sb, Plan* plan, const float* grad, const float* hp,
                      const float* d_hp, int update_slots, uint16_t today, cudaStream_t st,
                      float* d_adv) {
  constexpr int K = 0;
  ApplyParams p{};
  if (d_hp == nullptr) p = derive_params<K>(hp, var->dim, update_slots);
  p.update_slots = update_slots;
  return dispatch_apply_plan<K>(var, sa, Kind<K>::TWO ? sb : nullptr, plan, grad, p, d_hp, today, st,
                                d_adv);
}
int set_trace_plan_kind_0(unsigned long long* d_buf) { return set_trace_plan_local(d_buf); }

}  // namespace kvhbm
