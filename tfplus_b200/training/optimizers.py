"""AdamOptimizer, AdagradOptimizer, GroupAdamOptimizer, SparseGroupFtrlOptimizer (+ SGD) for
KvVariables — mirrors of tfplus/kv_variable/python/training/{adam,adagrad,group_adam,
sparse_group_ftrl,gradient_descent}.py.

apply_gradients takes [(IndexedSlices, KvVariable)] like the TF v1 optimizers.  As in TF
(`Optimizer._resource_apply_sparse_duplicate_indices` -> `_deduplicate_indexed_slices`) the
gradient is first deduplicated with unique + unsorted_segment_sum, then the fused op runs on
the unique ids.  Slot variables are KvVariables themselves, created like
variable_scope.py:1027-1093 does (a slot of a KvVariable is a KvVariable of dim
D * num_concat_opt_vars with a constant initializer).
"""
import zlib

import torch

from .. import ops
from ..kv_variable import IndexedSlices, KvVariable, PartitionedKvVariable


class _Optimizer:
  def __init__(self, name, use_locking=False):
    self._name = name
    self._use_locking = use_locking
    self._slots = {}

  # slot_creator wrapper, variable_scope.py:1027-1093
  def _zeros_slot(self, var, slot_name, width=1, value=0.0):
    key = (var.name, slot_name)
    if key not in self._slots:
      self._slots[key] = KvVariable("%s/%s/%s" % (var.name, self._name, slot_name),
                                    var.embedding_dim * width, initializer=float(value),
                                    device=var.device, init_rows=16,
                                    seed=zlib.crc32(str(key).encode()) % (2 ** 31) + 1)
    return self._slots[key]

  def get_slot(self, var, name):
    return self._slots.get((var.name, name))

  def _create_slots(self, var):
    pass

  def _finish(self):
    pass

  @staticmethod
  def _deduplicate_indexed_slices(values, indices):
    """TF training/optimizer.py: unique + unsorted_segment_sum."""
    indices = indices.reshape(-1)
    uniq, idx = ops.unique(indices)
    summed = ops.unsorted_segment_sum(values.reshape(indices.numel(), -1), idx, uniq.numel())
    return summed, uniq

  def apply_gradients(self, grads_and_vars, global_step=None, name=None):
    for grad, var in grads_and_vars:
      if isinstance(var, PartitionedKvVariable):
        raise NotImplementedError("apply per partition: pass (grad_i, part_i) pairs")
      if not isinstance(grad, IndexedSlices):
        raise TypeError("KvVariable gradients are IndexedSlices")
      self._create_slots(var)
      summed, uniq = self._deduplicate_indexed_slices(grad.values.to(var.device),
                                                      grad.indices.to(var.device))
      self._resource_apply_sparse(summed, var, uniq)
    self._finish()


class GradientDescentOptimizer(_Optimizer):
  """training/gradient_descent.py: var.scatter_sub(lr * grad)."""

  def __init__(self, learning_rate, use_locking=False, name="GradientDescent"):
    super().__init__(name, use_locking)
    self._lr = float(learning_rate)

  def _resource_apply_sparse(self, grad, var, indices):
    var.scatter_sub(IndexedSlices(grad * torch.tensor(self._lr, device=grad.device), indices))


class AdagradOptimizer(_Optimizer):
  """training/adagrad.py:33-45 -> KvVariableSparseApplyAdagrad."""

  def __init__(self, learning_rate, initial_accumulator_value=0.1, use_locking=False,
               name="Adagrad"):
    super().__init__(name, use_locking)
    if initial_accumulator_value <= 0.0:
      raise ValueError("initial_accumulator_value must be positive: %s" % initial_accumulator_value)
    self._lr, self._init_acc = float(learning_rate), float(initial_accumulator_value)

  def _create_slots(self, var):
    self._zeros_slot(var, "accumulator", value=self._init_acc)

  def _resource_apply_sparse(self, grad, var, indices):
    acc = self.get_slot(var, "accumulator")
    ops.kv_variable_sparse_apply_adagrad(var.handle, acc.handle, self._lr, grad, indices,
                                         use_locking=True)


class AdamOptimizer(_Optimizer):
  """training/adam.py.  version 1/2 with the concatenated m_v slot (adam.py:83-86): the
  reference issues gather(m_v) + elementwise TF ops + scatter_update(m_v) + scatter_sub(var);
  `fused=True` (default) runs the same separately-rounded arithmetic in one kernel."""

  def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8,
               use_locking=False, name="Adam", version=2, fused=True):
    super().__init__(name, use_locking)
    self._lr, self._beta1, self._beta2, self._epsilon = (float(learning_rate), float(beta1),
                                                         float(beta2), float(epsilon))
    self._fused = fused
    # non-slot variables beta1_power / beta2_power, fp32 like TF's (adam.py:66-75)
    self._beta1_power = torch.tensor(self._beta1, dtype=torch.float32)
    self._beta2_power = torch.tensor(self._beta2, dtype=torch.float32)

  def _create_slots(self, var):
    var.num_concat_opt_vars = 2
    self._zeros_slot(var, "m_v", width=2)

  def _resource_apply_sparse(self, grad, var, indices):
    m_v = self.get_slot(var, "m_v")
    b1p, b2p = float(self._beta1_power), float(self._beta2_power)
    if self._fused:
      ops.kv_variable_sparse_apply_adam(var.handle, m_v.handle, grad, indices, self._lr,
                                        self._beta1, self._beta2, self._epsilon, b1p, b2p)
      return
    d, dev, f = var.embedding_dim, grad.device, torch.float32
    c = lambda x: torch.tensor(x, dtype=f, device=dev)
    mv = m_v.sparse_read(indices)                                      # adam.py:100-101
    m, v = mv[:, :d], mv[:, d:]
    m_t = c(self._beta1) * m + grad * (c(1.0) - c(self._beta1))        # :115-117
    v_t = c(self._beta2) * v + (grad * grad) * (c(1.0) - c(self._beta2))
    m_v.scatter_update(IndexedSlices(torch.cat([m_t, v_t], 1), indices))
    lr = c(self._lr) * torch.sqrt(c(1.0) - c(b2p)) / (c(1.0) - c(b1p))  # :150
    var.scatter_sub(IndexedSlices(lr * m_t / (c(self._epsilon) + torch.sqrt(v_t)), indices))

  def _finish(self):
    # TF Adam._finish: beta_power *= beta, in fp32
    self._beta1_power = self._beta1_power * torch.tensor(self._beta1, dtype=torch.float32)
    self._beta2_power = self._beta2_power * torch.tensor(self._beta2, dtype=torch.float32)


class GroupAdamOptimizer(AdamOptimizer):
  """training/group_adam.py, version 4 (the default, :47) -> KvVariableGroupSparseApplyAdamV4
  with the m_v_linear slot of width 3 (:146-153)."""

  def __init__(self, learning_rate=0.001, initial_accumulator_value=0.0, beta1=0.9, beta2=0.999,
               epsilon=1e-8, l1_regularization_strength=0.0, l2_regularization_strength=0.0,
               l21_regularization_strength=0.0, use_locking=False, name="GroupAdam",
               accum_name=None, linear_name=None, version=4):
    super().__init__(learning_rate, beta1, beta2, epsilon, use_locking, name)
    for nm, val in [("initial_accumulator_value", initial_accumulator_value),
                    ("l1_regularization_strength", l1_regularization_strength),
                    ("l2_regularization_strength", l2_regularization_strength),
                    ("l21_regularization_strength", l21_regularization_strength)]:
      if val < 0.0:
        raise ValueError("%s %f needs to be positive or zero" % (nm, val))
    if version != 4:
      raise NotImplementedError("only GroupAdam version 4 is in scope (versions 2/3 are next)")
    self._l1, self._l2, self._l21 = (float(l1_regularization_strength),
                                     float(l2_regularization_strength),
                                     float(l21_regularization_strength))

  def _create_slots(self, var):
    var.num_concat_opt_vars = 3
    self._zeros_slot(var, "m_v_linear", width=3)

  def _resource_apply_sparse(self, grad, var, indices):
    mvl = self.get_slot(var, "m_v_linear")
    ops.kv_variable_group_sparse_apply_adam_v4(
        var.handle, mvl.handle, grad, indices, self._lr, float(self._beta1_power),
        float(self._beta2_power), self._beta1, self._beta2, self._epsilon, self._l1, self._l2,
        self._l21, use_locking=False)


class SparseGroupFtrlOptimizer(_Optimizer):
  """training/sparse_group_ftrl.py:75-96 -> KvVariableSparseGroupSparseApplyFtrlV2; l2 is TF
  Ftrl's adjusted l2 = l2 + beta / (2 * lr) with beta = 0."""

  def __init__(self, learning_rate, learning_rate_power=-0.5, initial_accumulator_value=0.1,
               l1_regularization_strength=0.0, l2_regularization_strength=0.0,
               l21_regularization_strength=0.0, use_locking=False, name="SparseGroupFtrl",
               accum_name=None, linear_name=None, l2_shrinkage_regularization_strength=0.0):
    super().__init__(name, use_locking)
    if initial_accumulator_value < 0.0:
      raise ValueError("initial_accumulator_value %f needs to be positive or zero" %
                       initial_accumulator_value)
    if learning_rate_power > 0.0:
      raise ValueError("learning_rate_power %f needs to be negative or zero" % learning_rate_power)
    for nm, val in [("l1_regularization_strength", l1_regularization_strength),
                    ("l2_regularization_strength", l2_regularization_strength),
                    ("l21_regularization_strength", l21_regularization_strength),
                    ("l2_shrinkage_regularization_strength", l2_shrinkage_regularization_strength)]:
      if val < 0.0:
        raise ValueError("%s %f needs to be positive or zero" % (nm, val))
    self._lr, self._lr_power = float(learning_rate), float(learning_rate_power)
    self._init_acc = float(initial_accumulator_value)
    self._l1, self._l2, self._l21 = (float(l1_regularization_strength),
                                     float(l2_regularization_strength),
                                     float(l21_regularization_strength))
    self._l2_shrinkage = float(l2_shrinkage_regularization_strength)

  def _create_slots(self, var):
    self._zeros_slot(var, "accum", value=self._init_acc)
    self._zeros_slot(var, "linear", value=0.0)

  def _resource_apply_sparse(self, grad, var, indices):
    ops.kv_variable_sparse_group_sparse_apply_ftrl_v2(
        var.handle, self.get_slot(var, "accum").handle, self.get_slot(var, "linear").handle, grad,
        indices, self._lr, self._l1, self._l2, self._l21, self._l2_shrinkage, self._lr_power,
        use_locking=True)
