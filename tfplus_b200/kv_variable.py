"""KvVariable / get_kv_variable for torch — mirror of
tfplus/kv_variable/python/ops/kv_variable_ops.py (class KvVariable, :539) and
variable_scope.py (get_kv_variable, :745-777).

The reference wraps a TF resource handle; here the handle is a device table (ops.KvHandle).
Method names and meanings follow the reference: sparse_read[_with_counts], scatter_*, export,
delete, delete_with_timestamp, total_count, total_freq, is_initialized.
"""
import threading
import zlib

import torch

from . import ops

# kv_variable_ops.py:95 — module global that switches gather between insert / zeros
IS_TRAINING = True
_FAKE_DIM0 = 10000  # variable_scope.py:229-231: the init table has 10000 rows


def set_training(value):
  global IS_TRAINING
  IS_TRAINING = bool(value)


class IndexedSlices:
  """tf.IndexedSlices stand-in: the gradient of a gather (kv_variable_ops.py:1829-1856)."""

  def __init__(self, values, indices):
    self.values, self.indices = values, indices


class KvVariable:
  """A dynamic embedding table: int64 key -> float32[embedding_dim] row in HBM."""

  def __init__(self, name, embedding_dim, initializer=None, key_dtype=torch.int64,
               value_dtype=torch.float32, enter_threshold=0, trainable=True, device=None,
               capacity_hint=0, seed=None, init_rows=_FAKE_DIM0, kv_options=None):
    self.name = name
    self.embedding_dim = int(embedding_dim)
    self._trainable = trainable
    self._enter_threshold = int(enter_threshold)
    self.num_concat_opt_vars = 1        # slot_creator widening, variable_scope.py:1039-1041
    self.kv_options = kv_options
    self.device = torch.device(device if device is not None else "cuda")
    if seed is None:
      seed = zlib.crc32(name.encode()) % (2 ** 31) + 1   # stable across processes and ranks
    self._handle = ops.kv_variable(key_dtype=key_dtype, value_dtype=value_dtype,
                                   value_shape=[self.embedding_dim],
                                   enter_threshold=enter_threshold, device=self.device,
                                   capacity_hint=capacity_hint, seed=seed, name=name)
    self.device = self._handle.device
    # kv_variable_ops.py:842-848: the initializer op feeds InitKvVariableV2 with [10000, D]
    init_val = _run_initializer(initializer, (int(init_rows), self.embedding_dim), self.device)
    self._initial_value = init_val
    ops.init_kv_variable_v2(self._handle, init_val)

  # -- properties ---------------------------------------------------------------
  @property
  def handle(self):
    return self._handle

  @property
  def enter_threshold(self):
    return self._enter_threshold

  @property
  def dtype(self):
    return torch.float32

  @property
  def key_dtype(self):
    return torch.int64

  @property
  def shape(self):
    return ops.kv_variable_shape_v2(self._handle)

  def total_count(self):
    """KvVariableSizeV2 (kv_variable_ops.py:992-997)."""
    return ops.kv_variable_size_v2(self._handle)

  def total_freq(self):
    return ops.kv_variable_frequency(self._handle)

  def is_initialized(self):
    return ops.kv_variable_is_initialized_v2(self._handle)

  # -- reads ----------------------------------------------------------------------
  def sparse_read(self, indices, name=None):
    """kv_variable_ops.py:1057-1080."""
    if IS_TRAINING:
      return ops.kv_variable_gather_or_insert_v2(self._handle, indices)
    return ops.kv_variable_gather_or_zeros_v2(self._handle, indices)

  def sparse_read_with_counts(self, indices, counts=None, name=None):
    """kv_variable_ops.py:1082-1113."""
    if IS_TRAINING:
      if counts is not None:
        return ops.kv_variable_gather_or_insert_with_counts(self._handle, indices, counts)
      return ops.kv_variable_gather_or_insert_v2(self._handle, indices)
    return ops.kv_variable_gather_or_zeros_v2(self._handle, indices)

  def read_value(self):
    """ReadKvVariableOpV2 -> (keys, values)."""
    return ops.read_kv_variable_op_v2(self._handle)

  value = read_value

  def get_counting(self, indices, name=None):
    return ops.kv_variable_get_count_v2(self._handle, indices)

  def increase_counting(self, indices, counts, name=None):
    return ops.kv_variable_increase_count_v2(self._handle, indices, counts)

  def get_timestamp(self, indices, name=None):
    return ops.kv_variable_get_time_stamp(self._handle, indices)

  # -- writes ---------------------------------------------------------------------
  def _scatter(self, fn, sparse_delta):
    if not isinstance(sparse_delta, IndexedSlices):
      raise TypeError("sparse_delta is not IndexedSlices: %s" % (sparse_delta,))
    fn(self._handle, sparse_delta.indices, sparse_delta.values)
    return self

  def scatter_sub(self, sparse_delta, use_locking=False, name=None):
    return self._scatter(ops.kv_variable_scatter_sub_v2, sparse_delta)

  def scatter_add(self, sparse_delta, use_locking=False, name=None):
    return self._scatter(ops.kv_variable_scatter_add_v2, sparse_delta)

  def scatter_mul(self, sparse_delta, use_locking=False, name=None):
    return self._scatter(ops.kv_variable_scatter_mul_v2, sparse_delta)

  def scatter_div(self, sparse_delta, use_locking=False, name=None):
    return self._scatter(ops.kv_variable_scatter_div_v2, sparse_delta)

  def scatter_update(self, sparse_delta, use_locking=False, name=None):
    return self._scatter(ops.kv_variable_scatter_update_v2, sparse_delta)

  def scatter_max(self, sparse_delta, use_locking=False, name=None):
    return self._scatter(ops.kv_variable_scatter_max_v2, sparse_delta)

  def scatter_min(self, sparse_delta, use_locking=False, name=None):
    return self._scatter(ops.kv_variable_scatter_min_v2, sparse_delta)

  def insert(self, indices, values):
    ops.kv_variable_insert_v2(self._handle, indices, values)

  # -- checkpoint / eviction ---------------------------------------------------------
  def export(self, name=None, first_n=None):
    """kv_variable_ops.py:1433-1498: first_n = 6 when training / 3 when predicting,
    enable_cutoff=True, cutoff_value=1e-20."""
    if first_n is None:
      first_n = 6 if IS_TRAINING else 3
    return ops.kv_variable_export(self._handle, first_n=first_n, enable_cutoff=True,
                                  cutoff_value=1e-20)

  def restore(self, tensors, first_n=6):
    """KvVariableSaveable.restore (kv_variable_ops.py:1586-1643): import, then re-run the
    init op (a no-op when the init table is already set)."""
    keys, values, init_table, blacklist, freq_keys, freq_values = tensors
    ops.kv_variable_import(self._handle, keys, values, init_table, blacklist, freq_keys,
                           freq_values, first_n=first_n)
    ops.init_kv_variable_v2(self._handle, self._initial_value)

  def delete(self, indices, name=None):
    ops.kv_variable_delete(self._handle, indices)

  def delete_with_timestamp(self, threshold, name=None):
    return ops.kv_variable_delete_with_timestamp(self._handle, threshold)

  def destroy(self):
    ops.destroy_kv_variable_op_v2(self._handle)


def _run_initializer(initializer, shape, device):
  """TF initializers are callables of a shape; accept those, tensors, scalars and None."""
  if initializer is None:
    # variable_scope.py:530-538 falls back to glorot-style uniform
    limit = (6.0 / (shape[0] + shape[1])) ** 0.5
    g = torch.Generator(device="cpu").manual_seed(0)
    return ((torch.rand(shape, generator=g) * 2 - 1) * limit).to(device)
  if isinstance(initializer, torch.Tensor):
    t = initializer.to(device=device, dtype=torch.float32)
    return t if t.dim() == 2 else t.reshape(-1, shape[1])
  if isinstance(initializer, (int, float)):
    return torch.full(shape, float(initializer), dtype=torch.float32, device=device)
  out = initializer(shape)
  return torch.as_tensor(out, dtype=torch.float32).to(device)


def ones_initializer(shape):
  return torch.ones(shape)


def zeros_initializer(shape):
  return torch.zeros(shape)


def constant_initializer(value):
  return lambda shape: torch.full(shape, float(value))


def random_normal_initializer(mean=0.0, stddev=0.05, seed=0):
  def init(shape):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(shape, generator=g) * stddev + mean
  return init


class PartitionedKvVariable:
  """variable_scope.py:292-296 + embedding_ops.py:121-204: a table split into `num_shards`
  KvVariables by `id % num_shards` (tf.fixed_size_partitioner) inside one process."""

  def __init__(self, name, embedding_dim, num_shards, **kw):
    self.name, self.embedding_dim = name, int(embedding_dim)
    self.parts = [KvVariable("%s/part_%d" % (name, i), embedding_dim, **kw)
                  for i in range(num_shards)]

  def __iter__(self):
    return iter(self.parts)

  def __len__(self):
    return len(self.parts)


# variable store: variable_scope.py:_KvVariableStore (get-or-create by name, reuse semantics)
_STORE = {}
_STORE_LOCK = threading.Lock()


def fixed_size_partitioner(num_shards):
  return int(num_shards)


def get_kv_variable(name, embedding_dim=None, key_dtype=torch.int64, value_dtype=torch.float32,
                    initializer=None, regularizer=None, trainable=None, collections=None,
                    partitioner=None, constraint=None, enter_threshold=0, kv_options=None,
                    device=None, **kw):
  """variable_scope.py:745-777 — same signature; returns the existing variable when `name`
  was created before (tf.get_variable reuse)."""
  if embedding_dim is None:
    raise ValueError("embedding_dim must be given for KvVariable %s" % name)
  with _STORE_LOCK:
    if name in _STORE:
      return _STORE[name]
    common = dict(initializer=initializer, key_dtype=key_dtype, value_dtype=value_dtype,
                  enter_threshold=enter_threshold, trainable=True if trainable is None else trainable,
                  device=device, kv_options=kv_options, **kw)
    if partitioner and int(partitioner) > 1:
      v = PartitionedKvVariable(name, embedding_dim, int(partitioner), **common)
    else:
      v = KvVariable(name, embedding_dim, **common)
    _STORE[name] = v
    return v


def reset_kv_variable_store():
  with _STORE_LOCK:
    for v in _STORE.values():
      for p in (v.parts if isinstance(v, PartitionedKvVariable) else [v]):
        try:
          p.destroy()
        except Exception:
          pass
    _STORE.clear()
