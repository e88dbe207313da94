"""Op-level mirror of TFPlus's `gen_kv_variable_ops` for torch CUDA tensors.

Same op names (snake_case of the REGISTER_OP names in
tfplus/kv_variable/ops/{kv_variable_ops,training_ops}.cc), same argument order
and meaning, same error behaviour; each function is a thin call into the C ABI
(include/kvhbm.h) on torch's current CUDA stream — exactly what the TF
DEVICE_GPU OpKernels in tf_ops/ do with `ctx->eigen_gpu_device().stream()`.
The resource handle is a `KvHandle` (TF: a DT_RESOURCE tensor in host memory).
"""
import threading
import time

import torch

from . import _lib
from ._lib import check

_TODAY_OVERRIDE = None


def set_today(day):
  """Inject the clock (days since the epoch, utility.cc:38-40); None = real time."""
  global _TODAY_OVERRIDE
  _TODAY_OVERRIDE = day


def today():
  if _TODAY_OVERRIDE is not None:
    return int(_TODAY_OVERRIDE) & 0xFFFF
  return int(time.time() // 86400) & 0xFFFF


def _stream(device):
  return torch.cuda.current_stream(device).cuda_stream


def _ptr(t):
  return None if t is None else t.data_ptr()


class KvHandle:
  """The KvVariable resource: owns one device table (kv_table*)."""

  def __init__(self, dim, enter_threshold, device, name, capacity_hint=0, seed=0):
    import ctypes as C
    self.name = name
    self.dim = int(dim)
    self.enter_threshold = int(enter_threshold)
    self.device = torch.device(device)
    if self.device.type != "cuda":
      raise RuntimeError("KvVariable lives in HBM: device must be a CUDA device "
                         "(there is no CPU fallback)")
    if self.device.index is None:
      self.device = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    out = C.c_void_p()
    with torch.cuda.device(self.device):
      check(lib.kv_create(self.dim, self.enter_threshold, int(capacity_hint), C.byref(out)))
    self.ptr = out.value
    self._lib = lib
    if seed:
      check(lib.kv_set_seed(self.ptr, int(seed)))

  @property
  def stream(self):
    return _stream(self.device)

  def destroy(self):
    if getattr(self, "ptr", None):
      self._lib.kv_destroy(self.ptr)
      self.ptr = None

  def __del__(self):
    try:
      self.destroy()
    except Exception:  # interpreter shutdown
      pass

  def _live(self):
    if not self.ptr:
      raise RuntimeError("NotFound: KvVariable %s has been destroyed" % self.name)
    return self.ptr


class Workspace:
  """Scratch memory of the dedup / routing kernels (one per device and stream user)."""
  _local = threading.local()

  def __init__(self, device):
    import ctypes as C
    self.device = torch.device(device)
    out = C.c_void_p()
    with torch.cuda.device(self.device):
      check(_lib.load().kv_workspace_create(C.byref(out)))
    self.ptr = out.value

  def __del__(self):
    try:
      if self.ptr:
        _lib.load().kv_workspace_destroy(self.ptr)
        self.ptr = None
    except Exception:
      pass

  @classmethod
  def get(cls, device):
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    cache = getattr(cls._local, "cache", None)
    if cache is None:
      cache = cls._local.cache = {}
    if idx not in cache:
      cache[idx] = cls(torch.device("cuda", idx))
    return cache[idx]


# ResourceMgr stand-in: (container, shared_name) -> handle, kv_variable_ops.cc:97-105
_RESOURCES = {}
_RES_LOCK = threading.Lock()
_ANON = [0]


def _ids(t, handle):
  if not isinstance(t, torch.Tensor):
    t = torch.as_tensor(t, dtype=torch.int64)
  if t.dtype != torch.int64:
    if t.dtype in (torch.int32,):
      raise NotImplementedError("Unimplemented: only int64 keys are supported on the device "
                                "table (the reference also registers int32/uint64)")
    raise TypeError("indices must be int64")
  if t.device != handle.device:
    t = t.to(handle.device, non_blocking=True)
  return t.contiguous()


def _vals(t, handle, what="values"):
  if not isinstance(t, torch.Tensor):
    t = torch.as_tensor(t, dtype=torch.float32)
  if t.dtype != torch.float32:
    raise NotImplementedError("Unimplemented: only float32 %s are supported" % what)
  if t.device != handle.device:
    t = t.to(handle.device, non_blocking=True)
  return t.contiguous()


# ---------------------------------------------------------------------------
# lifecycle
# ---------------------------------------------------------------------------
def kv_variable(container="", shared_name="", use_node_name_sharing=False,
                key_dtype=torch.int64, value_dtype=torch.float32, key_shape=(),
                value_shape=None, enter_threshold=0, name=None, device=None,
                capacity_hint=0, seed=0):
  """Op `KvVariable` (ops/kv_variable_ops.cc:37-74): creates or looks up the resource."""
  if key_dtype != torch.int64 or value_dtype != torch.float32:
    raise NotImplementedError("Unimplemented: the device table is int64 -> float32")
  if value_shape is None:
    raise ValueError("value_shape is required")
  dim = 1
  for d in value_shape:
    dim *= int(d)
  device = torch.device(device if device is not None else "cuda")
  key = None
  if shared_name or (use_node_name_sharing and name):
    key = (container, shared_name or name, str(device))
  with _RES_LOCK:
    if key is not None and key in _RESOURCES and _RESOURCES[key].ptr:
      return _RESOURCES[key]
    if name is None:
      _ANON[0] += 1
      name = shared_name or "KvVariable_%d" % _ANON[0]
    h = KvHandle(dim, enter_threshold, device, name, capacity_hint, seed)
    h.value_shape = list(value_shape)
    if key is not None:
      _RESOURCES[key] = h
    return h


def init_kv_variable_v2(table_handle, init_table):
  """Op `InitKvVariableV2` (ops/kv_variable_ops.cc:212-222)."""
  h = table_handle
  tbl = _vals(init_table, h, "init tables")
  if tbl.dim() < 2 or tbl.numel() // tbl.shape[0] != h.dim:
    raise ValueError("InvalidArgument: init table must be [rows, %d]" % h.dim)
  check(h._lib.kv_set_init_table(h._live(), tbl.data_ptr(), tbl.shape[0], h.stream))


def kv_variable_is_initialized_v2(table_handle):
  import ctypes as C
  h = table_handle
  if not h.ptr:
    return False  # kv_variable_ops.cc:226-229: lookup failure -> false
  out = C.c_int()
  check(h._lib.kv_is_initialized(h.ptr, C.byref(out)))
  return bool(out.value)


def _scalar(fn, h):
  import ctypes as C
  out = C.c_int64()
  with torch.cuda.device(h.device):
    check(fn(h._live(), h.stream, C.byref(out)))
  return out.value


def kv_variable_shape_v2(table_handle):
  h = table_handle
  return [_scalar(h._lib.kv_map_size, h)] + list(getattr(h, "value_shape", [h.dim]))


def kv_variable_size_v2(table_handle):
  return _scalar(table_handle._lib.kv_size, table_handle)


def kv_variable_frequency(table_handle):
  return _scalar(table_handle._lib.kv_sum_freq, table_handle)


def destroy_kv_variable_op_v2(table_handle, ignore_lookup_error=True):
  h = table_handle
  h._live()
  with _RES_LOCK:
    for k, v in list(_RESOURCES.items()):
      if v is h:
        del _RESOURCES[k]
  h.destroy()


def kv_variable_reserve(table_handle, n_keys):
  h = table_handle
  check(h._lib.kv_reserve(h._live(), int(n_keys), h.stream))


# ---------------------------------------------------------------------------
# lookups
# ---------------------------------------------------------------------------
def kv_variable_gather_or_zeros_v2(table_handle, indices):
  """Op `KvVariableGatherOrZerosV2`: output shape = indices.shape + value_shape."""
  h = table_handle
  ids = _ids(indices, h)
  out = torch.empty(tuple(ids.shape) + (h.dim,), dtype=torch.float32, device=h.device)
  if ids.numel():
    check(h._lib.kv_gather_or_zeros(h._live(), ids.data_ptr(), ids.numel(), out.data_ptr(),
                                    h.stream))
  return out


def kv_variable_gather_or_insert_v2(table_handle, indices, out=None):
  """Op `KvVariableGatherOrInsertV2` (ops/kv_variable_ops.cc:310-320)."""
  return kv_variable_gather_or_insert_with_counts(table_handle, indices, None, out=out)


def kv_variable_gather_or_insert_with_counts(table_handle, indices, counts, out=None,
                                             num_indices=None):
  """Op `KvVariableGatherOrInsertWithCounts` (ops/kv_variable_ops.cc:322-332).  num_indices: an
  int32 device scalar; only the first min(len, num_indices) ids are looked up (rows past it are
  left as they are) — the unique -> gather chain without a host read of the unique count."""
  h = table_handle
  ids = _ids(indices, h)
  if counts is not None:
    if not isinstance(counts, torch.Tensor):
      counts = torch.as_tensor(counts, dtype=torch.int32)
    if counts.dtype != torch.int32:
      raise ValueError("InvalidArgument: KvVariable %s: increment count, counts dtype must "
                       "be int32" % h.name)
    if tuple(counts.shape) != tuple(ids.shape):
      raise ValueError("InvalidArgument: KvVariable %s: increment count, indices shape %s does "
                       "not match with counts shape %s" %
                       (h.name, list(ids.shape), list(counts.shape)))
    counts = counts.to(h.device).contiguous()
  if out is None:
    out = torch.empty(tuple(ids.shape) + (h.dim,), dtype=torch.float32, device=h.device)
  if ids.numel() and num_indices is not None:
    check(h._lib.kv_gather_or_insert_n(h._live(), ids.data_ptr(), _ptr(counts), ids.numel(),
                                       num_indices.data_ptr(), out.data_ptr(), today(), h.stream))
  elif ids.numel():
    check(h._lib.kv_gather_or_insert(h._live(), ids.data_ptr(), _ptr(counts), ids.numel(),
                                     out.data_ptr(), today(), h.stream))
  return out


def batch_kv_variable_gather_or_zeros_v2(table_handles, indices_list):
  """Op `BatchKvVariableGatherOrZerosV2` (kernels/kv_variable_ops.cc:431-496): one read-only
  lookup per (table, indices) pair."""
  return [kv_variable_gather_or_zeros_v2(h, ids) for h, ids in zip(table_handles, indices_list)]


def kv_variable_gather_v2(table_handle, indices, use_init_value=True):
  """Legacy op `KvVariableGatherV2` (kernels/kv_variable_ops.cc:633-701): bool attr instead of two
  ops - use_init_value inserts missing keys from the initializer, otherwise zeros."""
  if use_init_value:
    return kv_variable_gather_or_insert_v2(table_handle, indices)
  return kv_variable_gather_or_zeros_v2(table_handle, indices)


def kv_variable_insert_v2(table_handle, indices, values, filter_out=None, blacklist=None):
  """Op `KvVariableInsertV2` -> KvVariable::InsertOrUpdate."""
  h = table_handle
  ids = _ids(indices, h)
  vals = _vals(values, h)
  if ids.numel() == 0:
    return
  if vals.numel() != ids.numel() * h.dim:
    raise ValueError("InvalidArgument: values must be [%d, %d]" % (ids.numel(), h.dim))
  f = None if filter_out is None else filter_out.to(h.device).to(torch.uint8).contiguous()
  b = None if blacklist is None else blacklist.to(h.device).to(torch.uint8).contiguous()
  check(h._lib.kv_insert_or_update(h._live(), ids.data_ptr(), vals.data_ptr(), ids.numel(),
                                   _ptr(f), _ptr(b), h.stream))


def kv_variable_increase_count_v2(table_handle, indices, counts):
  """Op `KvVariableIncreaseCountV2`: reserved op, empty body (kv_variable_ops.cc:754-756)."""
  return None


_SCATTER = {"update": 0, "add": 1, "sub": 2, "mul": 3, "div": 4, "min": 5, "max": 6}


def _scatter(op, table_handle, indices, updates, unique_indices=False):
  """indices may repeat (every occurrence is applied, in index order) unless the caller
  vouches for distinct ids with unique_indices=True (skips the internal dedup pass)."""
  h = table_handle
  ids = _ids(indices, h)
  upd = _vals(updates, h, "updates")
  if ids.numel() == 0:
    return
  if upd.numel() != ids.numel() * h.dim:
    raise ValueError("InvalidArgument: updates must be [%d, %d]" % (ids.numel(), h.dim))
  fn = h._lib.kv_scatter_unique if unique_indices else h._lib.kv_scatter
  check(fn(h._live(), _SCATTER[op], ids.data_ptr(), upd.data_ptr(), ids.numel(), h.stream))


def kv_variable_scatter_add_v2(table_handle, indices, updates, unique_indices=False):
  _scatter("add", table_handle, indices, updates, unique_indices)


def kv_variable_scatter_sub_v2(table_handle, indices, updates, unique_indices=False):
  _scatter("sub", table_handle, indices, updates, unique_indices)


def kv_variable_scatter_mul_v2(table_handle, indices, updates, unique_indices=False):
  _scatter("mul", table_handle, indices, updates, unique_indices)


def kv_variable_scatter_div_v2(table_handle, indices, updates, unique_indices=False):
  _scatter("div", table_handle, indices, updates, unique_indices)


def kv_variable_scatter_min_v2(table_handle, indices, updates, unique_indices=False):
  _scatter("min", table_handle, indices, updates, unique_indices)


def kv_variable_scatter_max_v2(table_handle, indices, updates, unique_indices=False):
  _scatter("max", table_handle, indices, updates, unique_indices)


def kv_variable_scatter_update_v2(table_handle, indices, updates, unique_indices=False):
  _scatter("update", table_handle, indices, updates, unique_indices)


def kv_variable_get_count_v2(table_handle, indices):
  h = table_handle
  ids = _ids(indices, h)
  out = torch.empty(ids.shape, dtype=torch.int32, device=h.device)
  if ids.numel():
    check(h._lib.kv_get_count(h._live(), ids.data_ptr(), ids.numel(), out.data_ptr(), h.stream))
  return out


def kv_variable_get_time_stamp(table_handle, indices):
  h = table_handle
  ids = _ids(indices, h)
  out = torch.empty(ids.shape, dtype=torch.int32, device=h.device)  # uint32 payload
  if ids.numel():
    check(h._lib.kv_get_timestamp(h._live(), ids.data_ptr(), ids.numel(), out.data_ptr(),
                                  today(), h.stream))
  return out


# ---------------------------------------------------------------------------
# checkpoint
# ---------------------------------------------------------------------------
def kv_variable_export(table_handle, first_n=3, enable_cutoff=False, cutoff_value=0.0,
                       freq_dtype=torch.uint16):
  """Op `KvVariableExport` (ops/kv_variable_ops.cc:421-462): returns the six tensors
  (keys, values, init_table, blacklist, freq_keys, freq_values); attr defaults are the op's."""
  import ctypes as C
  h = table_handle
  nk, nb, nf = C.c_int64(), C.c_int64(), C.c_int64()
  check(h._lib.kv_export_count(h._live(), first_n, int(bool(enable_cutoff)), float(cutoff_value),
                               h.stream, C.byref(nk), C.byref(nb), C.byref(nf)))
  dev = h.device
  keys = torch.empty(nk.value, dtype=torch.int64, device=dev)
  values = torch.empty((nk.value, h.dim), dtype=torch.float32, device=dev)
  blacklist = torch.empty(nb.value, dtype=torch.int64, device=dev)
  freq_keys = torch.empty(nf.value, dtype=torch.int64, device=dev)
  u32 = freq_dtype in (torch.uint32, torch.int32)
  freq_values = torch.empty(nf.value, dtype=torch.int32 if u32 else torch.uint16, device=dev)
  check(h._lib.kv_export_bounded(h.ptr, first_n, _ptr(keys), _ptr(values), nk.value,
                                 _ptr(blacklist), nb.value, _ptr(freq_keys), _ptr(freq_values),
                                 nf.value, int(u32), h.stream))
  if first_n > 3:
    rows = C.c_int64()
    check(h._lib.kv_init_table_rows(h.ptr, C.byref(rows)))
    init_table = torch.empty((rows.value, h.dim), dtype=torch.float32, device=dev)
    check(h._lib.kv_get_init_table(h.ptr, init_table.data_ptr(), h.stream))
  else:
    init_table = torch.empty((0, h.dim), dtype=torch.float32, device=dev)
  return keys, values, init_table, blacklist, freq_keys, freq_values


_COMBINERS = {"sum": 0, "mean": 1, "sqrtn": 2}


def sparse_combine(emb, idx, segment_ids, weights, n_rows, combiner):
  """The combiner half of embedding_lookup_sparse (embedding_ops.py:403-441) as one kernel:
  emb[U, D] rows of the distinct ids, idx[nnz] inverse index, segment_ids[nnz] sorted row ids."""
  if combiner not in _COMBINERS:
    raise ValueError("combiner must be one of 'mean', 'sqrtn' or 'sum'")
  emb = emb.contiguous()
  idx = idx.to(torch.int32).contiguous()
  seg = segment_ids.to(torch.int64).contiguous()
  w = None if weights is None else weights.to(torch.float32).contiguous()
  out = torch.empty((int(n_rows), emb.shape[1]), dtype=torch.float32, device=emb.device)
  with torch.cuda.device(emb.device):
    check(_lib.load().kv_sparse_combine(emb.data_ptr(), idx.data_ptr(), seg.data_ptr(), _ptr(w),
                                        idx.numel(), int(n_rows), emb.shape[1],
                                        _COMBINERS[combiner], out.data_ptr(), _stream(emb.device)))
  return out


def kv_variable_check_overflow(table_handle):
  """Raises if CUDA-graph replays inserted more keys than kv_variable_reserve made room for."""
  h = table_handle
  check(h._lib.kv_check_overflow(h._live(), h.stream))


def kv_variable_enable_delta_export(table_handle, support_prediction_delta=False):
  """SUPPORT_DELTA_EXPORT / SUPPORT_PREDICTION_DELTA_EXPORT (kernels/kv_variable.h:101-111): from
  now on the table records which keys change."""
  h = table_handle
  check(h._lib.kv_enable_delta_export(h._live(), int(bool(support_prediction_delta))))


def kv_variable_delta_size(table_handle):
  import ctypes as C
  h = table_handle
  out = C.c_int64()
  check(h._lib.kv_delta_size(h._live(), h.stream, C.byref(out)))
  return out.value


def kv_variable_full_or_delta_export(table_handle, first_n=6, do_full_export=False,
                                     enable_cutoff=False, cutoff_value=0.0):
  """Op `KvVariableFullOrDeltaExport` (ops/kv_variable_ops.cc:633-660): the 8-tensor format
  (keys, values, init_table, blacklist, freq_keys, freq_values uint32, need_full_import,
  delete_keys).  do_full_export routes to the full export (dynamic_save.hpp FullExport),
  otherwise KvVariable::DeltaExport (:197-449)."""
  import ctypes as C
  h = table_handle
  dev = h.device
  if do_full_export:
    k, v, it, bl, fk, fv = kv_variable_export(h, first_n=first_n, enable_cutoff=enable_cutoff,
                                              cutoff_value=cutoff_value, freq_dtype=torch.int32)
    return (k, v, it, bl, fk, fv, torch.ones(1, dtype=torch.bool),
            torch.empty(0, dtype=torch.int64, device=dev))
  nk, nb, nf, nd = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
  check(h._lib.kv_delta_export_count(h._live(), first_n, h.stream, C.byref(nk), C.byref(nb),
                                     C.byref(nf), C.byref(nd)))
  keys = torch.empty(nk.value, dtype=torch.int64, device=dev)
  values = torch.empty((nk.value, h.dim), dtype=torch.float32, device=dev)
  blacklist = torch.empty(nb.value, dtype=torch.int64, device=dev)
  freq_keys = torch.empty(nf.value, dtype=torch.int64, device=dev)
  freq_values = torch.empty(nf.value, dtype=torch.int32, device=dev)   # uint32 payload
  delete_keys = torch.empty(nd.value, dtype=torch.int64, device=dev)
  counts = (C.c_int64 * 4)()
  check(h._lib.kv_delta_export(h.ptr, first_n, _ptr(keys), _ptr(values), nk.value,
                               _ptr(blacklist), nb.value, _ptr(freq_keys), _ptr(freq_values),
                               nf.value, _ptr(delete_keys), nd.value, h.stream, counts))
  init_table = torch.empty((0, h.dim), dtype=torch.float32, device=dev)   # :312-325: empty
  return (keys, values, init_table, blacklist, freq_keys, freq_values,
          torch.zeros(1, dtype=torch.bool), delete_keys)


def kv_variable_full_or_delta_import_v2(table_handle, keys, values, init_table, blacklist,
                                        freq_keys, freq_values, need_full_import, delete_keys,
                                        first_n=6):
  """Op `KvVariableFullOrDeltaImportV2` (ops/kv_variable_ops.cc:604-631): need_full_import
  routes to ImportValues, otherwise KvVariable::DeltaImport (dynamic_restore.hpp:28-153)."""
  h = table_handle
  if bool(torch.as_tensor(need_full_import).reshape(-1)[0]):
    return kv_variable_import(h, keys, values, init_table, blacklist, freq_keys, freq_values,
                              first_n=first_n)
  keys = _ids(keys, h)
  values = _vals(values, h)
  bl = _ids(blacklist, h)
  fk = _ids(freq_keys, h)
  fv = torch.as_tensor(freq_values).to(h.device)
  if fv.dtype not in (torch.int32, torch.uint32):
    fv = fv.to(torch.int32)
  fv = fv.contiguous()
  dk = _ids(delete_keys, h)
  check(h._lib.kv_delta_import(h._live(), first_n, _ptr(keys), _ptr(values), keys.numel(),
                               _ptr(bl), bl.numel(), _ptr(fk), _ptr(fv), fk.numel(), _ptr(dk),
                               dk.numel(), h.stream))


def read_kv_variable_op_v2(table_handle):
  """Op `ReadKvVariableOpV2` = ExportValues(first_n=2) (kv_variable_ops.cc:325-346);
  like the reference it resets every under-threshold flag (enable_cutoff=false)."""
  keys, values = kv_variable_export(table_handle, first_n=2, enable_cutoff=False,
                                    cutoff_value=0.0)[:2]
  return keys, values


def kv_variable_import(table_handle, keys, values, init_table, blacklist, freq_keys,
                       freq_values, first_n=6):
  """Op `KvVariableImport` (ops/kv_variable_ops.cc:361-385; kernel kv_variable_ops.cc:779-851:
  the blacklist is dropped when first_n <= 3, the frequency table when first_n <= 4)."""
  h = table_handle
  keys = _ids(keys, h)
  values = _vals(values, h)
  init_table = None if init_table is None or init_table.numel() == 0 else _vals(init_table, h)
  if first_n <= 3:
    blacklist = None
  if first_n <= 4:
    freq_keys = freq_values = None
  bl = None if blacklist is None or len(blacklist) == 0 else _ids(blacklist, h)
  fk = None if freq_keys is None or len(freq_keys) == 0 else _ids(freq_keys, h)
  fv, u32 = None, 0
  if fk is not None:
    fv = freq_values if isinstance(freq_values, torch.Tensor) else torch.as_tensor(freq_values)
    if fv.dtype in (torch.int32, torch.uint32):
      u32 = 1
    elif fv.dtype == torch.uint16:
      u32 = 0
    else:  # python lists / int64: the op's dtype is uint16
      fv = fv.to(torch.int32).to(torch.uint16)
    fv = fv.to(h.device).contiguous()
  check(h._lib.kv_import(h._live(), _ptr(keys), _ptr(values), keys.numel(), _ptr(init_table),
                         0 if init_table is None else init_table.shape[0], _ptr(bl),
                         0 if bl is None else bl.numel(), _ptr(fk), _ptr(fv),
                         0 if fk is None else fk.numel(), u32, h.stream))


def kv_variable_delete(table_handle, indices):
  h = table_handle
  ids = _ids(indices, h)
  if ids.numel():
    check(h._lib.kv_delete(h._live(), ids.data_ptr(), ids.numel(), h.stream))


def kv_variable_delete_with_timestamp(table_handle, threshold):
  """Op `KvVariableDeleteWithTimestamp`: returns the deleted keys."""
  import ctypes as C
  h = table_handle
  cap = max(1, _scalar(h._lib.kv_map_size, h))
  out = torch.empty(cap, dtype=torch.int64, device=h.device)
  n = C.c_int64()
  check(h._lib.kv_delete_with_timestamp(h._live(), int(threshold), today(), out.data_ptr(), cap,
                                        h.stream, C.byref(n)))
  return out[:n.value]


# ---------------------------------------------------------------------------
# fused sparse optimizer applies (ops/training_ops.cc)
# ---------------------------------------------------------------------------
def _apply_args(var, grad, indices, num_indices):
  ids = _ids(indices, var)
  g = _vals(grad, var, "gradients")
  if ids.dim() != 1:
    raise ValueError("InvalidArgument: indices must be one-dimensional")
  if g.shape[0] != ids.shape[0]:
    raise ValueError("InvalidArgument: grad must be the same size as indices in the first "
                     "dimension.")
  if g.numel() != ids.numel() * var.dim:
    raise ValueError("InvalidArgument: var and grad must match in dimension 1")
  dn = None
  if num_indices is not None:
    dn = num_indices.to(var.device).to(torch.int32).contiguous()
  return ids, g, dn


def kv_variable_sparse_apply_adagrad(var, accum, lr, grad, indices, use_locking=False,
                                     update_slots=True, num_indices=None):
  """Op `KvVariableSparseApplyAdagrad` (ops/training_ops.cc:214-226)."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  check(var._lib.kv_apply_adagrad(var._live(), accum._live(), ids.data_ptr(), g.data_ptr(),
                                  ids.numel(), _ptr(dn), float(lr), int(update_slots), today(),
                                  var.stream))


def kv_variable_sparse_group_sparse_apply_ftrl_v2(var, accum, linear, grad, indices, lr, l1, l2,
                                                  l21, l2_shrinkage, lr_power,
                                                  use_locking=False, num_indices=None):
  """Op `KvVariableSparseGroupSparseApplyFtrlV2` (ops/training_ops.cc:135-150)."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  check(var._lib.kv_apply_sparse_group_ftrl(
      var._live(), accum._live(), linear._live(), ids.data_ptr(), g.data_ptr(), ids.numel(),
      _ptr(dn), float(lr), float(l1), float(l2), float(l21), float(l2_shrinkage),
      float(lr_power), today(), var.stream))


def kv_variable_group_sparse_apply_adam_v4(var, m_v_linear, grad, indices, lr, beta1_power,
                                           beta2_power, beat1, beta2, epsilon, l1, l2, l21,
                                           use_locking=False, num_indices=None):
  """Op `KvVariableGroupSparseApplyAdamV4` (ops/training_ops.cc:1266-1285; `beat1` is the
  reference's own spelling of the input name)."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  check(var._lib.kv_apply_group_adam_v4(
      var._live(), m_v_linear._live(), ids.data_ptr(), g.data_ptr(), ids.numel(), _ptr(dn),
      float(lr), float(beta1_power), float(beta2_power), float(beat1), float(beta2),
      float(epsilon), float(l1), float(l2), float(l21), today(), var.stream))


def kv_variable_sparse_apply_adam(var, m_v, grad, indices, lr, beta1, beta2, epsilon,
                                  beta1_power, beta2_power, num_indices=None):
  """Fused tfplus-Adam (python/training/adam.py:93-163 in one pass); not an op of the
  reference, which issues Gather + ScatterUpdate + ScatterSub."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  check(var._lib.kv_apply_adam(var._live(), m_v._live(), ids.data_ptr(), g.data_ptr(),
                               ids.numel(), _ptr(dn), float(lr), float(beta1), float(beta2),
                               float(epsilon), float(beta1_power), float(beta2_power), today(),
                               var.stream))


def _hp(var, hparams, n):
  if hparams.dtype != torch.float32 or hparams.numel() < n or hparams.device != var.device:
    raise ValueError("InvalidArgument: hparams must be %d float32 scalars on %s" % (n, var.device))
  return hparams.contiguous()


def kv_variable_sparse_apply_adagrad_dev(var, accum, hparams, grad, indices, update_slots=True,
                                         num_indices=None):
  """KvVariableSparseApplyAdagrad with hparams = [lr] in device memory."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  hp = _hp(var, hparams, 1)
  check(var._lib.kv_apply_adagrad_dev(var._live(), accum._live(), ids.data_ptr(), g.data_ptr(),
                                      ids.numel(), _ptr(dn), hp.data_ptr(), int(update_slots),
                                      today(), var.stream))


def kv_variable_group_sparse_apply_adam_v4_dev(var, m_v_linear, grad, indices, hparams,
                                               num_indices=None, advance_powers=False):
  """KvVariableGroupSparseApplyAdamV4 with hparams = [lr, beta1_power, beta2_power, beta1,
  beta2, epsilon, l1, l2, l21] in device memory (graph-capturable).  advance_powers: the same
  launch then multiplies the two powers by their betas in `hparams` (AdamOptimizer._finish)."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  hp = _hp(var, hparams, 9)
  if advance_powers:
    if not hparams.is_contiguous():
      raise ValueError("InvalidArgument: advance_powers needs contiguous hparams")
    fn = var._lib.kv_apply_group_adam_v4_dev_advance
  else:
    fn = var._lib.kv_apply_group_adam_v4_dev
  check(fn(var._live(), m_v_linear._live(), ids.data_ptr(), g.data_ptr(), ids.numel(), _ptr(dn),
           hp.data_ptr(), today(), var.stream))


def kv_variable_sparse_group_sparse_apply_ftrl_v2_dev(var, accum, linear, grad, indices, hparams,
                                                      num_indices=None):
  """KvVariableSparseGroupSparseApplyFtrlV2 with hparams = [lr, l1, l2, l21, l2_shrinkage,
  lr_power] in device memory."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  hp = _hp(var, hparams, 6)
  check(var._lib.kv_apply_sparse_group_ftrl_dev(var._live(), accum._live(), linear._live(),
                                                ids.data_ptr(), g.data_ptr(), ids.numel(),
                                                _ptr(dn), hp.data_ptr(), today(), var.stream))


def kv_variable_sparse_apply_adam_dev(var, m_v, grad, indices, hparams, num_indices=None,
                                      advance_powers=False):
  """Fused tfplus-Adam with hparams = [lr, beta1, beta2, epsilon, beta1_power, beta2_power];
  advance_powers as in kv_variable_group_sparse_apply_adam_v4_dev."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  hp = _hp(var, hparams, 6)
  if advance_powers and not hparams.is_contiguous():
    raise ValueError("InvalidArgument: advance_powers needs contiguous hparams")
  fn = var._lib.kv_apply_adam_dev_advance if advance_powers else var._lib.kv_apply_adam_dev
  check(fn(var._live(), m_v._live(), ids.data_ptr(), g.data_ptr(), ids.numel(), _ptr(dn),
           hp.data_ptr(), today(), var.stream))


# ---------------------------------------------------------------------------
# dedup plan: unique_with_counts + occurrence lists of one batch (kvhbm.h kv_plan_*)
# ---------------------------------------------------------------------------
OPT_ADAGRAD, OPT_GROUP_ADAM_V4, OPT_SPARSE_GROUP_FTRL, OPT_ADAM = 0, 1, 2, 3
OPT_GROUP_ADAM_V3, OPT_SPARSE_FTRL_V2, OPT_GROUP_SPARSE_FTRL_V2 = 4, 5, 6
_N_HP = {0: 1, 1: 9, 2: 6, 3: 6, 4: 9, 5: 6, 6: 6}


class Plan:
  """Device-resident dedup plan of up to `max_ids` ids (buffers allocated once).  build(ids)
  runs tf.unique_with_counts plus the per-id occurrence lists on the current stream; the
  arrays are views of the plan's own memory (valid until the next build)."""

  def __init__(self, max_ids, device):
    import ctypes as C
    self.device = torch.device(device)
    if self.device.index is None:
      self.device = torch.device("cuda", torch.cuda.current_device())
    self.max_ids = int(max_ids)
    out = C.c_void_p()
    with torch.cuda.device(self.device):
      check(_lib.load().kv_plan_create(self.max_ids, C.byref(out)))
    self.ptr = out.value
    self.n = 0

  def __del__(self):
    try:
      if self.ptr:
        _lib.load().kv_plan_destroy(self.ptr)
        self.ptr = None
    except Exception:
      pass

  def build(self, ids, ws=None):
    if ids.device.type != "cuda":
      raise RuntimeError("plan: ids must live on a CUDA device")
    ids = ids.contiguous()
    if ids.dtype != torch.int64:
      raise NotImplementedError("Unimplemented: only int64 keys")
    self.ids = ids   # the kernels of this batch read them later: keep them alive
    self.n = ids.numel()
    ws = ws or Workspace.get(self.device)
    with torch.cuda.device(self.device):
      check(_lib.load().kv_plan_build(self.ptr, ws.ptr, ids.data_ptr(), self.n,
                                      _stream(self.device)))
    return self

  def arrays(self):
    """(uniq i64[n], idx i32[n], counts i32[n], num_unique i32[1], seg_off i32[n], pos i32[n]):
    zero-copy views of the plan's device arrays; entries past num_unique are undefined and the
    contents change with the next build."""
    import ctypes as C
    ptrs = [C.c_void_p() for _ in range(6)]
    check(_lib.load().kv_plan_arrays(self.ptr, *[C.byref(p) for p in ptrs]))
    n = self.n
    specs = [("<i8", n, torch.int64), ("<i4", n, torch.int32), ("<i4", n, torch.int32),
             ("<i4", 1, torch.int32), ("<i4", n, torch.int32), ("<i4", n, torch.int32)]
    out = []
    for p, (ts, cnt, dt) in zip(ptrs, specs):
      if cnt == 0:
        out.append(torch.empty(0, dtype=dt, device=self.device))
      else:
        out.append(torch.as_tensor(_DevArray(p.value, cnt, ts, self), device=self.device))
    return tuple(out)


class _DevArray:
  """__cuda_array_interface__ view of device memory owned by `owner`."""

  def __init__(self, ptr, n, typestr, owner):
    self.owner = owner
    self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                     "version": 2}


def kv_variable_gather_or_insert_plan(table_handle, plan, out=None):
  """KvVariableGatherOrInsertV2 over the batch `plan` was built from: every distinct id is
  looked up once with its occurrence count, rows are expanded to all positions."""
  h = table_handle
  if out is None:
    out = torch.empty((plan.n, h.dim), dtype=torch.float32, device=h.device)
  with torch.cuda.device(h.device):
    check(h._lib.kv_gather_or_insert_plan(h._live(), plan.ptr, out.data_ptr(), today(), h.stream))
  return out


def kv_variable_gather_or_zeros_plan(table_handle, plan, out=None):
  h = table_handle
  if out is None:
    out = torch.empty((plan.n, h.dim), dtype=torch.float32, device=h.device)
  with torch.cuda.device(h.device):
    check(h._lib.kv_gather_or_zeros_plan(h._live(), plan.ptr, out.data_ptr(), h.stream))
  return out


def segment_sum_plan(plan, data, out=None):
  """tf.math.unsorted_segment_sum(data, plan.idx, num_unique): rows [0, num_unique) of `out`
  ([n, dim]) are written, every segment summed in increasing position (TF's CPU order)."""
  data = data.contiguous()
  dim = data.shape[1]
  if out is None:
    out = torch.empty((plan.n, dim), dtype=torch.float32, device=data.device)
  with torch.cuda.device(data.device):
    check(_lib.load().kv_segment_sum_plan(plan.ptr, data.data_ptr(), dim, out.data_ptr(),
                                          _stream(data.device)))
  return out


def kv_variable_apply_plan(kind, var, slot_a, slot_b, plan, grad, hparams, update_slots=True,
                           advance_powers=False):
  """UnsortedSegmentSum + the KvVariable*Apply* op `kind` (OPT_*) in one launch.  grad[n, dim]
  has one row per id occurrence of the plan's batch.  hparams: a python sequence of the op's
  scalar inputs in op order (host path, validated as the reference validates them) or a float32
  device tensor of them (graph-capturable; advance_powers then also does Adam's _finish)."""
  import ctypes as C
  g = grad.contiguous()
  if g.dtype != torch.float32:
    raise NotImplementedError("Unimplemented: only float32 gradients")
  if g.shape[0] != plan.n:
    raise ValueError("InvalidArgument: grad must be the same size as indices in the first "
                     "dimension.")
  if g.numel() != plan.n * var.dim:
    raise ValueError("InvalidArgument: var and grad must match in dimension 1")
  lib = var._lib
  b = slot_b._live() if slot_b is not None else None
  with torch.cuda.device(var.device):
    if isinstance(hparams, torch.Tensor):
      hp = _hp(var, hparams, _N_HP[kind])
      check(lib.kv_apply_plan_dev(kind, var._live(), slot_a._live(), b, plan.ptr, g.data_ptr(),
                                  hp.data_ptr(), int(bool(advance_powers)),
                                  int(bool(update_slots)), today(), var.stream))
    else:
      arr = (C.c_float * len(hparams))(*[float(x) for x in hparams])
      check(lib.kv_apply_plan(kind, var._live(), slot_a._live(), b, plan.ptr, g.data_ptr(), arr,
                              len(hparams), int(bool(update_slots)), today(), var.stream))


def kv_variable_group_sparse_apply_adam_v3(var, m_v_linear, grad, indices, lr, beta1_power,
                                           beta2_power, beta1, beta2, epsilon, l1, l2, l21,
                                           use_locking=False, num_indices=None):
  """ops/training_ops.cc:1086-1105."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  check(var._lib.kv_apply_group_adam_v3(var._live(), m_v_linear._live(), ids.data_ptr(),
                                        g.data_ptr(), ids.numel(), _ptr(dn), lr, beta1_power,
                                        beta2_power, beta1, beta2, epsilon, l1, l2, l21, today(),
                                        var.stream))


def kv_variable_sparse_apply_ftrl_v2(var, accum, linear, grad, indices, lr, l1, l2, l2_shrinkage,
                                     lr_power, use_locking=False, num_indices=None):
  """ops/training_ops.cc:103-117."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  check(var._lib.kv_apply_sparse_ftrl_v2(var._live(), accum._live(), linear._live(),
                                         ids.data_ptr(), g.data_ptr(), ids.numel(), _ptr(dn), lr,
                                         l1, l2, l2_shrinkage, lr_power, today(), var.stream))


def kv_variable_group_sparse_apply_ftrl_v2(var, accum, linear, grad, indices, lr, l1, l2,
                                           l2_shrinkage, lr_power, use_locking=False,
                                           num_indices=None):
  """ops/training_ops.cc:119-133."""
  ids, g, dn = _apply_args(var, grad, indices, num_indices)
  check(var._lib.kv_apply_group_sparse_ftrl_v2(var._live(), accum._live(), linear._live(),
                                               ids.data_ptr(), g.data_ptr(), ids.numel(),
                                               _ptr(dn), lr, l1, l2, l2_shrinkage, lr_power,
                                               today(), var.stream))


# ---------------------------------------------------------------------------
# stock TF ops on the path
# ---------------------------------------------------------------------------
def unique(x, with_counts=False, sync=True):
  """tf.unique / tf.unique_with_counts.  With sync=False the outputs keep the input length
  and the number of unique ids stays on the device (last return value)."""
  if x.device.type != "cuda":
    raise RuntimeError("unique: ids must live on a CUDA device")
  ids = x.contiguous()
  n = ids.numel()
  dev = ids.device
  uniq = torch.empty(n, dtype=torch.int64, device=dev)
  idx = torch.empty(n, dtype=torch.int32, device=dev)
  counts = torch.empty(n, dtype=torch.int32, device=dev) if with_counts else None
  num = torch.empty(1, dtype=torch.int32, device=dev)
  ws = Workspace.get(dev)
  with torch.cuda.device(dev):
    check(_lib.load().kv_unique(ws.ptr, ids.data_ptr(), n, uniq.data_ptr(), idx.data_ptr(),
                                _ptr(counts), num.data_ptr(), _stream(dev)))
  if not sync:
    return (uniq, idx, counts, num) if with_counts else (uniq, idx, num)
  u = int(num.item())
  if with_counts:
    return uniq[:u], idx, counts[:u]
  return uniq[:u], idx


def unique_with_counts(x):
  return unique(x, with_counts=True)


def zero_rows(out, num_rows=None):
  """out[:num_rows] = 0 (num_rows: None = all rows, or a 1-element int32 device tensor)."""
  with torch.cuda.device(out.device):
    check(_lib.load().kv_zero_rows(out.data_ptr(), out.shape[0], _ptr(num_rows),
                                   out.numel() // max(1, out.shape[0]), _stream(out.device)))
  return out


def unsorted_segment_sum(data, segment_ids, num_segments, out=None, accumulate=False):
  """tf.math.unsorted_segment_sum; num_segments may be an int or a 1-element int32 device
  tensor (then `out` keeps data.shape[0] rows and only the first num_segments are defined)."""
  data = data.contiguous()
  seg = segment_ids.contiguous()
  if seg.dtype != torch.int32:
    seg = seg.to(torch.int32)
  n = data.shape[0]
  dim = data.numel() // max(n, 1) if n else (data.shape[1] if data.dim() > 1 else 1)
  dev = data.device
  if isinstance(num_segments, torch.Tensor):
    max_seg, dnum = n, num_segments
  else:
    max_seg, dnum = int(num_segments), None
  if out is None:
    out = torch.empty((max_seg, dim), dtype=torch.float32, device=dev)
  ws = Workspace.get(dev)
  with torch.cuda.device(dev):
    check(_lib.load().kv_segment_sum(ws.ptr, data.data_ptr(), seg.data_ptr(), n, dim, max_seg,
                                     _ptr(dnum), out.data_ptr(), int(bool(accumulate)),
                                     _stream(dev)))
  return out


def partition_ids(ids, num_shards, mode="hash", num_ids=None):
  """Group ids by owner shard: (sorted_ids, perm, shard_counts)."""
  ids = ids.contiguous()
  n = ids.numel()
  dev = ids.device
  sorted_ids = torch.empty(n, dtype=torch.int64, device=dev)
  perm = torch.empty(n, dtype=torch.int32, device=dev)
  counts = torch.empty(num_shards, dtype=torch.int32, device=dev)
  ws = Workspace.get(dev)
  with torch.cuda.device(dev):
    check(_lib.load().kv_partition_ids(ws.ptr, ids.data_ptr(), n, _ptr(num_ids), num_shards,
                                       1 if mode == "mod" else 0, sorted_ids.data_ptr(),
                                       perm.data_ptr(), counts.data_ptr(), _stream(dev)))
  return sorted_ids, perm, counts


def route_ids(ids, occ, num_shards, capacity, mode="hash", num_ids=None, out=None):
  """Fixed-capacity routing (kv_route_ids).  `out` = dict of preallocated send_ids
  [num_shards*capacity] i64, send_occ (same, i32), perm [n] i32, counts [num_shards] i32,
  overflow [1] i32 (all reused across calls so the exchange can be graph-captured)."""
  n = ids.numel()
  dev = ids.device
  if out is None:
    out = {"send_ids": torch.empty(num_shards * capacity, dtype=torch.int64, device=dev),
           "send_occ": torch.empty(num_shards * capacity, dtype=torch.int32, device=dev),
           "perm": torch.empty(n, dtype=torch.int32, device=dev),
           "counts": torch.empty(num_shards, dtype=torch.int32, device=dev),
           "overflow": torch.zeros(1, dtype=torch.int32, device=dev)}
  ws = Workspace.get(dev)
  with torch.cuda.device(dev):
    check(_lib.load().kv_route_ids(ws.ptr, ids.data_ptr(), _ptr(occ), n, _ptr(num_ids),
                                   num_shards, 1 if mode == "mod" else 0, capacity,
                                   out["send_ids"].data_ptr(), out["send_occ"].data_ptr(),
                                   out["perm"].data_ptr(), out["counts"].data_ptr(),
                                   out["overflow"].data_ptr(), _stream(dev)))
  return out


def route_id_pairs(ids, occ, num_shards, capacity, mode, num_ids, out):
  """kv_route_id_pairs: out = dict(send_pairs [num_shards*capacity*2] i64, perm, counts,
  overflow)."""
  ws = Workspace.get(ids.device)
  with torch.cuda.device(ids.device):
    check(_lib.load().kv_route_id_pairs(ws.ptr, ids.data_ptr(), _ptr(occ), ids.numel(),
                                        _ptr(num_ids), num_shards, 1 if mode == "mod" else 0,
                                        capacity, out["send_pairs"].data_ptr(),
                                        out["perm"].data_ptr(), out["counts"].data_ptr(),
                                        out["overflow"].data_ptr(), _stream(ids.device)))
  return out


def unzip_pairs(pairs, ids, occ):
  with torch.cuda.device(pairs.device):
    check(_lib.load().kv_unzip_pairs(pairs.data_ptr(), ids.numel(), ids.data_ptr(),
                                     occ.data_ptr(), _stream(pairs.device)))


def expand_rows(src, perm, idx, n, out):
  """out[i] = src[perm[idx[i]]] (perm / idx may be None)."""
  with torch.cuda.device(src.device):
    check(_lib.load().kv_expand_rows(src.data_ptr(), _ptr(perm), _ptr(idx), n, src.shape[1],
                                     out.data_ptr(), _stream(src.device)))
  return out


def scatter_rows_n(src, perm, n, num, out):
  """out[perm[i]] = src[i] for i < min(n, num[0])."""
  with torch.cuda.device(src.device):
    check(_lib.load().kv_scatter_rows_n(src.data_ptr(), perm.data_ptr(), n, _ptr(num),
                                        src.shape[1], out.data_ptr(), _stream(src.device)))
  return out


# ---- peer-memory variants (include/kvhbm.h "Peer-memory"): `seg` is an int64 device tensor of
# num_shards device pointers, normally into the peers' symmetric buffers ----
def route_ids_peer(ids, occ, num_shards, capacity, mode, num_ids, seg_ids, seg_occ, out):
  """kv_route_ids_peer: ids / counts go to seg_ids[g][0..capacity) / seg_occ[g][0..capacity);
  out = dict(perm, counts, overflow)."""
  ws = Workspace.get(ids.device)
  with torch.cuda.device(ids.device):
    check(_lib.load().kv_route_ids_peer(ws.ptr, ids.data_ptr(), _ptr(occ), ids.numel(),
                                        _ptr(num_ids), num_shards, 1 if mode == "mod" else 0,
                                        capacity, seg_ids.data_ptr(), seg_occ.data_ptr(),
                                        out["perm"].data_ptr(), out["counts"].data_ptr(),
                                        out["overflow"].data_ptr(), _stream(ids.device)))
  return out


def unique_route_peer(ids, uniq, idx, counts, num, num_shards, capacity, mode, seg_ids, seg_occ,
                      out, ws=None):
  """kv_unique_route_peer: unique_into + route_ids_peer in the same launches;
  out = dict(perm, counts (per shard, zeroed by route_fill_peer), overflow)."""
  ws = ws or Workspace.get(ids.device)
  with torch.cuda.device(ids.device):
    check(_lib.load().kv_unique_route_peer(ws.ptr, ids.data_ptr(), ids.numel(), uniq.data_ptr(),
                                           idx.data_ptr(), _ptr(counts), num.data_ptr(),
                                           num_shards, 1 if mode == "mod" else 0, capacity,
                                           seg_ids.data_ptr(), seg_occ.data_ptr(),
                                           out["perm"].data_ptr(), out["counts"].data_ptr(),
                                           out["overflow"].data_ptr(), _stream(ids.device)))
  return out


def route_fill_peer(num_shards, capacity, seg_ids, seg_occ, shard_counts):
  """kv_route_fill_peer: pad the owners' rows and zero the shard counters for the next step."""
  with torch.cuda.device(shard_counts.device):
    check(_lib.load().kv_route_fill_peer(num_shards, capacity, seg_ids.data_ptr(),
                                         seg_occ.data_ptr(), shard_counts.data_ptr(),
                                         _stream(shard_counts.device)))


def kv_variable_gather_or_insert_peer(table_handle, indices, counts, seg, capacity):
  """KvVariableGatherOrInsertWithCounts whose row r lands in seg[r // capacity][r % capacity]."""
  h = table_handle
  ids = _ids(indices, h)
  if ids.numel():
    check(h._lib.kv_gather_or_insert_peer(h._live(), ids.data_ptr(), _ptr(counts), ids.numel(),
                                          seg.data_ptr(), capacity, today(), h.stream))


def scatter_rows_n_peer(src, perm, n, num, seg, capacity):
  """seg[perm[i] // capacity][perm[i] % capacity] = src[i] for i < min(n, num[0])."""
  with torch.cuda.device(src.device):
    check(_lib.load().kv_scatter_rows_n_peer(src.data_ptr(), perm.data_ptr(), n, _ptr(num),
                                             src.shape[1], seg.data_ptr(), capacity,
                                             _stream(src.device)))


def peer_barrier(peer_flags, my_flags, state, rank, world, timeout_ms=2000):
  """kv_peer_barrier on the current stream of my_flags' device."""
  with torch.cuda.device(my_flags.device):
    check(_lib.load().kv_peer_barrier(peer_flags.data_ptr(), my_flags.data_ptr(),
                                      state.data_ptr(), rank, world, timeout_ms,
                                      _stream(my_flags.device)))


def unique_into(ids, uniq, idx, counts, num, ws=None):
  """kv_unique into caller-owned buffers (graph-capturable: nothing is allocated or read back).
  Calls that may run concurrently on one device need a Workspace each (`ws`)."""
  ws = ws or Workspace.get(ids.device)
  with torch.cuda.device(ids.device):
    check(_lib.load().kv_unique(ws.ptr, ids.data_ptr(), ids.numel(), uniq.data_ptr(),
                                idx.data_ptr(), _ptr(counts), num.data_ptr(),
                                _stream(ids.device)))


def permute_rows(src, perm, out=None):
  """out[i] = src[perm[i]]."""
  src = src.contiguous()
  n = perm.numel()
  dim = src.shape[1]
  if out is None:
    out = torch.empty((n, dim), dtype=torch.float32, device=src.device)
  with torch.cuda.device(src.device):
    check(_lib.load().kv_permute_rows(src.data_ptr(), perm.data_ptr(), n, dim, out.data_ptr(),
                                      _stream(src.device)))
  return out


def scatter_rows(src, perm, out):
  """out[perm[i]] = src[i]."""
  src = src.contiguous()
  with torch.cuda.device(src.device):
    check(_lib.load().kv_scatter_rows(src.data_ptr(), perm.data_ptr(), perm.numel(),
                                      src.shape[1], out.data_ptr(), _stream(src.device)))
  return out
