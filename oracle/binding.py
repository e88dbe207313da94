"""ctypes binding of the CPU oracle (oracle/kv_oracle.cc).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under tfplus_b200/
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkvoracle.so")

SCATTER_OPS = {"assign": 0, "update": 0, "add": 1, "sub": 2, "mul": 3,
               "div": 4, "min": 5, "max": 6}


def build(force=False):
  """Compile the oracle with the recipe in oracle/Makefile."""
  src = os.path.join(_HERE, "kv_oracle.cc")
  if (force or not os.path.exists(_SO)
      or os.path.getmtime(_SO) < os.path.getmtime(src)):
    subprocess.check_call(["make", "-C", _HERE, "-s"])
  return _SO


_lib = None


def lib():
  global _lib
  if _lib is not None:
    return _lib
  build()
  L = C.CDLL(_SO)
  vp, i64, i32, f32, u16, u32, u64 = (C.c_void_p, C.c_int64, C.c_int32,
                                      C.c_float, C.c_uint16, C.c_uint32,
                                      C.c_uint64)
  P = C.c_void_p  # raw array pointers
  sig = {
      "kvo_set_threads": (None, [C.c_int]),
      "kvo_get_threads": (C.c_int, []),
      "kvo_create": (vp, [C.c_int, C.c_int]),
      "kvo_destroy": (None, [vp]),
      "kvo_set_seed": (None, [vp, u64]),
      "kvo_set_init_table": (None, [vp, P, i64]),
      "kvo_is_initialized": (C.c_int, [vp]),
      "kvo_init_rows": (i64, [vp]),
      "kvo_get_init_table": (None, [vp, P]),
      "kvo_size": (i64, [vp]),
      "kvo_sum_freq": (i64, [vp]),
      "kvo_map_size": (i64, [vp]),
      "kvo_gather_or_insert": (None, [vp, P, P, i64, P, u16]),
      "kvo_gather_or_zeros": (None, [vp, P, i64, P]),
      "kvo_insert_or_update": (None, [vp, P, P, i64, P, P]),
      "kvo_scatter": (None, [vp, C.c_int, P, P, i64]),
      "kvo_apply_adagrad": (None, [vp, vp, P, P, i64, f32, C.c_int, u16]),
      "kvo_apply_group_adam_v4": (None, [vp, vp, P, P, i64] + [f32] * 9 + [u16]),
      "kvo_apply_sparse_group_ftrl": (None, [vp, vp, vp, P, P, i64] + [f32] * 6 + [u16]),
      "kvo_enable_delta_export": (None, [vp, C.c_int]),
      "kvo_delta_size": (i64, [vp]),
      "kvo_delta_export": (None, [vp, C.c_int, P, P, P, P]),
      "kvo_delta_export_fetch_deleted": (None, [vp, P]),
      "kvo_delta_import": (None, [vp, C.c_int, P, P, i64, P, i64, P, P, i64, P, i64]),
      "kvo_apply_group_adam_v3": (None, [vp, vp, P, P, i64] + [f32] * 9 + [u16]),
      "kvo_apply_sparse_ftrl_v2": (None, [vp, vp, vp, P, P, i64] + [f32] * 5 + [u16]),
      "kvo_apply_group_sparse_ftrl_v2": (None, [vp, vp, vp, P, P, i64] + [f32] * 5 + [u16]),
      "kvo_adam_step": (None, [vp, vp, P, P, i64] + [f32] * 6 + [u16]),
      "kvo_get_count": (None, [vp, P, i64, P]),
      "kvo_get_timestamp": (None, [vp, P, i64, P, u16]),
      "kvo_freq_word": (u32, [vp, i64]),
      "kvo_key_flags": (C.c_int, [vp, i64]),
      "kvo_delete": (None, [vp, P, i64]),
      "kvo_delete_with_timestamp": (i64, [vp, C.c_int, u16, P, i64]),
      "kvo_export": (None, [vp, C.c_int, C.c_int, f32, C.c_int, P, P, P]),
      "kvo_export_fetch": (None, [vp, P, P, P, P, P]),
      "kvo_import": (None, [vp, P, P, i64, P, i64, P, i64, P, P, i64]),
      "kvo_unique": (i64, [P, i64, P, P, P]),
      "kvo_segment_sum": (None, [P, P, i64, C.c_int, i64, P]),
  }
  for name, (res, args) in sig.items():
    fn = getattr(L, name)
    fn.restype = res
    fn.argtypes = args
  _lib = L
  return L


def _p(a):
  return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ids(a):
  return np.ascontiguousarray(np.asarray(a, dtype=np.int64).reshape(-1))


def _f32(a):
  return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def set_threads(n):
  lib().kvo_set_threads(int(n))


def unique(ids, with_counts=False):
  """tf.unique / tf.unique_with_counts on int64 ids."""
  ids = _ids(ids)
  n = ids.size
  uniq = np.empty(n, np.int64)
  idx = np.empty(n, np.int32)
  counts = np.empty(n, np.int32) if with_counts else None
  u = lib().kvo_unique(_p(ids), n, _p(uniq), _p(idx), _p(counts))
  if with_counts:
    return uniq[:u].copy(), idx, counts[:u].copy()
  return uniq[:u].copy(), idx


def segment_sum(data, idx, num_segments):
  data = _f32(data)
  idx = np.ascontiguousarray(np.asarray(idx, np.int32))
  d = data.shape[1]
  out = np.empty((num_segments, d), np.float32)
  lib().kvo_segment_sum(_p(data), _p(idx), data.shape[0], d, num_segments, _p(out))
  return out


class OracleTable:
  """One KvVariable<int64,float> of the reference, on the CPU."""

  def __init__(self, dim, enter_threshold=0, seed=0):
    self.dim = int(dim)
    self.enter_threshold = int(enter_threshold)
    self._h = lib().kvo_create(self.dim, self.enter_threshold)
    lib().kvo_set_seed(self._h, seed)

  def __del__(self):
    if getattr(self, "_h", None) and _lib is not None:
      _lib.kvo_destroy(self._h)
      self._h = None

  # lifecycle ---------------------------------------------------------------
  def set_init_table(self, tbl):
    tbl = _f32(tbl).reshape(-1, self.dim)
    lib().kvo_set_init_table(self._h, _p(tbl), tbl.shape[0])

  def is_initialized(self):
    return bool(lib().kvo_is_initialized(self._h))

  def size(self):
    return int(lib().kvo_size(self._h))

  def sum_freq(self):
    return int(lib().kvo_sum_freq(self._h))

  def map_size(self):
    return int(lib().kvo_map_size(self._h))

  def shape(self):
    return [self.map_size(), self.dim]

  # lookups -----------------------------------------------------------------
  def gather_or_insert(self, ids, counts=None, today=0):
    ids = _ids(ids)
    out = np.empty((ids.size, self.dim), np.float32)
    c = None if counts is None else np.ascontiguousarray(
        np.asarray(counts, np.int32).reshape(-1))
    lib().kvo_gather_or_insert(self._h, _p(ids), _p(c), ids.size, _p(out), today)
    return out

  def gather_or_zeros(self, ids):
    ids = _ids(ids)
    out = np.empty((ids.size, self.dim), np.float32)
    lib().kvo_gather_or_zeros(self._h, _p(ids), ids.size, _p(out))
    return out

  def insert_or_update(self, ids, values, filter_out=None, blacklist=None):
    ids = _ids(ids)
    values = _f32(values).reshape(ids.size, self.dim)
    f = None if filter_out is None else np.ascontiguousarray(
        np.asarray(filter_out, np.uint8))
    b = None if blacklist is None else np.ascontiguousarray(
        np.asarray(blacklist, np.uint8))
    lib().kvo_insert_or_update(self._h, _p(ids), _p(values), ids.size, _p(f), _p(b))

  def scatter(self, op, ids, updates):
    ids = _ids(ids)
    updates = _f32(updates).reshape(ids.size, self.dim)
    lib().kvo_scatter(self._h, SCATTER_OPS[op], _p(ids), _p(updates), ids.size)

  def get_count(self, ids):
    ids = _ids(ids)
    out = np.empty(ids.size, np.int32)
    lib().kvo_get_count(self._h, _p(ids), ids.size, _p(out))
    return out

  def get_timestamp(self, ids, today=0):
    ids = _ids(ids)
    out = np.empty(ids.size, np.uint32)
    lib().kvo_get_timestamp(self._h, _p(ids), ids.size, _p(out), today)
    return out

  # -- delta checkpoints (dynamic_save.hpp:197-449, dynamic_restore.hpp:28-153) --
  def enable_delta_export(self, support_prediction_delta=False):
    lib().kvo_enable_delta_export(self._h, int(bool(support_prediction_delta)))

  def delta_size(self):
    return int(lib().kvo_delta_size(self._h))

  def delta_export(self, first_n=6):
    """KvVariableFullOrDeltaExport in delta mode -> dict(keys, values, blacklist, freq_keys,
    freq_values (uint32 words), delete_keys)."""
    nk, nb, nf, nd = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    lib().kvo_delta_export(self._h, first_n, C.byref(nk), C.byref(nb), C.byref(nf), C.byref(nd))
    keys = np.empty(nk.value, np.int64)
    vals = np.empty((nk.value, self.dim), np.float32)
    black = np.empty(nb.value, np.int64)
    fk = np.empty(nf.value, np.int64)
    fv = np.empty(nf.value, np.uint32)
    dk = np.empty(nd.value, np.int64)
    lib().kvo_export_fetch(self._h, _p(keys), _p(vals), _p(black), _p(fk), _p(fv))
    lib().kvo_delta_export_fetch_deleted(self._h, _p(dk))
    return dict(keys=keys, values=vals, blacklist=black, freq_keys=fk, freq_values=fv,
                delete_keys=dk)

  def delta_import(self, keys, values, blacklist, freq_keys, freq_values, delete_keys,
                   first_n=6):
    keys = _ids(keys)
    values = _f32(values).reshape(keys.size, self.dim)
    blacklist, freq_keys, delete_keys = _ids(blacklist), _ids(freq_keys), _ids(delete_keys)
    fv = np.ascontiguousarray(freq_values, np.uint32)
    lib().kvo_delta_import(self._h, first_n, _p(keys), _p(values), keys.size, _p(blacklist),
                           blacklist.size, _p(freq_keys), _p(fv), freq_keys.size,
                           _p(delete_keys), delete_keys.size)

  def freq_word(self, key):
    return int(lib().kvo_freq_word(self._h, int(key)))

  def key_flags(self, key):
    return int(lib().kvo_key_flags(self._h, int(key)))

  # eviction ----------------------------------------------------------------
  def delete(self, ids):
    ids = _ids(ids)
    lib().kvo_delete(self._h, _p(ids), ids.size)

  def delete_with_timestamp(self, threshold, today):
    cap = max(1, self.map_size())
    out = np.empty(cap, np.int64)
    n = lib().kvo_delete_with_timestamp(self._h, int(threshold), today, _p(out), cap)
    return out[:n].copy()

  # checkpoint --------------------------------------------------------------
  def export(self, first_n=6, enable_cutoff=True, cutoff_value=1e-20,
             freq_u32=False):
    """KvVariableExport -> dict of the 6 output tensors."""
    nk, nb, nf = C.c_int64(), C.c_int64(), C.c_int64()
    lib().kvo_export(self._h, first_n, int(enable_cutoff), cutoff_value,
                     int(freq_u32), C.byref(nk), C.byref(nb), C.byref(nf))
    keys = np.empty(nk.value, np.int64)
    vals = np.empty((nk.value, self.dim), np.float32)
    black = np.empty(nb.value, np.int64)
    fk = np.empty(nf.value, np.int64)
    fv = np.empty(nf.value, np.uint32)
    lib().kvo_export_fetch(self._h, _p(keys), _p(vals), _p(black), _p(fk), _p(fv))
    if first_n > 3:
      rows = int(lib().kvo_init_rows(self._h))
      init = np.empty((rows, self.dim), np.float32)
      if rows:
        lib().kvo_get_init_table(self._h, _p(init))
    else:
      init = np.empty((0, self.dim), np.float32)
    if not freq_u32:
      fv = fv.astype(np.uint16)
    out = {"keys": keys, "values": vals}
    if first_n > 2:
      out.update({"init_table": init, "blacklist": black, "freq_keys": fk,
                  "freq_values": fv})
    return out

  def import_(self, keys, values, init_table=None, blacklist=None,
              freq_keys=None, freq_values=None):
    keys = _ids(keys)
    values = _f32(values).reshape(keys.size, self.dim)
    it = None if init_table is None or np.size(init_table) == 0 else _f32(
        init_table).reshape(-1, self.dim)
    bl = _ids(blacklist if blacklist is not None else [])
    fk = _ids(freq_keys if freq_keys is not None else [])
    fv = np.ascontiguousarray(np.asarray(
        freq_values if freq_values is not None else [], dtype=np.uint32).reshape(-1))
    lib().kvo_import(self._h, _p(keys), _p(values), keys.size, _p(it),
                     0 if it is None else it.shape[0], _p(bl), bl.size, _p(fk),
                     _p(fv), fk.size)


# fused optimizer applies (free functions: they span several tables) ----------
def apply_adagrad(var, accum, ids, grad, lr, update_slots=True, today=0):
  ids = _ids(ids)
  grad = _f32(grad).reshape(ids.size, var.dim)
  lib().kvo_apply_adagrad(var._h, accum._h, _p(ids), _p(grad), ids.size, lr,
                          int(update_slots), today)


def apply_group_adam_v4(var, m_v_linear, ids, grad, lr, beta1_power, beta2_power,
                        beta1, beta2, epsilon, l1, l2, l21, today=0):
  ids = _ids(ids)
  grad = _f32(grad).reshape(ids.size, var.dim)
  lib().kvo_apply_group_adam_v4(var._h, m_v_linear._h, _p(ids), _p(grad), ids.size,
                                lr, beta1_power, beta2_power, beta1, beta2, epsilon,
                                l1, l2, l21, today)


def apply_sparse_group_ftrl(var, accum, linear, ids, grad, lr, l1, l2, l21,
                            l2_shrinkage=0.0, lr_power=-0.5, today=0):
  ids = _ids(ids)
  grad = _f32(grad).reshape(ids.size, var.dim)
  lib().kvo_apply_sparse_group_ftrl(var._h, accum._h, linear._h, _p(ids), _p(grad),
                                    ids.size, lr, l1, l2, l21, l2_shrinkage,
                                    lr_power, today)


def apply_group_adam_v3(var, m_v_linear, ids, grad, lr, beta1_power, beta2_power,
                        beta1, beta2, epsilon, l1, l2, l21, today=0):
  ids = _ids(ids)
  grad = _f32(grad).reshape(ids.size, var.dim)
  lib().kvo_apply_group_adam_v3(var._h, m_v_linear._h, _p(ids), _p(grad), ids.size,
                                lr, beta1_power, beta2_power, beta1, beta2, epsilon,
                                l1, l2, l21, today)


def apply_sparse_ftrl_v2(var, accum, linear, ids, grad, lr, l1, l2, l2_shrinkage=0.0,
                         lr_power=-0.5, today=0):
  ids = _ids(ids)
  grad = _f32(grad).reshape(ids.size, var.dim)
  lib().kvo_apply_sparse_ftrl_v2(var._h, accum._h, linear._h, _p(ids), _p(grad), ids.size,
                                 lr, l1, l2, l2_shrinkage, lr_power, today)


def apply_group_sparse_ftrl_v2(var, accum, linear, ids, grad, lr, l1, l2, l2_shrinkage=0.0,
                               lr_power=-0.5, today=0):
  ids = _ids(ids)
  grad = _f32(grad).reshape(ids.size, var.dim)
  lib().kvo_apply_group_sparse_ftrl_v2(var._h, accum._h, linear._h, _p(ids), _p(grad),
                                       ids.size, lr, l1, l2, l2_shrinkage, lr_power, today)


def adam_step(var, m_v, ids, grad, lr, beta1, beta2, epsilon, beta1_power,
              beta2_power, today=0):
  ids = _ids(ids)
  grad = _f32(grad).reshape(ids.size, var.dim)
  lib().kvo_adam_step(var._h, m_v._h, _p(ids), _p(grad), ids.size, lr, beta1, beta2,
                      epsilon, beta1_power, beta2_power, today)
