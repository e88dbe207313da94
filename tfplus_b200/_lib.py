"""ctypes loader of libkvhbm.so, the C ABI declared in include/kvhbm.h.

There is no CPU fallback: if the library has not been built, or no CUDA device
is present when a table is created, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkvhbm.so")

KV_OK = 0
_STATUS_EXC = {1: ValueError, 2: RuntimeError, 3: NotImplementedError, 4: MemoryError,
               5: RuntimeError}
_STATUS_NAME = {1: "InvalidArgument", 2: "FailedPrecondition", 3: "Unimplemented",
                4: "ResourceExhausted", 5: "Internal"}

vp, i64, i32, f32, u16, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint16, C.c_uint64

# name -> argtypes (every function returns int unless listed in _RESTYPE)
SIGNATURES = {
    "kv_last_error": [],
    "kv_launch_count": [],
    "kv_create": [i32, i32, i64, C.POINTER(vp)],
    "kv_destroy": [vp],
    "kv_dim": [vp],
    "kv_enter_threshold": [vp],
    "kv_set_seed": [vp, u64],
    "kv_reserve": [vp, i64, vp],
    "kv_set_init_table": [vp, vp, i64, vp],
    "kv_is_initialized": [vp, C.POINTER(i32)],
    "kv_init_table_rows": [vp, C.POINTER(i64)],
    "kv_get_init_table": [vp, vp, vp],
    "kv_size": [vp, vp, C.POINTER(i64)],
    "kv_sum_freq": [vp, vp, C.POINTER(i64)],
    "kv_map_size": [vp, vp, C.POINTER(i64)],
    "kv_gather_or_insert": [vp, vp, vp, i64, vp, u16, vp],
    "kv_gather_or_insert_n": [vp, vp, vp, i64, vp, vp, u16, vp],
    "kv_gather_or_zeros": [vp, vp, i64, vp, vp],
    "kv_insert_or_update": [vp, vp, vp, i64, vp, vp, vp],
    "kv_scatter": [vp, i32, vp, vp, i64, vp],
    "kv_scatter_unique": [vp, i32, vp, vp, i64, vp],
    "kv_get_count": [vp, vp, i64, vp, vp],
    "kv_get_timestamp": [vp, vp, i64, vp, u16, vp],
    "kv_apply_adagrad": [vp, vp, vp, vp, i64, vp, f32, i32, u16, vp],
    "kv_apply_group_adam_v4": [vp, vp, vp, vp, i64, vp] + [f32] * 9 + [u16, vp],
    "kv_apply_sparse_group_ftrl": [vp, vp, vp, vp, vp, i64, vp] + [f32] * 6 + [u16, vp],
    "kv_apply_adam": [vp, vp, vp, vp, i64, vp] + [f32] * 6 + [u16, vp],
    "kv_apply_adagrad_dev": [vp, vp, vp, vp, i64, vp, vp, i32, u16, vp],
    "kv_apply_group_adam_v4_dev": [vp, vp, vp, vp, i64, vp, vp, u16, vp],
    "kv_apply_sparse_group_ftrl_dev": [vp, vp, vp, vp, vp, i64, vp, vp, u16, vp],
    "kv_apply_adam_dev": [vp, vp, vp, vp, i64, vp, vp, u16, vp],
    "kv_apply_group_adam_v4_dev_advance": [vp, vp, vp, vp, i64, vp, vp, u16, vp],
    "kv_apply_adam_dev_advance": [vp, vp, vp, vp, i64, vp, vp, u16, vp],
    "kv_apply_group_adam_v3": [vp, vp, vp, vp, i64, vp] + [f32] * 9 + [u16, vp],
    "kv_apply_sparse_ftrl_v2": [vp, vp, vp, vp, vp, i64, vp] + [f32] * 5 + [u16, vp],
    "kv_apply_group_sparse_ftrl_v2": [vp, vp, vp, vp, vp, i64, vp] + [f32] * 5 + [u16, vp],
    "kv_plan_create": [i64, C.POINTER(vp)],
    "kv_plan_destroy": [vp],
    "kv_plan_build": [vp, vp, vp, i64, vp],
    "kv_plan_arrays": [vp] + [C.POINTER(vp)] * 6,
    "kv_gather_or_insert_plan": [vp, vp, vp, u16, vp],
    "kv_gather_or_zeros_plan": [vp, vp, vp, vp],
    "kv_segment_sum_plan": [vp, vp, i32, vp, vp],
    "kv_apply_plan": [i32, vp, vp, vp, vp, vp, C.POINTER(f32), i32, i32, u16, vp],
    "kv_apply_plan_dev": [i32, vp, vp, vp, vp, vp, vp, i32, i32, u16, vp],
    "kv_workspace_create": [C.POINTER(vp)],
    "kv_workspace_destroy": [vp],
    "kv_unique": [vp, vp, i64, vp, vp, vp, vp, vp],
    "kv_segment_sum": [vp, vp, vp, i64, i32, i64, vp, vp, i32, vp],
    "kv_zero_rows": [vp, i64, vp, i32, vp],
    "kv_export_count": [vp, i32, i32, f32, vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)],
    "kv_export": [vp, i32, vp, vp, vp, vp, vp, i32, vp],
    "kv_export_bounded": [vp, i32, vp, vp, i64, vp, i64, vp, vp, i64, i32, vp],
    "kv_import": [vp, vp, vp, i64, vp, i64, vp, i64, vp, vp, i64, i32, vp],
    "kv_delete": [vp, vp, i64, vp],
    "kv_check_overflow": [vp, vp],
    "kv_sparse_combine": [vp, vp, vp, vp, i64, i64, i32, i32, vp, vp],
    "kv_enable_delta_export": [vp, i32],
    "kv_delta_size": [vp, vp, C.POINTER(i64)],
    "kv_delta_export_count": [vp, i32, vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64),
                              C.POINTER(i64)],
    "kv_delta_export": [vp, i32, vp, vp, i64, vp, i64, vp, vp, i64, vp, i64, vp, C.POINTER(i64)],
    "kv_delta_import": [vp, i32, vp, vp, i64, vp, i64, vp, vp, i64, vp, i64, vp],
    "kv_delete_with_timestamp": [vp, i32, u16, vp, i64, vp, C.POINTER(i64)],
    "kv_partition_ids": [vp, vp, i64, vp, i32, i32, vp, vp, vp, vp],
    "kv_route_ids": [vp, vp, vp, i64, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "kv_route_id_pairs": [vp, vp, vp, i64, vp, i32, i32, i32, vp, vp, vp, vp, vp],
    "kv_unzip_pairs": [vp, i64, vp, vp, vp],
    "kv_expand_rows": [vp, vp, vp, i64, i32, vp, vp],
    "kv_scatter_rows_n": [vp, vp, i64, vp, i32, vp, vp],
    "kv_permute_rows": [vp, vp, i64, i32, vp, vp],
    "kv_scatter_rows": [vp, vp, i64, i32, vp, vp],
    "kv_route_ids_peer": [vp, vp, vp, i64, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "kv_unique_route_peer": [vp, vp, i64, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "kv_route_fill_peer": [i32, i32, vp, vp, vp, vp],
    "kv_gather_or_insert_peer": [vp, vp, vp, i64, vp, i64, u16, vp],
    "kv_scatter_rows_n_peer": [vp, vp, i64, vp, i32, vp, i64, vp],
    "kv_peer_barrier": [vp, vp, vp, i32, i32, i64, vp],
}
_RESTYPE = {"kv_last_error": C.c_char_p, "kv_launch_count": i64}

_lib = None


class KvError(RuntimeError):
  pass


def load():
  """Returns the loaded library; raises if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        "tfplus_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; "
        "g.build()'` (or python tfplus_b200/build.py); there is no CPU fallback." % LIB_PATH)
  lib = C.CDLL(LIB_PATH)
  for name, args in SIGNATURES.items():
    fn = getattr(lib, name)
    fn.argtypes = args
    fn.restype = _RESTYPE.get(name, C.c_int)
  _lib = lib
  return lib


def check(status):
  if status == KV_OK:
    return
  msg = load().kv_last_error().decode("utf-8", "replace")
  exc = _STATUS_EXC.get(status, RuntimeError)
  raise exc("%s: %s" % (_STATUS_NAME.get(status, "status %d" % status), msg))


def launch_count():
  return int(load().kv_launch_count())
