"""CPU-only checks of the drop-in boundary: libkvhbm.so loads without a GPU, exports exactly
what include/kvhbm.h declares, the ctypes binding covers it, and the product path fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kvhbm.h")


def declared_functions():
  src = open(HEADER).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  names = re.findall(r"\b(?:int|int64_t|const char\*)\s+(kv_[a-z0-9_]+)\s*\(", src)
  return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
  from tfplus_b200 import build
  build.build()
  from tfplus_b200 import _lib
  return _lib.load()


def test_header_declares_the_whole_surface():
  names = declared_functions()
  for must in ["kv_create", "kv_destroy", "kv_set_init_table", "kv_gather_or_insert",
               "kv_gather_or_zeros", "kv_insert_or_update", "kv_scatter", "kv_apply_adagrad",
               "kv_apply_group_adam_v4", "kv_apply_sparse_group_ftrl", "kv_apply_adam",
               "kv_apply_group_adam_v4_dev", "kv_unique", "kv_segment_sum", "kv_export_count",
               "kv_export", "kv_import", "kv_delete", "kv_delete_with_timestamp", "kv_size",
               "kv_sum_freq", "kv_map_size", "kv_partition_ids", "kv_last_error"]:
    assert must in names, must
  assert len(names) >= 40


def test_library_exports_every_declared_symbol(lib):
  raw = ctypes.CDLL(os.path.join(ROOT, "tfplus_b200", "libkvhbm.so"))
  missing = [n for n in declared_functions() if not hasattr(raw, n)]
  assert not missing, "declared in kvhbm.h but not exported: %s" % missing


def test_ctypes_binding_covers_the_header():
  from tfplus_b200 import _lib
  missing = [n for n in declared_functions() if n not in _lib.SIGNATURES]
  assert not missing, "no ctypes signature for: %s" % missing


def test_header_cites_the_reference():
  src = open(HEADER).read()
  for cite in ["kernels/kv_variable_ops.cc:498-631", "kernels/training_ops.cc:6980-7235",
               "kernels/training_ops.cc:532-801", "kernels/training_ops.cc:1372-1520",
               "kernels/dynamic_save.hpp:48-195", "kernels/dynamic_restore.hpp:156-262"]:
    assert cite in src, cite


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
  out = ctypes.c_void_p()
  rc = lib.kv_create(8, 0, 0, ctypes.byref(out))
  assert rc != 0 and out.value is None
  assert b"CUDA device" in lib.kv_last_error()
  ws = ctypes.c_void_p()
  assert lib.kv_workspace_create(ctypes.byref(ws)) != 0
  from tfplus_b200 import ops
  with pytest.raises(RuntimeError):
    ops.kv_variable(value_shape=[8], device="cuda")
  with pytest.raises(RuntimeError, match="CUDA device"):
    ops.kv_variable(value_shape=[8], device="cpu")


def test_product_code_never_touches_the_oracle():
  pkg = os.path.join(ROOT, "tfplus_b200")
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
        text = open(os.path.join(dirpath, f)).read()
        assert "oracle" not in text.replace("oracle/kv_oracle.cc", "").replace(
            "see oracle", ""), (dirpath, f)


def test_argument_checks_need_no_gpu(lib):
  """Null / out-of-range arguments are refused with InvalidArgument and a message before any
  CUDA call, as the reference's OP_REQUIRES checks are (no device work happens here)."""
  null = ctypes.c_void_p()
  INVALID = 1
  cases = [
      ("kv_peer_barrier", (null, null, null, 0, 2, 1000, null), b"peer_barrier"),
      ("kv_route_fill_peer", (4, 128, null, null, null, null), b"route_fill_peer"),
      ("kv_unique_route_peer", (null, null, 0, null, null, null, null, 4, 0, 128, null, null,
                                null, null, null, null), b"unique_route_peer"),
      ("kv_route_ids_peer", (null, null, null, 0, null, 4, 0, 128, null, null, null, null, null,
                             null), b"route_ids_peer"),
      ("kv_route_id_pairs", (null, null, null, 0, null, 4, 0, 128, null, null, null, null, null),
       b"route_id_pairs"),
  ]
  for name, args, needle in cases:
    rc = getattr(lib, name)(*args)
    assert rc == INVALID, (name, rc)
    assert needle in lib.kv_last_error(), (name, lib.kv_last_error())
