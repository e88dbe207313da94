"""Delta checkpoints on the device table against the oracle's restatement of
KvVariable::DeltaExport / DeltaImport (dynamic_save.hpp:197-449, dynamic_restore.hpp:28-153):
which keys get marked, the four variable-length outputs as key-sorted sets, both export modes,
and the import on top of an existing table."""
import numpy as np
import pytest
import torch

from oracle import binding as ob
from tfplus_b200 import ops

from kvtest_util import DEV, TODAY, Pair, t

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _clock():
  ops.set_today(TODAY)
  yield
  ops.set_today(None)


def _enable(p, pred=False):
  ops.kv_variable_enable_delta_export(p.gpu, pred)
  p.cpu.enable_delta_export(pred)


def _compare_delta(p, first_n):
  got = ops.kv_variable_full_or_delta_export(p.gpu, first_n=first_n)
  want = p.cpu.delta_export(first_n=first_n)
  keys, values, init_table, blacklist, fk, fv, need_full, delete_keys = [x.cpu().numpy() for x in got]
  assert init_table.shape == (0, p.dim) and need_full.tolist() == [False]
  o = np.argsort(keys)
  ow = np.argsort(want["keys"])
  np.testing.assert_array_equal(keys[o], want["keys"][ow])
  np.testing.assert_array_equal(values[o], want["values"][ow])
  np.testing.assert_array_equal(np.sort(blacklist), np.sort(want["blacklist"]))
  np.testing.assert_array_equal(np.sort(delete_keys), np.sort(want["delete_keys"]))
  o, ow = np.argsort(fk), np.argsort(want["freq_keys"])
  np.testing.assert_array_equal(fk[o], want["freq_keys"][ow])
  np.testing.assert_array_equal(fv.view(np.uint32)[o], want["freq_values"][ow])
  return dict(keys=keys, values=values, blacklist=blacklist, freq_keys=fk,
              freq_values=fv.view(np.uint32), delete_keys=delete_keys)


@pytest.mark.parametrize("pred", [False, True])
def test_delta_export_marks_and_modes(pred):
  dim = 16
  var, acc = Pair(dim, enter_threshold=2), Pair(dim, init=0.1)
  for p in (var, acc):
    _enable(p, pred)
  rng = np.random.default_rng(3)
  # before any export: everything touched is in the delta
  ids = rng.integers(0, 300, size=400).astype(np.int64)
  var.gather_or_insert(ids)
  u = np.unique(ids)
  g = rng.normal(size=(u.size, dim)).astype(np.float32)
  ops.kv_variable_sparse_apply_adagrad(var.gpu, acc.gpu, 0.05, t(g), t(u))      # low-frequency ids skipped
  ob.apply_adagrad(var.cpu, acc.cpu, u, g, 0.05, True, today=TODAY)
  assert ops.kv_variable_delta_size(var.gpu) == var.cpu.delta_size() == u.size
  assert ops.kv_variable_delta_size(acc.gpu) == acc.cpu.delta_size() < u.size   # only the applied ids
  d1 = _compare_delta(var, 6)
  _compare_delta(acc, 6)
  assert d1["freq_keys"].size == u.size and 0 < d1["keys"].size < u.size        # low-frequency keys: freq only
  assert ops.kv_variable_delta_size(var.gpu) == 0
  # second round: scatter, insert, delete, eviction, a blacklisting apply
  var.scatter("add", np.array([5, 7, 100001], np.int64), np.ones((3, dim), np.float32))
  var.insert(np.array([7, 100002], np.int64), np.full((2, dim), 2.0, np.float32))
  dk = np.array([int(u[0]), int(u[1]), 999999], np.int64)
  ops.kv_variable_delete(var.gpu, t(dk))
  var.cpu.delete(dk)
  d2 = _compare_delta(var, 6)
  assert set(dk.tolist()) <= set(d2["delete_keys"].tolist())
  # inference-mode export: train (now empty) + what the training exports handed over
  d3 = _compare_delta(var, 3)
  if pred:
    assert d3["keys"].size > 0
  else:
    assert d3["keys"].size == 0 and d3["delete_keys"].size == 0
  _compare_delta(var, 3)      # the prediction list was cleared
  var.check_state()


def test_delta_import_on_top_of_a_table():
  dim = 8
  src = Pair(dim, enter_threshold=0)
  dst = Pair(dim, enter_threshold=0)
  _enable(src)
  rng = np.random.default_rng(5)
  base = np.arange(200, dtype=np.int64)
  src.gather_or_insert(base)
  dst.gather_or_insert(base)
  _compare_delta(src, 6)                       # checkpoint 0: both sides hold `base`
  # changes on the source: updates, new keys, a blacklisted key, deletions
  src.scatter("add", np.arange(0, 50, dtype=np.int64), rng.normal(size=(50, dim)).astype(np.float32))
  src.gather_or_insert(np.arange(200, 230, dtype=np.int64))
  src.insert(np.array([3], np.int64), np.zeros((1, dim), np.float32), blacklist=np.array([1], np.uint8))
  dk = np.array([10, 11, 12], np.int64)
  ops.kv_variable_delete(src.gpu, t(dk))
  src.cpu.delete(dk)
  d = _compare_delta(src, 6)
  assert d["blacklist"].tolist() == [3] and set(d["delete_keys"].tolist()) == {10, 11, 12}
  ops.kv_variable_full_or_delta_import_v2(dst.gpu, t(d["keys"]), t(d["values"]), None, t(d["blacklist"]),
                                          t(d["freq_keys"]), t(d["freq_values"].view(np.int32)),
                                          torch.zeros(1, dtype=torch.bool), t(d["delete_keys"]), first_n=6)
  dst.cpu.delta_import(d["keys"], d["values"], d["blacklist"], d["freq_keys"], d["freq_values"],
                       d["delete_keys"], first_n=6)
  dst.check_state()
  # and the replica now equals the source, key for key
  ids = np.arange(0, 230, dtype=np.int64)
  np.testing.assert_array_equal(ops.kv_variable_gather_or_zeros_v2(dst.gpu, t(ids)).cpu().numpy(),
                                ops.kv_variable_gather_or_zeros_v2(src.gpu, t(ids)).cpu().numpy())
  assert ops.kv_variable_size_v2(dst.gpu) == ops.kv_variable_size_v2(src.gpu)


def test_delta_export_needs_enabling():
  p = Pair(4)
  with pytest.raises(Exception):
    ops.kv_variable_full_or_delta_export(p.gpu, first_n=6)
