"""Pins the CPU oracle on the reference's own known-answer tests.

Every test names the reference test it restates (paths relative to
/root/reference).  The reference cannot run here (no TensorFlow), so these
KATs plus the TF-optimizer equivalences below are what "the oracle is right"
rests on; areas with no reference assertion are listed as PARITY UNPINNED in
oracle/kv_oracle.cc.
"""
import numpy as np
import pytest

from oracle import binding as ob
from oracle.binding import OracleTable

D = 8
R = 1024


def _table(dim=D, thr=0, init=None, seed=0):
  t = OracleTable(dim, thr, seed)
  if init is None:
    init = np.random.default_rng(0).normal(size=(R, dim)).astype(np.float32)
  t.set_init_table(init)
  return t


def test_shape_empty():
  # py_ut/tests/test_kv_variable_ops.py:78-96
  t = OracleTable(D, 0)
  assert t.shape() == [0, D]


def test_is_initialized():
  # py_ut/tests/test_kv_variable_ops.py:98-123
  t = OracleTable(D, 0)
  assert not t.is_initialized()
  t.set_init_table(np.ones((R, D), np.float32))
  assert t.is_initialized()


def test_size_after_init():
  # py_ut/tests/test_kv_variable_ops.py:125-148
  assert _table().size() == 0


def test_frequency_kat():
  # py_ut/tests/test_kv_variable_ops.py:150-189, enter_threshold=2
  t = _table(thr=2)
  t.gather_or_insert([0, 1, 2, 3, 4])
  assert t.sum_freq() == 0
  t.gather_or_insert([2, 3, 4, 5, 6])
  assert t.sum_freq() == 6
  t.gather_or_zeros([0, 1, 2, 3, 4])
  assert t.sum_freq() == 6


def test_gather_zeros_vs_ones():
  # py_ut/tests/test_kv_variable_ops.py:234-268: predict gather of unknown
  # keys = zeros, train gather with a ones init table = ones
  t = _table(init=np.ones((R, D), np.float32))
  ids = [0, 1, 2, 3, 4]
  assert np.array_equal(t.gather_or_zeros(ids), np.zeros((5, D), np.float32))
  assert np.array_equal(t.gather_or_insert(ids), np.ones((5, D), np.float32))
  assert np.array_equal(t.gather_or_zeros(ids), np.ones((5, D), np.float32))


def test_random_init_range():
  # py_ut/tests/test_kv_variable_ops.py:270-308: uniform [0.01,1] init table
  init = np.random.default_rng(1).uniform(0.01, 1, size=(R, D)).astype(np.float32)
  t = _table(init=init)
  out = t.gather_or_insert(np.arange(100))
  assert (out >= 0.01).all() and (out <= 1.0).all()


@pytest.mark.parametrize("first_n,exp", [(3, (0, 5)), (4, (1, 6)), (6, (1, 6))])
def test_import_export_shape_kat(first_n, exp):
  # py_ut/tests/test_kv_variable_ops.py:345-435.  The import op drops the
  # blacklist for first_n<=3 and the frequency table for first_n<=4
  # (kernels/kv_variable_ops.cc:806-822); the export op runs with
  # first_n=6 and its attr defaults enable_cutoff=false, cutoff_value=0.0
  # (ops/kv_variable_ops.cc:430-432).
  t = _table(thr=1)
  init = np.random.default_rng(2).normal(size=(R, D)).astype(np.float32)
  ids = np.arange(5)
  vals = np.stack([np.full(D, float(x), np.float32) for x in range(5)])
  t.import_(ids, vals, init_table=init,
            blacklist=[7] if first_n > 3 else None,
            freq_keys=[1, 2, 3, 4, 5] if first_n > 4 else None,
            freq_values=[1, 2, 3, 4, 5] if first_n > 4 else None)
  out = t.export(first_n=6, enable_cutoff=False, cutoff_value=0.0)
  assert len(out) == 6
  n_black, n_freq = exp
  assert out["keys"].shape == (5,)
  assert out["values"].shape == (5, D)
  assert out["init_table"].shape == (R, D)
  assert out["blacklist"].shape == (n_black,)
  assert out["freq_keys"].shape == (n_freq,)
  assert out["freq_values"].shape == (n_freq,)
  assert out["freq_values"].dtype == np.uint16


def test_find_or_zeros_unknown_key():
  # kernels/kv_variable_test.cc:121-129
  t = _table(dim=64)
  assert not t.gather_or_zeros(np.arange(10)).any()


def test_find_or_insert_in_init_range():
  # kernels/kv_variable_test.cc:132-140
  init = np.random.default_rng(3).uniform(-1, 1, size=(R, 64)).astype(np.float32)
  t = _table(dim=64, init=init)
  out = t.gather_or_insert(np.arange(10))
  assert (out >= init.min()).all() and (out <= init.max()).all()


def test_insert_or_update_size():
  # kernels/kv_variable_test.cc:183-201
  t = _table(dim=64)
  t.insert_or_update(np.arange(10), np.ones((10, 64), np.float32))
  assert t.size() == 10
  mask = np.zeros(10, np.uint8)
  mask[:5] = 1
  t2 = _table(dim=64)
  t2.insert_or_update(np.arange(10), np.ones((10, 64), np.float32), filter_out=mask)
  assert t2.size() == 5


def test_scatter_exact_answers():
  # kernels/kv_variable_test.cc:272-356: exact 1.0 / 2.0 answers
  t = _table(dim=64)
  ids = np.arange(10)
  one = np.ones((10, 64), np.float32)
  two = np.full((10, 64), 2.0, np.float32)
  for op, upd, want in [("assign", one, 1.0), ("add", one, 2.0), ("sub", one, 1.0),
                        ("mul", two, 2.0), ("div", two, 1.0), ("min", two, 1.0),
                        ("max", two, 2.0)]:
    t.scatter(op, ids, upd)
    assert np.array_equal(t.gather_or_zeros(ids), np.full((10, 64), want, np.float32)), op


def test_freq_word_packing():
  # kernels/kv_variable_test.cc:359-382: lo16 = frequency, hi16 = day
  t = _table()
  t.gather_or_insert([42], counts=[65535], today=65534)
  w = t.freq_word(42)
  assert w & 0xFFFF == 65535 and w >> 16 == 65534
  assert w == (65534 << 16) | 65535
  t.gather_or_insert([42], counts=[7], today=65534)  # saturates, utility.h:65-70
  assert t.freq_word(42) & 0xFFFF == 65535


def test_delete():
  # kernels/kv_variable_test.cc:439-449
  t = _table(dim=64)
  t.gather_or_insert(np.arange(10))
  assert t.size() == 10
  t.delete(np.arange(5))
  assert t.size() == 5


# --- optimizer equivalences the reference asserts against TF's own optimizers --
def _np_adam_step(var, m, v, g, lr, b1, b2, eps, b1p, b2p):
  """TF 2.13 Adam sparse apply for unique indices (training/adam.py)."""
  f = np.float32
  lr_t = f(lr) * np.sqrt(f(1) - f(b2p)) / (f(1) - f(b1p))
  m[:] = m * f(b1) + g * (f(1) - f(b1))
  v[:] = v * f(b2) + (g * g) * (f(1) - f(b2))
  var -= lr_t * m / (np.sqrt(v) + f(eps))


@pytest.mark.parametrize("dim", [64, 1])
def test_group_adam_v4_zero_reg_equals_adam(dim):
  # py_ut/tests/test_training_ops.py:437-473: GroupAdam v4, l1=l2=l21=0,
  # one step == TF Adam, atol 1e-8, D=64 and D=1
  h = 10
  rng = np.random.default_rng(4)
  g = rng.random((h, dim), dtype=np.float32)
  var = _table(dim=dim, init=np.ones((R, dim), np.float32))
  slot = _table(dim=3 * dim, init=np.zeros((R, 3 * dim), np.float32))
  ids = np.arange(h)
  var.gather_or_insert(ids)
  lr, b1, b2, eps = 0.1, 0.9, 0.999, 1e-8
  ob.apply_group_adam_v4(var, slot, ids, g, lr, b1, b2, b1, b2, eps, 0., 0., 0.)
  ref = np.ones((h, dim), np.float32)
  m = np.zeros_like(ref)
  v = np.zeros_like(ref)
  _np_adam_step(ref, m, v, g, lr, b1, b2, eps, b1, b2)
  np.testing.assert_allclose(var.gather_or_zeros(ids), ref, rtol=0, atol=1e-7)


def test_adagrad_equals_tf_adagrad():
  # py_ut/tests/test_training_ops.py:418-435: atol 1e-8 vs TF Adagrad
  h, dim = 10, 64
  g = np.random.default_rng(5).random((h, dim), dtype=np.float32)
  var = _table(dim=dim, init=np.ones((R, dim), np.float32))
  acc = _table(dim=dim, init=np.full((R, dim), 0.1, np.float32))
  ids = np.arange(h)
  lr = np.float32(0.1)
  ob.apply_adagrad(var, acc, ids, g, lr)
  a = np.float32(0.1) + g * g
  ref = np.float32(1.0) - lr * g / np.sqrt(a)
  np.testing.assert_allclose(var.gather_or_zeros(ids), ref, rtol=0, atol=2e-7)
  np.testing.assert_array_equal(acc.gather_or_zeros(ids), a)


def _np_ftrl_v2(var, accum, linear, g, lr, l1, l2, l2s, lr_power):
  """TF ResourceSparseApplyFtrlV2 for unique indices (core/kernels/training_ops.cc)."""
  f = np.float32
  gs = g + f(2) * f(l2s) * var
  new_accum = accum + g * g
  p = -f(lr_power)
  linear += gs - (new_accum**p - accum**p) / f(lr) * var
  quad = new_accum**p / f(lr) + f(2) * f(l2)
  var[:] = np.where(np.abs(linear) > f(l1),
                    (np.sign(linear) * f(l1) - linear) / quad, f(0))
  accum[:] = new_accum


def test_sparse_group_ftrl_l21_zero_equals_ftrl():
  # py_ut/tests/test_training_ops.py:68-205 pins KvVariableSparseApplyFtrlV2 ==
  # TF ftrl_v2 at atol 1e-8 (300x64 N(0,1) grads, var 0.03, accum 0.1,
  # linear 0.0, lr 0.01); with l21 = 0 the group op computes the same update.
  n, dim = 300, 64
  g = np.random.default_rng(6).normal(size=(n, dim)).astype(np.float32)
  var = _table(dim=dim, init=np.full((R, dim), 0.03, np.float32))
  acc = _table(dim=dim, init=np.full((R, dim), 0.1, np.float32))
  lin = _table(dim=dim, init=np.zeros((R, dim), np.float32))
  ids = np.arange(n)
  lr = 0.01
  ob.apply_sparse_group_ftrl(var, acc, lin, ids, g, lr, 0., 0., 0., 0., -0.5)
  rv = np.full((n, dim), 0.03, np.float32)
  ra = np.full((n, dim), 0.1, np.float32)
  rl = np.zeros((n, dim), np.float32)
  _np_ftrl_v2(rv, ra, rl, g, lr, 0., 0., 0., -0.5)
  np.testing.assert_allclose(var.gather_or_zeros(ids), rv, rtol=0, atol=1e-6)
  np.testing.assert_allclose(acc.gather_or_zeros(ids), ra, rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(lin.gather_or_zeros(ids), rl, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("l1,l2", [(0.0, 0.0), (0.05, 0.1)])
def test_sparse_apply_ftrl_v2_equals_tf_ftrl(l1, l2):
  # py_ut/tests/test_training_ops.py:68-205: KvVariableSparseApplyFtrlV2 == TF ftrl_v2 at atol
  # 1e-8, three steps of 300x64 N(0,1) grads on var 0.03, accum 0.1, linear 0.0, lr 0.01
  # (l2_shrinkage 0: the two formulas only differ in what they square into accum otherwise)
  n, dim = 300, 64
  rng = np.random.default_rng(16)
  var = _table(dim=dim, init=np.full((R, dim), 0.03, np.float32))
  acc = _table(dim=dim, init=np.full((R, dim), 0.1, np.float32))
  lin = _table(dim=dim, init=np.zeros((R, dim), np.float32))
  ids = np.arange(n)
  rv = np.full((n, dim), 0.03, np.float32)
  ra = np.full((n, dim), 0.1, np.float32)
  rl = np.zeros((n, dim), np.float32)
  for _ in range(3):
    g = rng.normal(size=(n, dim)).astype(np.float32)
    ob.apply_sparse_ftrl_v2(var, acc, lin, ids, g, 0.01, l1, l2, 0.0, -0.5)
    _np_ftrl_v2(rv, ra, rl, g, 0.01, l1, l2, 0.0, -0.5)
  np.testing.assert_allclose(var.gather_or_zeros(ids), rv, rtol=2e-6, atol=1e-6)
  np.testing.assert_allclose(acc.gather_or_zeros(ids), ra, rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(lin.gather_or_zeros(ids), rl, rtol=2e-6, atol=2e-6)
  # no blacklist, no filter side effects: every key is still there
  assert var.size() == n


def test_group_adam_v3_zero_reg_first_step_is_adam_shaped():
  # V3 (training_ops.cc:5893-5925) with l1=l2=l21=0 on a fresh slot: linear = alpha*m -
  # (sqrt(nv)+eps)/lr*var and var = -linear / ((sqrt(nv)+eps)/lr)  =>  var' = var - lr*alpha*m /
  # (sqrt(nv)+eps): Adam's first step with alpha = sqrt(1-b2^t)/(1-b1^t)
  h, dim = 10, 64
  g = np.random.default_rng(17).random((h, dim), dtype=np.float32)
  var = _table(dim=dim, init=np.ones((R, dim), np.float32))
  slot = _table(dim=3 * dim, init=np.zeros((R, 3 * dim), np.float32))
  ids = np.arange(h)
  var.gather_or_insert(ids)
  lr, b1, b2, eps = 0.1, 0.9, 0.999, 1e-8
  ob.apply_group_adam_v3(var, slot, ids, g, lr, b1, b2, b1, b2, eps, 0., 0., 0.)
  ref = np.ones((h, dim), np.float32)
  m = np.zeros_like(ref)
  v = np.zeros_like(ref)
  _np_adam_step(ref, m, v, g, lr, b1, b2, eps, b1, b2)
  np.testing.assert_allclose(var.gather_or_zeros(ids), ref, rtol=0, atol=2e-6)
  got = slot.gather_or_zeros(ids)
  np.testing.assert_allclose(got[:, :dim], m, rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(got[:, dim:2 * dim], v, rtol=1e-6, atol=1e-9)


def test_group_sparse_ftrl_v2_thresholds_on_the_norm_of_linear():
  # training_ops.cc:985-1008: ||linear|| <= l1 blacklists the key, otherwise
  # var = (l1 - ||linear||) / ((sqrt(na)/lr + 2*l2) * ||linear||) * linear; accum gets g^2 twice
  dim = 8
  var = _table(dim=dim, init=np.full((R, dim), 0.03, np.float32))
  acc = _table(dim=dim, init=np.full((R, dim), 0.1, np.float32))
  lin = _table(dim=dim, init=np.zeros((R, dim), np.float32))
  ids = np.arange(2)
  g = np.stack([np.full(dim, 1e-4, np.float32), np.full(dim, 2.0, np.float32)])
  ob.apply_group_sparse_ftrl_v2(var, acc, lin, ids, g, 0.1, 0.5, 0.01, 0.0, -0.5)
  f = np.float32
  assert np.array_equal(var.gather_or_zeros(ids)[0], np.zeros(dim, f))   # blacklisted: reads zeros
  assert var.size() == 1
  na = f(0.1) + g[1] * g[1]
  linear = g[1] - (np.sqrt(na) - np.sqrt(f(0.1))) / f(0.1) * f(0.03)
  nrm = np.sqrt(np.sum(linear * linear, dtype=np.float32))
  want = (f(0.5) - nrm) / ((np.sqrt(na) / f(0.1) + f(0.02)) * nrm) * linear
  np.testing.assert_allclose(var.gather_or_zeros(ids)[1], want, rtol=2e-6)
  np.testing.assert_allclose(acc.gather_or_zeros(ids)[1], f(0.1) + 2 * g[1] * g[1], rtol=1e-6)


def test_sparse_group_ftrl_differs_with_regularisers():
  # py_ut/tests/test_training_ops.py:475-543 only asserts inequality
  n, dim = 50, 64
  g = np.random.default_rng(7).normal(size=(n, dim)).astype(np.float32)
  outs = []
  for l1, l2, l21 in [(0., 0., 0.), (0.01, 0.05, 0.05)]:
    var = _table(dim=dim, init=np.full((R, dim), 0.03, np.float32))
    acc = _table(dim=dim, init=np.full((R, dim), 0.1, np.float32))
    lin = _table(dim=dim, init=np.zeros((R, dim), np.float32))
    ob.apply_sparse_group_ftrl(var, acc, lin, np.arange(n), g, 0.01, l1, l2, l21)
    outs.append(var.gather_or_zeros(np.arange(n)))
  assert not np.array_equal(outs[0], outs[1])


def test_tfplus_adam_equals_dense_adam():
  # py_ut/tests/test_training_ops.py:395-416: tfplus Adam on a KvVariable is
  # bitwise the same update as on a dense variable (same op sequence)
  h, dim = 10, 64
  g = np.random.default_rng(8).random((h, dim), dtype=np.float32)
  var = _table(dim=dim, init=np.ones((R, dim), np.float32))
  mv = _table(dim=2 * dim, init=np.zeros((R, 2 * dim), np.float32))
  ids = np.arange(h)
  var.gather_or_insert(ids)
  lr, b1, b2, eps = 0.001, 0.9, 0.999, 1e-8
  ob.adam_step(var, mv, ids, g, lr, b1, b2, eps, b1, b2)
  f = np.float32
  m = g * (f(1) - f(b1))
  v = (g * g) * (f(1) - f(b2))
  lr_t = (f(lr) * np.sqrt(f(1) - f(b2))) / (f(1) - f(b1))
  ref = np.ones((h, dim), f) - (lr_t * m) / (np.sqrt(v) + f(eps))
  np.testing.assert_array_equal(var.gather_or_zeros(ids), ref)
  np.testing.assert_array_equal(mv.gather_or_zeros(ids), np.concatenate([m, v], 1))


# --- delta checkpoints (dynamic_save.hpp:197-449, dynamic_restore.hpp:28-153) ---
def test_delta_export_classes_and_list_handling():
  dim = 4
  tb = ob.OracleTable(dim, 2, seed=3)          # enter_threshold 2
  tb.set_init_table(np.ones((R, dim), np.float32))
  tb.enable_delta_export(support_prediction_delta=True)
  tb.gather_or_insert(np.array([1, 2, 2, 3, 3], np.int64))        # 1 stays low-frequency
  tb.insert_or_update(np.array([3], np.int64), np.zeros((1, dim), np.float32),
                      blacklist=np.array([1], np.uint8))          # 3 blacklisted
  tb.delete(np.array([9], np.int64))                              # absent -> delete_keys
  assert tb.delta_size() == 4
  d = tb.delta_export(first_n=6)
  assert d["keys"].tolist() == [2] and np.array_equal(d["values"], np.ones((1, dim), np.float32))
  assert d["blacklist"].tolist() == [3] and d["delete_keys"].tolist() == [9]
  assert sorted(d["freq_keys"].tolist()) == [1, 2, 3, 9]           # all_delta, 0 for the absent key
  fw = dict(zip(d["freq_keys"].tolist(), d["freq_values"].tolist()))
  assert fw[9] == 0 and (fw[2] & 0xffff) == 2 and (fw[1] & 0xffff) == 1
  assert tb.delta_size() == 0
  # training-mode export handed its keys to the prediction list: the inference-mode export
  # sees them once, with blacklisted keys turned into deletions (:333-339), and clears the list
  p = tb.delta_export(first_n=3)
  assert p["keys"].tolist() == [2] and p["blacklist"].size == 0
  assert sorted(p["delete_keys"].tolist()) == [3, 9] and p["freq_keys"].size == 0
  assert tb.delta_export(first_n=3)["keys"].size == 0
  # import on a replica
  rep = ob.OracleTable(dim, 2, seed=3)
  rep.set_init_table(np.ones((R, dim), np.float32))
  rep.gather_or_insert(np.array([2, 9], np.int64))
  rep.delta_import(d["keys"], d["values"], d["blacklist"], d["freq_keys"], d["freq_values"],
                   d["delete_keys"], first_n=6)
  assert rep.map_size() == 2                                       # 2 and the row-less 3; 9 deleted
  assert rep.freq_word(2) == fw[2]
  assert np.array_equal(rep.gather_or_zeros(np.array([3, 9])), np.zeros((2, dim), np.float32))


# --- TF dedup semantics (TensorFlow 2.13 Unique / UnsortedSegmentSum) ----------
def test_unique_first_occurrence_order():
  ids = np.array([7, 3, 7, 9, 3, 3, -1, 9], np.int64)
  u, idx, c = ob.unique(ids, with_counts=True)
  assert u.tolist() == [7, 3, 9, -1]
  assert idx.tolist() == [0, 1, 0, 2, 1, 1, 3, 2]
  assert c.tolist() == [2, 3, 2, 1]
  assert idx.dtype == np.int32
  u0, idx0 = ob.unique(np.array([], np.int64))
  assert u0.size == 0 and idx0.size == 0


def test_segment_sum():
  data = np.arange(12, dtype=np.float32).reshape(4, 3)
  out = ob.segment_sum(data, [1, 0, 1, 1], 3)
  assert np.array_equal(out, [[3, 4, 5], [0 + 6 + 9, 1 + 7 + 10, 2 + 8 + 11], [0, 0, 0]])
