"""The Python mirror (get_kv_variable / embedding_lookup / optimizers) on the GPU, written the
way the reference's own py_ut tests are: same scenarios, same expected numbers."""
import numpy as np
import pytest
import torch

import tfplus_b200 as tfp
from oracle import binding as ob
from tfplus_b200 import ops
from tfplus_b200.training import (AdagradOptimizer, AdamOptimizer, GradientDescentOptimizer,
                                  GroupAdamOptimizer, SparseGroupFtrlOptimizer)

from kvtest_util import DEV, TODAY

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fresh():
  ops.set_today(TODAY)
  tfp.set_training(True)
  yield
  tfp.reset_kv_variable_store()
  ops.set_today(None)


def _const_weights(num_shards, dim=64, enter_threshold=0, name="kv_embedding"):
  # py_ut/tests/test_embedding_ops.py:88-141
  w = tfp.get_kv_variable(name, embedding_dim=dim, initializer=tfp.ones_initializer,
                          partitioner=tfp.fixed_size_partitioner(num_shards),
                          enter_threshold=enter_threshold, device=DEV)
  parts = list(w) if isinstance(w, tfp.PartitionedKvVariable) else [w]
  def scatter():
    for i, p in enumerate(parts):
      keys = [k for k in range(100) if k % num_shards == i]
      vals = torch.tensor([[float(k)] * dim for k in keys])
      p.scatter_update(tfp.IndexedSlices(vals, torch.tensor(keys)))
  return w, scatter


def test_embedding_lookup_one_and_many_shards():
  # py_ut/tests/test_embedding_ops.py:160-212
  no_shard, scatter1 = _const_weights(1, name="a")
  shards, scatter2 = _const_weights(10, name="b")
  ids = torch.arange(100)
  r1, r2 = tfp.embedding_lookup(no_shard, ids), tfp.embedding_lookup(shards, ids)
  assert tuple(r1.shape) == (100, 64) and (r1 == 1.0).all() and (r2 == 1.0).all()
  scatter1(); scatter2()
  want = torch.tensor([[float(i)] * 64 for i in range(100)], device=DEV)
  assert torch.equal(tfp.embedding_lookup(no_shard, ids), want)
  assert torch.equal(tfp.embedding_lookup(shards, ids), want)


def test_embedding_lookup_sparse_sum_mean():
  # py_ut/tests/test_embedding_ops.py:259-298: one row holding ids 0..99
  params, scatter = _const_weights(10)
  sp = (torch.stack([torch.zeros(100, dtype=torch.int64), torch.arange(100)], 1),
        torch.arange(100), (1, 100))
  e1 = tfp.embedding_lookup_sparse(params, sp, None, combiner="sum")
  e2 = tfp.embedding_lookup_sparse(params, sp, None, combiner="mean")
  assert tuple(e1.shape) == (1, 64) and (e1 == 100.0).all() and (e2 == 1.0).all()
  scatter()
  e1 = tfp.embedding_lookup_sparse(params, sp, None, combiner="sum")
  e2 = tfp.embedding_lookup_sparse(params, sp, None, combiner="mean")
  assert (e1 == 4950.0).all() and (e2 == 49.5).all()


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("dim", [1, 16, 300])
def test_sparse_combine_kernel_is_the_sequential_reduction(combiner, weighted, dim):
  # embedding_ops.py:403-441: segment_sum of (weighted) rows, / segment_sum(w) or / sqrt(segment_sum(w^2));
  # entries added in increasing position, rows without entries zero
  rng = np.random.default_rng(dim)
  n_rows, U = 37, 50
  seg = np.sort(rng.integers(0, n_rows, size=400)).astype(np.int64)
  seg = seg[seg != 5]                                  # an empty row
  idx = rng.integers(0, U, size=seg.size).astype(np.int32)
  emb = rng.normal(size=(U, dim)).astype(np.float32)
  w = rng.random(seg.size).astype(np.float32) + 0.5 if weighted else None
  got = ops.sparse_combine(torch.from_numpy(emb).cuda(), torch.from_numpy(idx).cuda(),
                           torch.from_numpy(seg).cuda(),
                           None if w is None else torch.from_numpy(w).cuda(), n_rows,
                           combiner).cpu().numpy()
  want = np.zeros((n_rows, dim), np.float32)
  den = np.zeros(n_rows, np.float32)
  for i in range(seg.size):
    wi = np.float32(1.0) if w is None else w[i]
    want[seg[i]] += emb[idx[i]] * wi if w is not None else emb[idx[i]]
    den[seg[i]] += wi * wi if combiner == "sqrtn" else wi
  if combiner == "sqrtn":
    den = np.sqrt(den)
  if combiner != "sum":
    nz = den > 0
    want[nz] = want[nz] / den[nz, None]
  np.testing.assert_array_equal(got, want)
  assert not got[5].any()


def test_safe_embedding_lookup_sparse_negative_id_is_valid():
  # py_ut/tests/test_embedding_ops.py:301-337
  params, _ = _const_weights(2, dim=8)
  sp = (torch.tensor([[0, 0], [0, 1], [2, 0]]), torch.tensor([-1, 5, -1]), (4, 2))
  out = tfp.safe_embedding_lookup_sparse(params, sp, combiner="sum", default_id=None)
  assert out[:, 0].tolist() == [2.0, 0.0, 1.0, 0.0]
  out = tfp.safe_embedding_lookup_sparse(params, sp, combiner="sum", default_id=7)
  assert out[:, 0].tolist() == [2.0, 1.0, 1.0, 1.0]


def test_predict_mode_reads_zeros_and_counts():
  # py_ut/tests/test_kv_variable_ops.py:234-268 through the variable API
  v = tfp.get_kv_variable("v", embedding_dim=8, initializer=tfp.ones_initializer,
                          enter_threshold=2, device=DEV)
  tfp.set_training(False)
  assert not v.sparse_read(torch.arange(5)).any()
  tfp.set_training(True)
  assert (v.sparse_read(torch.arange(5)) == 1).all()
  assert v.total_count() == 0 and v.total_freq() == 0          # freq 1 < enter_threshold 2
  v.sparse_read_with_counts(torch.arange(5), torch.full((5,), 3, dtype=torch.int32))
  assert v.total_count() == 5 and v.total_freq() == 20
  assert v.get_counting(torch.tensor([0, 9])).tolist() == [4, 0]
  assert v.shape == [5, 8]


def _data(h=10, w=64, seed=0):
  g = np.random.default_rng(seed).random((h, w), dtype=np.float32)   # test_training_ops.py:207-236
  return torch.from_numpy(g).to(DEV), torch.arange(h, device=DEV)


def test_adam_bitwise_equals_dense_adam():
  # py_ut/tests/test_training_ops.py:395-416: tfplus Adam on a KvVariable == on a dense variable
  g, ids = _data()
  for fused in (True, False):
    kv = tfp.get_kv_variable("kv%d" % fused, embedding_dim=64, initializer=tfp.ones_initializer,
                             device=DEV)
    kv.sparse_read(ids)
    opt = AdamOptimizer(1e-3, fused=fused)
    dense, m, v = torch.ones(10, 64, device=DEV), torch.zeros(10, 64, device=DEV), torch.zeros(10, 64, device=DEV)
    b1p, b2p = torch.tensor(0.9), torch.tensor(0.999)
    c = lambda x: torch.tensor(x, dtype=torch.float32, device=DEV)
    for _ in range(3):
      opt.apply_gradients([(tfp.IndexedSlices(g, ids), kv)])
      m = c(0.9) * m + g * (c(1.0) - c(0.9))
      v = c(0.999) * v + (g * g) * (c(1.0) - c(0.999))
      lr = c(1e-3) * torch.sqrt(c(1.0) - b2p.to(DEV)) / (c(1.0) - b1p.to(DEV))
      dense = dense - lr * m / (c(1e-8) + torch.sqrt(v))
      b1p, b2p = b1p * torch.tensor(0.9), b2p * torch.tensor(0.999)
    tfp.set_training(False)
    assert torch.equal(kv.sparse_read(ids), dense)
    tfp.set_training(True)


@pytest.mark.parametrize("dim", [64, 1])
def test_group_adam_zero_reg_equals_adam(dim):
  # py_ut/tests/test_training_ops.py:437-473, atol 1e-8 in the reference (one step)
  g, ids = _data(w=dim)
  kv = tfp.get_kv_variable("kv", embedding_dim=dim, initializer=tfp.ones_initializer, device=DEV)
  kv.sparse_read(ids)
  GroupAdamOptimizer(0.1).apply_gradients([(tfp.IndexedSlices(g, ids), kv)])
  m, v = g * 0.1, g * g * np.float32(1 - np.float32(0.999))
  lr = np.float32(0.1) * np.sqrt(np.float32(1) - np.float32(0.999)) / (np.float32(1) - np.float32(0.9))
  want = 1.0 - lr * (g * np.float32(1 - np.float32(0.9))) / (torch.sqrt(v) + 1e-8)
  tfp.set_training(False)
  torch.testing.assert_close(kv.sparse_read(ids), want, rtol=0, atol=2e-7)


def test_adagrad_equals_tf_adagrad():
  # py_ut/tests/test_training_ops.py:418-435
  g, ids = _data()
  kv = tfp.get_kv_variable("kv", embedding_dim=64, initializer=tfp.ones_initializer, device=DEV)
  kv.sparse_read(ids)
  AdagradOptimizer(0.1, initial_accumulator_value=0.1).apply_gradients([(tfp.IndexedSlices(g, ids), kv)])
  want = 1.0 - 0.1 * g / torch.sqrt(0.1 + g * g)
  tfp.set_training(False)
  torch.testing.assert_close(kv.sparse_read(ids), want, rtol=0, atol=2e-7)


def test_sparse_group_ftrl_differs_from_ftrl():
  # py_ut/tests/test_training_ops.py:475-543 asserts only inequality
  g, ids = _data()
  outs = []
  for i, (l1, l2, l21) in enumerate([(0.0, 0.0, 0.0), (0.01, 0.05, 0.05)]):
    kv = tfp.get_kv_variable("kv%d" % i, embedding_dim=64, initializer=tfp.ones_initializer, device=DEV)
    kv.sparse_read(ids)
    SparseGroupFtrlOptimizer(0.1, l1_regularization_strength=l1, l2_regularization_strength=l2,
                             l21_regularization_strength=l21).apply_gradients(
                                 [(tfp.IndexedSlices(g, ids), kv)])
    outs.append(kv.read_value()[1])
  assert not torch.equal(outs[0], outs[1])


def test_optimizer_dedups_like_tf_and_matches_oracle():
  # duplicates in the gradient: TF's _deduplicate_indexed_slices path, several steps vs oracle
  dim = 16
  rng = np.random.default_rng(5)
  kv = tfp.get_kv_variable("kv", embedding_dim=dim, initializer=0.25, enter_threshold=2, device=DEV)
  opt = GroupAdamOptimizer(0.01, l1_regularization_strength=1e-5, l2_regularization_strength=1e-5,
                           l21_regularization_strength=1e-5)
  o_var = ob.OracleTable(dim, 2, seed=1); o_var.set_init_table(np.full((4, dim), 0.25, np.float32))
  o_slot = ob.OracleTable(3 * dim, 0, seed=1); o_slot.set_init_table(np.zeros((4, 3 * dim), np.float32))
  b1p, b2p = np.float32(0.9), np.float32(0.999)
  for _ in range(4):
    ids = (rng.zipf(1.3, size=2000) % 300).astype(np.int64)
    g = rng.integers(-4, 5, size=(2000, dim)).astype(np.float32) / 8   # exact sums in any order
    rows = tfp.embedding_lookup(kv, torch.from_numpy(ids))
    want = o_var.gather_or_insert(ids, today=TODAY)
    np.testing.assert_allclose(rows.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
    opt.apply_gradients([(tfp.IndexedSlices(torch.from_numpy(g), torch.from_numpy(ids)), kv)])
    u, idx = ob.unique(ids)
    ob.apply_group_adam_v4(o_var, o_slot, u, ob.segment_sum(g, idx, u.size), 0.01, float(b1p),
                           float(b2p), 0.9, 0.999, 1e-8, 1e-5, 1e-5, 1e-5, today=TODAY)
    b1p, b2p = b1p * np.float32(0.9), b2p * np.float32(0.999)
  keys, vals = kv.read_value()
  e = o_var.export(first_n=2, enable_cutoff=False, cutoff_value=0.0)
  got = {int(k): v for k, v in zip(keys.cpu().numpy(), vals.cpu().numpy())}
  assert set(got) == set(int(k) for k in e["keys"])
  for k, v in zip(e["keys"], e["values"]):
    np.testing.assert_allclose(got[int(k)], v, rtol=1e-6, atol=1e-7)


def test_sgd_and_checkpoint_round_trip():
  kv = tfp.get_kv_variable("kv", embedding_dim=8, initializer=tfp.ones_initializer, device=DEV)
  ids = torch.arange(50)
  kv.sparse_read(ids)
  GradientDescentOptimizer(0.5).apply_gradients(
      [(tfp.IndexedSlices(torch.ones(50, 8), ids), kv)])
  tensors = kv.export()
  assert len(tensors) == 6 and tensors[0].numel() == 50
  kv2 = tfp.get_kv_variable("kv2", embedding_dim=8, initializer=tfp.zeros_initializer, device=DEV)
  kv2.restore(tensors)
  tfp.set_training(False)
  assert (kv2.sparse_read(ids) == 0.5).all()
  kv2.delete(torch.arange(10))
  assert kv2.total_count() == 40


def test_ncf_and_dcn_shaped_steps():
  # BASELINE configs 1 and 3 in miniature: NCF (two tables, dim 32, tfplus Adam, ids deduped in
  # the model) and DCN (26 tables, dim 16, SparseGroupFtrl, duplicates reach the gather)
  rng = np.random.default_rng(0)
  user = tfp.get_kv_variable("user", embedding_dim=32, initializer=tfp.random_normal_initializer(), device=DEV)
  item = tfp.get_kv_variable("item", embedding_dim=32, initializer=tfp.random_normal_initializer(), device=DEV)
  adam = AdamOptimizer(1e-3)
  for _ in range(3):
    u = torch.from_numpy(rng.integers(1, 944, size=256))
    i = torch.from_numpy(rng.integers(1, 1683, size=256))
    ue, ie = tfp.embedding_lookup(user, u), tfp.embedding_lookup(item, i)
    loss_g = (ue * ie).sum(1, keepdim=True).sigmoid() - 0.5
    adam.apply_gradients([(tfp.IndexedSlices(loss_g * ie, u), user),
                          (tfp.IndexedSlices(loss_g * ue, i), item)])
  assert 0 < user.total_count() <= 943 and 0 < item.total_count() <= 1682
  ftrl = SparseGroupFtrlOptimizer(0.1, l1_regularization_strength=1e-5,
                                  l2_regularization_strength=1e-5, l21_regularization_strength=1e-5)
  fields = [tfp.get_kv_variable("C%d" % f, embedding_dim=16, initializer=0.01, device=DEV) for f in range(26)]
  for _ in range(2):
    pairs = []
    for f, var in enumerate(fields):
      ids = torch.from_numpy((rng.zipf(1.05, size=8192) % (1000 + 50 * f)).astype(np.int64))
      rows = tfp.embedding_lookup(var, ids)
      pairs.append((tfp.IndexedSlices(torch.tanh(rows) * 0.1, ids), var))
    ftrl.apply_gradients(pairs)
  assert all(v.total_count() > 0 for v in fields)
  assert torch.isfinite(fields[0].read_value()[1]).all()
