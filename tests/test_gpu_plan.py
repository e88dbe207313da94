"""GPU parity of the plan-driven path (kv_plan_build -> kv_gather_or_insert_plan ->
kv_apply_plan) against the CPU oracle running the reference's op sequence on the raw ids:
GatherOrInsert, tf.unique, tf.unsorted_segment_sum, the apply op.

Bit-exact: uniq / idx / counts, the occurrence lists, gathered rows, frequency words, flags,
AND the duplicate-gradient sums (the kernel adds in TF's CPU order).  Optimizer state: 1e-6."""
import numpy as np
import pytest
import torch

from oracle import binding as ob
from tfplus_b200 import ops

from kvtest_util import DEV, TODAY, Pair, t

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-6, 1e-7


@pytest.fixture(autouse=True)
def _clock():
  ops.set_today(TODAY)
  yield
  ops.set_today(None)


def zipf_ids(n, universe, seed, s=1.1):
  rng = np.random.default_rng(seed)
  return (rng.zipf(s, size=n) % universe).astype(np.int64)


def check_plan(plan, ids):
  uniq, idx, counts, num, seg_off, pos = [x.cpu().numpy() for x in plan.arrays()]
  ou, oi, oc = ob.unique(ids, with_counts=True)
  U = int(num[0])
  assert U == ou.size
  np.testing.assert_array_equal(uniq[:U], ou)
  np.testing.assert_array_equal(idx, oi)
  np.testing.assert_array_equal(counts[:U], oc)
  np.testing.assert_array_equal(seg_off[:U], np.concatenate([[0], np.cumsum(oc)[:-1]]))
  # occurrence lists: positions of every distinct id, increasing
  order = np.argsort(oi, kind="stable")
  np.testing.assert_array_equal(pos, order.astype(np.int32))
  return ou, oi, oc


@pytest.mark.parametrize("n,universe", [(1, 10), (5, 3), (33, 7), (1000, 1), (4097, 50),
                                        (65536, 10**7), (200001, 1000), (70000, 3)])
def test_plan_arrays_bit_exact(n, universe):
  ids = zipf_ids(n, universe, seed=n) if universe > 3 else (
      np.random.default_rng(n).integers(0, universe, size=n).astype(np.int64))
  plan = ops.Plan(n, DEV).build(t(ids))
  check_plan(plan, ids)
  # the buffers are reused by the next batch
  ids2 = np.roll(ids, 3) + 1
  plan.build(t(ids2))
  check_plan(plan, ids2)


@pytest.mark.parametrize("dim", [1, 3, 16, 64, 100, 128, 256])
def test_segment_sum_plan_is_tf_order_bit_exact(dim):
  n = 30000
  ids = zipf_ids(n, 5000, seed=dim)          # head id occurs thousands of times: heavy path
  data = np.random.default_rng(dim).normal(size=(n, dim)).astype(np.float32)
  plan = ops.Plan(n, DEV).build(t(ids))
  ou, oi, _ = check_plan(plan, ids)
  got = ops.segment_sum_plan(plan, t(data)).cpu().numpy()[:ou.size]
  want = ob.segment_sum(data, oi, ou.size)   # out[idx[i]] += data[i], increasing i
  np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("dim", [1, 8, 64, 192])
def test_gather_plan_equals_gather_on_raw_ids(dim):
  p = Pair(dim, enter_threshold=3)
  n = 20000
  plan = ops.Plan(n, DEV)
  for step in range(3):
    ids = zipf_ids(n, 6000 + 4000 * step, seed=10 * dim + step)
    plan.build(t(ids))
    got = ops.kv_variable_gather_or_insert_plan(p.gpu, plan).cpu().numpy()
    want = p.cpu.gather_or_insert(ids, today=TODAY)
    np.testing.assert_array_equal(got, want.reshape(got.shape))
    p.check_state()
  ids = np.concatenate([zipf_ids(500, 20000, seed=99), [-7, 2**41]]).astype(np.int64)
  plan.build(t(ids))
  got = ops.kv_variable_gather_or_zeros_plan(p.gpu, plan).cpu().numpy()
  np.testing.assert_array_equal(got, p.cpu.gather_or_zeros(ids).reshape(got.shape))
  p.check_state()


def test_gather_plan_saturates_counts():
  p = Pair(16)
  ids = np.concatenate([np.full(70000, 5), np.full(3, 9), [11]]).astype(np.int64)
  plan = ops.Plan(ids.size, DEV).build(t(ids))
  got = ops.kv_variable_gather_or_insert_plan(p.gpu, plan).cpu().numpy()
  want = p.cpu.gather_or_insert(ids, today=TODAY)
  np.testing.assert_array_equal(got, want.reshape(got.shape))
  p.check_state()


def _oracle_step(kind, tables, ids, grad, hp, today=TODAY):
  u, idx = ob.unique(ids)
  gs = ob.segment_sum(grad, idx, u.size)
  cpu = [x.cpu for x in tables]
  if kind == ops.OPT_GROUP_ADAM_V4:
    ob.apply_group_adam_v4(cpu[0], cpu[1], u, gs, *hp, today=today)
  elif kind == ops.OPT_ADAGRAD:
    ob.apply_adagrad(cpu[0], cpu[1], u, gs, hp[0], True, today=today)
  elif kind == ops.OPT_SPARSE_GROUP_FTRL:
    ob.apply_sparse_group_ftrl(cpu[0], cpu[1], cpu[2], u, gs, *hp, today=today)
  elif kind == ops.OPT_ADAM:
    ob.adam_step(cpu[0], cpu[1], u, gs, *hp, today=today)
  else:
    raise AssertionError(kind)


@pytest.mark.parametrize("dim,reg", [(64, 1e-5), (64, 0.0), (16, 1e-3), (8, 0.02), (256, 1e-5),
                                     (1, 0.0), (12, 1e-4)])
def test_apply_plan_group_adam(dim, reg):
  var = Pair(dim, enter_threshold=2)
  slot = Pair(3 * dim, init=0.0)
  n = 12000
  plan = ops.Plan(n, DEV)
  b1, b2, lr, eps = 0.9, 0.999, 1e-2, 1e-8
  b1p, b2p = b1, b2
  rng = np.random.default_rng(dim)
  for step in range(6):
    ids = zipf_ids(n, 3000 + 500 * step, seed=dim * 7 + step, s=1.15)
    grad = rng.normal(size=(n, dim)).astype(np.float32)
    plan.build(t(ids))
    rows = ops.kv_variable_gather_or_insert_plan(var.gpu, plan).cpu().numpy()
    want = var.cpu.gather_or_insert(ids, today=TODAY)
    np.testing.assert_allclose(rows, want.reshape(rows.shape), rtol=RTOL, atol=ATOL)
    hp = (lr, b1p, b2p, b1, b2, eps, reg, reg, reg)
    ops.kv_variable_apply_plan(ops.OPT_GROUP_ADAM_V4, var.gpu, slot.gpu, None, plan, t(grad), hp)
    _oracle_step(ops.OPT_GROUP_ADAM_V4, [var, slot], ids, grad, hp)
    b1p *= b1
    b2p *= b2
    var.check_state(rtol=RTOL, atol=ATOL)
    slot.check_state(rtol=RTOL, atol=ATOL)


def test_apply_plan_without_prior_lookup_and_with_device_hparams():
  """The apply must not depend on the lookup's hints: plan -> apply directly, new keys
  included, hyper-parameters on the device with the beta-power advance in the launch."""
  dim = 32
  var = Pair(dim)
  slot = Pair(3 * dim, init=0.0)
  n = 5000
  plan = ops.Plan(n, DEV)
  hp_host = [1e-2, 0.9, 0.999, 0.9, 0.999, 1e-8, 1e-5, 1e-5, 1e-5]
  hp_dev = torch.tensor(hp_host, dtype=torch.float32, device=DEV)
  rng = np.random.default_rng(3)
  b1p, b2p = np.float32(0.9), np.float32(0.999)
  for step in range(4):
    ids = zipf_ids(n, 2000, seed=step)
    grad = rng.normal(size=(n, dim)).astype(np.float32)
    plan.build(t(ids))
    ops.kv_variable_apply_plan(ops.OPT_GROUP_ADAM_V4, var.gpu, slot.gpu, None, plan, t(grad),
                               hp_dev, advance_powers=True)
    hp = (1e-2, float(b1p), float(b2p), 0.9, 0.999, 1e-8, 1e-5, 1e-5, 1e-5)
    _oracle_step(ops.OPT_GROUP_ADAM_V4, [var, slot], ids, grad, hp)
    b1p = np.float32(b1p * np.float32(0.9))
    b2p = np.float32(b2p * np.float32(0.999))
  var.check_state(rtol=RTOL, atol=ATOL)
  slot.check_state(rtol=RTOL, atol=ATOL)
  got = hp_dev.cpu().numpy()
  assert got[1] == b1p and got[2] == b2p


@pytest.mark.parametrize("kind", ["adagrad", "ftrl", "adam"])
def test_apply_plan_other_optimizers(kind):
  dim = 32
  n = 8000
  plan = ops.Plan(n, DEV)
  rng = np.random.default_rng(11)
  if kind == "adagrad":
    tabs = [Pair(dim, enter_threshold=2), Pair(dim, init=0.1)]
    k, hp = ops.OPT_ADAGRAD, (0.05,)
  elif kind == "ftrl":
    tabs = [Pair(dim, enter_threshold=1), Pair(dim, init=0.1), Pair(dim, init=0.0)]
    k, hp = ops.OPT_SPARSE_GROUP_FTRL, (0.1, 1e-4, 1e-4, 1e-4, 0.0, -0.5)
  else:
    tabs = [Pair(dim), Pair(2 * dim, init=0.0)]
    k, hp = ops.OPT_ADAM, None
  b1p, b2p = 0.9, 0.999
  for step in range(4):
    ids = zipf_ids(n, 2500, seed=40 + step)
    grad = rng.normal(size=(n, dim)).astype(np.float32)
    plan.build(t(ids))
    ops.kv_variable_gather_or_insert_plan(tabs[0].gpu, plan)
    tabs[0].cpu.gather_or_insert(ids, today=TODAY)
    if kind == "adam":
      hp = (1e-3, 0.9, 0.999, 1e-8, b1p, b2p)
      b1p *= 0.9
      b2p *= 0.999
    ops.kv_variable_apply_plan(k, tabs[0].gpu, tabs[1].gpu, tabs[2].gpu if len(tabs) > 2 else None,
                               plan, t(grad), hp)
    _oracle_step(k, tabs, ids, grad, hp)
  for tb in tabs:
    tb.check_state(rtol=RTOL, atol=ATOL)


def test_apply_plan_blacklist_and_revive():
  dim = 32
  var = Pair(dim, init=0.01)
  slot = Pair(3 * dim, init=0.0)
  rng = np.random.default_rng(9)
  n = 3000
  plan = ops.Plan(n, DEV)
  b1p, b2p = 0.9, 0.999
  n_black = 0
  for step in range(5):
    ids = rng.integers(0, 500, size=n).astype(np.int64)
    scale = 10.0 if step % 2 else 0.001
    grad = (rng.normal(size=(n, dim)) * scale).astype(np.float32)
    plan.build(t(ids))
    rows = ops.kv_variable_gather_or_insert_plan(var.gpu, plan).cpu().numpy()
    want = var.cpu.gather_or_insert(ids, today=TODAY)
    np.testing.assert_allclose(rows, want.reshape(rows.shape), rtol=RTOL, atol=ATOL)
    hp = (0.05, b1p, b2p, 0.9, 0.999, 1e-8, 0.0, 0.0, 0.5)
    ops.kv_variable_apply_plan(ops.OPT_GROUP_ADAM_V4, var.gpu, slot.gpu, None, plan, t(grad), hp)
    _oracle_step(ops.OPT_GROUP_ADAM_V4, [var, slot], ids, grad, hp)
    b1p *= 0.9
    b2p *= 0.999
    _, cs = var.check_state(rtol=RTOL, atol=ATOL)
    n_black = max(n_black, len(cs["black"]))
    slot.check_state(rtol=RTOL, atol=ATOL)
  assert n_black > 50
