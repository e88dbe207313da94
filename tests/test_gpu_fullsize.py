"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle only
checks a sample), plus the streaming configuration (configs[4]) end to end against the oracle."""
import numpy as np
import pytest
import torch

from oracle import binding as ob
from tfplus_b200 import ops

import bench
from kvtest_util import DEV, TODAY, Pair, t

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _clock():
  ops.set_today(TODAY)
  yield
  ops.set_today(None)


@pytest.fixture(scope="module")
def big():
  """configs[1]: 10 M-key table, dim 64, populated exactly as bench.py does."""
  ops.set_today(TODAY)
  st = bench.LocalStepper(bench.KEYS, bench.DIM, bench.BATCH, torch.device(DEV))
  st.overlap = False
  st.populate()
  yield st
  ops.destroy_kv_variable_op_v2(st.var)
  ops.destroy_kv_variable_op_v2(st.slot)


def test_full_size_membership_and_rows(big):
  assert ops.kv_variable_size_v2(big.var) == bench.KEYS == ops.kv_variable_shape_v2(big.var)[0]
  # every row is the deterministic initializer of its key: check a sample against the oracle
  o = ob.OracleTable(bench.DIM, 0, seed=1)
  o.set_init_table(bench.init_table(bench.DIM))
  rng = np.random.default_rng(0)
  sample = rng.integers(0, bench.KEYS, size=4096).astype(np.int64)
  got = ops.kv_variable_gather_or_zeros_v2(big.var, t(sample)).cpu().numpy()
  np.testing.assert_array_equal(got, o.gather_or_insert(sample, today=TODAY))
  # keys outside the populated range are absent (predict lookup = zeros), nothing was inserted
  miss = ops.kv_variable_gather_or_zeros_v2(big.var, t(sample + bench.KEYS)).cpu().numpy()
  assert not miss.any()
  assert ops.kv_variable_size_v2(big.var) == bench.KEYS


def test_full_size_step_properties(big):
  ids_np, grads_np = bench.make_batches(2, bench.KEYS, bench.BATCH, bench.DIM)
  ids, grad = t(ids_np[0]), t(grads_np[0])
  # unique: no duplicates, first-occurrence order, inverse index round trip, counts add up
  uniq, idx, counts = ops.unique(ids, with_counts=True)
  u = uniq.cpu().numpy()
  assert np.unique(u).size == u.size
  np.testing.assert_array_equal(u[idx.cpu().numpy()], ids_np[0])
  seen = {}
  for p, k in enumerate(ids_np[0]):
    if int(k) not in seen:
      seen[int(k)] = len(seen)
  np.testing.assert_array_equal(u, np.fromiter(seen.keys(), np.int64, len(seen)))
  assert int(counts.sum()) == bench.BATCH
  # segment sum: total mass is conserved (linearity), duplicates really merged
  gsum = ops.unsorted_segment_sum(grad, idx, uniq.numel())
  np.testing.assert_allclose(gsum.sum(0).cpu().numpy(), grads_np[0].sum(0), rtol=1e-4, atol=1e-2)
  k_hot = int(np.argmax(counts.cpu().numpy()))
  np.testing.assert_allclose(gsum[k_hot].cpu().numpy(),
                             grads_np[0][ids_np[0] == u[k_hot]].sum(0), rtol=1e-4, atol=1e-3)
  # gather is idempotent on rows (only frequencies move) and duplicates get identical rows
  r1 = ops.kv_variable_gather_or_insert_v2(big.var, ids)
  r2 = ops.kv_variable_gather_or_insert_v2(big.var, ids)
  assert torch.equal(r1, r2)
  rows_u = ops.kv_variable_gather_or_zeros_v2(big.var, uniq)
  assert torch.equal(rows_u.index_select(0, idx.long()), r1)
  # one GroupAdam step with zero regularisers == Adam (the reference's own check, full batch)
  before = rows_u.clone()
  ops.kv_variable_group_sparse_apply_adam_v4(big.var, big.slot, gsum, uniq, 1e-3, 0.9, 0.999, 0.9,
                                             0.999, 1e-8, 0.0, 0.0, 0.0)
  after = ops.kv_variable_gather_or_zeros_v2(big.var, uniq)
  c = lambda x: torch.tensor(x, dtype=torch.float32, device=DEV)
  m = (c(1.0) - c(0.9)) * gsum
  v = (c(1.0) - c(0.999)) * (gsum * gsum)
  lr = c(1e-3) * torch.sqrt(c(1.0) - c(0.999)) / (c(1.0) - c(0.9))
  want = before - lr * m / (torch.sqrt(v) + c(1e-8))
  torch.testing.assert_close(after, want, rtol=1e-5, atol=1e-7)
  assert ops.kv_variable_size_v2(big.var) == bench.KEYS       # nothing inserted, nothing blacklisted


def test_streaming_config_against_oracle():
  """configs[4]: 10 % unseen keys per batch, low-frequency filter (enter_threshold 3), Adagrad,
  eviction by timestamp every few steps and an export -> import round trip; every step is
  mirrored on the oracle."""
  dim, B, base_keys = 16, 4000, 20000
  var = Pair(dim, enter_threshold=3, init_seed=3)
  acc = Pair(dim, init=0.1)
  rng = np.random.default_rng(11)
  warm = np.arange(base_keys, dtype=np.int64)
  var.gather_or_insert(warm)
  day = TODAY
  for step in range(12):
    if step % 4 == 3:
      day += 3
      ops.set_today(day)
    old = (rng.zipf(1.1, size=B - B // 10) % base_keys).astype(np.int64)
    new = base_keys + step * (B // 10) + np.arange(B // 10, dtype=np.int64)
    ids = rng.permutation(np.concatenate([old, new]))
    g = (rng.integers(-8, 9, size=(B, dim)) / 16).astype(np.float32)
    got = ops.kv_variable_gather_or_insert_v2(var.gpu, t(ids)).cpu().numpy()
    want = var.cpu.gather_or_insert(ids, today=day)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
    u, idx = ob.unique(ids)
    gs = ob.segment_sum(g, idx, u.size)
    ops.kv_variable_sparse_apply_adagrad(var.gpu, acc.gpu, 0.05, t(gs), t(u))
    ob.apply_adagrad(var.cpu, acc.cpu, u, gs, 0.05, today=day)
    if step % 4 == 3:
      gone = ops.kv_variable_delete_with_timestamp(var.gpu, 2)
      want_gone = var.cpu.delete_with_timestamp(2, day)
      assert sorted(gone.cpu().tolist()) == sorted(want_gone.tolist())
    if step == 7:
      tensors = ops.kv_variable_export(var.gpu, first_n=6, enable_cutoff=True, cutoff_value=1e-20)
      e = var.cpu.export(first_n=6, enable_cutoff=True, cutoff_value=1e-20)
      ops.kv_variable_import(var.gpu, *tensors, first_n=6)
      var.cpu.import_(e["keys"], e["values"], e["init_table"], e["blacklist"], e["freq_keys"],
                      e["freq_values"].astype(np.uint32))
  import kvtest_util
  kvtest_util.TODAY = day          # check_state's oracle export does not need the clock
  var.check_state(rtol=1e-6, atol=1e-7)
  acc.check_state(rtol=1e-6, atol=1e-7)
  kvtest_util.TODAY = TODAY


def _export_state(st):
  k, v, *_rest = ops.kv_variable_export(st.var, first_n=6, enable_cutoff=False, cutoff_value=0.0,
                                        freq_dtype=torch.int32)
  ks, vs, *_ = ops.kv_variable_export(st.slot, first_n=6, enable_cutoff=False, cutoff_value=0.0,
                                      freq_dtype=torch.int32)
  a = dict(zip(k.cpu().numpy().tolist(), v.cpu().numpy()))
  b = dict(zip(ks.cpu().numpy().tolist(), vs.cpu().numpy()))
  return a, b, _rest


def test_rotation_graph_equals_step_by_step_and_oracle():
  """bench.py's timed schedule (one graph per rotation: lookup -> fused gradient sum + apply ->
  next lookup on one stream, the dedup plan of the next batch built beside it) leaves the
  tables EXACTLY as running the steps strictly one after the other does - the gradient sums
  are added in TF's order, nothing depends on scheduling - and both equal the oracle run over
  the same steps at the north-star's 1e-6, slot rows (m, v, linear) included."""
  keys, D, B, nb = 30000, 64, 4096, bench.N_BATCHES
  dev = torch.device(DEV)
  ids_np, grads_np = bench.make_batches(nb, keys, B, D, seed_ids=5, seed_grad=6)
  states, rows = [], []
  for mode in ("strict", "rotation"):
    st = bench.LocalStepper(keys, D, B, dev)
    st.populate()
    st.prepare([t(x) for x in ids_np], [t(x) for x in grads_np])   # = one eager pass (16 steps)
    assert st.rotation is not None
    if mode == "strict":
      for i in range(2 * nb + 3):
        st.step(i)
    else:
      st.run_steps(2 * nb + 3)                                      # 2 rotations + 3 single steps
    torch.cuda.synchronize()
    states.append(_export_state(st))
    rows.append([r.cpu().numpy().copy() for r in st.rows])
    hp = st.hp.cpu().numpy().copy()
    ops.destroy_kv_variable_op_v2(st.var)
    ops.destroy_kv_variable_op_v2(st.slot)
  (va, sa, _), (vb, sb, _) = states
  assert va.keys() == vb.keys() and sa.keys() == sb.keys()
  ks = sorted(va)
  np.testing.assert_array_equal(np.stack([va[k] for k in ks]), np.stack([vb[k] for k in ks]))
  np.testing.assert_array_equal(np.stack([sa[k] for k in ks]), np.stack([sb[k] for k in ks]))
  for ra, rb in zip(*rows):
    np.testing.assert_array_equal(ra, rb)
  # and both equal the oracle run over the same 3 * nb + 3 steps
  var = ob.OracleTable(D, 0, seed=1)
  var.set_init_table(bench.init_table(D))
  slot = ob.OracleTable(3 * D, 0, seed=1)
  slot.set_init_table(np.zeros((bench.INIT_ROWS, 3 * D), np.float32))
  allk = np.arange(keys, dtype=np.int64)
  var.gather_or_insert(allk, today=TODAY)
  slot.gather_or_insert(allk, today=TODAY)
  h = bench.HP
  b1p, b2p = np.float32(h["beta1"]), np.float32(h["beta2"])
  last_rows = None
  for i in range(3 * nb + 3):
    ids, grad = ids_np[i % nb], grads_np[i % nb]
    last_rows = var.gather_or_insert(ids, today=TODAY)
    u, idx = ob.unique(ids)
    ob.apply_group_adam_v4(var, slot, u, ob.segment_sum(grad, idx, u.size), h["lr"], float(b1p),
                           float(b2p), h["beta1"], h["beta2"], h["epsilon"], h["l1"], h["l2"],
                           h["l21"], today=TODAY)
    b1p, b2p = b1p * np.float32(h["beta1"]), b2p * np.float32(h["beta2"])
  assert hp[1] == b1p and hp[2] == b2p
  for got_map, otab, name in ((vb, var, "var"), (sb, slot, "m_v_linear")):
    ref = otab.export(first_n=6, enable_cutoff=False, cutoff_value=0.0, freq_u32=True)
    ref_rows = dict(zip(ref["keys"].tolist(), ref["values"]))
    assert ref_rows.keys() == got_map.keys()
    got = np.stack([got_map[k] for k in sorted(got_map)])
    want = np.stack([ref_rows[k] for k in sorted(got_map)])
    print("rotation vs oracle, %s: max |d| %.3g" % (name, np.abs(got - want).max()))
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7, err_msg=name)
  np.testing.assert_allclose(rows[1][(3 * nb + 2) % nb], last_rows.reshape(B, D), rtol=1e-6,
                             atol=1e-7)
