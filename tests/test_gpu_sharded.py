"""GPU side of key-hash sharding: routing kernels vs numpy, a world-1 router vs a plain table,
and (when the box has >= 2 GPUs) two NCCL ranks vs one unsharded oracle table."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding as ob
from tfplus_b200 import ops

from kvtest_util import DEV, TODAY, Pair, t
from test_sharded_gloo import D, _free_port, batches, make_table, owner_of

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _clock():
  ops.set_today(TODAY)
  yield
  ops.set_today(None)


@pytest.mark.parametrize("shards,mode", [(1, "hash"), (2, "hash"), (8, "hash"), (8, "mod"), (256, "hash")])
def test_partition_ids_kernel(shards, mode):
  rng = np.random.default_rng(shards)
  ids = rng.integers(-2**40, 2**40, size=50000).astype(np.int64)
  s_ids, perm, counts = ops.partition_ids(t(ids), shards, mode)
  s_ids, perm, counts = s_ids.cpu().numpy(), perm.cpu().numpy(), counts.cpu().numpy()
  own = owner_of(ids, shards, mode)
  np.testing.assert_array_equal(counts, np.bincount(own, minlength=shards))
  np.testing.assert_array_equal(s_ids[perm], ids)                       # perm: input pos -> sorted pos
  assert sorted(perm.tolist()) == list(range(ids.size))
  np.testing.assert_array_equal(owner_of(s_ids, shards, mode), np.repeat(np.arange(shards), counts))


def test_permute_and_scatter_rows():
  rng = np.random.default_rng(0)
  src = rng.normal(size=(1000, 24)).astype(np.float32)
  perm = rng.permutation(1000).astype(np.int32)
  got = ops.permute_rows(t(src), t(perm)).cpu().numpy()
  np.testing.assert_array_equal(got, src[perm])
  out = torch.empty(1000, 24, device=DEV)
  ops.scatter_rows(t(src), t(perm), out)
  want = np.empty_like(src)
  want[perm] = src
  np.testing.assert_array_equal(out.cpu().numpy(), want)


def test_world1_router_equals_plain_table():
  from tfplus_b200.sharded import ShardedKvVariable
  p = Pair(16, enter_threshold=2, init=0.25)
  tbl = ShardedKvVariable(16, 1, 0, DEV, enter_threshold=2, seed=11)
  ops.init_kv_variable_v2(tbl.var, t(p.init))
  rng = np.random.default_rng(3)
  for _ in range(3):
    ids = (rng.zipf(1.2, size=3000) % 700).astype(np.int64)
    rows = tbl.lookup(t(ids).reshape(30, 100))
    want = p.cpu.gather_or_insert(ids, today=TODAY)
    np.testing.assert_array_equal(rows.reshape(-1, 16).cpu().numpy(), want)
  p.gpu = tbl.var
  p.check_state()   # frequencies: occurrence counts travel with the deduped ids


def _nccl_worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  from tfplus_b200.sharded import ShardedKvVariable
  ops.set_today(TODAY)
  tbl = ShardedKvVariable(D, world, rank, dev, slot_dims=(D,), enter_threshold=2, seed=5)
  ops.init_kv_variable_v2(tbl.var, torch.full((16, D), 0.5, device=dev))
  ops.init_kv_variable_v2(tbl.slots[0], torch.full((16, D), 0.1, device=dev))
  looked = []
  for step in range(3):
    ids_all, grads_all = batches(world, step)
    ids = torch.from_numpy(ids_all[rank]).to(dev)
    grad = torch.from_numpy(grads_all[rank]).to(dev)
    looked.append(tbl.lookup(ids).cpu().numpy())
    owner_ids, owner_grads = tbl.owner_gradients(grad)
    ops.kv_variable_sparse_apply_adagrad(tbl.var, tbl.slots[0], 0.5, owner_grads, owner_ids)
  k, v, _, _, fk, fv = ops.kv_variable_export(tbl.var, first_n=6, enable_cutoff=True,
                                              cutoff_value=1e-20, freq_dtype=torch.int32)
  q.put((rank, looked, k.cpu().numpy(), v.cpu().numpy(), fk.cpu().numpy(),
         fv.cpu().numpy().view(np.uint32)))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_nccl_ranks_equal_one_table():
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  results = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  var, acc = make_table(D, 2, 0.5), make_table(D, 0, 0.1)
  for step in range(3):
    ids_all, grads_all = batches(world, step)
    ids, grad = np.concatenate(ids_all), np.concatenate(grads_all)
    rows = var.gather_or_insert(ids, today=TODAY)
    off = 0
    for r in range(world):
      n = ids_all[r].size
      np.testing.assert_array_equal(results[r][1][step], rows[off:off + n])
      off += n
    u, idx = ob.unique(ids)
    ob.apply_adagrad(var, acc, u, ob.segment_sum(grad, idx, u.size), 0.5, today=TODAY)
  ref = var.export(first_n=6, enable_cutoff=True, cutoff_value=1e-20, freq_u32=True)
  got_rows, got_freq = {}, {}
  for _, _, k, v, fk, fv in results:
    got_rows.update({int(a): b for a, b in zip(k, v)})
    got_freq.update({int(a): int(b) for a, b in zip(fk, fv)})
  assert got_freq == {int(a): int(b) for a, b in zip(ref["freq_keys"], ref["freq_values"])}
  assert set(got_rows) == set(int(a) for a in ref["keys"])
  for a, b in zip(ref["keys"], ref["values"]):
    np.testing.assert_allclose(got_rows[int(a)], b, rtol=1e-6, atol=1e-7)


PAD = -2**63 + 1


def test_route_ids_padded_layout():
  rng = np.random.default_rng(7)
  ids = np.unique(rng.integers(-2**40, 2**40, size=6000).astype(np.int64))
  occ = rng.integers(1, 9, size=ids.size).astype(np.int32)
  G, cap = 4, 2048
  num = torch.tensor([ids.size - 100], dtype=torch.int32, device=DEV)   # only a prefix is valid
  r = ops.route_ids(t(ids), t(occ), G, cap, "hash", num_ids=num)
  n = ids.size - 100
  own = owner_of(ids[:n], G)
  counts = r["counts"].cpu().numpy()
  np.testing.assert_array_equal(counts, np.bincount(own, minlength=G))
  send_ids = r["send_ids"].cpu().numpy().reshape(G, cap)
  send_occ = r["send_occ"].cpu().numpy().reshape(G, cap)
  perm = r["perm"].cpu().numpy()[:n]
  assert int(r["overflow"].item()) == 0
  for g in range(G):
    assert set(send_ids[g, :counts[g]].tolist()) == set(ids[:n][own == g].tolist())
    assert (send_ids[g, counts[g]:] == PAD).all() and (send_occ[g, counts[g]:] == 0).all()
  np.testing.assert_array_equal(send_ids.reshape(-1)[perm], ids[:n])
  np.testing.assert_array_equal(send_occ.reshape(-1)[perm], occ[:n])
  # overflow: capacity too small is reported, nothing is written out of bounds
  r2 = ops.route_ids(t(ids), t(occ), G, 64, "hash")
  assert int(r2["overflow"].item()) == 1
  assert (r2["perm"].cpu().numpy() < G * 64).all()


def test_pad_ids_are_ignored_and_row_helpers():
  p = Pair(16, init=0.5)
  acc = Pair(16, init=0.1)
  ids = np.array([5, PAD, 7, PAD], np.int64)
  rows = ops.kv_variable_gather_or_insert_v2(p.gpu, t(ids)).cpu().numpy()
  assert (rows[[1, 3]] == 0).all() and (rows[[0, 2]] == 0.5).all()
  assert ops.kv_variable_shape_v2(p.gpu)[0] == 2
  ops.kv_variable_sparse_apply_adagrad(p.gpu, acc.gpu, 0.1, torch.ones(3, 16, device=DEV),
                                       t(np.array([5, PAD, 9], np.int64)))
  assert ops.kv_variable_shape_v2(p.gpu)[0] == 3 and ops.kv_variable_shape_v2(acc.gpu)[0] == 2
  src = torch.arange(40, dtype=torch.float32, device=DEV).reshape(10, 4)
  perm = torch.tensor([9, -1, 0, 3], dtype=torch.int32, device=DEV)
  idx = torch.tensor([3, 3, 0, 1, 2], dtype=torch.int32, device=DEV)
  out = ops.expand_rows(src, perm, idx, 5, torch.empty(5, 4, device=DEV)).cpu().numpy()
  np.testing.assert_array_equal(out, src.cpu().numpy()[[3, 3, 9, 0, 0]] * np.array([1, 1, 1, 0, 1])[:, None])
  dst = torch.zeros(10, 4, device=DEV)
  ops.scatter_rows_n(src[:4], perm, 4, torch.tensor([3], dtype=torch.int32, device=DEV), dst)
  want = np.zeros((10, 4), np.float32)
  want[9], want[0] = src[0].cpu().numpy(), src[2].cpu().numpy()
  np.testing.assert_array_equal(dst.cpu().numpy(), want)


def _segments(buf, G):
  """Device array of G pointers to the G equal row-blocks of `buf` (what a peer table is,
  with every "peer" living on this GPU)."""
  step = buf.numel() // G * buf.element_size()
  return torch.tensor([buf.data_ptr() + g * step for g in range(G)], dtype=torch.int64, device=DEV)


def test_peer_variants_equal_the_local_ones_on_one_gpu():
  # the _peer entry points only change WHERE a row is stored: same kernels, segment pointers
  rng = np.random.default_rng(17)
  G, cap, D = 4, 1024, 16
  ids = np.unique(rng.integers(-2**40, 2**40, size=3000).astype(np.int64))
  occ = rng.integers(1, 9, size=ids.size).astype(np.int32)
  num = torch.tensor([ids.size - 50], dtype=torch.int32, device=DEV)
  ref = ops.route_ids(t(ids), t(occ), G, cap, "hash", num_ids=num)
  ids_in = torch.zeros(G * cap, dtype=torch.int64, device=DEV)
  occ_in = torch.full((G * cap,), -1, dtype=torch.int32, device=DEV)
  out = {"perm": torch.empty(ids.size, dtype=torch.int32, device=DEV),
         "counts": torch.empty(G, dtype=torch.int32, device=DEV),
         "overflow": torch.zeros(1, dtype=torch.int32, device=DEV)}
  ops.route_ids_peer(t(ids), t(occ), G, cap, "hash", num, _segments(ids_in, G),
                     _segments(occ_in, G), out)
  np.testing.assert_array_equal(out["counts"].cpu().numpy(), ref["counts"].cpu().numpy())
  got_ids, got_occ = ids_in.cpu().numpy().reshape(G, cap), occ_in.cpu().numpy().reshape(G, cap)
  want_ids = ref["send_ids"].cpu().numpy().reshape(G, cap)
  want_occ = ref["send_occ"].cpu().numpy().reshape(G, cap)
  for g in range(G):      # order inside a shard row is launch-dependent: compare as multisets
    assert sorted(zip(got_ids[g].tolist(), got_occ[g].tolist())) == \
           sorted(zip(want_ids[g].tolist(), want_occ[g].tolist()))
  perm = out["perm"].cpu().numpy()[:ids.size - 50]
  np.testing.assert_array_equal(got_ids.reshape(-1)[perm], ids[:ids.size - 50])
  assert int(out["overflow"].item()) == 0

  # dedup + route fused: same shard rows as unique followed by route
  raw = rng.choice(ids[:2000], size=9000).astype(np.int64)
  B = raw.size
  uq = {k: torch.empty(B, dtype=d, device=DEV) for k, d in
        [("uniq", torch.int64), ("idx", torch.int32), ("cnt", torch.int32)]}
  num2 = torch.zeros(1, dtype=torch.int32, device=DEV)
  ids2 = torch.zeros(G * cap, dtype=torch.int64, device=DEV)
  occ2 = torch.full((G * cap,), -1, dtype=torch.int32, device=DEV)
  out2 = {"perm": torch.full((B,), -7, dtype=torch.int32, device=DEV),
          "counts": torch.full((G,), 99, dtype=torch.int32, device=DEV),
          "overflow": torch.zeros(1, dtype=torch.int32, device=DEV)}
  ops.route_fill_peer(G, cap, _segments(ids2, G), _segments(occ2, G), out2["counts"])
  assert (ids2 == PAD).all() and (occ2 == 0).all() and (out2["counts"] == 0).all()
  ops.unique_route_peer(t(raw), uq["uniq"], uq["idx"], uq["cnt"], num2, G, cap, "hash",
                        _segments(ids2, G), _segments(occ2, G), out2)
  u_ref, idx_ref = ob.unique(raw)
  n_u = int(num2.item())
  assert n_u == u_ref.size
  np.testing.assert_array_equal(uq["uniq"].cpu().numpy()[:n_u], u_ref)
  np.testing.assert_array_equal(uq["idx"].cpu().numpy(), idx_ref)
  c_ref = np.bincount(idx_ref, minlength=n_u)
  np.testing.assert_array_equal(uq["cnt"].cpu().numpy()[:n_u], c_ref)
  own = owner_of(u_ref, G)
  np.testing.assert_array_equal(out2["counts"].cpu().numpy(), np.bincount(own, minlength=G))
  perm2 = out2["perm"].cpu().numpy()[:n_u]
  np.testing.assert_array_equal(perm2 // cap, own)
  np.testing.assert_array_equal(ids2.cpu().numpy()[perm2], u_ref)
  np.testing.assert_array_equal(occ2.cpu().numpy()[perm2], c_ref)
  assert (ids2 != PAD).sum().item() == n_u and int(out2["overflow"].item()) == 0

  # owner lookup into segments == plain lookup; padding rows are left untouched
  p, q = Pair(D), Pair(D)
  recv = ids_in.clone()
  counts = t(occ_in.clamp(min=0).cpu().numpy())
  rows_seg = torch.full((G * cap, D), 7.0, device=DEV)
  ops.kv_variable_gather_or_insert_peer(p.gpu, recv, counts, _segments(rows_seg, G), cap)
  rows_ref = ops.kv_variable_gather_or_insert_with_counts(q.gpu, recv, counts)
  pad = (recv == PAD).cpu().numpy()
  np.testing.assert_array_equal(rows_seg.cpu().numpy()[~pad], rows_ref.cpu().numpy()[~pad])
  assert (rows_seg.cpu().numpy()[pad] == 7.0).all()
  p.cpu.gather_or_insert(recv.cpu().numpy()[~pad], counts.cpu().numpy()[~pad], today=TODAY)
  p.check_state()

  # gradient rows into segments == scatter_rows_n
  src = torch.randn(ids.size, D, device=DEV)
  a, b = torch.zeros(G * cap, D, device=DEV), torch.zeros(G * cap, D, device=DEV)
  ops.scatter_rows_n(src, out["perm"], ids.size, num, a)
  ops.scatter_rows_n_peer(src, out["perm"], ids.size, num, _segments(b, G), cap)
  np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy())

  # a one-GPU "world" needs no barrier; bad arguments are refused
  flags = torch.zeros(1, dtype=torch.int32, device=DEV)
  state = torch.zeros(2, dtype=torch.int32, device=DEV)
  ops.peer_barrier(_segments(flags, 1), flags, state, 0, 1)
  with pytest.raises(Exception):
    ops.peer_barrier(_segments(flags, 1), flags, state, 3, 2)


NSTEPS = 4


def _padded_worker(rank, world, port, q, use_graph, exchange="nccl"):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  from tfplus_b200.sharded import PaddedShardedStep, PeerShardedStep
  Step = PeerShardedStep if exchange.startswith("peer") else PaddedShardedStep
  ops.set_today(TODAY)
  var = ops.kv_variable(value_shape=[D], enter_threshold=0, device=dev, seed=5, capacity_hint=4096)
  slot = ops.kv_variable(value_shape=[3 * D], device=dev, seed=5, capacity_hint=4096)
  ops.init_kv_variable_v2(var, torch.full((16, D), 0.5, device=dev))
  ops.init_kv_variable_v2(slot, torch.zeros(16, 3 * D, device=dev))
  hp = torch.tensor([0.05, 0.9, 0.999, 0.9, 0.999, 1e-8, 1e-5, 1e-5, 1e-5], device=dev)
  betas = torch.tensor([0.9, 0.999], device=dev)
  step = Step(var, slot, D, 400, world, rank, dev, hp, betas, cap=256)
  data = [batches(world, s) for s in range(NSTEPS)]
  ids = [torch.from_numpy(d[0][rank]).to(dev) for d in data]
  grads = [torch.from_numpy(d[1][rank]).to(dev) for d in data]
  looked = {}
  if exchange == "peer_rotation":
    # the pipelined schedule: steps issued together, only true dependencies between them
    first = 0
    if use_graph:
      for s in (0, 1):                    # strict steps first: the two schedules must mix
        looked[s] = step.run(ids[s], grads[s]).cpu().numpy().copy()
      first = 2
      torch.cuda.synchronize()
      side = torch.cuda.Stream(device=dev)
      side.wait_stream(torch.cuda.current_stream(dev))
      with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
          step.run_rotation(ids[first:], grads[first:])
      torch.cuda.current_stream(dev).wait_stream(side)
      g.replay()
      torch.cuda.synchronize()
      del g
    else:
      step.run_rotation(ids, grads)
    looked[NSTEPS - 1] = step.out.cpu().numpy().copy()
  elif use_graph:
    step.run(ids[0], grads[0])            # eager warm-up = step 0
    looked[0] = step.out.cpu().numpy().copy()
    torch.cuda.synchronize()
    graphs = []
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      for s in range(1, NSTEPS):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
          step.run(ids[s], grads[s])
        graphs.append(g)
    torch.cuda.current_stream(dev).wait_stream(side)
    for s, g in enumerate(graphs, 1):
      g.replay()
      torch.cuda.synchronize()
      looked[s] = step.out.cpu().numpy().copy()
    del graphs, g          # captured NCCL work must be gone before the process group is torn down
    torch.cuda.synchronize()
  else:
    for s in range(NSTEPS):
      looked[s] = step.run(ids[s], grads[s]).cpu().numpy().copy()
  assert not step.overflowed()
  if exchange.startswith("peer"):
    assert step.barrier_timeouts() == 0
  k, v, _, _, fk, fv = ops.kv_variable_export(var, first_n=6, enable_cutoff=True,
                                              cutoff_value=1e-20, freq_dtype=torch.int32)
  q.put((rank, looked, k.cpu().numpy(), v.cpu().numpy(), fk.cpu().numpy(),
         fv.cpu().numpy().view(np.uint32)))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["nccl", "peer", "peer_rotation"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_padded_sharded_step_equals_one_table(use_graph, exchange):
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_padded_worker, args=(r, world, port, q, use_graph, exchange))
           for r in range(world)]
  for p in procs:
    p.start()
  results = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  var = make_table(D, 0, 0.5)
  slot = ob.OracleTable(3 * D, 0, seed=5)
  slot.set_init_table(np.zeros((16, 3 * D), np.float32))
  b1p, b2p = np.float32(0.9), np.float32(0.999)
  for step in range(NSTEPS):
    ids_all, grads_all = batches(world, step)
    ids, grad = np.concatenate(ids_all), np.concatenate(grads_all)
    rows = var.gather_or_insert(ids, today=TODAY)
    off = 0
    for r in range(world):
      n = ids_all[r].size
      if step in results[r][1]:
        np.testing.assert_allclose(results[r][1][step], rows[off:off + n], rtol=1e-6, atol=1e-7)
      off += n
    u, idx = ob.unique(ids)
    ob.apply_group_adam_v4(var, slot, u, ob.segment_sum(grad, idx, u.size), 0.05, float(b1p),
                           float(b2p), 0.9, 0.999, 1e-8, 1e-5, 1e-5, 1e-5, today=TODAY)
    b1p, b2p = b1p * np.float32(0.9), b2p * np.float32(0.999)
  ref = var.export(first_n=6, enable_cutoff=True, cutoff_value=1e-20, freq_u32=True)
  got_rows, got_freq = {}, {}
  for _, _, k, v, fk, fv in results:
    got_rows.update({int(a): b for a, b in zip(k, v)})
    got_freq.update({int(a): int(b) for a, b in zip(fk, fv)})
  assert got_freq == {int(a): int(b) for a, b in zip(ref["freq_keys"], ref["freq_values"])}
  assert set(got_rows) == set(int(a) for a in ref["keys"])
  for a, b in zip(ref["keys"], ref["values"]):
    np.testing.assert_allclose(got_rows[int(a)], b, rtol=1e-6, atol=1e-7)
