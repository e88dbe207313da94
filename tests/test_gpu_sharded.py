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
