"""Host-side logic of key-hash sharding (tfplus_b200/sharded.py) with world_size 2 over gloo on
the CPU: the router is backend-agnostic, so the test plugs in a CPU backend built on the oracle
and checks that two shards together behave exactly like one unsharded reference table."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding as ob

D = 8
TODAY = 19000
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x):
  x = x.astype(np.uint64)
  with np.errstate(over="ignore"):
    x ^= x >> np.uint64(30); x *= np.uint64(0xbf58476d1ce4e5b9)
    x ^= x >> np.uint64(27); x *= np.uint64(0x94d049bb133111eb)
    x ^= x >> np.uint64(31)
  return x


def owner_of(ids, world, mode="hash"):
  ids = np.asarray(ids, np.int64)
  if mode == "mod":
    return np.mod(ids, world).astype(np.int64)
  return (mix64(ids.astype(np.uint64) ^ np.uint64(0x5446534d)) % np.uint64(world)).astype(np.int64)


class OracleBackend:
  """CPU stand-in for tfplus_b200.sharded.CudaBackend (test infrastructure)."""

  def unique_with_counts(self, ids):
    u, idx, c = ob.unique(ids.numpy(), with_counts=True)
    return torch.from_numpy(u), torch.from_numpy(idx.copy()), torch.from_numpy(c)

  def unique(self, ids):
    u, idx = ob.unique(ids.numpy())
    return torch.from_numpy(u), torch.from_numpy(idx.copy())

  def partition_ids(self, ids, world, mode):
    a = ids.numpy()
    own = owner_of(a, world, mode)
    order = np.argsort(own, kind="stable")
    perm = np.empty(a.size, np.int32)
    perm[order] = np.arange(a.size, dtype=np.int32)
    counts = np.bincount(own, minlength=world).astype(np.int32)
    return torch.from_numpy(a[order].copy()), torch.from_numpy(perm), torch.from_numpy(counts)

  def permute_rows(self, src, perm):
    return src[perm.long()]

  def scatter_rows(self, src, perm, out):
    out[perm.long()] = src
    return out

  def segment_sum(self, data, idx, num):
    return torch.from_numpy(ob.segment_sum(data.numpy(), idx.numpy(), int(num)))

  def gather_or_insert(self, table, ids, counts):
    return torch.from_numpy(table.gather_or_insert(ids.numpy(), counts.numpy(), today=TODAY))

  def gather_or_zeros(self, table, ids):
    return torch.from_numpy(table.gather_or_zeros(ids.numpy()))


def make_table(dim, thr, init_value):
  t = ob.OracleTable(dim, thr, seed=5)
  t.set_init_table(np.full((16, dim), init_value, np.float32))
  return t


def batches(world, step):
  rng = np.random.default_rng(100 + step)
  return [(rng.zipf(1.3, size=400) % 150 - 3).astype(np.int64) for _ in range(world)], \
         [rng.integers(-3, 4, size=(400, D)).astype(np.float32) for _ in range(world)]


def _worker(rank, world, port, mode, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from tfplus_b200.sharded import ShardedKvVariable
  tbl = ShardedKvVariable(D, world, rank, "cpu", slot_dims=(D,), enter_threshold=2,
                          backend=OracleBackend(), mode=mode,
                          table_factory=lambda d, thr: make_table(d, thr, 0.5 if thr else 0.1))
  looked = []
  for step in range(3):
    ids_all, grads_all = batches(world, step)
    ids, grad = torch.from_numpy(ids_all[rank]), torch.from_numpy(grads_all[rank])
    rows = tbl.lookup(ids.reshape(20, 20))
    assert tuple(rows.shape) == (20, 20, D)
    looked.append(rows.reshape(-1, D).numpy().copy())
    owner_ids, owner_grads = tbl.owner_gradients(grad)
    assert len(set(owner_ids.tolist())) == owner_ids.numel()           # merged across ranks
    assert (owner_of(owner_ids.numpy(), world, mode) == rank).all()    # only keys this rank owns
    ob.apply_adagrad(tbl.var, tbl.slots[0], owner_ids.numpy(), owner_grads.numpy(), 0.5, today=TODAY)
  e = tbl.var.export(first_n=6, enable_cutoff=True, cutoff_value=1e-20, freq_u32=True)
  q.put((rank, looked, {k: v for k, v in e.items() if k != "init_table"}))
  dist.barrier()
  dist.destroy_process_group()


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


@pytest.mark.parametrize("mode", ["hash", "mod"])
def test_two_shards_equal_one_table(mode):
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
  for p in procs:
    p.start()
  results = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0

  # reference: ONE table fed the concatenated batches (what the reference computes unsharded)
  var, acc = make_table(D, 2, 0.5), make_table(D, 0, 0.1)
  for step in range(3):
    ids_all, grads_all = batches(world, step)
    ids, grad = np.concatenate(ids_all), np.concatenate(grads_all)
    rows = var.gather_or_insert(ids, today=TODAY)
    off = 0
    for r in range(world):
      n = ids_all[r].size
      np.testing.assert_array_equal(results[r][1][step], rows[off:off + n])  # looked-up rows
      off += n
    u, idx = ob.unique(ids)
    ob.apply_adagrad(var, acc, u, ob.segment_sum(grad, idx, u.size), 0.5, today=TODAY)
  ref = var.export(first_n=6, enable_cutoff=True, cutoff_value=1e-20, freq_u32=True)
  got_rows, got_freq = {}, {}
  for _, _, e in results:
    got_rows.update({int(k): v for k, v in zip(e["keys"], e["values"])})
    got_freq.update({int(k): int(v) for k, v in zip(e["freq_keys"], e["freq_values"])})
  assert got_freq == {int(k): int(v) for k, v in zip(ref["freq_keys"], ref["freq_values"])}
  assert set(got_rows) == set(int(k) for k in ref["keys"])
  for k, v in zip(ref["keys"], ref["values"]):
    np.testing.assert_array_equal(got_rows[int(k)], v)   # integer gradients: sums are exact


def test_owner_hash_matches_device_rule():
  # kv_partition_ids mode 0: mix64(id ^ 0x5446534d) % shards (dedup.cu owner_of)
  ids = np.array([0, 1, -1, 2**40, -2**62], np.int64)
  assert owner_of(ids, 1).tolist() == [0] * 5
  o8 = owner_of(ids, 8)
  assert ((0 <= o8) & (o8 < 8)).all()
  assert owner_of(np.array([-3, -1, 5]), 4, "mod").tolist() == [1, 3, 1]   # floormod
