"""Key-hash sharding of a KvVariable over the GPUs of one box.

The reference scales a table by `PartitionedVariable` shards placed on parameter servers:
`p = ids % np`, dynamic_partition -> per-shard gather -> dynamic_stitch
(tfplus/kv_variable/python/ops/embedding_ops.py:121-204), over TF gRPC.  Here every rank owns
the keys with `mix64(key) % world == rank` (hash, not raw mod, so Zipf-hot small ids spread;
mode="mod" keeps the reference's rule for checkpoints saved as `<var>/part_i`) in its own
HBM-resident table, the batch is data-parallel (every rank brings its own ids), and one step
is three exchanges over NVLink:

  forward   ids (+ occurrence counts) to their owners        all_to_all
            rows back to the requesters                      all_to_all
  backward  per-id summed gradients to the owners            all_to_all
            owners merge what several ranks sent for the same key and run the fused apply.

The routing logic is backend-agnostic torch code: the product backend below is the CUDA one
(tfplus_b200.ops -> C ABI); tests drive the same logic with world_size-2 gloo on CPU.
"""
import torch
import torch.distributed as dist

from . import ops


class CudaBackend:
  """The device ops the router needs, bound to tfplus_b200.ops."""

  def unique_with_counts(self, ids):
    return ops.unique(ids, with_counts=True)

  def unique(self, ids):
    return ops.unique(ids)

  def partition_ids(self, ids, world, mode):
    return ops.partition_ids(ids, world, mode)

  def permute_rows(self, src, perm):
    return ops.permute_rows(src, perm)

  def scatter_rows(self, src, perm, out):
    return ops.scatter_rows(src, perm, out)

  def segment_sum(self, data, idx, num):
    return ops.unsorted_segment_sum(data, idx, num)

  def gather_or_insert(self, table, ids, counts):
    return ops.kv_variable_gather_or_insert_with_counts(table, ids, counts)

  def gather_or_zeros(self, table, ids):
    return ops.kv_variable_gather_or_zeros_v2(table, ids)


class Route:
  """Everything the backward pass needs to retrace one lookup."""
  __slots__ = ("n", "uniq", "idx", "perm", "send_counts", "recv_counts", "recv_ids", "timing")


class ShardedRouter:
  """Routes ids / rows / gradients between requesters and owners."""

  def __init__(self, world, rank, group=None, backend=None, mode="hash"):
    self.world, self.rank, self.group = world, rank, group
    self.be = backend or CudaBackend()
    self.mode = mode
    self.bytes_sent = 0  # NVLink payload this rank put on the wire (bench bookkeeping)

  # -- collectives ------------------------------------------------------------
  def _a2a(self, inp, in_splits, out_splits):
    out = inp.new_empty((sum(out_splits),) + tuple(inp.shape[1:]))
    if self.world == 1:
      out.copy_(inp)
      return out
    dist.all_to_all_single(out, inp.contiguous(), out_splits, in_splits, group=self.group)
    if inp.shape[0]:
      per_row = inp.numel() // inp.shape[0] * inp.element_size()
      self.bytes_sent += (inp.shape[0] - in_splits[self.rank]) * per_row
    return out

  def _exchange_counts(self, send_counts):
    if self.world == 1:
      return send_counts.clone()
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts, group=self.group)
    return recv

  # -- forward ----------------------------------------------------------------
  def route(self, ids):
    """Dedup locally (with occurrence counts, so owners keep exact frequencies), group the
    unique ids by owner and ship them.  Returns (Route, counts received by this owner)."""
    r = Route()
    ids = ids.reshape(-1)
    r.n = ids.numel()
    uniq, idx, counts = self.be.unique_with_counts(ids)
    r.uniq, r.idx = uniq, idx
    sorted_ids, perm, shard_counts = self.be.partition_ids(uniq, self.world, self.mode)
    r.perm = perm
    sorted_counts = torch.empty_like(counts)
    sorted_counts[perm.long()] = counts
    recv_counts = self._exchange_counts(shard_counts)
    r.send_counts = shard_counts.cpu().tolist()   # the one host sync of the forward pass
    r.recv_counts = recv_counts.cpu().tolist()
    r.recv_ids = self._a2a(sorted_ids, r.send_counts, r.recv_counts)
    recv_occ = self._a2a(sorted_counts, r.send_counts, r.recv_counts)
    return r, recv_occ

  def return_rows(self, route, owner_rows):
    """Owner rows [sum(recv), D] -> rows for the original ids [n, D]."""
    rows_sorted = self._a2a(owner_rows, route.recv_counts, route.send_counts)
    rows_uniq = self.be.permute_rows(rows_sorted, route.perm)
    return rows_uniq.index_select(0, route.idx.long())

  # -- backward ---------------------------------------------------------------
  def send_grads(self, route, grad):
    """grad [n, D] for the original ids -> (owner_unique_ids, owner_summed_grads)."""
    u = route.uniq.numel()
    gsum = self.be.segment_sum(grad.reshape(route.n, -1), route.idx, u)
    g_sorted = torch.empty_like(gsum)
    self.be.scatter_rows(gsum, route.perm, g_sorted)
    recv = self._a2a(g_sorted, route.send_counts, route.recv_counts)
    # several ranks may have sent the same key: merge before the (unique-id) apply
    owner_ids, oidx = self.be.unique(route.recv_ids)
    owner_grads = self.be.segment_sum(recv, oidx, owner_ids.numel())
    return owner_ids, owner_grads


class ShardedKvVariable:
  """A KvVariable (plus its slot variables) sharded by key hash over `world` ranks."""

  def __init__(self, dim, world, rank, device, slot_dims=(), enter_threshold=0, group=None,
               backend=None, mode="hash", capacity_hint=0, seed=0, table_factory=None):
    self.dim, self.world, self.rank = dim, world, rank
    self.router = ShardedRouter(world, rank, group, backend, mode)
    make = table_factory or (lambda d, thr: ops.kv_variable(
        value_shape=[d], enter_threshold=thr, device=device, capacity_hint=capacity_hint,
        seed=seed))
    self.var = make(dim, enter_threshold)
    self.slots = [make(d, 0) for d in slot_dims]

  def lookup(self, ids, training=True):
    """embedding rows for `ids` (any shape) -> ids.shape + [dim]; keeps the route for backward."""
    shape = tuple(ids.shape)
    route, occ = self.router.route(ids)
    be = self.router.be
    if training:
      rows = be.gather_or_insert(self.var, route.recv_ids, occ)
    else:
      rows = be.gather_or_zeros(self.var, route.recv_ids)
    out = self.router.return_rows(route, rows)
    self.last_route = route
    return out.reshape(shape + (self.dim,))

  def owner_gradients(self, grad, route=None):
    return self.router.send_grads(route or self.last_route, grad)


# ---------------------------------------------------------------------------
# bench driver (bench.py --gpus N)
# ---------------------------------------------------------------------------
class ShardedStepper:
  """Weak-scaling microbench: `keys` keys per GPU, every rank brings its own batch."""

  STAGES = ["route+lookup", "grads+apply"]

  def __init__(self, keys_per_gpu, dim, batch, hp, dev, rank, world):
    import bench
    self.torch = torch
    self.keys, self.dim, self.batch, self.dev = keys_per_gpu, dim, batch, dev
    self.rank, self.world, self.hp = rank, world, hp
    cap = int(keys_per_gpu * 1.15) + batch
    self.tbl = ShardedKvVariable(dim, world, rank, dev, slot_dims=(3 * dim,),
                                 capacity_hint=cap, seed=1)
    ops.init_kv_variable_v2(self.tbl.var, torch.from_numpy(bench.init_table(dim)).to(dev))
    ops.init_kv_variable_v2(self.tbl.slots[0], torch.zeros(bench.INIT_ROWS, 3 * dim, device=dev))
    self.hpt = torch.tensor([hp["lr"], hp["beta1"], hp["beta2"], hp["beta1"], hp["beta2"],
                             hp["epsilon"], hp["l1"], hp["l2"], hp["l21"]], dtype=torch.float32,
                            device=dev)
    self.betas = torch.tensor([hp["beta1"], hp["beta2"]], dtype=torch.float32, device=dev)
    self.launches_per_step = 0
    self.steps_done = 0

  def populate(self):
    """Insert the keys this rank owns out of the global id range [0, keys * world)."""
    t = self.torch
    total = self.keys * self.world
    chunk = 1 << 20
    for s in range(0, total, chunk):
      ids = t.arange(s, min(total, s + chunk), dtype=t.int64, device=self.dev)
      sorted_ids, _, counts = ops.partition_ids(ids, self.world, "hash")
      c = counts.cpu().tolist()
      lo = sum(c[:self.rank])
      mine = sorted_ids[lo:lo + c[self.rank]]
      if mine.numel():
        ops.kv_variable_gather_or_insert_v2(self.tbl.var, mine)
        ops.kv_variable_gather_or_insert_v2(self.tbl.slots[0], mine)
    t.cuda.synchronize()

  def step_eager(self, ids, grad):
    rows = self.tbl.lookup(ids)
    owner_ids, owner_grads = self.tbl.owner_gradients(grad)
    if owner_ids.numel():
      ops.kv_variable_group_sparse_apply_adam_v4_dev(self.tbl.var, self.tbl.slots[0], owner_grads,
                                                     owner_ids, self.hpt)
    self.hpt[1:3].mul_(self.betas)
    self.steps_done += 1
    return rows

  def prepare(self, ids_d, grads_d):
    from . import _lib
    self.ids_d, self.grads_d = ids_d, grads_d
    l0 = _lib.launch_count()
    for i in range(2):
      self.step_eager(ids_d[i], grads_d[i])
    self.launches_per_step = (_lib.launch_count() - l0) // 2
    self.torch.cuda.synchronize()

  def step(self, i):
    k = i % len(self.ids_d)
    return self.step_eager(self.ids_d[k], self.grads_d[k])

  def stage_times(self, steps):
    t = self.torch
    acc = [0.0, 0.0]
    for i in range(steps):
      k = i % len(self.ids_d)
      e = [t.cuda.Event(enable_timing=True) for _ in range(3)]
      e[0].record()
      self.tbl.lookup(self.ids_d[k])
      e[1].record()
      owner_ids, owner_grads = self.tbl.owner_gradients(self.grads_d[k])
      if owner_ids.numel():
        ops.kv_variable_group_sparse_apply_adam_v4_dev(self.tbl.var, self.tbl.slots[0],
                                                       owner_grads, owner_ids, self.hpt)
      e[2].record()
      t.cuda.synchronize()
      acc[0] += e[0].elapsed_time(e[1])
      acc[1] += e[1].elapsed_time(e[2])
    return {self.STAGES[0]: acc[0] / steps, self.STAGES[1]: acc[1] / steps}

  def prepare_host(self, ids_h, grads_h, rows_h):
    t = self.torch
    self.ids_h, self.grads_h, self.rows_h = ids_h, grads_h, rows_h
    self.h_ids = t.empty(self.batch, dtype=t.int64, device=self.dev)
    self.h_grad = t.empty((self.batch, self.dim), dtype=t.float32, device=self.dev)

  def finish_host(self):
    pass

  def step_host(self, i):
    k = i % len(self.ids_h)
    self.h_ids.copy_(self.ids_h[k], non_blocking=True)
    self.h_grad.copy_(self.grads_h[k], non_blocking=True)
    rows = self.step_eager(self.h_ids, self.h_grad)
    self.rows_h.copy_(rows, non_blocking=True)
