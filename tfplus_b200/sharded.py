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
import math
import os

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


class PaddedShardedStep:
  """The sync-free sharded step: every buffer has a data-independent shape, so the whole
  forward + backward (kernels and the NCCL all-to-alls between them) is captured in one CUDA
  graph and replayed; no host round trip anywhere.

  Exchange layout: each rank sends each peer a fixed `cap` ids (padded with the pad id, which
  lookups answer with zeros and applies skip).  `cap` defaults to batch / (2 * world) + 1024:
  ~1.5x the expected unique ids per peer for the Zipf(1.1) batch; if a row overflows, the
  overflow flag is raised on the device and the caller must rerun the step on the exact
  (variable-size) path of ShardedKvVariable.
  """

  def __init__(self, var, slot, dim, batch, world, rank, dev, hpt, betas, cap=None, group=None,
               mode="hash"):
    t = torch
    self.var, self.slot, self.dim, self.batch = var, slot, dim, batch
    self.world, self.rank, self.dev, self.group, self.mode = world, rank, dev, group, mode
    self.hpt, self.betas = hpt, betas
    self.cap = cap or (batch // (2 * world) + 1024) // 2 * 2
    B, G, C, D = batch, world, self.cap, dim
    i64 = dict(dtype=t.int64, device=dev)
    i32 = dict(dtype=t.int32, device=dev)
    f32 = dict(dtype=t.float32, device=dev)
    self.uniq, self.idx = t.empty(B, **i64), t.empty(B, **i32)
    self.cnt, self.num = t.empty(B, **i32), t.zeros(1, **i32)
    self.pairs = os.environ.get("KVHBM_SHARDED_PAIRS", "1") != "0"
    self.route = {"send_pairs": t.empty(G * C * 2, **i64), "perm": t.empty(B, **i32),
                  "counts": t.empty(G, **i32), "overflow": t.zeros(1, **i32),
                  "send_ids": t.empty(G * C, **i64), "send_occ": t.empty(G * C, **i32)}
    self.recv_pairs = t.empty(G * C * 2, **i64)
    self.recv_ids, self.recv_occ = t.empty(G * C, **i64), t.empty(G * C, **i32)
    self.rows_owner, self.rows_recv = t.empty(G * C, D, **f32), t.empty(G * C, D, **f32)
    self.out = t.empty(B, D, **f32)
    self.gsum = t.empty(B, D, **f32)
    self.g_send, self.g_recv = t.zeros(G * C, D, **f32), t.empty(G * C, D, **f32)
    self.o_uniq, self.o_idx = t.empty(G * C, **i64), t.empty(G * C, **i32)
    self.o_num = t.zeros(1, **i32)
    self.o_gsum = t.empty(G * C, D, **f32)
    self.wire_bytes = (G - 1) * C * (16 + 2 * 4 * D)  # per step, sent by this rank
    self.side = t.cuda.Stream(device=dev)

  def _a2a(self, out, inp):
    if self.world == 1:
      out.copy_(inp)
    else:
      dist.all_to_all_single(out, inp, group=self.group)

  def run(self, ids, grad, out=None):
    """All NCCL calls stay on the calling stream in a fixed order; work that only depends on
    the ids (the owner-side dedup) runs on a side stream underneath the exchanges and is
    joined where its result is needed.  The gradient is a function of the gathered rows, so
    nothing of the backward half starts before the rows have been expanded."""
    B, G, C, D = self.batch, self.world, self.cap, self.dim
    t = torch
    main = t.cuda.current_stream(self.dev)
    side = self.side
    # ---- forward: dedup, route, exchange, owner lookup, rows back ----
    ops.unique_into(ids, self.uniq, self.idx, self.cnt, self.num)
    if self.pairs:
      ops.route_id_pairs(self.uniq, self.cnt, G, C, self.mode, self.num, self.route)
    else:
      ops.route_ids(self.uniq, self.cnt, G, C, self.mode, num_ids=self.num, out=self.route)
    if self.pairs:
      self._a2a(self.recv_pairs, self.route["send_pairs"])   # ids + occurrence counts together
      ops.unzip_pairs(self.recv_pairs, self.recv_ids, self.recv_occ)
    else:
      self._a2a(self.recv_ids, self.route["send_ids"])
      self._a2a(self.recv_occ, self.route["send_occ"])
    ev_ids = t.cuda.Event()
    ev_ids.record(main)
    with t.cuda.stream(side):      # backward, owner half: dedup what the peers sent
      side.wait_event(ev_ids)
      ops.unique_into(self.recv_ids, self.o_uniq, self.o_idx, None, self.o_num)
      ops.zero_rows(self.o_gsum)
    ops.kv_variable_gather_or_insert_with_counts(self.var, self.recv_ids, self.recv_occ,
                                                 out=self.rows_owner)
    self._a2a(self.rows_recv, self.rows_owner)
    out = self.out if out is None else out
    ops.expand_rows(self.rows_recv, self.route["perm"], self.idx, B, out)
    # ---- [the model runs here: `grad` is a function of `out`] ----
    # ---- backward: sum duplicate gradients, route them, owner merge, fused apply ----
    ops.unsorted_segment_sum(grad, self.idx, self.num, out=self.gsum)
    ops.scatter_rows_n(self.gsum, self.route["perm"], B, self.num, self.g_send)
    main.wait_stream(side)
    self._a2a(self.g_recv, self.g_send)
    ops.unsorted_segment_sum(self.g_recv, self.o_idx, self.o_num, out=self.o_gsum,
                             accumulate=True)
    self._apply(self.o_gsum, self.o_uniq, self.o_num)
    return out

  optimizer = "group_adam"   # or "adam": tfplus-Adam on an [m | v] slot (BASELINE config 4)

  def _apply(self, gsum, uniq, num):
    """The owner's fused apply over the gradient sums of the ids it owns; hpt holds the op's
    scalar inputs in op order, beta^t advanced in the launch."""
    if self.optimizer == "adam":
      ops.kv_variable_sparse_apply_adam_dev(self.var, self.slot, gsum, uniq, self.hpt,
                                            num_indices=num, advance_powers=True)
    else:
      ops.kv_variable_group_sparse_apply_adam_v4_dev(self.var, self.slot, gsum, uniq, self.hpt,
                                                     num_indices=num, advance_powers=True)

  def overflowed(self):
    return bool(self.route["overflow"].item())


class PeerMemory:
  """One symmetric allocation per GPU of the group, each mapped into every rank's address space
  (torch symmetric memory: CUDA VMM handles exchanged through the process-group store).  Only
  plumbing: the library sees raw device pointers."""

  def __init__(self, nbytes, dev, group=None):
    import torch.distributed._symmetric_memory as symm
    group = group or dist.group.WORLD
    self.buf = symm.empty(int(nbytes), dtype=torch.uint8, device=dev)
    self.buf.zero_()
    self.hdl = symm.rendezvous(self.buf, group.group_name)
    self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
    self.dev = dev
    torch.cuda.synchronize(dev)
    dist.barrier(group=group)       # every rank's zero-fill is done before anybody stores

  def local(self, offset, shape, dtype):
    n = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
    return self.buf[offset:offset + n].view(dtype).view(*shape)

  def table(self, offset):
    """Device array of every rank's `base + offset` (rank order)."""
    return torch.tensor([p + int(offset) for p in self.ptrs], dtype=torch.int64, device=self.dev)


class PeerShardedStep(PaddedShardedStep):
  """PaddedShardedStep without NCCL in the data path: every exchange is the stores of the
  kernel that produces the data, written straight into the destination GPU's buffer over
  NVLink (kv_unique_route_peer, kv_gather_or_insert_peer, kv_scatter_rows_n_peer), and the only
  cross-GPU synchronisation is three kv_peer_barrier kernels per step, one per exchange of the
  reference's data flow (ids -> rows -> [model] -> gradients):

    unique+route ==ids==> |A| owner lookup ==rows==> |B| expand rows -> [model] ->
                          | | owner-side dedup       | | local grad sums ==grads==> |C| owner sum, apply

  The gradient of a batch is a function of its gathered rows, so the gradient sums leave a
  requester only after barrier B has delivered the rows and they have been expanded.  Buffer
  reuse across steps needs no extra barrier: a rank stores into a peer's inbox only after
  barrier B of an earlier step, which the peer reaches after its last read of that inbox; into
  a peer's row buffer only after barrier A, which the peer reaches after the expand of the
  step that used the buffer before; into a peer's gradient buffer only after barrier B, which
  the peer reaches after the owner sum that read the buffer before.

  `run` is one step, strictly after the previous one.  `run_rotation` issues a list of steps
  and lets only what depends on the ids alone run ahead: the requester-side dedup + routing of
  step t+1 (the id exchange), barrier A(t+1) and the owner-side dedup(t+1) run under step t.
  Every exchange buffer and both dedup buffer sets exist twice and alternate, barriers A, B
  and C use separate flag channels, everything else is ordered by events and the barriers."""

  def __init__(self, *a, **kw):
    super().__init__(*a, **kw)
    t = torch
    B, G, C, D, r = self.batch, self.world, self.cap, self.dim, self.rank
    al = lambda x: (x + 255) // 256 * 256
    # one symmetric allocation: 2 barrier channels, and two of every exchange buffer (the
    # pipelined schedule alternates them; `run` uses the first of each)
    off, cur = {}, 0
    for name, nbytes in (("flags", 4 * G), ("ids", G * C * 8), ("occ", G * C * 4),
                         ("rows", G * C * D * 4), ("grads", G * C * D * 4)):
      off[name] = []
      for _ in range(3 if name == "flags" else 2):
        off[name].append(cur)
        cur += al(nbytes)
    self.peer = PeerMemory(cur, self.dev, self.group)
    pm = self.peer
    self.flags = [pm.local(o, (G,), t.int32) for o in off["flags"]]
    self.ids_in = [pm.local(o, (G * C,), t.int64) for o in off["ids"]]
    self.occ_in = [pm.local(o, (G * C,), t.int32) for o in off["occ"]]
    self.rows_in = [pm.local(o, (G * C, D), t.float32) for o in off["rows"]]
    self.grads_in = [pm.local(o, (G * C, D), t.float32) for o in off["grads"]]
    self.seg_flags = [pm.table(o) for o in off["flags"]]
    self.seg_ids = [pm.table(o + r * C * 8) for o in off["ids"]]
    self.seg_occ = [pm.table(o + r * C * 4) for o in off["occ"]]
    self.seg_rows = [pm.table(o + r * C * D * 4) for o in off["rows"]]
    self.seg_grads = [pm.table(o + r * C * D * 4) for o in off["grads"]]
    self.bstate = [t.zeros(2, dtype=t.int32, device=self.dev) for _ in range(3)]
    self.barrier_ms = int(os.environ.get("KVHBM_PEER_TIMEOUT_MS", "2000"))
    self.side2 = t.cuda.Stream(device=self.dev)
    self.side3 = t.cuda.Stream(device=self.dev)
    self.side4 = t.cuda.Stream(device=self.dev)
    self.side5 = t.cuda.Stream(device=self.dev)
    self.wire_bytes = (G - 1) * C * (12 + 2 * 4 * D)      # capacity (padded) bytes per step
    # requester-side and owner-side dedup buffers, two sets (set 0 = the ones `run` uses)
    i64 = dict(dtype=t.int64, device=self.dev)
    i32 = dict(dtype=t.int32, device=self.dev)
    self.sets = [dict(uniq=self.uniq, idx=self.idx, cnt=self.cnt, num=self.num, gsum=self.gsum,
                      route=self.route),
                 dict(uniq=t.empty(B, **i64), idx=t.empty(B, **i32), cnt=t.empty(B, **i32),
                      num=t.zeros(1, **i32), gsum=t.empty(B, D, dtype=t.float32, device=self.dev),
                      route={"perm": t.empty(B, **i32), "counts": t.empty(G, **i32),
                             "overflow": self.route["overflow"]})]
    self.osets = [dict(uniq=self.o_uniq, idx=self.o_idx, num=self.o_num),
                  dict(uniq=t.empty(G * C, **i64), idx=t.empty(G * C, **i32),
                       num=t.zeros(1, **i32))]
    self.ws_owner = ops.Workspace(self.dev)     # owner-side dedup runs beside the requester's
    # dedup and routing share their launches (kv_unique_route_peer); the owners' inboxes are
    # padded ahead of time, under the tail of an earlier step
    self.fused_route = os.environ.get("KVHBM_PEER_FUSED_ROUTE", "1") != "0"
    for p in range(2):
      ops.route_fill_peer(G, C, self.seg_ids[p], self.seg_occ[p], self.sets[p]["route"]["counts"])
    t.cuda.synchronize(self.dev)
    dist.barrier(group=self.group)

  def _barrier(self, channel=1):
    """Barriers of one channel must be issued in the same order on every rank; the pipelined
    schedule runs its A barriers (channel 0) on another stream than its B barriers (1)."""
    ops.peer_barrier(self.seg_flags[channel], self.flags[channel], self.bstate[channel],
                     self.rank, self.world, self.barrier_ms)

  def barrier_timeouts(self):
    return sum(int(b[1].item()) for b in self.bstate)

  def _owner_update(self, p=0):
    O = self.osets[p]
    ops.unsorted_segment_sum(self.grads_in[p], O["idx"], O["num"], out=self.o_gsum,
                             accumulate=True)
    self._apply(self.o_gsum, O["uniq"], O["num"])

  def run(self, ids, grad, out=None):
    B, G, C, D = self.batch, self.world, self.cap, self.dim
    t = torch
    out = self.out if out is None else out
    main = t.cuda.current_stream(self.dev)
    s1, s2 = self.side, self.side2
    if self.fused_route:           # dedup + route = the id exchange, in the same launches
      ops.unique_route_peer(ids, self.uniq, self.idx, self.cnt, self.num, G, C, self.mode,
                            self.seg_ids[0], self.seg_occ[0], self.route)
    else:
      ops.unique_into(ids, self.uniq, self.idx, self.cnt, self.num)
    if not self.fused_route:       # route = the id exchange: ids / counts land in the owners' inboxes
      ops.route_fill_peer(G, C, self.seg_ids[0], self.seg_occ[0], self.route["counts"])
      ops.route_ids_peer(self.uniq, self.cnt, G, C, self.mode, self.num, self.seg_ids[0],
                         self.seg_occ[0], self.route)
    self._barrier()                # A: every peer's ids are in my inbox
    ev_a = t.cuda.Event()
    ev_a.record(main)
    with t.cuda.stream(s1):        # owner-side dedup of what the peers sent
      s1.wait_event(ev_a)
      ops.unique_into(self.ids_in[0], self.o_uniq, self.o_idx, None, self.o_num,
                      ws=self.ws_owner)
      ops.zero_rows(self.o_gsum)
    # owner lookup = the row exchange: rows land in the requesters' buffers
    ops.kv_variable_gather_or_insert_peer(self.var, self.ids_in[0], self.occ_in[0],
                                          self.seg_rows[0], C)
    main.wait_stream(s1)           # arriving at B also says: I am done reading my inbox
    self._barrier()                # B: my rows have arrived
    ops.expand_rows(self.rows_in[0], self.route["perm"], self.idx, B, out)
    # ---- [the model runs here: `grad` is a function of `out`] ----
    ops.unsorted_segment_sum(grad, self.idx, self.num, out=self.gsum)
    # gradient exchange: the sums go to the owners' buffers
    ops.scatter_rows_n_peer(self.gsum, self.route["perm"], B, self.num, self.seg_grads[0], C)
    self._barrier()                # C: every peer's gradient sums have arrived
    if self.fused_route:           # every owner has read its inbox (B): pad it for the next step,
      with t.cuda.stream(s1):      # beside the owner update
        s1.wait_stream(main)
        ops.route_fill_peer(G, C, self.seg_ids[0], self.seg_occ[0], self.route["counts"])
    self._owner_update()
    main.wait_stream(s1)
    return out

  def run_rotation(self, ids_list, grad_list, out=None):
    """len(ids_list) (even) consecutive steps; only what depends on the ids alone runs ahead.

    The chain that must run step after step is  lookup(t) -> barrier B(t) -> expand(t) ->
    [model] -> local gradient sum(t) -> gradient exchange(t) -> barrier C(t) -> owner sum(t) ->
    apply(t) -> lookup(t+1); all of it is issued on the calling stream.  Fed to it from the side:
      s_r  requester: dedup+route(t) [= id exchange], step after step;
      s_o  barrier A(t) (channel 0: every peer's ids(t) are in my inbox) and the owner-side
           dedup(t), which therefore run under step t-1;
      s_z  the zero-fill of the owner sums' destination;
      s_e  after barrier B(t): pad the inbox for step t+2.
    Step t uses buffer set t % 2 of everything exchanged.  Who may overwrite what, and why it
    is safe, is argued hazard by hazard in DESIGN.md (section 6)."""
    assert len(ids_list) % 2 == 0 and self.fused_route
    B, G, C, D = self.batch, self.world, self.cap, self.dim
    t = torch
    out = self.out if out is None else out
    main = t.cuda.current_stream(self.dev)
    s_r, s_o, s_e, s_z = self.side, self.side2, self.side3, self.side4
    for S in self.sets:       # the local gradient sums accumulate into cleared buffers
      ops.zero_rows(S["gsum"])
    for s in (s_r, s_o, s_e, s_z):
      s.wait_stream(main)
    ev_g = [None, None]       # expand + gradient exchange of the last step on buffer set p
    ev_apply = [None, None]   # apply of the last step on buffer set p
    for step, (ids, grad) in enumerate(zip(ids_list, grad_list)):
      p = step & 1
      S, O = self.sets[p], self.osets[p]
      with t.cuda.stream(s_r):     # requester: nothing here depends on the table
        if ev_g[p] is not None:
          s_r.wait_event(ev_g[p])  # set p's perm / idx / sums were read, and inbox p padded, by then
        ops.unique_route_peer(ids, S["uniq"], S["idx"], S["cnt"], S["num"], G, C, self.mode,
                              self.seg_ids[p], self.seg_occ[p], S["route"])
        ev_r1 = t.cuda.Event()
        ev_r1.record(s_r)
      with t.cuda.stream(s_o):
        s_o.wait_event(ev_r1)      # my ids are stored before I say so
        self._barrier(0)           # A(t)
        ev_a = t.cuda.Event()
        ev_a.record(s_o)
        if ev_apply[p] is not None:
          s_o.wait_event(ev_apply[p])   # owner set p was read by that apply
        ops.unique_into(self.ids_in[p], O["uniq"], O["idx"], None, O["num"], ws=self.ws_owner)
        ev_ub = t.cuda.Event()
        ev_ub.record(s_o)
      with t.cuda.stream(s_z):     # the sums' destination: free once the previous apply is done
        if ev_apply[1 - p] is not None:
          s_z.wait_event(ev_apply[1 - p])
        ops.zero_rows(self.o_gsum)
        ev_z = t.cuda.Event()
        ev_z.record(s_z)
      # ---- the serial chain ----
      main.wait_event(ev_a)
      ops.kv_variable_gather_or_insert_peer(self.var, self.ids_in[p], self.occ_in[p],
                                            self.seg_rows[p], C)
      main.wait_event(ev_ub)       # arriving at B also says: I am done reading inbox p
      self._barrier(1)             # B(t): my rows have arrived
      ev_b = t.cuda.Event()
      ev_b.record(main)
      # the chain continues on this stream (a hop to another stream costs microseconds)
      ops.expand_rows(self.rows_in[p], S["route"]["perm"], S["idx"], B, out)
      # [the model runs here: `grad` is a function of `out`]
      ops.unsorted_segment_sum(grad, S["idx"], S["num"], out=S["gsum"], accumulate=True)
      ops.scatter_rows_n_peer(S["gsum"], S["route"]["perm"], B, S["num"], self.seg_grads[p], C)
      ev_sc = t.cuda.Event()
      ev_sc.record(main)           # set p's perm / idx / sums have been read
      with t.cuda.stream(s_e):     # off the chain: every owner has read inbox p (they passed B):
        s_e.wait_event(ev_b)       # pad it for step t+2 ...
        ops.route_fill_peer(G, C, self.seg_ids[p], self.seg_occ[p], S["route"]["counts"])
        s_e.wait_event(ev_sc)      # ... and clear set p's sums for step t+2 (the sum accumulates)
        ops.zero_rows(S["gsum"], S["num"])
        ev_g[p] = t.cuda.Event()
        ev_g[p].record(s_e)
      main.wait_event(ev_z)
      self._barrier(2)             # C(t): every peer's gradient sums have arrived
      self._owner_update(p)
      ev_apply[p] = t.cuda.Event()
      ev_apply[p].record(main)
    for s in (s_r, s_o, s_e, s_z):
      main.wait_stream(s)
    return out


def make_padded_step(*a, **kw):
  """The sync-free step over peer memory when the GPUs can map each other (KVHBM_SHARDED_P2P,
  default on), else over NCCL all-to-alls.  Both are device paths; the choice is logged."""
  import sys
  want = os.environ.get("KVHBM_SHARDED_P2P", "1") != "0"
  world = a[4] if len(a) > 4 else kw.get("world", 1)
  if want and world > 1:
    try:
      return PeerShardedStep(*a, **kw)
    except Exception as e:    # no P2P mapping on this box: say so, use the NCCL exchange
      if os.environ.get("KVHBM_SHARDED_P2P") == "1":
        raise
      sys.stderr.write("[kvhbm] peer-memory exchange unavailable (%s: %s); using NCCL\n"
                       % (type(e).__name__, e))
  return PaddedShardedStep(*a, **kw)
