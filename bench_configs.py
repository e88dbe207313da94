"""BASELINE.json configs 1, 3, 4 and 5 behind `bench.py --config {ncf,dcn,sharded200m,streaming}`
(the default, config 2, lives in bench.py).  Same contract as the default: W untimed warm-up
steps, K timed steps bracketed by synchronisation, CUDA events, ONE JSON line on rank 0 with the
config's keys/s, the whole step's algorithmic bytes against the measured HBM peak, and the CPU
arm (oracle port) on a bounded sample of the same workload.

  ncf          example/NCFModel/train.py:36-113: user (943) and item (1682) KvVariables, dim 32,
               batch 256, tfplus-Adam on an [m | v] slot (python/training/adam.py:83-163).
  dcn          example/dcn/train.py:45-72,400-409: 26 KvVariables of the Criteo bucket sizes,
               dim 16, batch 8192 per field, forward lookups NOT deduplicated, SparseGroupFtrl
               (lr 0.1, accumulator 0.1, l1 = l2 = l21 = 1e-5).
  sharded200m  config 4: 25 M keys per GPU (200 M on 8), dim 64, tfplus-Adam slots, the sharded
               step of bench.py with Adam as the owner's apply; on one GPU the plan-driven step.
  streaming    config 5: the microbench table with 10 % never-seen ids per batch,
               enter_threshold 3, Adagrad, every 50 steps DeleteWithTimestamp with the day moved
               on, every 100 steps export -> import round trip compared as key-sorted sets.
"""
import json
import os
import time

import numpy as np

import bench

DCN_BUCKETS = [2500, 2000, 300000, 250000, 1000, 100, 20000, 4000, 20, 100000, 10000, 250000,
               40000, 100, 100, 200000, 50, 10000, 4000, 20, 250000, 100, 100, 250000, 400,
               100000]


def _zipf_ids(n_keys, s, size, rng):
  w = np.arange(1, n_keys + 1, dtype=np.float64) ** (-s)
  cdf = np.cumsum(w)
  cdf /= cdf[-1]
  r = np.searchsorted(cdf, rng.random(size)).astype(np.int64)
  return ((r + 1) * bench.PERM_A) % n_keys


def _peak():
  try:
    return float(json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
  except Exception:
    return 6650.0


def _time_graph_steps(torch, graphs, K, W):
  for i in range(W):
    graphs[i % len(graphs)].replay()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for i in range(K):
    graphs[i % len(graphs)].replay()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / K


def _capture(torch, fn):
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    fn()
  return g


def _emit(args, name, workload, keys_per_step, ms, step_bytes, extra, launches):
  peak = _peak()
  line = {
      "metric": bench.METRIC.replace("GroupAdam", extra.get("optimizer", "GroupAdam")),
      "value": keys_per_step / (ms * 1e-3), "unit": bench.UNIT, "n_gpus": extra.get("n_gpus", 1),
      "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic", "config": dict(workload=workload, name=name),
      "gpu_launches": int(launches),
      "roofline": {"bound": "hbm", "kernel": "whole step", "unit": "GB/s", "peak": peak,
                   "achieved": step_bytes / (ms * 1e-3) / 1e9,
                   "frac": step_bytes / (ms * 1e-3) / 1e9 / peak,
                   "step_algorithmic_bytes_per_gpu": step_bytes, "traffic": None},
  }
  line.update({k: v for k, v in extra.items() if k not in ("optimizer", "n_gpus")})
  bench.emit_json(line)


# ----------------------------------------------------------------------------- config 1
def run_ncf(args):
  import torch
  from tfplus_b200 import ops
  dev = torch.device("cuda", 0)
  torch.cuda.set_device(dev)
  ops.set_today(bench.TODAY)
  D, B, nb = 32, 256, 16
  cards = {"user": 943, "item": 1682}
  rng = np.random.Generator(np.random.PCG64(11))
  tabs = {}
  for name, card in cards.items():
    var = ops.kv_variable(value_shape=[D], device=dev, capacity_hint=4 * card, seed=1)
    mv = ops.kv_variable(value_shape=[2 * D], device=dev, capacity_hint=4 * card, seed=1)
    ops.init_kv_variable_v2(var, torch.from_numpy(bench.init_table(D)).to(dev))
    ops.init_kv_variable_v2(mv, torch.zeros(bench.INIT_ROWS, 2 * D, device=dev))
    hp = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.9, 0.999], dtype=torch.float32, device=dev)
    ids = [torch.from_numpy(rng.integers(1, card + 1, size=B).astype(np.int64)).to(dev)
           for _ in range(nb)]
    grads = [torch.from_numpy(rng.standard_normal((B, D), dtype=np.float32)).to(dev)
             for _ in range(nb)]
    tabs[name] = dict(var=var, mv=mv, hp=hp, ids=ids, grads=grads,
                      plans=[ops.Plan(B, dev) for _ in range(nb)],
                      rows=torch.empty((B, D), dtype=torch.float32, device=dev))

  def step(i):
    for tb in tabs.values():
      tb["plans"][i].build(tb["ids"][i])
      ops.kv_variable_gather_or_insert_plan(tb["var"], tb["plans"][i], out=tb["rows"])
      ops.kv_variable_apply_plan(ops.OPT_ADAM, tb["var"], tb["mv"], None, tb["plans"][i],
                                 tb["grads"][i], tb["hp"], advance_powers=True)

  l0 = ops._lib.launch_count()
  for i in range(nb):
    step(i)
  per_step = (ops._lib.launch_count() - l0) // nb
  torch.cuda.synchronize()
  for tb in tabs.values():
    ops.kv_variable_reserve(tb["var"], 2 * B)
    ops.kv_variable_reserve(tb["mv"], 2 * B)
  graphs = [_capture(torch, lambda i=i: step(i)) for i in range(nb)]
  ms = _time_graph_steps(torch, graphs, args.steps, max(3, args.warmup))
  u = float(np.mean([np.unique(x.cpu().numpy()).size for tb in tabs.values() for x in tb["ids"]]))
  step_bytes = 2 * ((8 + 12 + 4 + 8 * D) * B + 12 * B + 8 * u + B * (4 * D + 4) + 4 * D * u +
                    (8 + 24 + 4 * D + 8 * D + 2 * 2 * 4 * D) * u)
  cpu = _cpu_ncf(D, B, cards)
  _emit(args, "ncf", "example/NCFModel: user (943 keys) + item (1682 keys) KvVariables, dim 32, "
        "batch 256 ids per table, lookup + tfplus-Adam ([m | v] slot) per table per step; "
        "launch-latency bound (%d launches per step in one CUDA graph)" % per_step, 2 * B, ms,
        step_bytes, {"optimizer": "Adam", "cpu_baseline": cpu}, per_step * args.steps)


def _cpu_ncf(D, B, cards, steps=200):
  from oracle import binding as ob
  rng = np.random.Generator(np.random.PCG64(11))
  ob.set_threads(os.cpu_count() or 1)
  tabs = []
  for card in cards.values():
    var = ob.OracleTable(D, 0, seed=1)
    mv = ob.OracleTable(2 * D, 0, seed=1)
    var.set_init_table(bench.init_table(D))
    mv.set_init_table(np.zeros((bench.INIT_ROWS, 2 * D), np.float32))
    tabs.append((var, mv, card))
  b1p, b2p = 0.9, 0.999
  t0 = time.perf_counter()
  for _ in range(steps):
    for var, mv, card in tabs:
      ids = rng.integers(1, card + 1, size=B).astype(np.int64)
      g = rng.standard_normal((B, D), dtype=np.float32)
      var.gather_or_insert(ids, today=bench.TODAY)
      u, idx = ob.unique(ids)
      ob.adam_step(var, mv, u, ob.segment_sum(g, idx, u.size), 1e-3, 0.9, 0.999, 1e-8, b1p, b2p,
                   today=bench.TODAY)
    b1p *= 0.9
    b2p *= 0.999
  dt = time.perf_counter() - t0
  return {"value": 2 * B * steps / dt, "unit": bench.UNIT, "cores": os.cpu_count() or 1,
          "kind": "port", "sample": "%d steps of the same workload (input generation included), "
                                     "oracle port" % steps}


# ----------------------------------------------------------------------------- config 3
def run_dcn(args):
  import torch
  from tfplus_b200 import ops
  dev = torch.device("cuda", 0)
  torch.cuda.set_device(dev)
  ops.set_today(bench.TODAY)
  D, B, nb = 16, 8192, 8
  rng = np.random.Generator(np.random.PCG64(12))
  hp = torch.tensor([0.1, 1e-5, 1e-5, 1e-5, 0.0, -0.5], dtype=torch.float32, device=dev)
  fields = []
  u_sum = 0.0
  for card in DCN_BUCKETS:
    var = ops.kv_variable(value_shape=[D], device=dev, capacity_hint=2 * card + B, seed=1)
    acc = ops.kv_variable(value_shape=[D], device=dev, capacity_hint=2 * card + B, seed=1)
    lin = ops.kv_variable(value_shape=[D], device=dev, capacity_hint=2 * card + B, seed=1)
    ops.init_kv_variable_v2(var, torch.from_numpy(bench.init_table(D)).to(dev))
    ops.init_kv_variable_v2(acc, torch.full((16, D), 0.1, device=dev))
    ops.init_kv_variable_v2(lin, torch.zeros(16, D, device=dev))
    ids_np = [_zipf_ids(card, 1.05, B, rng) for _ in range(nb)]
    u_sum += float(np.mean([np.unique(x).size for x in ids_np]))
    fields.append(dict(var=var, acc=acc, lin=lin,
                       ids=[torch.from_numpy(x).to(dev) for x in ids_np],
                       plans=[ops.Plan(B, dev) for _ in range(nb)]))
  grads = [torch.from_numpy(rng.standard_normal((B, D), dtype=np.float32)).to(dev)
           for _ in range(nb)]
  # the 26 fields are independent tables: four streams inside the graph (a field's dedup plan
  # is built on its stream ahead of the backward half; each stream has its own dedup scratch)
  NS = 4
  streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
  wss = [ops.Workspace(dev) for _ in range(NS)]
  rows_s = [torch.empty((B, D), dtype=torch.float32, device=dev) for _ in range(NS)]

  def step(i):
    main = torch.cuda.current_stream(dev)
    for k, st_ in enumerate(streams):
      st_.wait_stream(main)
      with torch.cuda.stream(st_):
        mine = fields[k::NS]
        for f in mine:   # forward: raw ids, duplicates hit the lookup (no dedup in the example)
          ops.kv_variable_gather_or_insert_v2(f["var"], f["ids"][i], out=rows_s[k])
        for f in mine:   # backward: Unique + UnsortedSegmentSum + SparseGroupFtrl, fused
          f["plans"][i].build(f["ids"][i], ws=wss[k])
          ops.kv_variable_apply_plan(ops.OPT_SPARSE_GROUP_FTRL, f["var"], f["acc"], f["lin"],
                                     f["plans"][i], grads[i], hp)
    for st_ in streams:
      main.wait_stream(st_)

  l0 = ops._lib.launch_count()
  for i in range(nb):
    step(i)
  per_step = (ops._lib.launch_count() - l0) // nb
  torch.cuda.synchronize()
  for f in fields:
    for tb in (f["var"], f["acc"], f["lin"]):
      ops.kv_variable_reserve(tb, 2 * B)
  graphs = [_capture(torch, lambda i=i: step(i)) for i in range(nb)]
  ms = _time_graph_steps(torch, graphs, args.steps, max(3, args.warmup))
  nf = len(DCN_BUCKETS)
  step_bytes = nf * ((8 + 12 + 4 + 8 * D) * B + 12 * B + B * (4 * D + 4)) + u_sum * (
      8 + 4 * D + (8 + 3 * 12 + 4 * D + 8 * D + 2 * 2 * 4 * D))
  cpu = _cpu_dcn(D, B, 3)
  _emit(args, "dcn", "example/dcn: 26 KvVariables (Criteo hash-bucket sizes, 1.7 M keys), dim 16, "
        "batch 8192 Zipf(1.05) ids per field, forward lookups on the raw ids (no dedup), "
        "backward Unique + UnsortedSegmentSum + SparseGroupFtrl (lr 0.1, l1=l2=l21=1e-5) fused per "
        "field; %d launches per step in one CUDA graph" % per_step, nf * B, ms, step_bytes,
        {"optimizer": "SparseGroupFtrl", "cpu_baseline": cpu, "unique_per_step": u_sum},
        per_step * args.steps)


def _cpu_dcn(D, B, steps):
  from oracle import binding as ob
  rng = np.random.Generator(np.random.PCG64(12))
  ob.set_threads(os.cpu_count() or 1)
  fields = []
  for card in DCN_BUCKETS:
    var, acc, lin = (ob.OracleTable(D, 0, seed=1) for _ in range(3))
    var.set_init_table(bench.init_table(D))
    acc.set_init_table(np.full((16, D), 0.1, np.float32))
    lin.set_init_table(np.zeros((16, D), np.float32))
    fields.append((var, acc, lin, [_zipf_ids(card, 1.05, B, rng) for _ in range(steps + 1)]))
  g = rng.standard_normal((B, D), dtype=np.float32)
  def one(i):
    for var, acc, lin, ids in fields:
      var.gather_or_insert(ids[i], today=bench.TODAY)
    for var, acc, lin, ids in fields:
      u, idx = ob.unique(ids[i])
      ob.apply_sparse_group_ftrl(var, acc, lin, u, ob.segment_sum(g, idx, u.size), 0.1, 1e-5,
                                 1e-5, 1e-5, 0.0, -0.5, today=bench.TODAY)
  one(0)
  t0 = time.perf_counter()
  for i in range(1, steps + 1):
    one(i)
  dt = time.perf_counter() - t0
  return {"value": len(DCN_BUCKETS) * B * steps / dt, "unit": bench.UNIT,
          "cores": os.cpu_count() or 1, "kind": "port",
          "sample": "%d steps of the same workload after 1 warm-up, oracle port" % steps}


# ----------------------------------------------------------------------------- config 4
def run_sharded200m(args):
  """25 M keys per GPU with tfplus-Adam slots: (256 + 512) B x 25 M = 19.2 GB of rows per GPU
  (153.6 GB at 8 GPUs) + 0.8 GB of slots.  --keys overrides the keys per GPU."""
  import torch
  import torch.distributed as dist
  from tfplus_b200 import ops
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  ops.set_today(bench.TODAY)
  keys = args.keys if args.keys != bench.KEYS else 25_000_000
  D, B, nb = bench.DIM, bench.BATCH, bench.N_BATCHES
  K, W = max(1, args.steps), max(3, args.warmup)
  ids_np, grads_np = bench.make_batches(nb, keys * world, B, D, seed_ids=2024 + rank,
                                        seed_grad=7 + rank)
  ids_d = [torch.from_numpy(x).to(dev) for x in ids_np]
  grads_d = [torch.from_numpy(x).to(dev) for x in grads_np]
  u_meas = float(np.mean([np.unique(x).size for x in ids_np]))
  hp_adam = [1e-3, 0.9, 0.999, 1e-8, 0.9, 0.999]
  t_pop = time.perf_counter()
  if world > 1:
    st = bench.ShardedStepper(keys, D, B, bench.HP, dev, rank, world, slot_mult=2)
    st.hpt = torch.tensor(hp_adam, dtype=torch.float32, device=dev)
    st.optimizer = "adam"
    st.populate()
    st.prepare(ids_d, grads_d)
    run, launches_per_step = st.run_steps, st.launches_per_step
  else:
    var = ops.kv_variable(value_shape=[D], device=dev, capacity_hint=keys + B, seed=1)
    mv = ops.kv_variable(value_shape=[2 * D], device=dev, capacity_hint=keys + B, seed=1)
    ops.init_kv_variable_v2(var, torch.from_numpy(bench.init_table(D)).to(dev))
    ops.init_kv_variable_v2(mv, torch.zeros(bench.INIT_ROWS, 2 * D, device=dev))
    for s in range(0, keys, 1 << 20):
      ids = torch.arange(s, min(keys, s + (1 << 20)), dtype=torch.int64, device=dev)
      ops.kv_variable_gather_or_insert_v2(var, ids)
      ops.kv_variable_gather_or_insert_v2(mv, ids)
    hp = torch.tensor(hp_adam, dtype=torch.float32, device=dev)
    plans = [ops.Plan(B, dev) for _ in range(nb)]
    rows = torch.empty((B, D), dtype=torch.float32, device=dev)
    def step(i):
      plans[i].build(ids_d[i])
      ops.kv_variable_gather_or_insert_plan(var, plans[i], out=rows)
      ops.kv_variable_apply_plan(ops.OPT_ADAM, var, mv, None, plans[i], grads_d[i], hp,
                                 advance_powers=True)
    l0 = ops._lib.launch_count()
    for i in range(nb):
      step(i)
    launches_per_step = (ops._lib.launch_count() - l0) // nb
    torch.cuda.synchronize()
    ops.kv_variable_reserve(var, 2 * B)
    ops.kv_variable_reserve(mv, 2 * B)
    graphs = [_capture(torch, lambda i=i: step(i)) for i in range(nb)]
    def run(k):
      for i in range(k):
        graphs[i % nb].replay()
  torch.cuda.synchronize()
  t_pop = time.perf_counter() - t_pop
  run(W)
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  run(K)
  b.record()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / K
  if world > 1:
    tt = torch.tensor([ms], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
  free, total = torch.cuda.mem_get_info(dev)
  if rank == 0:
    step_bytes = ((8 + 12 + 4 + 8 * D) * B + 12 * B + 8 * u_meas + B * (4 * D + 4) +
                  4 * D * u_meas + (8 + 24 + 4 * D + 8 * D + 2 * 2 * 4 * D) * u_meas)
    _emit(args, "sharded200m", "%d-key int64 table per GPU x %d GPU(s) (%.0f M keys in all), dim "
          "%d, tfplus-Adam [m | v] slots: %.1f GB of rows per GPU; batch %d Zipf(1.1) ids per GPU; "
          "%s" % (keys, world, keys * world / 1e6, D, keys * 3 * D * 4 / 1e9, B,
                  "key-hash sharding, ids / rows / gradients over NVLink" if world > 1
                  else "single GPU, plan-driven step"),
          B * world, ms, step_bytes,
          {"optimizer": "Adam", "n_gpus": world, "unique_per_step": u_meas,
           "hbm_used_gb": (total - free) / 1e9, "populate_and_capture_s": t_pop,
           "cpu_baseline": None}, launches_per_step * K)
  if world > 1:
    st.release()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


# ----------------------------------------------------------------------------- config 5
def run_streaming(args):
  import torch
  from tfplus_b200 import ops
  dev = torch.device("cuda", 0)
  torch.cuda.set_device(dev)
  keys = args.keys
  D, B = bench.DIM, bench.BATCH
  K = max(1, args.steps)
  new_per_step = B // 10
  day = bench.TODAY
  ops.set_today(day)
  var = ops.kv_variable(value_shape=[D], enter_threshold=3, device=dev, capacity_hint=keys + B,
                        seed=1)
  acc = ops.kv_variable(value_shape=[D], device=dev, capacity_hint=keys + B, seed=1)
  ops.init_kv_variable_v2(var, torch.from_numpy(bench.init_table(D)).to(dev))
  ops.init_kv_variable_v2(acc, torch.full((16, D), 0.1, device=dev))
  for s in range(0, keys, 1 << 20):
    ids = torch.arange(s, min(keys, s + (1 << 20)), dtype=torch.int64, device=dev)
    for _ in range(3):                       # seen three times: above the entry threshold
      ops.kv_variable_gather_or_insert_v2(var, ids)
    ops.kv_variable_gather_or_insert_v2(acc, ids)
  z = bench.ZipfSampler(keys, bench.ZIPF_S, 2024)
  rng = np.random.Generator(np.random.PCG64(7))
  nb = 16
  old = [z.ids(B - new_per_step) for _ in range(nb)]
  grads = [torch.from_numpy(rng.standard_normal((B, D), dtype=np.float32)).to(dev)
           for _ in range(nb)]
  plan = ops.Plan(B, dev)
  rows = torch.empty((B, D), dtype=torch.float32, device=dev)
  hp = (0.05,)
  evicted = checked = 0

  def batch(step):
    fresh = keys + step * new_per_step + np.arange(new_per_step, dtype=np.int64)
    ids = np.concatenate([old[step % nb], fresh])
    return torch.from_numpy(ids).to(dev, non_blocking=True)

  def one(step):
    nonlocal day, evicted, checked
    ids = batch(step)
    plan.build(ids)
    ops.kv_variable_gather_or_insert_plan(var, plan, out=rows)
    ops.kv_variable_apply_plan(ops.OPT_ADAGRAD, var, acc, None, plan, grads[step % nb], hp)
    if step % 50 == 49:                       # eviction: the day moves on, stale keys go
      day += 1
      ops.set_today(day)
      if step % 100 == 99:
        evicted += int(ops.kv_variable_delete_with_timestamp(var, 2).numel())
    if step % 100 == 99:                      # checkpoint round trip, compared as sorted sets
      ex = ops.kv_variable_export(var, first_n=6, enable_cutoff=True, cutoff_value=1e-20,
                                  freq_dtype=torch.int32)
      ops.kv_variable_import(var, *ex, first_n=6)
      ex2 = ops.kv_variable_export(var, first_n=6, enable_cutoff=True, cutoff_value=1e-20,
                                   freq_dtype=torch.int32)
      o1, o2 = torch.argsort(ex[0]), torch.argsort(ex2[0])
      assert torch.equal(ex[0][o1], ex2[0][o2]) and torch.equal(ex[1][o1], ex2[1][o2])
      assert torch.equal(torch.sort(ex[3]).values, torch.sort(ex2[3]).values)
      # frequency table: keys without a row (low-frequency, all-zero) are not restored by the
      # import (dynamic_restore.hpp:217-245 only updates keys that exist); the survivors keep
      # their words
      f1, f2 = torch.argsort(ex[4]), torch.argsort(ex2[4])
      k1, k2 = ex[4][f1], ex2[4][f2]
      at = torch.searchsorted(k1, k2)
      assert torch.equal(k1[at], k2) and torch.equal(ex[5][f1][at], ex2[5][f2])
      checked += 1

  l0 = ops._lib.launch_count()
  for s in range(max(3, args.warmup)):
    one(s)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for s in range(K):
    one(max(3, args.warmup) + s)
  b.record()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / K
  wall = (time.perf_counter() - t0) / K * 1e3
  u = float(np.unique(np.concatenate([old[0], keys + np.arange(new_per_step)])).size)
  step_bytes = ((8 + 12 + 4 + 8 * D) * B + 12 * B + 8 * u + B * (4 * D + 4) + 4 * D * u +
                (8 + 24 + 4 * D + 8 * D + 2 * 4 * D) * u)
  size = ops.kv_variable_size_v2(var)
  _emit(args, "streaming", "microbench table (%d keys, dim %d, enter_threshold 3) with %d never-seen "
        "ids in every batch of %d, lookup + Adagrad (lr 0.05); every 50 steps the day moves on, "
        "every 100 steps DeleteWithTimestamp(2 days) and an export -> import -> export round trip "
        "compared as key-sorted sets; eager launches (the table grows, no CUDA graph), host -> "
        "device copy of the ids inside the timed region" % (keys, D, new_per_step, B),
        B, ms, step_bytes,
        {"optimizer": "Adagrad", "unique_per_step": u, "wall_ms_per_step": wall,
         "checkpoint_round_trips": checked, "keys_evicted": evicted, "table_size_after": size,
         "cpu_baseline": None}, ops._lib.launch_count() - l0)


RUNNERS = {"ncf": run_ncf, "dcn": run_dcn, "sharded200m": run_sharded200m,
           "streaming": run_streaming}
