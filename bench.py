#!/usr/bin/env python
"""KvVariable microbench (BASELINE.json configs[1]): lookup + GroupAdam apply keys/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of B = 65 536 Zipf(1.1) int64 ids on a
10 M-key, dim-64 table: KvVariableGatherOrInsertV2 on all B ids (forward, not deduped), then
what TF's optimizer does with the IndexedSlices gradient: Unique + UnsortedSegmentSum, then
KvVariableGroupSparseApplyAdamV4 on the unique ids.  One JSON line on stdout (rank 0).

N > 1: the table is sharded by key hash, every rank brings its own B ids (weak scaling), ids /
rows / gradients cross NVLink by all-to-all (tfplus_b200/sharded.py).

--impl reference times the reference's CPU algorithm (the oracle port: TensorFlow and the
reference's Bazel build are not available) on the host cores, same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

KEYS = 10_000_000
DIM = 64
BATCH = 65536
INIT_ROWS = 10000
ZIPF_S = 1.1
PERM_A = 7368787            # rank -> id: (rank * A) mod KEYS, A coprime to 10^7
N_BATCHES = 16              # rotating input batches (16 x 17.3 MB = 277 MB > 126 MB L2)
HP = dict(lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8, l1=1e-5, l2=1e-5, l21=1e-5)
TODAY = 20000
METRIC = "KvVariable lookup+GroupAdam apply keys/s"
UNIT = "keys/s"


# ----------------------------------------------------------------------------- workload
class ZipfSampler:
  """Inverse-CDF Zipf(s) over ranks 1..n, numpy PCG64."""

  def __init__(self, n, s, seed):
    w = np.arange(1, n + 1, dtype=np.float64) ** (-s)
    self.cdf = np.cumsum(w)
    self.cdf /= self.cdf[-1]
    self.n = n
    self.rng = np.random.Generator(np.random.PCG64(seed))

  def ids(self, b, offset=0):
    r = np.searchsorted(self.cdf, self.rng.random(b)).astype(np.int64)
    return ((r + 1) * PERM_A) % self.n + offset


def make_batches(n_batches, keys, batch, dim, seed_ids=2024, seed_grad=7, offset=0):
  z = ZipfSampler(keys, ZIPF_S, seed_ids)
  g = np.random.Generator(np.random.PCG64(seed_grad))
  ids = [z.ids(batch, offset) for _ in range(n_batches)]
  grads = [g.standard_normal((batch, dim), dtype=np.float32) for _ in range(n_batches)]
  return ids, grads


def init_table(dim):
  return np.random.Generator(np.random.PCG64(1234)).normal(0, 0.05, (INIT_ROWS, dim)).astype(
      np.float32)


def algorithmic_bytes(b, u, d):
  """SURVEY.md 8(d) / BASELINE.md 3 per-stage algorithmic bytes."""
  return {
      "gather": (8 + 12 + 4 + 4 * d + 4 * d) * b,
      "unique": 8 * b + 8 * u + 4 * b,
      "segment_sum": b * (4 * d + 4) + 4 * d * u,
      "apply": (8 + 2 * 12 + 4 * d + 2 * 4 * d + 2 * 3 * 4 * d) * u,
  }


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
  """Samples SM clock and throttle reasons through NVML while the timed region runs."""

  def __init__(self, index, enabled=True, period_s=0.002):
    self.samples, self.reasons, self.max_mhz = [], set(), None
    self._stop = threading.Event()
    self._t = None
    self.period_s = period_s
    self.nv = None
    if not enabled:
      return
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:
      self.nv = None

  def _once(self):
    nv = self.nv
    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
    names = {
        "hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
        "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
        "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
        "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
    }
    for k, bit in names.items():
      if r & bit:
        self.reasons.add(k)

  def start(self):
    if not self.nv:
      return
    def run():
      while not self._stop.is_set():
        try:
          self._once()
        except Exception:
          return
        time.sleep(self.period_s)
    self._t = threading.Thread(target=run, daemon=True)
    self._t.start()

  def stop(self):
    if not self.nv:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
    try:
      self._once()
    except Exception:
      pass
    self._stop.set()
    if self._t:
      self._t.join()
    med = float(np.median(self.samples)) if self.samples else None
    return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
            "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU arm
def run_cpu(steps, warmup, keys, batch, dim, threads=None, quiet=False):
  """The reference algorithm on the host cores (oracle port).  Returns (keys/s, info)."""
  from oracle import binding as ob
  threads = threads or os.cpu_count() or 1
  ob.set_threads(threads)
  var = ob.OracleTable(dim, 0, seed=1)
  slot = ob.OracleTable(3 * dim, 0, seed=1)
  var.set_init_table(init_table(dim))
  slot.set_init_table(np.zeros((INIT_ROWS, 3 * dim), np.float32))
  for s in range(0, keys, 1 << 20):
    ids = np.arange(s, min(keys, s + (1 << 20)), dtype=np.int64)
    var.gather_or_insert(ids, today=TODAY)
    slot.gather_or_insert(ids, today=TODAY)
  nb = min(N_BATCHES, steps + warmup)
  ids_l, grads_l = make_batches(nb, keys, batch, dim)
  b1p, b2p = HP["beta1"], HP["beta2"]
  t0 = None
  for i in range(warmup + steps):
    if i == warmup:
      t0 = time.perf_counter()
    ids, g = ids_l[i % nb], grads_l[i % nb]
    var.gather_or_insert(ids, today=TODAY)
    u, idx = ob.unique(ids)                      # TF Unique (single-threaded CPU kernel)
    gs = ob.segment_sum(g, idx, u.size)          # TF UnsortedSegmentSum
    ob.apply_group_adam_v4(var, slot, u, gs, HP["lr"], b1p, b2p, HP["beta1"], HP["beta2"],
                           HP["epsilon"], HP["l1"], HP["l2"], HP["l21"], today=TODAY)
    b1p *= HP["beta1"]
    b2p *= HP["beta2"]
  dt = time.perf_counter() - t0
  return batch * steps / dt, {"seconds": dt, "threads": threads}


def main_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return 0
  steps = max(1, args.steps)
  cores = os.cpu_count() or 1
  val, info = run_cpu(steps, args.warmup, args.keys, args.batch, args.dim, cores)
  sample = "%d-key table, %d timed steps of B=%d (whole workload, no subsampling)" % (
      args.keys, steps, args.batch)
  line = {
      "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
      "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * info["seconds"] / steps,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic", "config": workload_config(args, 1),
      "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": sample},
      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  emit_json(line)
  return 0


def workload_config(args, n_gpus):
  return {
      "workload": "KvVariable microbench: %d-key int64 table per GPU x %d GPU(s), dim %d, batch %d "
                  "Zipf(%.1f) ids per GPU, gather_or_insert + unique + segment_sum + "
                  "GroupAdam v4 apply (lr 1e-3, l1=l2=l21=1e-5)" % (
                      args.keys, n_gpus, args.dim, args.batch, ZIPF_S),
      "keys": args.keys * n_gpus, "dim": args.dim, "batch_per_gpu": args.batch,
      "global_batch": args.batch * n_gpus,
      "parallelism": "single GPU" if n_gpus == 1 else "key-hash sharding x%d" % n_gpus,
      "l2": "table working set %.1f GB >> 126 MB L2; %d rotating input batches (%.0f MB)" % (
          args.keys * args.dim * 4 * 4 / 1e9, N_BATCHES,
          N_BATCHES * args.batch * (8 + 4 * args.dim) / 1e6),
  }


# ----------------------------------------------------------------------------- GPU arm
def main_ours(args):
  import torch
  import torch.distributed as dist
  from tfplus_b200 import _lib, ops

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))

  def note(msg):
    if os.environ.get("KVHBM_BENCH_VERBOSE"):
      sys.stderr.write("[rank %d] %s\n" % (rank, msg))
      sys.stderr.flush()

  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); "
                     "use --impl reference for the CPU arm")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  ops.set_today(TODAY)
  K, W = max(1, args.steps), max(3, args.warmup)
  keys, B, D = args.keys, args.batch, args.dim

  if world > 1:
    from tfplus_b200 import sharded
    stepper = sharded.ShardedStepper(keys, D, B, HP, dev, rank, world)
  else:
    stepper = LocalStepper(keys, D, B, dev)
  note('created')
  stepper.populate()
  note('populated')

  nb = N_BATCHES
  ids_np, grads_np = make_batches(nb, keys * world, B, D, seed_ids=2024 + rank,
                                  seed_grad=7 + rank)
  ids_d = [torch.from_numpy(x).to(dev) for x in ids_np]
  grads_d = [torch.from_numpy(x).to(dev) for x in grads_np]
  u_meas = float(np.mean([np.unique(x).size for x in ids_np]))

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- warm-up + timed region (inputs resident in HBM) ----
  stepper.prepare(ids_d, grads_d)
  note('prepared')
  for i in range(W):
    stepper.step(i)
  pipelined = getattr(stepper, "rotation", None) is not None
  if pipelined:
    stepper.run_steps(N_BATCHES)   # first replay of the rotation graph (upload) is warm-up too
  # NVML is initialised BEFORE the barrier: a rank still inside nvmlInit when its peers start
  # their timed loop shows up in their step time through the first exchange (this cost 20-60 %
  # at N = 2..8 until it was found; the polling itself is harmless, scripts/sampler_probe.py).
  # One sampler per node is enough.
  clocks = ClockSampler(local_rank, enabled=(rank == 0),
                        period_s=float(os.environ.get("KVHBM_BENCH_CLOCK_PERIOD_MS", "2")) * 1e-3)
  barrier()
  clocks.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  h0 = time.perf_counter()
  if hasattr(stepper, "run_steps"):
    stepper.run_steps(K)
  else:
    for i in range(K):
      stepper.step(i)
  host_us = (time.perf_counter() - h0) / K * 1e6   # host time to enqueue one step
  e1.record()
  barrier()
  launches = K * stepper.launches_per_step   # graph replays: counted at capture time
  ms = e0.elapsed_time(e1)
  clk = clocks.stop()
  if world > 1:
    tt = torch.tensor([ms], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
  value = B * world * K / (ms * 1e-3)

  # ---- the same K steps strictly one after the other (every step waits for the previous
  # one to finish entirely): reported next to the pipelined schedule ----
  strict = None
  if pipelined:
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(K):
      stepper.step(i)
    g1.record()
    barrier()
    sms = g0.elapsed_time(g1)
    if world > 1:
      tt = torch.tensor([sms], device=dev)
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      sms = float(tt.item())
    strict = {"ms_per_step": sms / K, "value": B * world * K / (sms * 1e-3), "unit": UNIT,
              "note": "one CUDA graph per step, step t+1 starts when step t has finished"}

  # ---- per-stage device times (same steps, events between the stages) ----
  note('timed %.3f ms/step' % (ms / K))
  stage_ms = stepper.stage_times(K)
  note('stages')

  # ---- end to end: host buffers in, host rows out, copies inside the timed region ----
  ids_h = [torch.from_numpy(x).pin_memory() for x in ids_np]
  grads_h = [torch.from_numpy(x).pin_memory() for x in grads_np]
  rows_h = torch.empty((B, D), dtype=torch.float32).pin_memory()
  stepper.prepare_host(ids_h, grads_h, rows_h)
  for i in range(3):
    stepper.step_host(i)
  stepper.finish_host()
  barrier()
  f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  f0.record()
  for i in range(K):
    stepper.step_host(i)
  stepper.finish_host()
  f1.record()
  barrier()
  e2e_ms = f0.elapsed_time(f1)
  if world > 1:
    tt = torch.tensor([e2e_ms], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_ms = float(tt.item())
  e2e_val = B * world * K / (e2e_ms * 1e-3)
  note('e2e')

  if rank == 0:
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
      pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
    ab = algorithmic_bytes(B, u_meas, D)
    if world > 1:
      # every rank does the single-GPU step's table traffic on its shard
      ab = {"step": sum(ab.values()) * world}
      stage_ms = dict(stage_ms, step=ms / K)
    kern = {}
    for name, t_ms in stage_ms.items():
      if name in ab and t_ms > 0:
        kern[name] = {"ms": t_ms, "algorithmic_bytes": ab[name],
                      "achieved_gbs": ab[name] / (t_ms * 1e-3) / 1e9,
                      "frac": ab[name] / (t_ms * 1e-3) / 1e9 / peak}
    # the dominant KERNEL: stages are 1 (gather, apply), 2 (zero + segment-sum) or 3 (unique)
    # launches, so compare per-launch time, not stage time
    for name in kern:
      kern[name]["launches"] = STAGE_LAUNCHES.get(name, 1)
    top = max((k for k in kern), key=lambda k: kern[k]["ms"] / kern[k]["launches"]) if kern else None
    roof = None
    if top:
      roof = {"bound": "hbm", "kernel": top, "achieved": kern[top]["achieved_gbs"], "peak": peak,
              "unit": "GB/s", "frac": kern[top]["frac"], "traffic": TRAFFIC.get(top),
              "peak_source": peak_src, "stages": kern,
              "step_algorithmic_bytes": sum(ab.values()),
              "step_frac_of_peak": sum(ab.values()) / (ms / K * 1e-3) / 1e9 / peak,
              "step_frac_of_8TBs": sum(ab.values()) / (ms / K * 1e-3) / 1e9 / 8000.0}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "unique_per_step": u_meas, "clocks": clk, "gpu_launches": int(launches),
        "host_enqueue_us_per_step": host_us,
        "schedule": ("rotation graphs of %d steps with only the true dependencies between "
                     "consecutive steps (dedup + gradient sum of batch t+1 run under the "
                     "lookup/apply of batch t); remainder step by step" % N_BATCHES)
                    if pipelined else "one CUDA graph per step",
        "strict_per_step": strict,
        "e2e": {"value": e2e_val, "unit": UNIT,
                "h2d_bytes_per_step": int(B * 8 + B * D * 4),
                "d2h_bytes_per_step": int(B * D * 4), "ms_per_step": e2e_ms / K},
        "roofline": roof,
    }
    if world > 1:
      sent = stepper.padded.wire_bytes
      peer = hasattr(stepper.padded, "barrier_timeouts")
      if peer and stepper.padded.barrier_timeouts():
        raise SystemExit("peer barrier timed out %d times: the step is invalid"
                         % stepper.padded.barrier_timeouts())
      line["nvlink"] = {"exchange": "peer-memory stores fused into the producing kernels + "
                                    "2 barrier kernels" if peer else "NCCL all_to_all_single x3",
                        "bytes_sent_per_gpu_per_step": sent,
                        "bus_gbs_per_gpu": sent / (ms / K * 1e-3) / 1e9,
                        "peak_gbs_per_direction": 900.0, "measured_peer_copy_gbs": 770.0,
                        "exchanges_per_step": 3, "barriers_per_step": 2 if peer else None,
                        "capacity_per_peer": stepper.padded.cap,
                        "overflowed": stepper.padded.overflowed(),
                        "note": "fixed-capacity exchange of {id, occurrence count} pairs, rows, "
                                "gradients (bytes = capacity upper bound), captured with the "
                                "kernels in one CUDA graph"}
      line["stage_ms"] = stage_ms
    if world == 1 and not args.no_cpu:
      cores = os.cpu_count() or 1
      cs = max(3, min(20, K))
      v, info = run_cpu(cs, 2, keys, B, D, cores)
      line["cpu_baseline"] = {
          "value": v, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": "same workload (%d-key table), %d timed steps of B=%d after 2 warm-up, "
                    "oracle port of the reference algorithm, %d threads" % (keys, cs, B, cores)}
    emit_json(line)
  if world > 1:
    stepper.release()      # captured NCCL work must be gone before the process group
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
  return 0


# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full
# captures of round 1 (profiles/r01_ncu_kernels.txt); None where no capture exists
STAGE_LAUNCHES = {"gather": 1, "unique": 3, "segment_sum": 2, "apply": 1, "step": 1}
TRAFFIC = {"apply": 31.52e6, "gather": 8.33e6, "segment_sum": 22.23e6, "unique": 4.3e6}


class LocalStepper:
  """One GPU: var + m_v_linear tables and the four-stage step.

  The optimizer's scalar inputs live in device memory (`hp`, as TF would hand them to a GPU
  kernel) and beta^t is advanced on the device by every apply launch (TF Adam's `_finish`), so
  whole steps are capturable: the timed loops replay CUDA graphs (one per rotation of the
  batches, or one per step) instead of paying ~10 host launches per step.
  """

  STAGES = ["gather", "unique", "segment_sum", "apply"]

  def __init__(self, keys, dim, batch, dev):
    import torch
    from tfplus_b200 import ops
    self.torch, self.ops = torch, ops
    self.keys, self.dim, self.batch, self.dev = keys, dim, batch, dev
    self.var = ops.kv_variable(value_shape=[dim], device=dev, capacity_hint=keys + batch, seed=1)
    self.slot = ops.kv_variable(value_shape=[3 * dim], device=dev, capacity_hint=keys + batch,
                                seed=1)
    ops.init_kv_variable_v2(self.var, torch.from_numpy(init_table(dim)).to(dev))
    ops.init_kv_variable_v2(self.slot, torch.zeros(INIT_ROWS, 3 * dim, device=dev))
    self.hp = torch.tensor([HP["lr"], HP["beta1"], HP["beta2"], HP["beta1"], HP["beta2"],
                            HP["epsilon"], HP["l1"], HP["l2"], HP["l21"]], dtype=torch.float32,
                           device=dev)
    self.graphs = {}
    self.overlap = os.environ.get("KVHBM_BENCH_OVERLAP", "1") != "0"
    self.side = torch.cuda.Stream(device=dev)
    self.side2 = torch.cuda.Stream(device=dev)

  def populate(self):
    torch, ops = self.torch, self.ops
    for s in range(0, self.keys, 1 << 20):
      ids = torch.arange(s, min(self.keys, s + (1 << 20)), dtype=torch.int64, device=self.dev)
      ops.kv_variable_gather_or_insert_v2(self.var, ids)
      ops.kv_variable_gather_or_insert_v2(self.slot, ids)
    torch.cuda.synchronize()

  # ---- the step, stage by stage (eager; also what gets captured) ----
  def new_buffers(self):
    t, B, D, dev = self.torch, self.batch, self.dim, self.dev
    return {"rows": t.empty((B, D), dtype=t.float32, device=dev),
            "uniq": t.empty(B, dtype=t.int64, device=dev),
            "idx": t.empty(B, dtype=t.int32, device=dev),
            "num": t.zeros(1, dtype=t.int32, device=dev),
            "gsum": t.empty((B, D), dtype=t.float32, device=dev)}

  def stage(self, name, ids, grad, buf):
    ops, lib, C = self.ops, self.ops._lib.load(), self.ops.check
    st = self.torch.cuda.current_stream(self.dev).cuda_stream
    if name == "gather":
      ops.kv_variable_gather_or_insert_v2(self.var, ids, out=buf["rows"])
    elif name == "unique":
      ws = ops.Workspace.get(self.dev)
      C(lib.kv_unique(ws.ptr, ids.data_ptr(), ids.numel(), buf["uniq"].data_ptr(),
                      buf["idx"].data_ptr(), None, buf["num"].data_ptr(), st))
    elif name == "zero":
      ops.zero_rows(buf["gsum"])
    elif name == "segment_sum":
      ops.unsorted_segment_sum(grad, buf["idx"], buf["num"], out=buf["gsum"],
                               accumulate=self.overlap)
    elif name == "apply":
      # beta1_power *= beta1, beta2_power *= beta2 happen in the same launch (last block)
      ops.kv_variable_group_sparse_apply_adam_v4_dev(self.var, self.slot, buf["gsum"], buf["uniq"],
                                                     self.hp, num_indices=buf["num"],
                                                     advance_powers=True)

  def step_eager(self, ids, grad, buf):
    """gather and (unique -> segment_sum) only depend on the ids / gradients, so they run on
    two streams and join before the apply (the fork/join is captured into the step graph)."""
    torch = self.torch
    if not self.overlap:
      for name in self.STAGES:
        self.stage(name, ids, grad, buf)
      return buf["rows"]
    # (the eager warm-up also goes through here, so the side streams exist before capture)
    main = torch.cuda.current_stream(self.dev)
    self.side.wait_stream(main)
    self.side2.wait_stream(main)
    with torch.cuda.stream(self.side2):
      self.stage("zero", ids, grad, buf)       # the sums' destination, off the critical path
    with torch.cuda.stream(self.side):
      self.stage("unique", ids, grad, buf)
      self.side.wait_stream(self.side2)
      self.stage("segment_sum", ids, grad, buf)
    self.stage("gather", ids, grad, buf)
    main.wait_stream(self.side)
    self.stage("apply", ids, grad, buf)
    return buf["rows"]

  def _capture(self, fn):
    torch = self.torch
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
      fn()
    return g

  def prepare(self, ids_d, grads_d):
    """Warm every batch once eagerly (sizes the workspace, no allocation later), then capture
    one full-step graph and four single-stage graphs per batch."""
    self.ids_d, self.grads_d = ids_d, grads_d
    self.bufs = [self.new_buffers() for _ in ids_d]
    l0 = self.ops._lib.launch_count()
    for ids, grad, buf in zip(ids_d, grads_d, self.bufs):
      self.step_eager(ids, grad, buf)
    # kernels of this library per step (the beta^t advance is inside the apply launch)
    self.launches_per_step = (self.ops._lib.launch_count() - l0) // len(ids_d)
    self.torch.cuda.synchronize()
    # captured work may not grow the tables: make the room now (also refreshes the exact counts)
    self.ops.kv_variable_reserve(self.var, 2 * self.batch)
    self.ops.kv_variable_reserve(self.slot, 2 * self.batch)
    self.full = [self._capture(lambda i=i: self.step_eager(ids_d[i], grads_d[i], self.bufs[i]))
                 for i in range(len(ids_d))]
    def one_stage(n, i):
      if n == "segment_sum" and self.overlap:   # timed alone it includes its zeroing pass
        self.stage("zero", ids_d[i], grads_d[i], self.bufs[i])
      self.stage(n, ids_d[i], grads_d[i], self.bufs[i])
    self.stage_graphs = {
        n: [self._capture(lambda i=i, n=n: one_stage(n, i)) for i in range(len(ids_d))]
        for n in self.STAGES}
    self.rotation = None
    if self.overlap and os.environ.get("KVHBM_BENCH_PIPELINE", "1") != "0":
      self.rotation = self._capture(self.rotation_eager)
    self.torch.cuda.synchronize()

  def step(self, i):
    self.full[i % len(self.full)].replay()

  def rotation_eager(self):
    """All rotating batches, in order, with only the TRUE dependencies between consecutive
    steps: lookup(t+1) and apply(t+1) wait for apply(t) (they read what it wrote), the dedup
    and gradient sum of batch t+1 depend on nothing but their inputs, so they run under the
    lookup/apply of batch t.  Same kernels, same table updates in the same order as step by
    step (tests/test_gpu_fullsize.py compares the two, and both with the oracle)."""
    torch = self.torch
    main = torch.cuda.current_stream(self.dev)
    s_u, s_s = self.side, self.side2
    s_u.wait_stream(main)
    s_s.wait_stream(main)
    for ids, grad, buf in zip(self.ids_d, self.grads_d, self.bufs):
      with torch.cuda.stream(s_s):
        self.stage("zero", ids, grad, buf)
      with torch.cuda.stream(s_u):             # one dedup scratch per device: a chain
        self.stage("unique", ids, grad, buf)
        ev_u = torch.cuda.Event()
        ev_u.record(s_u)
      with torch.cuda.stream(s_s):
        s_s.wait_event(ev_u)
        self.stage("segment_sum", ids, grad, buf)
        ev_s = torch.cuda.Event()
        ev_s.record(s_s)
      self.stage("gather", ids, grad, buf)     # after apply(t-1): stream order
      main.wait_event(ev_s)
      self.stage("apply", ids, grad, buf)
    main.wait_stream(s_u)
    main.wait_stream(s_s)

  def run_steps(self, K):
    """K consecutive steps: whole rotations as one graph each, the rest step by step."""
    n, i = len(self.full), 0
    if self.rotation is not None:
      while K - i >= n:
        self.rotation.replay()
        i += n
    while i < K:
      self.full[i % n].replay()
      i += 1

  def stage_times(self, steps):
    """Average device time of each stage: K back-to-back replays of that stage's graph over the
    rotating batches, bracketed by CUDA events on the launching stream."""
    torch = self.torch
    out = {}
    for n in self.STAGES:
      gs = self.stage_graphs[n]
      for i in range(3):
        gs[i % len(gs)].replay()
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for i in range(steps):
        gs[i % len(gs)].replay()
      b.record()
      torch.cuda.synchronize()
      out[n] = a.elapsed_time(b) / steps
    return out

  # ---- end to end: pinned host buffers in, pinned host rows out ----
  def prepare_host(self, ids_h, grads_h, rows_h):
    """End-to-end stepper: every step copies its ids + gradients from pinned host memory and
    copies the gathered rows back.  The three legs run on three streams, double-buffered, so
    the H2D of step i+1 and the D2H of step i-1 overlap the kernels of step i (PCIe is full
    duplex); every copy still happens, once per step, inside the timed region."""
    t = self.torch
    self.ids_h, self.grads_h, self.rows_h = ids_h, grads_h, rows_h
    self.s_h2d, self.s_d2h = t.cuda.Stream(device=self.dev), t.cuda.Stream(device=self.dev)
    self.h_ids = [t.empty(self.batch, dtype=t.int64, device=self.dev) for _ in range(2)]
    self.h_grad = [t.empty((self.batch, self.dim), dtype=t.float32, device=self.dev)
                   for _ in range(2)]
    self.h_buf = [self.new_buffers() for _ in range(2)]
    self.rows_host = [rows_h, t.empty_like(rows_h).pin_memory()]
    self.ev_in = [t.cuda.Event() for _ in range(2)]     # inputs of slot k are on the device
    self.ev_done = [t.cuda.Event() for _ in range(2)]   # kernels of slot k finished
    self.ev_out = [t.cuda.Event() for _ in range(2)]    # rows of slot k reached the host
    for k in range(2):
      self.h_ids[k].copy_(ids_h[k])
      self.h_grad[k].copy_(grads_h[k])
      self.step_eager(self.h_ids[k], self.h_grad[k], self.h_buf[k])
    t.cuda.synchronize()
    self.e2e = [self._capture(lambda k=k: self.step_eager(self.h_ids[k], self.h_grad[k],
                                                          self.h_buf[k])) for k in range(2)]
    main = t.cuda.current_stream(self.dev)
    for k in range(2):
      self.ev_done[k].record(main)
      self.ev_out[k].record(main)
    self.e2e_i = 0

  def step_host(self, i):
    t = self.torch
    k = self.e2e_i & 1
    self.e2e_i += 1
    j = i % len(self.ids_h)
    main = t.cuda.current_stream(self.dev)
    with t.cuda.stream(self.s_h2d):
      self.s_h2d.wait_event(self.ev_done[k])       # slot k's previous kernels no longer read it
      self.h_ids[k].copy_(self.ids_h[j], non_blocking=True)
      self.h_grad[k].copy_(self.grads_h[j], non_blocking=True)
      self.ev_in[k].record(self.s_h2d)
    main.wait_event(self.ev_in[k])
    main.wait_event(self.ev_out[k])                # slot k's rows buffer has been drained
    self.e2e[k].replay()
    self.ev_done[k].record(main)
    with t.cuda.stream(self.s_d2h):
      self.s_d2h.wait_event(self.ev_done[k])
      self.rows_host[k].copy_(self.h_buf[k]["rows"], non_blocking=True)
      self.ev_out[k].record(self.s_d2h)

  def finish_host(self):
    main = self.torch.cuda.current_stream(self.dev)
    main.wait_stream(self.s_h2d)
    main.wait_stream(self.s_d2h)


class _QuietStdout:
  """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on
  stdout.  Everything written to fd 1 while this is active goes to stderr instead; emit() writes
  to the real stdout."""

  def __init__(self):
    sys.stdout.flush()
    self._real = os.dup(1)
    os.dup2(2, 1)

  def emit(self, text):
    sys.stdout.flush()
    os.write(self._real, (text + "\n").encode())


_OUT = None


def emit_json(line):
  text = json.dumps(line)
  if _OUT is not None:
    _OUT.emit(text)
  else:
    print(text)


def main():
  global _OUT
  _OUT = _QuietStdout()
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=20)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--keys", type=int, default=KEYS)
  ap.add_argument("--batch", type=int, default=BATCH)
  ap.add_argument("--dim", type=int, default=DIM)
  ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
  args = ap.parse_args()
  if args.impl == "reference":
    return main_reference(args)
  return main_ours(args)


if __name__ == "__main__":
  sys.exit(main())
