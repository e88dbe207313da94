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


def pick_rotation(steps):
  """Rotating batches for a K-step run: the largest even n in 8..16 that divides K, so that all K
  timed steps are whole rotation graphs (8 x 17.3 MB = 138 MB is still more than the 126 MB L2);
  16 when there is none (the remainder then runs as strict steps, and the line says so)."""
  for n in range(16, 7, -2):
    if steps >= n and steps % n == 0:
      return n
  return 16
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
  seg = b * (4 * d + 4) + 4 * d * u
  app = (8 + 2 * 12 + 4 * d + 2 * 4 * d + 2 * 3 * 4 * d) * u
  return {
      "gather": (8 + 12 + 4 + 4 * d + 4 * d) * b,
      "unique": 8 * b + 8 * u + 4 * b,
      # one fused pass here; the survey's figures for the two ops it replaces, added
      "segment_sum+apply": seg + app,
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


def schedule_text(world, pipelined, K):
  if not pipelined:
    return "one CUDA graph per step, nothing runs ahead"
  split = "CUDA graphs of %d steps; %d of the timed steps ran that way and %d strictly" % (
      N_BATCHES, K - K % N_BATCHES, K % N_BATCHES)
  if world > 1:
    return ("sharded: lookup(t) -> barrier B -> expand -> local gradient sums -> gradient "
            "exchange -> barrier C -> owner sum + apply(t) -> lookup(t+1); only the "
            "requester-side dedup + id exchange, barrier A and the owner-side dedup of step t+1 "
            "(functions of the ids alone) run beside step t; " + split)
  return ("gather(t) -> [segment_sum + apply](t) -> gather(t+1) on one stream (the gradient "
          "depends on the gathered rows, the next lookup on the apply); only the dedup plan of "
          "batch t+1 (a function of its ids alone) is built beside step t; " + split)


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
def bind_to_gpu_numa_node(torch, local_rank):
  """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers
  of the end-to-end leg are first-touched there and their PCIe traffic stays on that socket
  (torchrun pins nothing; eight ranks' buffers on one node cost the 8-GPU e2e leg its scaling).
  Returns a description for the JSON line, or None when the topology is not exposed."""
  try:
    pr = torch.cuda.get_device_properties(local_rank)
    bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    if node < 0:
      return None
    cpus = set()
    for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
      lo, _, hi = part.partition("-")
      cpus.update(range(int(lo), int(hi or lo) + 1))
    cpus &= os.sched_getaffinity(0)
    if not cpus:
      return None
    os.sched_setaffinity(0, cpus)
    return {"gpu": bdf, "numa_node": node, "cpus": len(cpus)}
  except (OSError, ValueError, AttributeError, RuntimeError):
    return None


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
  # (N = 1 keeps every host core: the cpu_baseline leg runs in this process)
  numa = bind_to_gpu_numa_node(torch, local_rank) if (
      world > 1 and os.environ.get("KVHBM_BENCH_NUMA", "1") != "0") else None
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  ops.set_today(TODAY)
  K, W = max(1, args.steps), max(3, args.warmup)
  keys, B, D = args.keys, args.batch, args.dim

  if world > 1:
    stepper = ShardedStepper(keys, D, B, HP, dev, rank, world)
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
              "note": "plan -> lookup -> apply of one batch per CUDA graph; nothing of batch t+1 "
                      "(not even its dedup plan) starts before step t has finished"}

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

  check = None
  if not args.no_check:
    check = parity_check(world, rank, dev, dist)
    note('parity check')

  if rank == 0:
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
      pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
    ab = algorithmic_bytes(B, u_meas, D)
    step_bytes = sum(ab.values())      # per GPU: every rank does the single-GPU step on its shard
    if world > 1:
      ab = {"step": step_bytes}
      stage_ms = dict(stage_ms, step=ms / K)
    kern = {}
    for name, t_ms in stage_ms.items():
      if name in ab and t_ms > 0:
        kern[name] = {"ms": t_ms, "algorithmic_bytes": ab[name],
                      "achieved_gbs": ab[name] / (t_ms * 1e-3) / 1e9,
                      "frac": ab[name] / (t_ms * 1e-3) / 1e9 / peak,
                      "launches": STAGE_LAUNCHES.get(name, 1)}
    # the dominant kernel = the stage with the longest launch (the fused apply's staging pass is
    # a fraction of its stage, so the stage time is charged to the main kernel: conservative)
    top = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
    roof = None
    if top:
      tr = measured_traffic(top)
      roof = {"bound": "hbm", "kernel": top, "achieved": kern[top]["achieved_gbs"], "peak": peak,
              "unit": "GB/s", "frac": kern[top]["frac"],
              "traffic": tr["bytes"] if tr else None, "traffic_source": tr,
              "peak_source": peak_src, "stages": kern,
              "step_algorithmic_bytes_per_gpu": step_bytes,
              "step_frac_of_peak": step_bytes / (ms / K * 1e-3) / 1e9 / peak,
              "step_frac_of_8TBs": step_bytes / (ms / K * 1e-3) / 1e9 / 8000.0,
              "note": "fractions are per GPU (one GPU's algorithmic bytes over one GPU's peak)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "unique_per_step": u_meas, "clocks": clk, "gpu_launches": int(launches),
        "host_enqueue_us_per_step": host_us,
        "schedule": schedule_text(world, pipelined, K),
        "strict_per_step": strict,
        "e2e": {"value": e2e_val, "unit": UNIT,
                "h2d_bytes_per_step": int(B * 8 + B * D * 4),
                "d2h_bytes_per_step": int(B * D * 4), "ms_per_step": e2e_ms / K,
                "host_numa_binding": numa},
        "roofline": roof,
        "parity_check": check,
    }
    if world > 1:
      sent = stepper.padded.wire_bytes
      peer = hasattr(stepper.padded, "barrier_timeouts")
      if peer and stepper.padded.barrier_timeouts():
        raise SystemExit("peer barrier timed out %d times: the step is invalid"
                         % stepper.padded.barrier_timeouts())
      # bytes that actually cross NVLink per GPU and step: the unique ids + counts
      # routed to OTHER ranks, their rows back and their gradient sums out
      u_remote = u_meas * (world - 1) / world
      actual = int(u_remote * (12 + 2 * 4 * D))
      line["nvlink"] = {"exchange": "peer-memory stores fused into the producing kernels + "
                                    "3 barrier kernels" if peer else "NCCL all_to_all_single x3",
                        "bytes_sent_per_gpu_per_step": actual,
                        "bus_gbs_per_gpu": actual / (ms / K * 1e-3) / 1e9,
                        "frac_of_nvlink_peak": actual / (ms / K * 1e-3) / 1e9 / 900.0,
                        "capacity_bytes_per_gpu_per_step": sent,
                        "peak_gbs_per_direction": 900.0, "measured_peer_copy_gbs": 770.0,
                        "exchanges_per_step": 3, "barriers_per_step": 3 if peer else None,
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
                    "oracle port of the reference algorithm, %d threads (the port's per-shard "
                    "locking does not scale past ~16 threads: 32 cores measured slower than 16, "
                    "so this is a stated baseline, not a tuned one)" % (keys, cs, B, cores)}
    emit_json(line)
  failed = check is not None and not check["ok"]
  if world > 1:
    stepper.release()      # captured NCCL work must be gone before the process group
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
  return 1 if failed else 0


# kernels of this library per stage (the fused apply is the staging pass + the chain/apply kernel)
STAGE_LAUNCHES = {"gather": 1, "unique": 5, "segment_sum+apply": 2, "step": 1}


def measured_traffic(kernel):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, as the last committed
  `ncu --set full` capture recorded it (profiles/ncu_traffic.json, written by
  scripts/ncu_summary.py --json together with the git revision it was taken at); None when no
  capture of this kernel exists."""
  try:
    d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    k = d["kernels"].get(kernel)
    return None if k is None else {"bytes": k["dram_bytes"], "git": d.get("git"), "kernel": k.get("name")}
  except Exception:
    return None


class LocalStepper:
  """One GPU: var + m_v_linear tables and the step  plan -> lookup -> fused gradient sum + apply.

  plan    kv_plan_build: tf.unique_with_counts of the batch's ids + occurrence lists.  It is a
          function of the ids ALONE, so the input pipeline may build it ahead of the step.
  gather  kv_gather_or_insert_plan: KvVariableGatherOrInsertV2 over all B ids.
  apply   kv_apply_plan_dev: UnsortedSegmentSum (TF's order, bit-exact) + GroupAdam v4 in one
          pass, beta^t advanced in the same launch (TF Adam's `_finish`).

  The gradient of a real model is a function of the gathered rows, so the apply of a batch may
  not start before its lookup has finished; the lookup of the next batch reads what this apply
  wrote.  gather(t) -> apply(t) -> gather(t+1) is therefore ONE chain on one stream, in every
  schedule timed here.  Only plan(t+1) runs beside it (second stream).  The optimizer's scalar
  inputs live in device memory (`hp`), so steps are capturable and the timed loops replay CUDA
  graphs.
  """

  STAGES = ["unique", "gather", "segment_sum+apply"]

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
    # the chain runs on a high-priority stream, the look-ahead plan build on a low-priority one:
    # when both have blocks to place, the chain's go first
    prio = os.environ.get("KVHBM_BENCH_PRIORITY", "1") != "0"
    self.side = torch.cuda.Stream(device=dev, priority=0)
    self.chain = torch.cuda.Stream(device=dev, priority=-1 if prio else 0)
    self.lookahead = os.environ.get("KVHBM_BENCH_PLAN_AHEAD", "1") != "0"
    self.prebuilt = os.environ.get("KVHBM_BENCH_PLAN_PREBUILT", "0") == "1"

  def populate(self):
    torch, ops = self.torch, self.ops
    for s in range(0, self.keys, 1 << 20):
      ids = torch.arange(s, min(self.keys, s + (1 << 20)), dtype=torch.int64, device=self.dev)
      ops.kv_variable_gather_or_insert_v2(self.var, ids)
      ops.kv_variable_gather_or_insert_v2(self.slot, ids)
    torch.cuda.synchronize()

  # ---- the step, stage by stage (eager; also what gets captured) ----
  def stage(self, name, ids, grad, plan, rows):
    ops = self.ops
    if name == "unique":
      plan.build(ids)
    elif name == "gather":
      ops.kv_variable_gather_or_insert_plan(self.var, plan, out=rows)
    elif name == "segment_sum+apply":
      ops.kv_variable_apply_plan(ops.OPT_GROUP_ADAM_V4, self.var, self.slot, None, plan, grad,
                                 self.hp, advance_powers=True)

  def step_eager(self, ids, grad, plan, rows):
    for name in self.STAGES:
      self.stage(name, ids, grad, plan, rows)
    return rows

  def _capture(self, fn):
    torch = self.torch
    g = torch.cuda.CUDAGraph()
    self.chain.wait_stream(torch.cuda.current_stream(self.dev))
    with torch.cuda.graph(g, stream=self.chain):
      fn()
    return g

  def prepare(self, ids_d, grads_d):
    """Warm every batch once eagerly (sizes every scratch buffer, no allocation later), then
    capture: one graph per strict step, one per stage and batch, and the rotation."""
    torch, ops = self.torch, self.ops
    self.ids_d, self.grads_d = ids_d, grads_d
    n = len(ids_d)
    self.plans = [ops.Plan(self.batch, self.dev) for _ in range(n)]
    self.rows = [torch.empty((self.batch, self.dim), dtype=torch.float32, device=self.dev)
                 for _ in range(n)]
    l0 = ops._lib.launch_count()
    for i in range(n):
      self.step_eager(ids_d[i], grads_d[i], self.plans[i], self.rows[i])
    self.launches_per_step = (ops._lib.launch_count() - l0) // n
    torch.cuda.synchronize()
    # captured work may not grow the tables: make the room now (also refreshes the exact counts)
    ops.kv_variable_reserve(self.var, 2 * self.batch)
    ops.kv_variable_reserve(self.slot, 2 * self.batch)
    self.full = [self._capture(lambda i=i: self.step_eager(ids_d[i], grads_d[i], self.plans[i],
                                                           self.rows[i])) for i in range(n)]
    self.stage_graphs = {
        name: [self._capture(lambda i=i, name=name: self.stage(name, ids_d[i], grads_d[i],
                                                               self.plans[i], self.rows[i]))
               for i in range(n)] for name in self.STAGES}
    self.rotation = self._capture(self.rotation_eager) if self.lookahead else None
    if self.rotation is not None:
      self.plans[0].build(ids_d[0])   # the rotation expects the plan of its first batch
    torch.cuda.synchronize()

  def step(self, i):
    """Strict: plan -> lookup -> apply of one batch, nothing of the next batch before it ends."""
    self.full[i % len(self.full)].replay()

  def rotation_eager(self):
    """All rotating batches in order.  Main stream: gather(t) -> apply(t) -> gather(t+1) ...
    (the dependency chain of a training loop).  Second stream: the plan of batch t+1 — a
    function of its ids alone — is built while batch t trains; the rotation ends by building
    the plan of batch 0 for the next replay."""
    torch = self.torch
    main = torch.cuda.current_stream(self.dev)
    side = self.side
    n = len(self.ids_d)
    side.wait_stream(main)
    ev_plan, ev_apply = [None] * n, [None] * n
    for t in range(n):
      if t > 0:
        main.wait_event(ev_plan[t])               # plan(t) was built beside step t-1
      self.stage("gather", self.ids_d[t], self.grads_d[t], self.plans[t], self.rows[t])
      self.stage("segment_sum+apply", self.ids_d[t], self.grads_d[t], self.plans[t], self.rows[t])
      ev_apply[t] = torch.cuda.Event()
      ev_apply[t].record(main)
      nxt = (t + 1) % n
      with torch.cuda.stream(side):
        if nxt == 0:
          side.wait_event(ev_apply[0])            # the last reader of plan(0)'s buffers
        if not self.prebuilt:                     # (diagnostic: the chain alone, plans left from warm-up)
          self.stage("unique", self.ids_d[nxt], None, self.plans[nxt], None)
        ev_plan[nxt] = torch.cuda.Event()
        ev_plan[nxt].record(side)
    main.wait_stream(side)

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
    if self.rotation is not None and K % n:
      self.plans[0].build(self.ids_d[0])          # leave the state a rotation expects

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
    t, ops = self.torch, self.ops
    self.ids_h, self.grads_h, self.rows_h = ids_h, grads_h, rows_h
    self.s_h2d, self.s_d2h = t.cuda.Stream(device=self.dev), t.cuda.Stream(device=self.dev)
    self.h_ids = [t.empty(self.batch, dtype=t.int64, device=self.dev) for _ in range(2)]
    self.h_grad = [t.empty((self.batch, self.dim), dtype=t.float32, device=self.dev)
                   for _ in range(2)]
    self.h_plan = [ops.Plan(self.batch, self.dev) for _ in range(2)]
    self.h_rows = [t.empty((self.batch, self.dim), dtype=t.float32, device=self.dev)
                   for _ in range(2)]
    self.rows_host = [rows_h, t.empty_like(rows_h).pin_memory()]
    self.ev_in = [t.cuda.Event() for _ in range(2)]     # inputs of slot k are on the device
    self.ev_done = [t.cuda.Event() for _ in range(2)]   # kernels of slot k finished
    self.ev_out = [t.cuda.Event() for _ in range(2)]    # rows of slot k reached the host
    for k in range(2):
      self.h_ids[k].copy_(ids_h[k])
      self.h_grad[k].copy_(grads_h[k])
      self.step_eager(self.h_ids[k], self.h_grad[k], self.h_plan[k], self.h_rows[k])
    t.cuda.synchronize()
    self.e2e = [self._capture(lambda k=k: self.step_eager(self.h_ids[k], self.h_grad[k],
                                                          self.h_plan[k], self.h_rows[k]))
                for k in range(2)]
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
      self.rows_host[k].copy_(self.h_rows[k], non_blocking=True)
      self.ev_out[k].record(self.s_d2h)

  def finish_host(self):
    main = self.torch.cuda.current_stream(self.dev)
    main.wait_stream(self.s_h2d)
    main.wait_stream(self.s_d2h)


class ShardedStepper:
  """Weak-scaling microbench: `keys` keys per GPU, every rank brings its own batch."""

  STAGES = ["route+lookup", "grads+apply"]

  def __init__(self, keys_per_gpu, dim, batch, hp, dev, rank, world, slot_mult=3):
    import torch
    from tfplus_b200 import ops, sharded
    self.optimizer = "group_adam"      # "adam": slot_mult = 2, hpt = Adam's scalar inputs
    self.torch, self.ops, self.sharded = torch, ops, sharded
    self.keys, self.dim, self.batch, self.dev = keys_per_gpu, dim, batch, dev
    self.rank, self.world, self.hp = rank, world, hp
    cap = int(keys_per_gpu * 1.15) + batch
    self.tbl = sharded.ShardedKvVariable(dim, world, rank, dev, slot_dims=(slot_mult * dim,),
                                         capacity_hint=cap, seed=1)
    self.ops.init_kv_variable_v2(self.tbl.var, torch.from_numpy(init_table(dim)).to(dev))
    self.ops.init_kv_variable_v2(self.tbl.slots[0],
                                 torch.zeros(INIT_ROWS, slot_mult * dim, device=dev))
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
      sorted_ids, _, counts = self.ops.partition_ids(ids, self.world, "hash")
      c = counts.cpu().tolist()
      lo = sum(c[:self.rank])
      mine = sorted_ids[lo:lo + c[self.rank]]
      if mine.numel():
        self.ops.kv_variable_gather_or_insert_v2(self.tbl.var, mine)
        self.ops.kv_variable_gather_or_insert_v2(self.tbl.slots[0], mine)
    t.cuda.synchronize()

  def step_exact(self, ids, grad):
    """The variable-size path (host reads the per-shard counts): reference behaviour for the
    padded path and its fallback on overflow."""
    rows = self.tbl.lookup(ids)
    owner_ids, owner_grads = self.tbl.owner_gradients(grad)
    if owner_ids.numel():
      self.ops.kv_variable_group_sparse_apply_adam_v4_dev(self.tbl.var, self.tbl.slots[0], owner_grads,
                                                     owner_ids, self.hpt)
    self.hpt[1:3].mul_(self.betas)
    return rows

  def step_eager(self, ids, grad):
    self.steps_done += 1
    return self.padded.run(ids, grad)

  def prepare(self, ids_d, grads_d):
    from tfplus_b200 import _lib
    t = self.torch
    self.ids_d, self.grads_d = ids_d, grads_d
    self.padded = self.sharded.make_padded_step(self.tbl.var, self.tbl.slots[0], self.dim,
                                                self.batch, self.world, self.rank, self.dev,
                                                self.hpt, self.betas)
    self.padded.optimizer = self.optimizer
    l0 = _lib.launch_count()
    for i in range(2):
      self.step_eager(ids_d[i], grads_d[i])
    self.launches_per_step = (_lib.launch_count() - l0) // 2
    t.cuda.synchronize()
    if self.padded.overflowed():
      raise RuntimeError("padded shard exchange overflowed: raise cap")
    self.ops.kv_variable_reserve(self.tbl.var, 2 * self.padded.cap * self.world)
    self.ops.kv_variable_reserve(self.tbl.slots[0], 2 * self.padded.cap * self.world)
    self.graphs = [self._capture(lambda i=i: self.padded.run(ids_d[i], grads_d[i]))
                   for i in range(len(ids_d))]
    if self.graphs and self.graphs[0] is None:
      self.graphs = []
    # the pipelined schedule: one graph per rotation of the batches (PeerShardedStep only)
    self.rotation = None
    if (self.graphs and hasattr(self.padded, "run_rotation") and self.padded.fused_route
        and len(ids_d) % 2 == 0 and os.environ.get("KVHBM_BENCH_PIPELINE", "1") != "0"):
      self.padded.run_rotation(ids_d, grads_d)        # once eagerly: errors surface here
      t.cuda.synchronize()
      self.rotation = self._capture(lambda: self.padded.run_rotation(ids_d, grads_d))
    t.cuda.synchronize()

  def run_steps(self, K):
    """K consecutive steps: whole rotations as one graph each, the rest step by step."""
    n, i = len(self.ids_d), 0
    if self.rotation is not None:
      while K - i >= n:
        self.rotation.replay()
        i += n
      self.steps_done += i
    while i < K:
      self.step(i)
      i += 1

  def step(self, i):
    k = i % len(self.ids_d)
    self.steps_done += 1
    if self.graphs:
      self.graphs[k].replay()
      return self.padded.out
    return self.padded.run(self.ids_d[k], self.grads_d[k])

  def release(self):
    self.graphs = []
    self.e2e = []
    self.rotation = None

  def stage_times(self, steps):
    t = self.torch
    a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    t.cuda.synchronize()
    a.record()
    for i in range(steps):
      self.step(i)
    b.record()
    t.cuda.synchronize()
    return {"sharded step": a.elapsed_time(b) / steps}

  def _capture(self, fn):
    t = self.torch
    if os.environ.get("KVHBM_SHARDED_GRAPH", "1") == "0":
      return None
    side = t.cuda.Stream(device=self.dev)
    side.wait_stream(t.cuda.current_stream(self.dev))
    with t.cuda.stream(side):
      g = t.cuda.CUDAGraph()
      with t.cuda.graph(g, stream=side):
        fn()
    t.cuda.current_stream(self.dev).wait_stream(side)
    return g

  def prepare_host(self, ids_h, grads_h, rows_h):
    """End to end: every step copies its ids + gradients from pinned host memory and its rows
    back.  Two input / output slots and two copy streams, so the H2D of step i+1 and the D2H
    of step i-1 run under the kernels and exchanges of step i; every copy still happens once
    per step inside the timed region."""
    t = self.torch
    self.ids_h, self.grads_h, self.rows_h = ids_h, grads_h, rows_h
    self.s_h2d, self.s_d2h = t.cuda.Stream(device=self.dev), t.cuda.Stream(device=self.dev)
    self.h_ids = [t.empty(self.batch, dtype=t.int64, device=self.dev) for _ in range(2)]
    self.h_grad = [t.empty((self.batch, self.dim), dtype=t.float32, device=self.dev)
                   for _ in range(2)]
    self.h_out = [t.empty((self.batch, self.dim), dtype=t.float32, device=self.dev)
                  for _ in range(2)]
    self.rows_host = [rows_h, t.empty_like(rows_h).pin_memory()]
    self.ev_in = [t.cuda.Event() for _ in range(2)]
    self.ev_done = [t.cuda.Event() for _ in range(2)]
    self.ev_out = [t.cuda.Event() for _ in range(2)]
    for k in range(2):
      self.h_ids[k].copy_(ids_h[k])
      self.h_grad[k].copy_(grads_h[k])
    t.cuda.synchronize()
    self.e2e = [self._capture(lambda k=k: self.padded.run(self.h_ids[k], self.h_grad[k],
                                                          out=self.h_out[k])) for k in range(2)]
    main = t.cuda.current_stream(self.dev)
    for k in range(2):
      self.ev_done[k].record(main)
      self.ev_out[k].record(main)
    self.e2e_i = 0

  def finish_host(self):
    main = self.torch.cuda.current_stream(self.dev)
    main.wait_stream(self.s_h2d)
    main.wait_stream(self.s_d2h)

  def step_host(self, i):
    t = self.torch
    k = self.e2e_i & 1
    self.e2e_i += 1
    j = i % len(self.ids_h)
    main = t.cuda.current_stream(self.dev)
    with t.cuda.stream(self.s_h2d):
      self.s_h2d.wait_event(self.ev_done[k])       # slot k's previous step no longer reads it
      self.h_ids[k].copy_(self.ids_h[j], non_blocking=True)
      self.h_grad[k].copy_(self.grads_h[j], non_blocking=True)
      self.ev_in[k].record(self.s_h2d)
    main.wait_event(self.ev_in[k])
    main.wait_event(self.ev_out[k])                # slot k's rows have been drained to the host
    if self.e2e[k] is not None:
      self.e2e[k].replay()
    else:
      self.padded.run(self.h_ids[k], self.h_grad[k], out=self.h_out[k])
    self.steps_done += 1
    self.ev_done[k].record(main)
    with t.cuda.stream(self.s_d2h):
      self.s_d2h.wait_event(self.ev_done[k])
      self.rows_host[k].copy_(self.h_out[k], non_blocking=True)
      self.ev_out[k].record(self.s_d2h)


# ----------------------------------------------------------------------------- self-check
def _check_batches(world, steps, keys, batch, dim, dyadic=False):
  """Deterministic small batches for the untimed parity check: 90 % uniform ids over the key
  range, 10 % out of 64 hot keys (duplicates within and across ranks), N(0,1) gradients - or,
  `dyadic`, gradients k/8 with |k| <= 16: every partial sum of those is exact in fp32, so the
  duplicate-gradient sums do not depend on the order in which they are added."""
  out = []
  for s in range(steps):
    per_rank = []
    for r in range(world):
      g = np.random.Generator(np.random.PCG64(1000 * s + r + 17))
      ids = g.integers(0, keys, size=batch, dtype=np.int64)
      hot = g.random(batch) < 0.1
      ids[hot] = (g.integers(0, 64, size=int(hot.sum()), dtype=np.int64) * 7919) % keys
      if dyadic:
        grad = (g.integers(-16, 17, size=(batch, dim)) / 8.0).astype(np.float32)
      else:
        grad = g.standard_normal((batch, dim), dtype=np.float32)
      per_rank.append((ids, grad))
    out.append(per_rank)
  return out


def parity_check(world, rank, dev, dist):
  """Untimed correctness evidence for the path the timed loops run: a small replay (fresh
  tables, same code path: strict steps, then the rotation schedule) whose final table state -
  every shard exported and gathered on rank 0 - is compared with ONE oracle table fed the
  concatenated batches: membership and frequency words bit-exactly, value and slot rows at
  1e-6.  Returns the dict that goes into the JSON line (rank 0) or None."""
  import torch
  from tfplus_b200 import ops
  keys_pg, B, D, steps = 20000, 2048, DIM, 6
  keys = keys_pg * world
  hp = dict(HP)
  # The sharded step adds a key's gradients per rank and then across ranks with float atomics,
  # not in the oracle's sequential order (the single-GPU step adds in TF's order, bit-exact).
  # Rounding then differs by an ulp of the SUM, which is not small against m or against a
  # group-lasso threshold.  The sharded check therefore feeds dyadic gradients (k/8): their sums
  # are exact whatever the order, and everything downstream - routing, the three exchanges,
  # the barriers, dedup across ranks, the apply - must then agree with the oracle at 1e-6.
  data = _check_batches(world, steps, keys, B, D, dyadic=world > 1)
  ids_d = [torch.from_numpy(data[s][rank][0]).to(dev) for s in range(steps)]
  grads_d = [torch.from_numpy(data[s][rank][1]).to(dev) for s in range(steps)]
  if world > 1:
    st = ShardedStepper(keys_pg, D, B, hp, dev, rank, world)
    st.populate()
    st.padded = st.sharded.make_padded_step(st.tbl.var, st.tbl.slots[0], D, B, world, rank, dev,
                                            st.hpt, st.betas)
    var, slot = st.tbl.var, st.tbl.slots[0]
    for s in range(2):
      st.padded.run(ids_d[s], grads_d[s])
    if hasattr(st.padded, "run_rotation") and st.padded.fused_route:
      st.padded.run_rotation(ids_d[2:], grads_d[2:])
      path = "PeerShardedStep.run x2 + run_rotation x%d" % (steps - 2)
    else:
      for s in range(2, steps):
        st.padded.run(ids_d[s], grads_d[s])
      path = "%s.run x%d" % (type(st.padded).__name__, steps)
    rows_last = st.padded.out.cpu().numpy().copy()
    torch.cuda.synchronize()
    bad = st.padded.overflowed() or (hasattr(st.padded, "barrier_timeouts")
                                     and st.padded.barrier_timeouts() > 0)
  else:
    st = LocalStepper(keys, D, B, dev)
    st.populate()
    var, slot = st.var, st.slot
    plan = ops.Plan(B, dev)
    rows = torch.empty((B, D), dtype=torch.float32, device=dev)
    for s in range(steps):
      st.step_eager(ids_d[s], grads_d[s], plan, rows)
    rows_last = rows.cpu().numpy().copy()
    path = "plan -> gather_or_insert_plan -> apply_plan x%d" % steps
    bad = False
  exp = {}
  for name, tb in (("var", var), ("slot", slot)):
    k, v, _, bl, fk, fv = ops.kv_variable_export(tb, first_n=6, enable_cutoff=False,
                                                 freq_dtype=torch.int32)
    exp[name] = [x.cpu().numpy() for x in (k, v, bl, fk, fv)]
  payload = (exp, bool(bad))
  if world > 1:
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
  else:
    gathered = [payload]
  if rank != 0:
    return None
  from oracle import binding as ob
  o_var = ob.OracleTable(D, 0, seed=1)
  o_slot = ob.OracleTable(3 * D, 0, seed=1)
  o_var.set_init_table(init_table(D))
  o_slot.set_init_table(np.zeros((INIT_ROWS, 3 * D), np.float32))
  all_ids = np.arange(keys, dtype=np.int64)
  o_var.gather_or_insert(all_ids, today=TODAY)
  o_slot.gather_or_insert(all_ids, today=TODAY)
  b1p, b2p = np.float32(HP["beta1"]), np.float32(HP["beta2"])
  want_rows = None
  for s in range(steps):
    ids = np.concatenate([data[s][r][0] for r in range(world)])
    grad = np.concatenate([data[s][r][1] for r in range(world)])
    want_rows = o_var.gather_or_insert(ids, today=TODAY)
    u, idx = ob.unique(ids)
    ob.apply_group_adam_v4(o_var, o_slot, u, ob.segment_sum(grad, idx, u.size), hp["lr"],
                           float(b1p), float(b2p), hp["beta1"], hp["beta2"], hp["epsilon"],
                           hp["l1"], hp["l2"], hp["l21"], today=TODAY)
    b1p, b2p = b1p * np.float32(HP["beta1"]), b2p * np.float32(HP["beta2"])
  res = {"ok": True, "keys": keys, "steps": steps, "batch_per_gpu": B, "path": path,
         "hyper_parameters": hp,
         "compared": "membership, blacklist and frequency words bit-exact; value and slot rows "
                     "rtol 1e-6 atol 1e-7; rank 0's looked-up rows of the last step",
         "against": "one oracle table fed the concatenated batches"}
  try:
    assert not any(g[1] for g in gathered), "exchange overflow or barrier timeout"
    np.testing.assert_allclose(rows_last, want_rows[:B].reshape(B, D), rtol=1e-6, atol=1e-7)
    for name, otab in (("var", o_var), ("slot", o_slot)):
      ref = otab.export(first_n=6, enable_cutoff=False, freq_u32=True)
      rows, freq, black = {}, {}, set()
      for g in gathered:
        k, v, bl, fk, fv = g[0][name]
        rows.update({int(a): b for a, b in zip(k, v)})
        freq.update({int(a): int(b) for a, b in zip(fk, fv.view(np.uint32))})
        black.update(int(a) for a in bl)
      assert freq == {int(a): int(b) for a, b in zip(ref["freq_keys"], ref["freq_values"])}, \
          name + ": frequency words differ"
      assert black == set(int(a) for a in ref["blacklist"]), name + ": blacklists differ"
      assert set(rows) == set(int(a) for a in ref["keys"]), name + ": key sets differ"
      got = np.stack([rows[int(a)] for a in ref["keys"]])
      np.testing.assert_allclose(got, ref["values"].reshape(got.shape), rtol=1e-6, atol=1e-7,
                                 err_msg=name + " rows")
  except AssertionError as e:
    res["ok"] = False
    res["error"] = str(e)[:400]
  return res


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
  ap.add_argument("--no-check", action="store_true", help="skip the untimed parity self-check")
  ap.add_argument("--config", default="microbench",
                  choices=["microbench", "ncf", "dcn", "sharded200m", "streaming"],
                  help="BASELINE.json config: microbench = configs[1] (the metric's own, default); "
                       "the others are measured by bench_configs.py")
  args = ap.parse_args()
  if args.impl == "reference":
    return main_reference(args)
  if args.config != "microbench":
    import bench_configs
    bench_configs.bench._OUT = _OUT      # (this file runs as __main__: share the real stdout)
    bench_configs.RUNNERS[args.config](args)
    return 0
  global N_BATCHES
  N_BATCHES = pick_rotation(args.steps)
  return main_ours(args)


if __name__ == "__main__":
  sys.exit(main())
