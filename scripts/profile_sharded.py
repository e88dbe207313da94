#!/usr/bin/env python
"""torchrun -n 2 scripts/profile_sharded.py : kernel-time table of the sharded step (rank 0)."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tfplus_b200 import ops, sharded
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
ops.set_today(bench.TODAY)
keys = int(os.environ.get("KEYS", "2000000"))
st = bench.ShardedStepper(keys, bench.DIM, bench.BATCH, bench.HP, dev, rank, world)
st.populate()
ids_np, g_np = bench.make_batches(4, keys * world, bench.BATCH, bench.DIM, seed_ids=2024 + rank, seed_grad=7 + rank)
ids = [torch.from_numpy(x).to(dev) for x in ids_np]
gr = [torch.from_numpy(x).to(dev) for x in g_np]
os.environ.setdefault("KVHBM_SHARDED_GRAPH", "0")
st.prepare(ids, gr)
for i in range(5): st.step(i)
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
PIPE = os.environ.get("PIPE", "0") == "1" and getattr(st, "rotation", None) is not None
if PIPE:
  st.run_steps(8)
  torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  if PIPE:
    st.run_steps(12)          # 3 rotations of the 4 batches; table below is per 10 -> scale by eye
  else:
    for i in range(10): st.step(i)
  torch.cuda.synchronize()
if rank == 0:
  rows = []
  for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t: rows.append((t / 10.0, e.count / 10.0, e.key[:70]))
  rows.sort(reverse=True)
  tot = sum(r[0] for r in rows)
  print("kernel time per step: %.1f us" % tot)
  for r in rows[:25]: print("%8.1f us  x%.1f  %s" % r)
if rank == 0:
  # timeline of the last two steps: start (us, relative), duration, stream, kernel
  import json, tempfile
  path = os.path.join(tempfile.gettempdir(), "kv_trace_%d.json" % os.getpid())
  prof.export_chrome_trace(path)
  ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
  ev.sort(key=lambda e: e["ts"])
  per_step = max(1, len(ev) // (12 if PIPE else 10))
  tail = ev[-(4 if PIPE else 2) * per_step:]
  t0 = tail[0]["ts"]
  print("timeline (last 2 steps, %d kernels per step):" % per_step)
  for e in tail:
    print("%8.1f +%6.1f us  s%-3s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream", "?"),
                                        e["name"][:60]))
st.release()
torch.cuda.synchronize()
dist.barrier(); dist.destroy_process_group()
