#!/usr/bin/env python
"""python scripts/profile_local.py : per-kernel timeline of the replayed single-GPU step."""
import json, os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tfplus_b200 import ops
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
ops.set_today(bench.TODAY)
keys = int(os.environ.get("KEYS", "10000000"))
st = bench.LocalStepper(keys, bench.DIM, bench.BATCH, dev)
st.populate()
ids_np, g_np = bench.make_batches(bench.N_BATCHES, keys, bench.BATCH, bench.DIM, seed_ids=2024, seed_grad=7)
st.prepare([torch.from_numpy(x).to(dev) for x in ids_np], [torch.from_numpy(x).to(dev) for x in g_np])
for i in range(20): st.step(i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
PIPE = os.environ.get("PIPE", "0") == "1" and st.rotation is not None
if PIPE:
  st.run_steps(bench.N_BATCHES)
  torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  if PIPE:
    st.run_steps(bench.N_BATCHES)
  else:
    for i in range(10): st.step(i)
  torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "kv_trace_local.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
per_step = max(1, len(ev) // (bench.N_BATCHES if PIPE else 10))
tail = ev[-(4 if PIPE else 2) * per_step:]
t0 = tail[0]["ts"]
print("timeline (last 2 steps, %d kernels per step):" % per_step)
for e in tail:
  print("%8.1f +%6.1f us  s%-3s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream", "?"), e["name"][:70]))
