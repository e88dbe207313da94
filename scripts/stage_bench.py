#!/usr/bin/env python
"""Times the four stages of the bench step (CUDA-graph replays, events) — for kernel tuning."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--keys", type=int, default=bench.KEYS)
ap.add_argument("--tag", default="")
ap.add_argument("--uniform", action="store_true")
a = ap.parse_args()
from tfplus_b200 import ops
ops.set_today(bench.TODAY)
dev = torch.device("cuda:0")
st = bench.LocalStepper(a.keys, bench.DIM, bench.BATCH, dev)
st.populate()
ids_np, grads_np = bench.make_batches(bench.N_BATCHES, a.keys, bench.BATCH, bench.DIM)
if a.uniform:
  import numpy as np
  rng = np.random.default_rng(0)
  ids_np = [rng.integers(0, a.keys, size=bench.BATCH).astype(np.int64) for _ in ids_np]
st.prepare([torch.from_numpy(x).to(dev) for x in ids_np], [torch.from_numpy(x).to(dev) for x in grads_np])
for i in range(20): st.step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps): st.step(i)
e1.record(); torch.cuda.synchronize()
res = {k: round(v * 1e3, 2) for k, v in st.stage_times(a.steps).items()}
res["step_us"] = round(e0.elapsed_time(e1) / a.steps * 1e3, 2)
print(a.tag, json.dumps(res))
