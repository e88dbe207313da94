#!/usr/bin/env python
"""Per-warp timeline of one gather launch (tuning aid)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tfplus_b200 import ops, _lib
keys = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
ops.set_today(bench.TODAY)
dev = torch.device("cuda:0")
st = bench.LocalStepper(keys, bench.DIM, bench.BATCH, dev)
st.populate()
ids_np, _ = bench.make_batches(2, keys, bench.BATCH, bench.DIM)
ids = [torch.from_numpy(x).to(dev) for x in ids_np]
out = torch.empty((bench.BATCH, bench.DIM), device=dev)
lib = _lib.load()
lib.kv_debug_set_trace.argtypes = [ctypes.c_void_p]
for i in range(3):
  ops.kv_variable_gather_or_insert_v2(st.var, ids[i % 2], out=out)
buf = torch.zeros(65536 * 4, dtype=torch.int64, device=dev)
lib.kv_debug_set_trace(buf.data_ptr())
torch.cuda.synchronize()
ops.kv_variable_gather_or_insert_v2(st.var, ids[1], out=out)
torch.cuda.synchronize()
lib.kv_debug_set_trace(None)
t = buf.cpu().numpy().reshape(-1, 4)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
start, end, sm, probe = t[:, 0] - t0, t[:, 1] - t0, t[:, 2], t[:, 3] - t0
print("warps", len(t), "kernel span ns", end.max())
print("start  pct 0/50/90/100:", np.percentile(start, [0, 50, 90, 100]))
print("end    pct 0/50/90/100:", np.percentile(end, [0, 50, 90, 100]))
dur = end - start
print("dur    pct 0/50/90/100:", np.percentile(dur, [0, 50, 90, 100]))
print("probe-phase (start->freq) pct 50/90/100:", np.percentile(probe - start, [50, 90, 100]))
print("copy-phase  (freq->end)   pct 50/90/100:", np.percentile(end - probe, [50, 90, 100]))
late = np.argsort(-end)[:8]
print("latest warps: idx, sm, start, end:", [(int(i), int(sm[i]), int(start[i]), int(end[i])) for i in late])
per_sm = {}
for i in range(len(t)):
  per_sm.setdefault(int(sm[i]), []).append(end[i])
ends = sorted((max(v), k, len(v)) for k, v in per_sm.items())
print("SM finish times (first 3, last 3):", ends[:3], ends[-3:])
