#!/usr/bin/env python
"""Per-warp timeline of one apply launch (tuning aid)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tfplus_b200 import ops, _lib
keys = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
ops.set_today(bench.TODAY)
dev = torch.device("cuda:0")
st = bench.LocalStepper(keys, bench.DIM, bench.BATCH, dev)
st.overlap = False
st.populate()
ids_np, g_np = bench.make_batches(3, keys, bench.BATCH, bench.DIM)
ids = [torch.from_numpy(x).to(dev) for x in ids_np]
gr = [torch.from_numpy(x).to(dev) for x in g_np]
buf = st.new_buffers()
for i in range(3):
  st.step_eager(ids[i], gr[i], buf)
lib = _lib.load()
lib.kv_debug_set_trace.argtypes = [ctypes.c_void_p]
for name in ("gather", "unique", "segment_sum"):
  st.stage(name, ids[1], gr[1], buf)
tb = torch.zeros(65536 * 4, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
lib.kv_debug_set_trace(tb.data_ptr())
st.stage("apply", ids[1], gr[1], buf)
torch.cuda.synchronize()
lib.kv_debug_set_trace(None)
t = tb.cpu().numpy().reshape(-1, 4)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
start, end, p2, p1 = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3] - t0
pc = lambda a: np.percentile(a, [0, 50, 90, 100]).round()
print("warps", len(t), "span ns", end.max())
print("start:", pc(start)); print("end:", pc(end)); print("dur:", pc(end - start))
print("phase1 (probe):", pc(p1 - start)); print("phase2 (rows):", pc(p2 - p1)); print("phase3:", pc(end - p2))
