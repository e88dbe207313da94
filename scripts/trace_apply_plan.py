#!/usr/bin/env python
"""Per-warp timeline of one fused segment-sum + apply launch (tuning aid; needs a build with
KVHBM_TRACE=1)."""
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
st.populate()
ids_np, g_np = bench.make_batches(3, keys, bench.BATCH, bench.DIM)
ids = [torch.from_numpy(x).to(dev) for x in ids_np]
gr = [torch.from_numpy(x).to(dev) for x in g_np]
plan = ops.Plan(bench.BATCH, dev)
def step(i):
  plan.build(ids[i])
  ops.kv_variable_gather_or_insert_plan(st.var, plan)
  ops.kv_variable_apply_plan(ops.OPT_GROUP_ADAM_V4, st.var, st.slot, None, plan, gr[i], st.hp,
                             advance_powers=True)
for i in range(3):
  step(i)
lib = _lib.load()
lib.kv_debug_set_trace.argtypes = [ctypes.c_void_p]
plan.build(ids[1])
ops.kv_variable_gather_or_insert_plan(st.var, plan)
tb = torch.zeros(65536 * 4, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
lib.kv_debug_set_trace(tb.data_ptr())
ops.kv_variable_apply_plan(ops.OPT_GROUP_ADAM_V4, st.var, st.slot, None, plan, gr[1], st.hp,
                           advance_powers=True)
torch.cuda.synchronize()
lib.kv_debug_set_trace(None)
raw = tb.cpu().numpy()
gt = raw[131072:131072 + 2960 * 16].reshape(-1, 16)
gt = gt[(gt[:, 0] > 0) & (gt[:, 15] > 0)]
if len(gt):
  d = lambda a, b: np.percentile(gt[:, b] - gt[:, a], [10, 50, 90]).round()
  print("light group (first per warp), ns p10/p50/p90: phase1", d(0, 1), "round0", d(1, 2), "round1", d(2, 3),
        "round2", d(3, 4), "round3", d(4, 5), "phase3", d(5, 15), "total", d(0, 15), "n", len(gt))
  print("  round 0: issue loads", d(1, 8), "loads land", d(8, 9), "math", d(9, 10), "stores+flags", d(10, 2))
raw[131072:] = 0
t = raw.reshape(-1, 4)
nw = 20
t = t[: (len(t) // nw) * nw]
live = t[:, 0] > 0
t0 = t[live, 0].min()
start, heavy, end, groups = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3]
pc = lambda a: np.percentile(a, [0, 10, 50, 90, 100]).round()
print("warps", live.sum(), "span ns", end[live].max())
print("start:", pc(start[live])); print("heavy done:", pc(heavy[live])); print("end:", pc(end[live]))
print("light dur:", pc((end - heavy)[live])); print("groups/warp:", pc(groups[live]), "total", groups[live].sum())
print("per light group ns:", pc(((end - heavy) / np.maximum(groups, 1))[live & (groups > 0)]))
hb = (heavy - start)[live].reshape(-1, nw)[:, 0]
order = np.argsort(-hb)[:12]
print("longest heavy phases (block, ns):", [(int(b), int(hb[b])) for b in order])
uniq, idx, counts, num, seg_off, pos = plan.arrays()
c = counts[: int(num[0])].cpu().numpy()
print("U", c.size, "max count", c.max(), "heavy ids", (c > 32).sum(), "heavy occ", c[c > 32].sum())
