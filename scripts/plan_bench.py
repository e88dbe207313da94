"""Stage timings of the plan-driven step (tuning aid; bench.py is the measurement of record).

  python scripts/plan_bench.py [--keys N] [--steps K] [--eager]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tfplus_b200 import ops  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--keys", type=int, default=bench.KEYS)
  ap.add_argument("--steps", type=int, default=64)
  ap.add_argument("--batch", type=int, default=bench.BATCH)
  ap.add_argument("--dim", type=int, default=bench.DIM)
  args = ap.parse_args()
  dev = torch.device("cuda:0")
  torch.cuda.set_device(dev)
  ops.set_today(bench.TODAY)
  st = bench.LocalStepper(args.keys, args.dim, args.batch, dev)
  st.populate()
  nb = bench.N_BATCHES
  ids_np, grads_np = bench.make_batches(nb, args.keys, args.batch, args.dim)
  ids_d = [torch.from_numpy(x).to(dev) for x in ids_np]
  grads_d = [torch.from_numpy(x).to(dev) for x in grads_np]
  plans = [ops.Plan(args.batch, dev) for _ in range(nb)]
  rows = [torch.empty((args.batch, args.dim), dtype=torch.float32, device=dev) for _ in range(nb)]
  hp = st.hp

  def s_plan(i): plans[i].build(ids_d[i])
  def s_gather(i): ops.kv_variable_gather_or_insert_plan(st.var, plans[i], out=rows[i])
  def s_apply(i):
    ops.kv_variable_apply_plan(ops.OPT_GROUP_ADAM_V4, st.var, st.slot, None, plans[i], grads_d[i],
                               hp, advance_powers=True)
  def s_step(i):
    s_gather(i)
    s_apply(i)

  for i in range(nb):       # eager warm-up: sizes every scratch buffer
    s_plan(i); s_gather(i); s_apply(i)
  torch.cuda.synchronize()
  ops.kv_variable_reserve(st.var, 2 * args.batch)
  ops.kv_variable_reserve(st.slot, 2 * args.batch)

  def graphs(fn):
    out = []
    for i in range(nb):
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g):
        fn(i)
      out.append(g)
    return out

  def timeit(gs, K):
    for i in range(3):
      gs[i % nb].replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
      gs[i % nb].replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K * 1e3

  res = {}
  for name, fn in [("plan_build", s_plan), ("gather_plan", s_gather), ("apply_plan", s_apply),
                   ("gather+apply", s_step)]:
    res[name] = timeit(graphs(fn), args.steps)
  print("us per stage:", {k: round(v, 2) for k, v in res.items()})
  print("keys/s (gather+apply chain): %.3f G" % (args.batch / res["gather+apply"] / 1e3))


if __name__ == "__main__":
  main()
