"""Times single stages of the plan-driven step under the current environment (tuning aid; the
env knobs KVHBM_APPLYP_* / KVHBM_PLAN_HEAVY are read once per process, so run one process per
setting).

  python scripts/plan_stage.py [--stages plan,gather,segsum,apply,chain] [--steps K] [--tag T]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tfplus_b200 import ops  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--keys", type=int, default=bench.KEYS)
  ap.add_argument("--steps", type=int, default=48)
  ap.add_argument("--stages", default="plan,gather,segsum,apply,chain")
  ap.add_argument("--tag", default="")
  ap.add_argument("--eager", action="store_true", help="no graphs (for ncu)")
  args = ap.parse_args()
  dev = torch.device("cuda:0")
  torch.cuda.set_device(dev)
  ops.set_today(bench.TODAY)
  B, D = bench.BATCH, bench.DIM
  st = bench.LocalStepper(args.keys, D, B, dev)
  st.populate()
  nb = bench.N_BATCHES
  ids_np, grads_np = bench.make_batches(nb, args.keys, B, D)
  ids_d = [torch.from_numpy(x).to(dev) for x in ids_np]
  grads_d = [torch.from_numpy(x).to(dev) for x in grads_np]
  plans = [ops.Plan(B, dev) for _ in range(nb)]
  rows = [torch.empty((B, D), dtype=torch.float32, device=dev) for _ in range(nb)]
  sums = torch.empty((B, D), dtype=torch.float32, device=dev)
  hp = st.hp

  fns = {
      "plan": lambda i: plans[i].build(ids_d[i]),
      "gather": lambda i: ops.kv_variable_gather_or_insert_plan(st.var, plans[i], out=rows[i]),
      "segsum": lambda i: ops.segment_sum_plan(plans[i], grads_d[i], out=sums),
      "apply": lambda i: ops.kv_variable_apply_plan(ops.OPT_GROUP_ADAM_V4, st.var, st.slot, None,
                                                    plans[i], grads_d[i], hp, advance_powers=True),
  }
  fns["chain"] = lambda i: (fns["gather"](i), fns["apply"](i))
  for i in range(nb):
    fns["plan"](i); fns["gather"](i); fns["segsum"](i); fns["apply"](i)
  torch.cuda.synchronize()
  ops.kv_variable_reserve(st.var, 2 * B)
  ops.kv_variable_reserve(st.slot, 2 * B)
  res = {}
  for name in args.stages.split(","):
    fn = fns[name]
    if args.eager:
      run = [lambda i=i: fn(i) for i in range(nb)]
    else:
      gs = []
      for i in range(nb):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
          fn(i)
        gs.append(g)
      run = [g.replay for g in gs]
    for i in range(3):
      run[i % nb]()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.steps):
      run[i % nb]()
    b.record()
    torch.cuda.synchronize()
    res[name] = round(a.elapsed_time(b) / args.steps * 1e3, 2)
  print("%s us: %s" % (args.tag or "default", res))


if __name__ == "__main__":
  main()
