#!/usr/bin/env python
"""Runs a few eager steps of the bench workload (no CUDA graphs) so that ncu can capture the
individual kernels:  ncu --set full -k regex:<kernel> -s <skip> -c <n> python scripts/profile_step.py"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--keys", type=int, default=bench.KEYS)
ap.add_argument("--batch", type=int, default=bench.BATCH)
ap.add_argument("--dim", type=int, default=bench.DIM)
args = ap.parse_args()

from tfplus_b200 import ops  # noqa: E402
ops.set_today(bench.TODAY)
dev = torch.device("cuda:0")
st = bench.LocalStepper(args.keys, args.dim, args.batch, dev)
st.populate()
ids_np, grads_np = bench.make_batches(min(4, args.steps), args.keys, args.batch, args.dim)
ids = [torch.from_numpy(x).to(dev) for x in ids_np]
grads = [torch.from_numpy(x).to(dev) for x in grads_np]
plan = ops.Plan(args.batch, dev)
rows = torch.empty((args.batch, args.dim), dtype=torch.float32, device=dev)
for i in range(args.steps):
  st.step_eager(ids[i % len(ids)], grads[i % len(ids)], plan, rows)
torch.cuda.synchronize()
print("done")
