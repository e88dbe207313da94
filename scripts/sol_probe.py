#!/usr/bin/env python
"""Speed-of-light reference points on this box: torch index_select (a plain row gather with
known row indices) and index_add_ at the bench shape, timed as CUDA-graph replays."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
N, D, B = 10_000_000, 64, 65536
rows = torch.randn(N, D, device=dev)
ids_np, _ = bench.make_batches(16, N, B, D)
idx = [torch.from_numpy(x).to(dev) for x in ids_np]
out = torch.empty(B, D, device=dev)
def timeit(fn, name, reps=100):
  gs = []
  for i in range(16):
    g = torch.cuda.CUDAGraph()
    fn(i)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
      fn(i)
    gs.append(g)
  for i in range(5): gs[i].replay()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for i in range(reps): gs[i % 16].replay()
  b.record(); torch.cuda.synchronize()
  print("%-40s %.2f us" % (name, a.elapsed_time(b) / reps * 1e3))
timeit(lambda i: torch.index_select(rows, 0, idx[i], out=out), "index_select 65536 x 256B (Zipf)")
u = [torch.unique(x) for x in idx]
outu = [torch.empty(x.numel(), D, device=dev) for x in u]
timeit(lambda i: torch.index_select(rows, 0, u[i], out=outu[i]), "index_select ~20K unique rows")
emp = torch.empty(1, device=dev)
timeit(lambda i: emp.add_(1.0), "trivial kernel (graph launch floor)")
big = torch.empty(B, D, device=dev)
timeit(lambda i: big.copy_(out), "copy 16.8 MB")
src3 = torch.randn(N, 3 * D, device=dev)
out3 = [torch.empty(x.numel(), 3 * D, device=dev) for x in u]
timeit(lambda i: torch.index_select(src3, 0, u[i], out=out3[i]), "index_select ~20K x 768B")
