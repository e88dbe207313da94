#!/usr/bin/env python
"""Generates tests/golden/*.npz: seeded inputs -> outputs of the CPU oracle (oracle/kv_oracle.cc),
so that the oracle itself is regression-pinned and the device path can be checked against
stored vectors without the oracle in the loop.

  python scripts/make_golden.py        # rewrites tests/golden/kv_golden_v1.npz

The reference cannot be run here (TensorFlow 2.13 + Bazel), so these are
vectors of the restatement, not of the reference binary: they pin behaviour against drift, the
reference's own known answers are restated one by one in tests/test_oracle_kat.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

TODAY = 19000
OUT = os.path.join(ROOT, "tests", "golden", "kv_golden_v1.npz")


def _export(tb):
  e = tb.export(first_n=6, enable_cutoff=False, cutoff_value=0.0, freq_u32=True)
  o = np.argsort(e["keys"])
  f = np.argsort(e["freq_keys"])
  return dict(keys=e["keys"][o], values=e["values"][o], blacklist=np.sort(e["blacklist"]),
              freq_keys=e["freq_keys"][f], freq_values=e["freq_values"][f])


def scenario(kind, dim, seed):
  """Four steps of lookup + dedup + duplicate-gradient sum + apply on Zipf ids."""
  rng = np.random.default_rng(seed)
  init = rng.normal(0, 0.05, size=(64, dim)).astype(np.float32)
  var = ob.OracleTable(dim, 2, seed=7)
  var.set_init_table(init)
  widths = {"adagrad": [dim], "group_adam_v4": [3 * dim], "group_adam_v3": [3 * dim],
            "sparse_group_ftrl": [dim, dim], "sparse_ftrl_v2": [dim, dim],
            "group_sparse_ftrl_v2": [dim, dim], "adam": [2 * dim]}[kind]
  inits = {"adagrad": [0.1], "sparse_group_ftrl": [0.1, 0.0], "sparse_ftrl_v2": [0.1, 0.0],
           "group_sparse_ftrl_v2": [0.1, 0.0]}.get(kind, [0.0])
  slots = []
  for w, v in zip(widths, inits):
    s = ob.OracleTable(w, 0, seed=7)
    s.set_init_table(np.full((8, w), v, np.float32))
    slots.append(s)
  out = {"init": init}
  b1p, b2p = 0.9, 0.999
  for step in range(4):
    ids = (rng.zipf(1.2, size=300) % 200).astype(np.int64)
    grad = rng.normal(size=(300, dim)).astype(np.float32)
    rows = var.gather_or_insert(ids, today=TODAY)
    u, idx, cnt = ob.unique(ids, with_counts=True)
    gs = ob.segment_sum(grad, idx, u.size)
    if kind == "adagrad":
      ob.apply_adagrad(var, slots[0], u, gs, 0.05, True, today=TODAY)
    elif kind == "group_adam_v4":
      ob.apply_group_adam_v4(var, slots[0], u, gs, 1e-2, b1p, b2p, 0.9, 0.999, 1e-8, 1e-4, 1e-4,
                             1e-3, today=TODAY)
    elif kind == "group_adam_v3":
      ob.apply_group_adam_v3(var, slots[0], u, gs, 1e-2, b1p, b2p, 0.9, 0.999, 1e-8, 1e-4, 1e-3,
                             1e-3, today=TODAY)
    elif kind == "sparse_group_ftrl":
      ob.apply_sparse_group_ftrl(var, slots[0], slots[1], u, gs, 0.1, 1e-3, 1e-3, 1e-2, 0.0, -0.5,
                                 today=TODAY)
    elif kind == "sparse_ftrl_v2":
      ob.apply_sparse_ftrl_v2(var, slots[0], slots[1], u, gs, 0.1, 1e-2, 1e-2, 1e-3, -0.5,
                              today=TODAY)
    elif kind == "group_sparse_ftrl_v2":
      ob.apply_group_sparse_ftrl_v2(var, slots[0], slots[1], u, gs, 0.1, 3.0, 1e-2, 0.0, -0.5,
                                    today=TODAY)
    elif kind == "adam":
      ob.adam_step(var, slots[0], u, gs, 1e-2, 0.9, 0.999, 1e-8, b1p, b2p, today=TODAY)
    b1p *= 0.9
    b2p *= 0.999
    out.update({"ids%d" % step: ids, "grad%d" % step: grad, "rows%d" % step: rows,
                "uniq%d" % step: u, "idx%d" % step: idx, "counts%d" % step: cnt,
                "gsum%d" % step: gs})
  for name, tb in [("var", var)] + [("slot%d" % i, s) for i, s in enumerate(slots)]:
    for k, v in _export(tb).items():
      out["%s_%s" % (name, k)] = v
  return out


def main():
  blob = {}
  for kind, dim, seed in [("adagrad", 8, 1), ("group_adam_v4", 16, 2), ("group_adam_v3", 8, 3),
                          ("sparse_group_ftrl", 8, 4), ("sparse_ftrl_v2", 4, 5),
                          ("group_sparse_ftrl_v2", 8, 6), ("adam", 8, 7)]:
    for k, v in scenario(kind, dim, seed).items():
      blob["%s/%d/%d/%s" % (kind, dim, seed, k)] = v
  np.savez_compressed(OUT, **blob)
  print("wrote %s: %d arrays, %.0f KB" % (OUT, len(blob), os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
  main()
