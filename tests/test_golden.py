"""tests/golden/kv_golden_v1.npz (scripts/make_golden.py): the oracle is re-run on the stored
inputs and must reproduce the stored outputs bit for bit (CPU, no GPU); on a GPU the device
path is checked against the same stored vectors with no oracle in the loop."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import make_golden  # noqa: E402

GOLD = np.load(make_golden.OUT)
CASES = sorted({tuple(k.split("/")[:3]) for k in GOLD.files})


@pytest.mark.parametrize("kind,dim,seed", CASES)
def test_oracle_reproduces_golden(kind, dim, seed):
  out = make_golden.scenario(kind, int(dim), int(seed))
  for k, v in out.items():
    g = GOLD["%s/%s/%s/%s" % (kind, dim, seed, k)]
    assert g.dtype == v.dtype and g.shape == v.shape, k
    np.testing.assert_array_equal(g, v, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,dim,seed", CASES)
def test_device_matches_golden(kind, dim, seed):
  import torch
  from tfplus_b200 import ops
  dim = int(dim)
  G = lambda k: GOLD["%s/%s/%s/%s" % (kind, dim, seed, k)]
  dev = torch.device("cuda:0")
  T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
  ops.set_today(make_golden.TODAY)
  var = ops.kv_variable(value_shape=[dim], enter_threshold=2, device=dev, seed=7)
  ops.init_kv_variable_v2(var, T(G("init")))
  widths = {"adagrad": [dim], "group_adam_v4": [3 * dim], "group_adam_v3": [3 * dim],
            "sparse_group_ftrl": [dim, dim], "sparse_ftrl_v2": [dim, dim],
            "group_sparse_ftrl_v2": [dim, dim], "adam": [2 * dim]}[kind]
  inits = {"adagrad": [0.1], "sparse_group_ftrl": [0.1, 0.0], "sparse_ftrl_v2": [0.1, 0.0],
           "group_sparse_ftrl_v2": [0.1, 0.0]}.get(kind, [0.0])
  slots = []
  for w, v in zip(widths, inits):
    s = ops.kv_variable(value_shape=[w], device=dev, seed=7)
    ops.init_kv_variable_v2(s, torch.full((8, w), v, device=dev))
    slots.append(s)
  opt = {"adagrad": ops.OPT_ADAGRAD, "group_adam_v4": ops.OPT_GROUP_ADAM_V4,
         "group_adam_v3": ops.OPT_GROUP_ADAM_V3, "sparse_group_ftrl": ops.OPT_SPARSE_GROUP_FTRL,
         "sparse_ftrl_v2": ops.OPT_SPARSE_FTRL_V2,
         "group_sparse_ftrl_v2": ops.OPT_GROUP_SPARSE_FTRL_V2, "adam": ops.OPT_ADAM}[kind]
  plan = ops.Plan(300, dev)
  b1p, b2p = 0.9, 0.999
  for step in range(4):
    plan.build(T(G("ids%d" % step)))
    uniq, idx, counts, num, _, _ = [x.cpu().numpy() for x in plan.arrays()]
    U = int(num[0])
    np.testing.assert_array_equal(uniq[:U], G("uniq%d" % step))
    np.testing.assert_array_equal(idx, G("idx%d" % step))
    np.testing.assert_array_equal(counts[:U], G("counts%d" % step))
    rows = ops.kv_variable_gather_or_insert_plan(var, plan).cpu().numpy()
    want = G("rows%d" % step).reshape(rows.shape)
    if step == 0:
      np.testing.assert_array_equal(rows, want)        # nothing but the initializer yet
    else:
      np.testing.assert_allclose(rows, want, rtol=1e-6, atol=1e-7)
    gsum = ops.segment_sum_plan(plan, T(G("grad%d" % step))).cpu().numpy()[:U]
    np.testing.assert_array_equal(gsum, G("gsum%d" % step))   # TF's order, bit for bit
    hp = {"adagrad": (0.05,),
          "group_adam_v4": (1e-2, b1p, b2p, 0.9, 0.999, 1e-8, 1e-4, 1e-4, 1e-3),
          "group_adam_v3": (1e-2, b1p, b2p, 0.9, 0.999, 1e-8, 1e-4, 1e-3, 1e-3),
          "sparse_group_ftrl": (0.1, 1e-3, 1e-3, 1e-2, 0.0, -0.5),
          "sparse_ftrl_v2": (0.1, 1e-2, 1e-2, 0.0, 1e-3, -0.5),
          "group_sparse_ftrl_v2": (0.1, 3.0, 1e-2, 0.0, 0.0, -0.5),
          "adam": (1e-2, 0.9, 0.999, 1e-8, b1p, b2p)}[kind]
    ops.kv_variable_apply_plan(opt, var, slots[0], slots[1] if len(slots) > 1 else None, plan,
                               T(G("grad%d" % step)), hp)
    b1p *= 0.9
    b2p *= 0.999
  for name, tb in [("var", var)] + [("slot%d" % i, s) for i, s in enumerate(slots)]:
    k, v, _, bl, fk, fv = ops.kv_variable_export(tb, first_n=6, enable_cutoff=False,
                                                 freq_dtype=torch.int32)
    k, v, bl, fk = (x.cpu().numpy() for x in (k, v, bl, fk))
    fv = fv.cpu().numpy().view(np.uint32)
    o, f = np.argsort(k), np.argsort(fk)
    np.testing.assert_array_equal(k[o], G(name + "_keys"))
    np.testing.assert_array_equal(np.sort(bl), G(name + "_blacklist"))
    np.testing.assert_array_equal(fk[f], G(name + "_freq_keys"))
    np.testing.assert_array_equal(fv[f], G(name + "_freq_values"))
    np.testing.assert_allclose(v[o], G(name + "_values").reshape(v.shape), rtol=1e-6, atol=1e-7,
                               err_msg=name)
  ops.set_today(None)
