"""Helpers shared by the GPU parity tests: run the same call on the CUDA table (through
tfplus_b200.ops -> C ABI) and on the CPU oracle, and compare whole-table state."""
import numpy as np
import torch

from oracle import binding as ob
from tfplus_b200 import ops

TODAY = 19000
DEV = "cuda:0"


def t(a, dtype=None):
  return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(DEV)


class Pair:
  """One KvVariable on the GPU and its oracle twin, created identically."""

  def __init__(self, dim, enter_threshold=0, init=None, rows=256, seed=11, init_seed=0,
               capacity_hint=0):
    self.dim = dim
    if init is None:
      init = np.random.default_rng(init_seed).normal(0, 0.05, size=(rows, dim)).astype(np.float32)
    elif np.isscalar(init):
      init = np.full((rows, dim), init, np.float32)
    self.init = init
    self.gpu = ops.kv_variable(value_shape=[dim], enter_threshold=enter_threshold, device=DEV,
                               seed=seed, capacity_hint=capacity_hint)
    ops.init_kv_variable_v2(self.gpu, t(init))
    self.cpu = ob.OracleTable(dim, enter_threshold, seed=seed)
    self.cpu.set_init_table(init)

  # paired calls ------------------------------------------------------------
  def gather_or_insert(self, ids, counts=None, exact=True):
    ids = np.asarray(ids, np.int64)
    c = None if counts is None else t(np.asarray(counts, np.int32))
    got = ops.kv_variable_gather_or_insert_with_counts(self.gpu, t(ids), c).cpu().numpy()
    want = self.cpu.gather_or_insert(ids, counts, today=TODAY)
    if exact:
      np.testing.assert_array_equal(got, want.reshape(got.shape))
    return got, want

  def gather_or_zeros(self, ids):
    ids = np.asarray(ids, np.int64)
    got = ops.kv_variable_gather_or_zeros_v2(self.gpu, t(ids)).cpu().numpy()
    want = self.cpu.gather_or_zeros(ids)
    return got, want.reshape(got.shape)

  def scatter(self, op, ids, upd):
    ids = np.asarray(ids, np.int64)
    upd = np.asarray(upd, np.float32)
    getattr(ops, "kv_variable_scatter_%s_v2" % op)(self.gpu, t(ids), t(upd))
    self.cpu.scatter(op, ids, upd)

  def insert(self, ids, vals, filter_out=None, blacklist=None):
    ids = np.asarray(ids, np.int64)
    vals = np.asarray(vals, np.float32)
    f = None if filter_out is None else t(np.asarray(filter_out, np.uint8))
    b = None if blacklist is None else t(np.asarray(blacklist, np.uint8))
    ops.kv_variable_insert_v2(self.gpu, t(ids), t(vals), f, b)
    self.cpu.insert_or_update(ids, vals, filter_out, blacklist)

  # state comparison -----------------------------------------------------------
  def state_gpu(self):
    k, v, _, b, fk, fv = ops.kv_variable_export(self.gpu, first_n=6, enable_cutoff=True,
                                                cutoff_value=1e-20, freq_dtype=torch.int32)
    k, v, b, fk = (x.cpu().numpy() for x in (k, v, b, fk))
    fv = fv.cpu().numpy().view(np.uint32)
    return _as_state(k, v, b, fk, fv)

  def state_cpu(self):
    e = self.cpu.export(first_n=6, enable_cutoff=True, cutoff_value=1e-20, freq_u32=True)
    return _as_state(e["keys"], e["values"], e["blacklist"], e["freq_keys"], e["freq_values"])

  def check_state(self, rtol=0.0, atol=0.0, skip=()):
    """Membership, blacklist and frequency words bit-exact; rows exact or within tolerance."""
    g, c = self.state_gpu(), self.state_cpu()
    assert set(g["freq"]) == set(c["freq"]), "table membership differs"
    for key in c["freq"]:
      assert g["freq"][key] == c["freq"][key], (
          "freq word of key %d: gpu %#x oracle %#x" % (key, g["freq"][key], c["freq"][key]))
    skip = set(int(s) for s in skip)
    assert g["black"] - skip == c["black"] - skip, "blacklist differs: %s vs %s" % (
        sorted(g["black"] ^ c["black"])[:10], "")
    gk, ck = set(g["rows"]) - skip, set(c["rows"]) - skip
    assert gk == ck, "exported key set differs (under-threshold / low-frequency flags): %s" % (
        sorted(gk ^ ck)[:10])
    if ck:
      keys = sorted(ck)
      G = np.stack([g["rows"][k] for k in keys])
      C = np.stack([c["rows"][k] for k in keys])
      if rtol == 0.0 and atol == 0.0:
        np.testing.assert_array_equal(G, C)
      else:
        np.testing.assert_allclose(G, C, rtol=rtol, atol=atol)
    assert ops.kv_variable_size_v2(self.gpu) - 0 == self.cpu.size() or skip
    assert ops.kv_variable_shape_v2(self.gpu)[0] == self.cpu.map_size()
    return g, c


def _as_state(keys, vals, black, fkeys, fvals):
  return {
      "rows": {int(k): vals[i] for i, k in enumerate(keys)},
      "black": set(int(k) for k in black),
      "freq": {int(k): int(fvals[i]) for i, k in enumerate(fkeys)},
  }
