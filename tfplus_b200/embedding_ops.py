"""embedding_lookup / embedding_lookup_sparse / safe_embedding_lookup_sparse over KvVariables —
mirror of tfplus/kv_variable/python/ops/embedding_ops.py."""
import torch

from . import ops
from .kv_variable import KvVariable, PartitionedKvVariable


def _parts(params):
  if isinstance(params, PartitionedKvVariable):
    return params.parts
  if isinstance(params, (list, tuple)):
    return list(params)
  return [params]


def embedding_lookup(params, ids, partition_strategy="mod", name=None, validate_indices=True,
                     max_norm=None, counts=None):
  """embedding_ops.py:242 -> _embedding_lookup_and_transform (:48-204).  One shard: a plain
  sparse_read[_with_counts]; N shards: p = ids % N, per-shard gather, stitch back in order."""
  parts = _parts(params)
  if not isinstance(ids, torch.Tensor):
    ids = torch.as_tensor(ids, dtype=torch.int64)
  ids = ids.to(parts[0].device)
  if len(parts) == 1:
    out = parts[0].sparse_read_with_counts(ids, counts)
  else:
    if partition_strategy != "mod":
      raise NotImplementedError("only the 'mod' partition strategy is used with KvVariable")
    flat = ids.reshape(-1)
    n = len(parts)
    # embedding_ops.py:121-127 floormod(ids, np); dynamic_partition / dynamic_stitch
    sorted_ids, perm, shard_counts = ops.partition_ids(flat, n, mode="mod")
    c = shard_counts.cpu().tolist()
    sorted_counts = None
    if counts is not None:
      sorted_counts = torch.empty_like(counts.reshape(-1))
      sorted_counts[perm.long()] = counts.reshape(-1).to(flat.device)
    rows, off = [], 0
    for p, k in zip(parts, c):
      sl = slice(off, off + k)
      rows.append(p.sparse_read_with_counts(
          sorted_ids[sl], None if sorted_counts is None else sorted_counts[sl]))
      off += k
    out = ops.permute_rows(torch.cat(rows, 0), perm).reshape(tuple(ids.shape) + (parts[0].embedding_dim,))
  if max_norm is not None:
    norm = out.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    out = out * torch.clamp(max_norm / norm, max=1.0)
  return out


def embedding_lookup_sparse(params, sp_ids, sp_weights=None, partition_strategy="mod", name=None,
                            combiner="mean", max_norm=None):
  """embedding_ops.py:320-441.  sp_ids = (indices[nnz, 2] or segment row ids [nnz], values[nnz],
  dense_shape); dedups with unique (unique_with_counts when the table has an enter_threshold,
  :365-372), gathers once per distinct id, combines per row with sum / mean / sqrtn."""
  indices, values, dense_shape = sp_ids
  parts = _parts(params)
  dev = parts[0].device
  values = torch.as_tensor(values, dtype=torch.int64).to(dev)
  seg = torch.as_tensor(indices).to(dev)
  if seg.dim() == 2:
    seg = seg[:, 0]
  seg = seg.to(torch.int64)
  n_rows = int(dense_shape[0])
  if parts[0].enter_threshold > 0:
    uniq, idx, counts = ops.unique(values, with_counts=True)
  else:
    (uniq, idx), counts = ops.unique(values), None
  emb = embedding_lookup(params, uniq, partition_strategy, max_norm=max_norm, counts=counts)
  w = None
  if sp_weights is not None:
    w = torch.as_tensor(sp_weights[1] if isinstance(sp_weights, tuple) else sp_weights,
                        dtype=torch.float32).to(dev)
  # expand through the inverse index, weight, reduce per row, normalise: one kernel
  return ops.sparse_combine(emb.reshape(uniq.numel(), -1), idx, seg, w, n_rows, combiner)


def safe_embedding_lookup_sparse(embedding_weights, sparse_ids, sparse_weights=None,
                                 combiner="mean", default_id=None, name=None,
                                 partition_strategy="mod", max_norm=None):
  """embedding_ops.py:443-560: rows with no ids get the default id (or zeros); negative ids are
  valid KvVariable keys (py_ut/tests/test_embedding_ops.py:301-337), so nothing is pruned."""
  indices, values, dense_shape = sparse_ids
  out = embedding_lookup_sparse(embedding_weights, sparse_ids, sparse_weights,
                                partition_strategy, combiner=combiner, max_norm=max_norm)
  if default_id is not None:
    seg = torch.as_tensor(indices)
    seg = seg[:, 0] if seg.dim() == 2 else seg
    present = torch.zeros(int(dense_shape[0]), dtype=torch.bool, device=out.device)
    present[seg.to(out.device).long()] = True
    if (~present).any():
      d = embedding_lookup(embedding_weights, torch.tensor([default_id], dtype=torch.int64))
      out = torch.where(present[:, None], out, d.reshape(1, -1).expand_as(out))
  return out
