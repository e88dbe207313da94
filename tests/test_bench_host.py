"""Host-side logic of bench.py that needs no GPU: how K steps are split into rotation replays
and single steps, the algorithmic-byte model, and the Zipf sampler's determinism."""
import types

import numpy as np
import pytest

import bench


class _Graph:
  def __init__(self, log, tag):
    self.log, self.tag = log, tag

  def replay(self):
    self.log.append(self.tag)


@pytest.mark.parametrize("K", [0, 1, 5, 15, 16, 17, 32, 200, 208])
@pytest.mark.parametrize("with_rotation", [True, False])
def test_local_run_steps_covers_exactly_k_steps_in_batch_order(K, with_rotation):
  n, log = bench.N_BATCHES, []
  st = types.SimpleNamespace(full=[_Graph(log, ("step", i)) for i in range(n)],
                             rotation=_Graph(log, ("rotation",)) if with_rotation else None,
                             plans=[types.SimpleNamespace(build=lambda ids: None)], ids_d=[None])
  bench.LocalStepper.run_steps(st, K)
  batches = []
  for ev in log:
    batches.extend(range(n) if ev[0] == "rotation" else [ev[1]])
  assert batches == [i % n for i in range(K)]
  if with_rotation:
    assert sum(1 for ev in log if ev[0] == "rotation") == K // n


@pytest.mark.parametrize("K", [0, 3, 16, 35, 200])
def test_sharded_run_steps_covers_exactly_k_steps(K):
  n, log = 16, []
  st = types.SimpleNamespace(ids_d=[None] * n, rotation=_Graph(log, ("rotation",)), steps_done=0)
  st.step = lambda i: log.append(("step", i % n))
  bench.ShardedStepper.run_steps(st, K)
  batches = []
  for ev in log:
    batches.extend(range(n) if ev[0] == "rotation" else [ev[1]])
  assert batches == [i % n for i in range(K)]


def test_algorithmic_bytes_match_the_survey_model():
  # 536 B per looked-up id, 2336 B per applied unique id (GroupAdam, D = 64)
  B, U, D = 65536, 20300, 64
  ab = bench.algorithmic_bytes(B, U, D)
  assert ab["gather"] == 536 * B
  assert ab["unique"] == 12 * B + 8 * U
  # the fused pass is charged what the survey charges the two ops it replaces
  assert ab["segment_sum+apply"] == (260 * B + 256 * U) + 2336 * U
  assert abs(sum(ab.values()) / 1e6 - 105.7) < 0.1


def test_parity_check_batches_are_deterministic_with_duplicates():
  a = bench._check_batches(2, 2, 40000, 2048, 4)
  b = bench._check_batches(2, 2, 40000, 2048, 4)
  for s in range(2):
    for r in range(2):
      np.testing.assert_array_equal(a[s][r][0], b[s][r][0])
      np.testing.assert_array_equal(a[s][r][1], b[s][r][1])
  ids = np.concatenate([a[0][0][0], a[0][1][0]])
  assert np.unique(ids).size < ids.size - 100      # hot keys repeat within and across ranks
  assert set(np.intersect1d(a[0][0][0], a[0][1][0])) != set()


def test_batches_are_deterministic_and_zipf_shaped():
  a_ids, a_g = bench.make_batches(2, 10_000_000, 65536, 8, seed_ids=2024, seed_grad=7)
  b_ids, b_g = bench.make_batches(2, 10_000_000, 65536, 8, seed_ids=2024, seed_grad=7)
  for x, y in zip(a_ids + a_g, b_ids + b_g):
    np.testing.assert_array_equal(x, y)
  ids = a_ids[0]
  assert ids.dtype == np.int64 and ids.min() >= 0 and ids.max() < 10_000_000
  u = np.unique(ids).size
  assert 19_000 < u < 23_000          # SURVEY: U ~ 20.3 K unique of 65 536 at Zipf(1.1)
  _, counts = np.unique(ids, return_counts=True)
  assert counts.max() > 5_000         # the head id takes ~10 % of the batch


@pytest.mark.parametrize("K,n", [(20, 10), (200, 10), (208, 16), (16, 16), (30, 10), (24, 12),
                                 (100, 10), (17, 16), (3, 16), (64, 16), (28, 14)])
def test_rotation_length_divides_the_timed_steps(K, n):
  # all K timed steps run as whole rotation graphs whenever an even n in 8..16 divides K
  assert bench.pick_rotation(K) == n
  assert n % 2 == 0 and 8 <= n <= 16
