"""Frame-dedup semantics (oracle/frame_stack.py): stacks rebuilt from a one-frame-per-slot ring are
bit-identical to the stacks a plain ring stores, for every sampleable id, including episode
starts and after the ring wrapped."""
import numpy as np
import pytest

from oracle import frame_stack as ofs
from oracle import replay as oreplay


@pytest.mark.parametrize('B_env,L,adds,K', [(1, 8, 5, 4), (3, 8, 8, 4), (2, 8, 21, 4), (4, 16, 50, 3), (2, 5, 13, 2)])
def test_rebuilt_stacks_equal_stored_stacks(B_env, L, adds, K):
  rng = np.random.RandomState(B_env * 100 + adds)
  H, W = 6, 5
  frames = rng.randint(0, 256, size=(B_env, adds, H, W)).astype(np.uint8)
  step_types = rng.choice([0, 1, 1, 1, 2], size=(B_env, adds)).astype(np.int32)
  step_types[:, 0] = 0
  for b in range(B_env):                                  # a LAST is followed by a FIRST
    for t in range(1, adds):
      if step_types[b, t - 1] == 2:
        step_types[b, t] = 0
  stacks = np.stack([ofs.stack_rule(frames[b], step_types[b], K) for b in range(B_env)])
  plain = oreplay.UniformReplayOracle([(), (H, W, K)], [np.int32, np.uint8], B_env, L)
  dedup = oreplay.UniformReplayOracle([(), (H, W)], [np.int32, np.uint8], B_env, L)
  for t in range(adds):
    plain.add_batch([step_types[:, t], stacks[:, t]])
    dedup.add_batch([step_types[:, t], frames[:, t]])
  lo, hi = ofs.valid_range_ids(dedup.last_id, L, 1, K)
  plo, phi = oreplay.valid_range_ids(plain.last_id, L, 1)
  assert hi == phi and lo == plo + (K - 1 if adds > L else 0)
  assert hi > lo
  for b in range(B_env):
    for id_ in range(lo, hi):
      want = plain.storage[1][b * L + id_ % L]
      got = ofs.rebuild(dedup.storage[1], dedup.storage[0], dedup.id_table, b, id_, K, L)
      np.testing.assert_array_equal(got, want)
  # storage: K x fewer observation bytes for the same history
  assert plain.storage[1].nbytes == K * dedup.storage[1].nbytes


def test_stack_rule_fills_with_first_frame():
  f = np.arange(5, dtype=np.uint8).reshape(5, 1, 1) + 10
  st = np.array([0, 1, 1, 0, 1], np.int32)
  s = ofs.stack_rule(f, st, 3)[:, 0, 0, :].tolist()
  assert s == [[10, 10, 10], [10, 10, 11], [10, 11, 12], [13, 13, 13], [13, 13, 14]]
