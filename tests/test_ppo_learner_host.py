"""PPOLearner host logic (no GPU): shuffle order of libb200rl vs the oracle restatement, the
stream/iteration arithmetic of train/ppo_learner.py:220-304, constructor checks (:172-192)."""
import types

import numpy as np
import pytest

from agents_b200.train import ppo_learner
from oracle import ppo_learner as opl


@pytest.mark.parametrize('n,buf', [(1, 1), (7, 1), (20, 5), (257, 64), (100, 100), (100, 1000), (4096, 333)])
def test_shuffle_order_matches_oracle_and_is_a_permutation(n, buf):
  got = ppo_learner.shuffle_order(n, buf, seed=1234, call=2)
  want = opl.shuffle_order(n, buf, 1234, 2)
  np.testing.assert_array_equal(got, want)
  assert sorted(got.tolist()) == list(range(n))
  # shuffle(buffer) can emit element e no earlier than position e - (buffer - 1)
  assert (got <= np.arange(n) + buf - 1).all()
  if buf == 1:
    np.testing.assert_array_equal(got, np.arange(n))


def test_shuffle_order_depends_on_seed_and_call():
  a = ppo_learner.shuffle_order(512, 512, 1, 0)
  assert (a != ppo_learner.shuffle_order(512, 512, 2, 0)).any()
  assert (a != ppo_learner.shuffle_order(512, 512, 1, 1)).any()
  np.testing.assert_array_equal(a, ppo_learner.shuffle_order(512, 512, 1, 0))


def test_minibatch_rows_and_iteration_count():
  rows = opl.minibatch_rows(num_frames=1000, num_epochs=10, minibatch_size=64,
                            shuffle_buffer_size=1000, seed=3, call=0)
  assert rows.shape == (156, 64) and rows.min() >= 0 and rows.max() < 1000
  # the learner consumes int(1000/64)*10 = 150 of the 156 minibatches per run (:283-292)
  assert opl.iterations_per_run(1000, 1, 10, 64, 1) == 150
  assert opl.iterations_per_run(1000, 1, 10, 64, 4) == 37
  assert opl.iterations_per_run(0, 3, 25, None, 1) == 75
  with pytest.raises(ValueError, match='Cannot distribute'):
    opl.iterations_per_run(10, 1, 1, 64, 1)


def _fake_agent(in_train, update_norm):
  return types.SimpleNamespace(_compute_value_and_advantage_in_train=in_train,
                               update_normalizers_in_train=update_norm)


def test_constructor_checks():
  ds = lambda: iter(())
  with pytest.raises(ValueError, match='shuffle_buffer_size must be provided'):
    ppo_learner.PPOLearner('/tmp/x', None, _fake_agent(False, False), ds, ds, 1, minibatch_size=8)
  with pytest.raises(ValueError, match='compute_value_and_advantage_in_train should be set to False'):
    ppo_learner.PPOLearner('/tmp/x', None, _fake_agent(True, False), ds, ds, 1, minibatch_size=8,
                           shuffle_buffer_size=16)
  with pytest.raises(ValueError, match='update_normalizers_in_train should be set to False'):
    ppo_learner.PPOLearner('/tmp/x', None, _fake_agent(False, True), ds, ds, 1)
