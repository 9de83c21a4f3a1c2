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


# ---- train/ppo_learner_test.py replayed on the oracle restatement -----------------------------------
def _reference_observations(n_time_steps, batch_size):
  """ppo_learner_test.py:83-96 `_create_trajectories`: obs[b, t] = 10 b + t."""
  return np.asarray([np.arange(n_time_steps) + 10 * i for i in range(batch_size)], np.float32)


@pytest.mark.parametrize('num_epochs,envs,mb,expected', [
    (1, 1, 10, 10), (2, 1, 10, 20), (2, 3, 10, 60), (1, 1, None, 1), (2, 1, None, 2), (2, 3, None, 2)])
def test_reference_one_element_dataset(num_epochs, envs, mb, expected):   # :193-253
  obs = _reference_observations(100, envs)
  n = obs.size
  assert opl.iterations_per_run(n, 1, num_epochs, mb, 1) == expected
  if mb:
    rows = opl.minibatch_rows(n, num_epochs, mb, shuffle_buffer_size=1, seed=0, call=0)
    stream = np.concatenate([obs] * num_epochs, 0).reshape(-1)      # _concat_and_flatten (:133-152)
    for i in range(expected):                                      # _get_expected_minibatch (:155-178)
      np.testing.assert_array_equal(obs.reshape(-1)[rows[i]], stream[mb * i:mb * (i + 1)])


@pytest.mark.parametrize('num_epochs,envs,mb,expected', [
    (1, 1, 10, 12), (2, 1, 10, 24), (2, 3, 10, 72), (1, 1, None, 3), (2, 1, None, 6), (2, 3, None, 6)])
def test_reference_multi_element_dataset(num_epochs, envs, mb, expected):  # :255-329
  episodes = 3
  obs = _reference_observations(40, envs)
  cache = np.concatenate([obs.reshape(-1)] * episodes)             # 3 cached samples, unbatched
  n = cache.size
  assert opl.iterations_per_run(n, episodes, num_epochs, mb, 1) == expected
  if mb:
    rows = opl.minibatch_rows(n, num_epochs, mb, shuffle_buffer_size=1, seed=0, call=0)
    stream = np.concatenate([obs] * (episodes * num_epochs), 0).reshape(-1)
    for i in range(expected):
      np.testing.assert_array_equal(cache[rows[i]], stream[mb * i:mb * (i + 1)])


def test_reference_parallel_iterations_count():                           # :331-376
  assert opl.iterations_per_run(3 * 40, 3, 4, 10, 1) == 48
