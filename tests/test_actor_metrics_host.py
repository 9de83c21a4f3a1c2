"""Host metrics and Actor (no GPU): replays tf_agents/metrics/py_metrics_test.py expectations and
runs train.Actor over the reference's mock env / policy."""
import json
import os

import numpy as np
import pytest

from agents_b200.metrics import py_metrics
from agents_b200.train import actor as actor_lib
from agents_b200.trajectories import trajectory
from py_env_mocks import PyEnvironmentMock, PyPolicyMock

F, M, L = 0, 1, 2


def _t(step_type, next_step_type, reward, discount=1.0):
  return trajectory.Trajectory(np.int32(step_type), (), (), (), np.int32(next_step_type),
                               np.float32(reward), np.float32(discount))


first = lambda r, d=1.0: _t(F, M, r, d)
mid = lambda r, d=1.0: _t(M, M, r, d)
last = lambda r, d=1.0: _t(M, L, r, d)
boundary = lambda r, d=1.0: _t(L, F, r, d)


def _stack(a, b):
  return trajectory.Trajectory(*[(np.stack([x, y]) if not isinstance(x, tuple) else ()) for x, y in zip(a, b)])


ALL = [py_metrics.AverageReturnMetric, py_metrics.AverageEpisodeLengthMetric, py_metrics.EnvironmentSteps,
       py_metrics.NumberOfEpisodes]


@pytest.mark.parametrize('cls,want', list(zip(ALL, [0.0, 0.0, 1.0, 0.0])))
def test_zero_episodes(cls, want):                        # py_metrics_test.py:60-77
  m = cls()
  m(boundary(0.0)); m(first(1.0))
  assert m.result() == want


@pytest.mark.parametrize('cls,want', list(zip(ALL, [6.0, 3.0, 3.0, 1.0])))
def test_average_one_episode(cls, want):                  # :79-96
  m = cls()
  m(boundary(0.0)); m(mid(1.0)); m(mid(2.0)); m(last(3.0, 0.0))
  assert m.result() == want


def test_average_one_episode_with_reset():                # :98-117
  m = py_metrics.AverageReturnMetric()
  m(first(0.0)); m(mid(1.0)); m(mid(2.0)); m(first(3.0)); m(last(4.0))
  assert m.result() == 7.0


@pytest.mark.parametrize('cls,want', list(zip(ALL, [0.0, 2.0, 4.0, 2.0])))
def test_average_two_episodes(cls, want):                 # :119-146
  m = cls()
  m(boundary(0.0)); m(first(1.0)); m(mid(2.0)); m(last(3.0, 0.0)); m(boundary(0.0)); m(_t(F, L, -6.0))
  assert m.result() == want


@pytest.mark.parametrize('batch_size', [None, 2])
@pytest.mark.parametrize('cls,want', [(py_metrics.AverageReturnMetric, 5.0),
                                      (py_metrics.AverageEpisodeLengthMetric, 2.5)])
def test_batch(cls, want, batch_size):                    # :148-232
  m = cls(batch_size=batch_size) if batch_size else cls()
  m(_stack(boundary(0.0), boundary(0.0)))
  m(_stack(first(1.0), first(1.0)))
  m(_stack(mid(2.0), last(3.0, 0.0)))
  m(_stack(last(3.0, 0.0), boundary(0.0)))
  m(_stack(boundary(0.0), first(1.0)))
  assert m.result() == want


def test_counter_and_deque():                             # :234-334
  c = py_metrics.CounterMetric()
  assert c.result() == 0
  c(); assert c.result() == 1
  c(); assert c.result() == 2
  c.reset(); assert c.result() == 0
  buf = py_metrics.NumpyDeque(maxlen=3, dtype=np.float64)
  buf.add(2); buf.add(6)
  assert buf.mean() == 4 and len(buf) == 2
  buf.add(4)
  assert buf.mean() == 4
  buf.extend([8, 9])                                      # past maxlen: keeps 4, 8, 9
  assert buf.mean() == 7 and buf.last == 9
  buf.clear(); buf.add(5)
  assert buf.mean() == 5
  unbounded = py_metrics.NumpyDeque(maxlen=np.inf, dtype=np.float64)
  unbounded.extend(range(101))
  assert unbounded.mean() == 50 and len(unbounded) == 101


def test_actor_runs_driver_and_updates_metrics(tmp_path):
  """train/actor.py: metrics observe every trajectory; run() continues where it stopped."""
  env, policy = PyEnvironmentMock(), PyPolicyMock()
  metrics = actor_lib.collect_metrics(buffer_size=10)
  seen = []
  act = actor_lib.Actor(env, policy, train_step=np.int64(0), episodes_per_run=2, observers=[seen.append],
                        metrics=metrics, summary_dir=str(tmp_path), summary_interval=1, name='collect')
  act.run()
  # two episodes of the mock: FIRST, LAST(+1 reward each), boundary  -> 6 trajectories, 4 steps
  assert len(seen) == 6
  episodes, steps, avg_return, avg_len = [m.result() for m in metrics]
  assert (episodes, steps, avg_return, avg_len) == (2, 4, 2.0, 2.0)
  act.run()
  assert metrics[0].result() == 4 and len(seen) == 12
  assert 'NumberOfEpisodes = 4' in act.log_metrics()
  recs = [json.loads(l) for l in open(os.path.join(str(tmp_path), 'metrics.jsonl'))]
  assert recs[0]['NumberOfEpisodes'] == 2.0 and recs[0]['actor'] == 'collect'
  with pytest.raises(ValueError, match='Unknown environment type'):
    actor_lib.Actor(object(), policy, train_step=0, steps_per_run=1)
