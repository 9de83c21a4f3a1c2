"""PyDriver on the host (no GPU): replays tf_agents/drivers/py_driver_test.py with the numpy
restatements of its PyEnvironmentMock / PyPolicyMock (tests/py_env_mocks.py)."""
import numpy as np
import pytest

from agents_b200.drivers import py_driver
from agents_b200.environments import batched_py_environment
from py_env_mocks import PyEnvironmentMock, PyPolicyMock


def _traj(step_type, obs, action, info, next_step_type, reward, discount):
  return dict(step_type=step_type, observation=obs, action=action, policy_info=info,
              next_step_type=next_step_type, reward=reward, discount=discount)


# py_driver_test.py:47-60: first(0,1,2,1,1) last(1,2,4,1,0) boundary(3,1,2,0,1) ...
EXPECTED = [_traj(0, 0, 1, 2, 1, 1., 1.), _traj(1, 1, 2, 4, 2, 1., 0.), _traj(2, 3, 1, 2, 0, 0., 1.)] * 2 + [
    _traj(0, 0, 1, 2, 1, 1., 1.)]


def _as_dict(t):
  return {k: np.asarray(v).tolist() for k, v in t._asdict().items()}


def _run(max_steps, max_episodes, runs=1, **kw):
  env, policy, seen = PyEnvironmentMock(), PyPolicyMock(), []
  drv = py_driver.PyDriver(env, policy, observers=[seen.append], max_steps=max_steps,
                           max_episodes=max_episodes, **kw)
  time_step, state = env.reset(), policy.get_initial_state()
  for _ in range(runs):
    time_step, state = drv.run(time_step, state)
  return seen, policy


@pytest.mark.parametrize('max_steps,max_episodes,expected', [
    (None, 1, 3), (None, 2, 6), (2, 2, 2), (4, 2, 5), (4, 1, 3), (4, None, 5)])
def test_run_once(max_steps, max_episodes, expected):               # :62-88
  seen, _ = _run(max_steps, max_episodes)
  assert [_as_dict(t) for t in seen] == EXPECTED[:expected]


@pytest.mark.parametrize('max_steps,max_episodes,expected', [
    (None, 1, 2), (None, 2, 5), (2, 2, 2), (4, 2, 5), (4, 1, 2), (4, None, 5)])
def test_run_once_transition_observer(max_steps, max_episodes, expected):   # :90-121
  env, policy, seen = PyEnvironmentMock(), PyPolicyMock(), []
  drv = py_driver.PyDriver(env, policy, observers=[], transition_observers=[seen.append],
                           max_steps=max_steps, max_episodes=max_episodes,
                           end_episode_on_boundary=False)
  drv.run(env.reset(), policy.get_initial_state())
  assert len(seen) == expected and len(seen[0]) == 3


def test_info_observer():                                           # :123-141
  env, policy, infos = PyEnvironmentMock(), PyPolicyMock(), []
  drv = py_driver.PyDriver(env, policy, observers=[], info_observers=[infos.append], max_steps=2)
  drv.run(env.reset(), policy.get_initial_state())
  assert infos == [{'mock': 1}, {'mock': 1}]


def test_multiple_runs():                                           # :143-191
  seen, _ = _run(1, None, runs=3)
  assert [_as_dict(t) for t in seen] == EXPECTED[:4]
  seen, _ = _run(None, 1, runs=2)
  assert [_as_dict(t) for t in seen] == EXPECTED[:6]


def test_policy_state_reset():                                      # :193-214
  seen, policy = _run(None, 2)
  assert [_as_dict(t) for t in seen] == EXPECTED[:6]
  assert policy.get_initial_state_call_count == 2


@pytest.mark.parametrize('max_steps,max_episodes', [(None, None), (0, None), (None, 0), (0, 0)])
def test_invalid_args(max_steps, max_episodes):                     # :216-233
  with pytest.raises(ValueError):
    py_driver.PyDriver(PyEnvironmentMock(), PyPolicyMock(), observers=[], max_steps=max_steps,
                       max_episodes=max_episodes)


@pytest.mark.parametrize('max_steps,max_episodes,expected', [
    (4, None, 2), (5, None, 3), (None, 2, 4), (2, 2, 1), (4, 2, 2)])
def test_batched_environment(max_steps, max_episodes, expected):    # :235-327
  want = [_traj([0, 0], [0, 0], [2, 1], [4, 2], [1, 1], [1., 1.], [1., 1.]),
          _traj([1, 1], [2, 1], [1, 2], [2, 4], [2, 1], [1., 1.], [0., 1.]),
          _traj([2, 1], [3, 3], [2, 1], [4, 2], [0, 2], [0., 1.], [1., 0.]),
          _traj([0, 2], [0, 4], [2, 2], [4, 4], [1, 0], [1., 0.], [1., 1.])]
  env = batched_py_environment.BatchedPyEnvironment([PyEnvironmentMock(3), PyEnvironmentMock(4)])
  policy, seen = PyPolicyMock(initial_policy_state=np.array([1, 2])), []
  drv = py_driver.PyDriver(env, policy, observers=[seen.append], max_steps=max_steps,
                           max_episodes=max_episodes)
  drv.run(env.reset(), policy.get_initial_state())
  assert [_as_dict(t) for t in seen] == want[:expected]
  env.close()


def test_parallel_py_environment_matches_batched():
  """environments/parallel_py_environment.py: worker processes give the same batched stream as
  the in-process BatchedPyEnvironment; worker exceptions surface in the parent."""
  from agents_b200.environments import parallel_py_environment
  ctors = [lambda: PyEnvironmentMock(3), lambda: PyEnvironmentMock(4), lambda: PyEnvironmentMock(5)]
  for blocking in (False, True):
    par = parallel_py_environment.ParallelPyEnvironment(ctors, start_serially=not blocking, blocking=blocking)
    ref = batched_py_environment.BatchedPyEnvironment([c() for c in ctors], multithreading=False)
    assert par.batched and par.batch_size == 3 and par.action_spec() == ref.action_spec()
    a, b = par.reset(), ref.reset()
    for step in range(12):
      for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
      act = np.array([1 + (step + i) % 2 for i in range(3)], np.int32)
      a, b = par.step(act), ref.step(act)
    par.close(); ref.close()
  with pytest.raises(TypeError, match='non-callable'):
    parallel_py_environment.ParallelPyEnvironment([PyEnvironmentMock()])

  def broken():
    raise ValueError('boom in worker')
  with pytest.raises(RuntimeError, match='boom in worker'):
    parallel_py_environment.ParallelPyEnvironment([broken])
