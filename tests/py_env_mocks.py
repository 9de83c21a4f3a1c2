"""Host (numpy) restatement of the reference's PyEnvironmentMock (drivers/test_utils.py:43-93):
state += action; FIRST(0) -> MID ... -> LAST once state >= final_state; a step after LAST resets."""
import numpy as np
import torch

from agents_b200.environments import py_environment
from agents_b200.specs import tensor_spec


class PyEnvironmentMock(py_environment.PyEnvironment):

  def __init__(self, final_state=3):
    super().__init__()
    self._state = 0
    self._final = final_state
    self.actions_taken = []
    self.steps = 0
    self.resets = 0

  def observation_spec(self):
    return tensor_spec.TensorSpec([], torch.int32, 'observation')

  def action_spec(self):
    return tensor_spec.BoundedTensorSpec([], torch.int32, 1, 2, 'action')

  def _reset(self):
    self._state = 0
    self.resets += 1
    return py_environment.restart(np.int32(0))

  def _step(self, action):
    if self._state >= self._final:
      return self.reset()
    self.actions_taken.append(int(action))
    self.steps += 1
    self._state += int(action)
    obs = np.int32(self._state)
    if self._state < self._final:
      return py_environment.transition(obs, 1.0)
    return py_environment.termination(obs, 1.0)

  def get_info(self):
    return {'mock': 1}


class PyPolicyMock(object):
  """drivers/test_utils.py:169-205 PyPolicyMock: actions 1, 2 alternating from the policy state
  (reset to `initial_policy_state` on FIRST steps); info = 2 * action."""

  def __init__(self, initial_policy_state=np.int32(2)):
    self._initial = initial_policy_state
    self.get_initial_state_call_count = 0

  def get_initial_state(self, batch_size=None):
    self.get_initial_state_call_count += 1
    return self._initial

  def action(self, time_step, policy_state=()):
    from agents_b200.trajectories import policy_step
    first = time_step.is_first()
    if np.ndim(first) == 0:
      if first:
        policy_state = self._initial
    else:
      policy_state = np.array(policy_state)
      policy_state[first] = self._initial[first]
    action = (policy_state % 2) + 1
    return policy_step.PolicyStep(np.int32(action), np.int32(policy_state + 1), np.int32(action * 2))


class CountingEnv(py_environment.PyEnvironment):
  """Restatement of the reference's environments/test_envs.py:32-74 CountingEnv: the observation
  counts the steps of the episode (+ 10 x finished episodes); an episode ends with reward 1 after
  `steps_per_episode` steps; auto-reset on the step after LAST."""

  def __init__(self, steps_per_episode=10, dtype=np.int32):
    super().__init__(handle_auto_reset=True)
    self._n = steps_per_episode
    self._dtype = np.dtype(dtype)
    self._episodes = 0
    self._t = 0

  def observation_spec(self):
    info = np.iinfo(self._dtype)
    return tensor_spec.BoundedTensorSpec((), self._dtype, info.min, info.max, 'observation')

  def action_spec(self):
    return tensor_spec.BoundedTensorSpec((), self._dtype, 0, 1, 'action')

  def _obs(self):
    return np.array(10 * self._episodes + self._t, dtype=self._dtype)

  def _reset(self):
    if self._current_time_step is not None and np.all(self._current_time_step.is_last()):
      self._episodes += 1
    self._t = 0
    return py_environment.restart(self._obs())

  def _step(self, action):
    del action
    self._t += 1
    if self._t < self._n:
      return py_environment.transition(self._obs(), 0.0)
    return py_environment.termination(self._obs(), 1.0)

  def get_info(self):
    return {}


class EpisodeCountingEnv(py_environment.PyEnvironment):
  """Reference environments/test_envs.py:78-117: observation = (episode index, step index)."""

  def __init__(self, steps_per_episode=10):
    super().__init__(handle_auto_reset=True)
    self._n = steps_per_episode
    self._episodes = 0
    self._steps = 0

  def observation_spec(self):
    big = np.iinfo(np.int32).max
    return (tensor_spec.BoundedTensorSpec((), np.int32, 0, big, 'episode'),
            tensor_spec.BoundedTensorSpec((), np.int32, 0, big, 'step'))

  def action_spec(self):
    return tensor_spec.BoundedTensorSpec((), np.int32, 0, 1, 'action')

  def _obs(self):
    return (np.array(self._episodes, np.int32), np.array(self._steps, np.int32))

  def _reset(self):
    if self._current_time_step is not None and np.all(self._current_time_step.is_last()):
      self._episodes += 1
      self._steps = 0
    return py_environment.restart(self._obs())

  def _step(self, action):
    del action
    self._steps += 1
    if self._steps < self._n:
      return py_environment.transition(self._obs(), 0.0)
    return py_environment.termination(self._obs(), 1.0)

  def get_info(self):
    return {}


class NumpyStepStore(object):
  """Test double of `reverb_local.HbmStepStore` (same protocol, host arrays): lets the table /
  writer / row-pool logic be checked without a GPU.  `commit` asserts what the device write
  relies on: no row appears twice in one launch."""

  def __init__(self, specs, capacity, stage=8):
    self._specs = list(specs)
    self.capacity, self.stage = int(capacity), int(stage)
    np_dt = [tensor_spec.as_numpy_dtype(s.dtype) for s in self._specs]
    self._data = [np.zeros((self.capacity,) + s.shape, dt) for s, dt in zip(self._specs, np_dt)]
    self._stage = [np.zeros((self.stage,) + s.shape, dt) for s, dt in zip(self._specs, np_dt)]
    self.commits = 0
    self.reads = 0

  def staging(self):
    return self._stage

  def commit(self, rows, n):
    rows = np.asarray(rows[:n])
    assert len(set(rows.tolist())) == n, 'a row is written twice by one launch'
    for d, s in zip(self._data, self._stage):
      d[rows] = s[:n]
    self.commits += 1

  def read(self, rows):
    self.reads += 1
    return [d[np.asarray(rows)] for d in self._data]

  def grow(self, capacity):
    self._data = [np.concatenate([d, np.zeros((capacity - d.shape[0],) + d.shape[1:], d.dtype)])
                  for d in self._data]
    self.capacity = int(capacity)
