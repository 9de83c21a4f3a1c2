"""Host-side (numpy) environment contract — PyEnvironment of the reference
(tf_agents/environments/py_environment.py:37-365).

Real simulators (gym, MuJoCo, ALE) live on the host; they are the only part of the collect path
that is not on the GPU.  `TFPyEnvironment` (tf_py_environment.py) stages their batched outputs
through pinned memory into device TimeSteps for `DynamicStepDriver` / `add_batch`.

Kept from the reference: `reset()` caches the TimeStep (:185-201); `step(action)` resets instead
of stepping when there is no current TimeStep or `should_reset` says so (:203-239);
`should_reset` = `handle_auto_reset and all(is_last)` (:106-118).  Specs are
`agents_b200.specs.tensor_spec` objects; TimeSteps hold numpy arrays.
"""
import abc

import numpy as np

from agents_b200.trajectories import time_step as ts


def restart(observation, batch_size=None, reward_dtype=np.float32):
  """numpy `ts.restart` (trajectories/time_step.py:160-211)."""
  shape = () if batch_size is None else (batch_size,)
  return ts.TimeStep(np.full(shape, ts.StepType.FIRST, np.int32), np.zeros(shape, reward_dtype),
                     np.ones(shape, np.float32), observation)


def transition(observation, reward, discount=1.0):
  """numpy `ts.transition` (:214-267); outer shape follows `reward`."""
  reward = np.asarray(reward, np.float32)
  return ts.TimeStep(np.full(reward.shape, ts.StepType.MID, np.int32), reward,
                     np.broadcast_to(np.asarray(discount, np.float32), reward.shape).copy(),
                     observation)


def termination(observation, reward):
  """numpy `ts.termination` (:270-317): LAST with discount 0."""
  reward = np.asarray(reward, np.float32)
  return ts.TimeStep(np.full(reward.shape, ts.StepType.LAST, np.int32), reward,
                     np.zeros(reward.shape, np.float32), observation)


def truncation(observation, reward, discount=1.0):
  """numpy `ts.truncation` (:320-371): LAST that keeps the discount."""
  reward = np.asarray(reward, np.float32)
  return ts.TimeStep(np.full(reward.shape, ts.StepType.LAST, np.int32), reward,
                     np.broadcast_to(np.asarray(discount, np.float32), reward.shape).copy(),
                     observation)


class PyEnvironment(abc.ABC):
  """Abstract base class for Python RL environments."""

  def __init__(self, handle_auto_reset=False):
    self._handle_auto_reset = handle_auto_reset
    self._current_time_step = None

  @property
  def batched(self):
    return False

  @property
  def batch_size(self):
    if self.batched:
      raise RuntimeError('Environment %s marked itself as batched but did not override the '
                         'batch_size property' % type(self))
    return None

  def should_reset(self, current_time_step):
    handle_auto_reset = getattr(self, '_handle_auto_reset', False)
    return bool(handle_auto_reset and np.all(current_time_step.step_type == ts.StepType.LAST))

  @abc.abstractmethod
  def observation_spec(self):
    pass

  @abc.abstractmethod
  def action_spec(self):
    pass

  def reward_spec(self):
    from agents_b200.specs import tensor_spec
    import torch
    return tensor_spec.TensorSpec((), torch.float32, 'reward')

  def time_step_spec(self):
    return ts.time_step_spec(self.observation_spec(), self.reward_spec())

  def current_time_step(self):
    return self._current_time_step

  def reset(self):
    self._current_time_step = self._reset()
    return self._current_time_step

  def step(self, action):
    if self._current_time_step is None or self.should_reset(self._current_time_step):
      return self.reset()
    self._current_time_step = self._step(action)
    return self._current_time_step

  def close(self):
    pass

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self.close()

  def get_info(self):
    raise NotImplementedError('No support of get_info for this environment.')

  @abc.abstractmethod
  def _step(self, action):
    pass

  @abc.abstractmethod
  def _reset(self):
    pass
