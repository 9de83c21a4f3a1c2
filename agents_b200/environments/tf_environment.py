"""TFEnvironment contract (tf_agents/environments/tf_environment.py:37-241): batched
environments whose state lives in device tensors; `step` honours the auto-reset rule (after a
LAST step the next `step` ignores the action and returns FIRST, reward 0, discount 1)."""
import abc

from agents_b200.trajectories import time_step as ts


class TFEnvironment(abc.ABC):

  def __init__(self, time_step_spec=None, action_spec=None, batch_size=1):
    self._time_step_spec = time_step_spec
    self._action_spec = action_spec
    self._batch_size = batch_size

  def time_step_spec(self):
    return self._time_step_spec

  def action_spec(self):
    return self._action_spec

  def observation_spec(self):
    return self._time_step_spec.observation

  def reward_spec(self):
    return self._time_step_spec.reward

  @property
  def batched(self):
    return True

  @property
  def batch_size(self):
    return self._batch_size

  def current_time_step(self):
    """Returns the current TimeStep (tf_environment.py:185-197)."""
    return self._current_time_step()

  def reset(self):
    """Resets every environment and returns FIRST time steps (tf_environment.py:199-209)."""
    return self._reset()

  def step(self, action):
    """Applies `action` ([batch_size, ...]) and returns the next TimeStep (:211-241)."""
    return self._step(action)

  @abc.abstractmethod
  def _current_time_step(self):
    pass

  @abc.abstractmethod
  def _reset(self):
    pass

  @abc.abstractmethod
  def _step(self, action):
    pass
