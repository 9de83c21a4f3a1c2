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
