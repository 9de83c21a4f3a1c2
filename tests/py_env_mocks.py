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
