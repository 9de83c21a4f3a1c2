"""TFPolicy base (tf_agents/policies/tf_policy.py:276 `action`), torch tensors on CUDA."""
import torch

from agents_b200.trajectories import policy_step
from agents_b200.trajectories import trajectory
from agents_b200.utils import nest


class TFPolicy(object):

  def __init__(self, time_step_spec, action_spec, policy_state_spec=(), info_spec=(),
               name=None):
    self._time_step_spec = time_step_spec
    self._action_spec = action_spec
    self._policy_state_spec = policy_state_spec
    self._info_spec = info_spec
    self._name = name or type(self).__name__

  @property
  def time_step_spec(self):
    return self._time_step_spec

  @property
  def action_spec(self):
    return self._action_spec

  @property
  def policy_state_spec(self):
    return self._policy_state_spec

  @property
  def info_spec(self):
    return self._info_spec

  @property
  def policy_step_spec(self):
    return policy_step.PolicyStep(self._action_spec, self._policy_state_spec, self._info_spec)

  @property
  def trajectory_spec(self):
    ts = self._time_step_spec
    return trajectory.Trajectory(
        step_type=ts.step_type, observation=ts.observation, action=self._action_spec,
        policy_info=self._info_spec, next_step_type=ts.step_type, reward=ts.reward,
        discount=ts.discount)

  collect_data_spec = trajectory_spec

  def get_initial_state(self, batch_size=None):
    return ()

  def variables(self):
    return []

  def action(self, time_step, policy_state=(), seed=None):
    """Returns PolicyStep(action, state, info) (tf_policy.py:276-381)."""
    return self._action(time_step, policy_state, seed)

  def _action(self, time_step, policy_state, seed):
    raise NotImplementedError
