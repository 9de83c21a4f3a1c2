"""RandomTFPolicy: uniform random discrete actions, optionally restricted by an action mask.

Reference: policies/random_tf_policy.py:80-161 (masked case: sample among the allowed actions,
:104-131; unmasked: `tensor_spec.sample_spec_nest`, :137).  Device work: the shared
b200rl_epsilon_greedy launch with epsilon = 2 (always the random branch) over zero Q values.
"""
import torch

from agents_b200.policies.greedy_policy import _Selecting
from agents_b200.trajectories import policy_step
from agents_b200.utils import nest


class RandomTFPolicy(_Selecting):
  """Uniform random discrete actions (random_tf_policy.py:137)."""

  def __init__(self, time_step_spec, action_spec, seed=0,
               observation_and_action_constraint_splitter=None, name=None):
    super().__init__(time_step_spec, action_spec, seed=seed, name=name)
    spec = nest.flatten(action_spec)[0]
    self._num_actions = int(spec.maximum - spec.minimum + 1)
    self._action_dtype = spec.dtype
    self._splitter = observation_and_action_constraint_splitter
    self._zeros = None

  def _action(self, time_step, policy_state, seed):
    b = time_step.step_type.shape[0]
    dev = time_step.step_type.device
    mask = None
    if self._splitter is not None:
      _, mask = self._splitter(time_step.observation)
    if self._zeros is None or self._zeros.shape[0] != b or self._zeros.device != dev:
      self._zeros = torch.zeros((b, self._num_actions), dtype=torch.float32, device=dev)
    act = self._select(self._zeros, mask, 2.0, self._action_dtype)
    return policy_step.PolicyStep(act, policy_state, ())
