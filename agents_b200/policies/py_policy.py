"""PyPolicy (tf_agents/policies/py_policy.py:33-210): host-side (numpy) policy contract used by
`PyDriver`; `PyTFEagerPolicy` runs a device policy of this package behind that contract
(policies/py_tf_eager_policy.py:44-197): numpy TimeStep -> device -> `policy.action` -> numpy."""
import abc

import numpy as np
import torch

from agents_b200.trajectories import policy_step
from agents_b200.utils import nest


class PyPolicy(abc.ABC):

  def __init__(self, time_step_spec, action_spec, policy_state_spec=(), info_spec=()):
    self._time_step_spec = time_step_spec
    self._action_spec = action_spec
    self._policy_state_spec = policy_state_spec
    self._info_spec = info_spec

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

  def get_initial_state(self, batch_size=None):
    return self._get_initial_state(batch_size)

  def action(self, time_step, policy_state=(), seed=None):
    return self._action(time_step, policy_state)

  def _get_initial_state(self, batch_size=None):
    return ()

  @abc.abstractmethod
  def _action(self, time_step, policy_state):
    pass


class PyTFEagerPolicy(PyPolicy):
  """Runs a device-resident policy on numpy inputs (one H2D + one D2H per call)."""

  def __init__(self, policy, device='cuda', batch_time_steps=False):
    super().__init__(policy.time_step_spec, policy.action_spec,
                     getattr(policy, 'policy_state_spec', ()), getattr(policy, 'info_spec', ()))
    self._policy = policy
    self._device = torch.device(device)
    self._batch_time_steps = batch_time_steps

  def _to_device(self, x):
    def conv(a):
      t = torch.as_tensor(np.asarray(a)).to(self._device)
      return t.unsqueeze(0) if self._batch_time_steps else t
    return nest.map_structure(conv, x)

  def _to_host(self, x):
    def conv(t):
      a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
      return a[0] if self._batch_time_steps else a
    return nest.map_structure(conv, x)

  def _get_initial_state(self, batch_size=None):
    return self._to_host(self._policy.get_initial_state(batch_size))

  def _action(self, time_step, policy_state):
    step = self._policy.action(self._to_device(time_step), self._to_device(policy_state))
    return policy_step.PolicyStep(self._to_host(step.action), self._to_host(step.state),
                                  self._to_host(step.info))
