"""PyTFEagerPolicy (tf_agents/policies/py_tf_eager_policy.py:44-197): a device-resident policy of
this package behind the host `PyPolicy` contract, for `PyDriver` loops over host environments:
numpy TimeStep -> one H2D copy -> `policy.action` on the GPU -> one D2H copy -> numpy PolicyStep.
`batch_time_steps` adds / strips the outer batch dimension for unbatched environments (:150-176).
"""
import numpy as np
import torch

from agents_b200.policies.py_policy import PyPolicy
from agents_b200.trajectories import policy_step
from agents_b200.utils import nest


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
