"""EpsilonGreedyPolicy: where(u >= epsilon, greedy action, uniform random allowed action).

Reference: policies/epsilon_greedy_policy.py:120-145 (`u ~ U[0, 1)` per batch row, the random
branch is `RandomTFPolicy` over the same action mask).  Device work: one b200rl_epsilon_greedy
launch (greedy_policy._Selecting).  `epsilon` may be a float or a callable evaluated on the host at
every call (:93-99).
"""
import warnings

import torch

from agents_b200.policies.greedy_policy import _Selecting
from agents_b200.trajectories import policy_step


class EpsilonGreedyPolicy(_Selecting):
  """where(u >= epsilon, greedy, uniform random) (epsilon_greedy_policy.py:120-145)."""

  def __init__(self, policy, epsilon, seed=0, name=None):
    super().__init__(policy.time_step_spec, policy.action_spec, seed=seed, name=name)
    self._wrapped_policy = policy
    self._epsilon = epsilon
    self._warned_capture = False

  @property
  def wrapped_policy(self):
    return self._wrapped_policy

  def variables(self):
    return self._wrapped_policy.variables()

  def _get_epsilon(self):
    if not callable(self._epsilon):
      return self._epsilon
    if (not self._warned_capture and torch.cuda.is_available() and
        torch.cuda.is_current_stream_capturing()):
      # the launch takes epsilon BY VALUE: inside common.function (a CUDA graph) the value seen
      # at capture time is replayed, unlike a tf.Variable read inside a tf.function
      warnings.warn('EpsilonGreedyPolicy: a callable epsilon is evaluated once when the collect '
                    'step is captured by common.function; call the policy outside the captured '
                    'function (or re-capture) to follow an epsilon schedule.')
      self._warned_capture = True
    return self._epsilon()

  def _action(self, time_step, policy_state, seed):
    q, mask = self._wrapped_policy.q_values(time_step)
    act = self._select(q, mask, self._get_epsilon(), self._wrapped_policy._action_dtype)
    return policy_step.PolicyStep(act, policy_state, ())
