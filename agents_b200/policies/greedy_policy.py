"""GreedyPolicy and the action-selection launch shared by the discrete policies.

Reference: policies/greedy_policy.py:70-89 (the mode of a Categorical over the Q values == the
first argmax; masked actions get dtype.min logits, q_policy.py:174-183).  `_Selecting._select` is
the ONE launch (b200rl_epsilon_greedy, csrc/env.cu) behind GreedyPolicy (epsilon = -1: never
random), EpsilonGreedyPolicy (epsilon_greedy_policy.py) and RandomTFPolicy (epsilon = 2: always
random, random_tf_policy.py): Philox draw, masked argmax and masked uniform choice per batch row.
"""
import torch

from agents_b200 import _lib
from agents_b200.policies import tf_policy
from agents_b200.trajectories import policy_step

_POLICY_SEED_TAG = 0x9E3779B97F4A7C15


class _Selecting(tf_policy.TFPolicy):
  """Shared launch of b200rl_epsilon_greedy for the three selection policies."""

  def __init__(self, time_step_spec, action_spec, seed=0, name=None):
    super().__init__(time_step_spec, action_spec, name=name)
    self._seed = (int(seed) ^ _POLICY_SEED_TAG) & 0xFFFFFFFFFFFFFFFF
    self._rng = None

  def _select(self, q, mask, eps, dtype):
    b, a = q.shape
    if self._rng is None or self._rng.device != q.device:
      self._rng = torch.zeros(2, dtype=torch.int64, device=q.device)
    out = torch.empty(b, dtype=torch.int32, device=q.device)
    if mask is not None:
      mask = mask.to(torch.int32).contiguous()
    _lib.call('b200rl_epsilon_greedy', _lib.ptr(q.contiguous()), _lib.ptr(mask), b, a,
              float(eps), self._seed, _lib.ptr(self._rng), None, None, _lib.ptr(out),
              _lib.stream())
    return out if dtype == torch.int32 else out.to(dtype)


class GreedyPolicy(_Selecting):
  """argmax_a Q(s,a) (greedy_policy.py:70-89)."""

  def __init__(self, policy, name=None):
    super().__init__(policy.time_step_spec, policy.action_spec, name=name)
    self._wrapped_policy = policy

  @property
  def wrapped_policy(self):
    return self._wrapped_policy

  def variables(self):
    return self._wrapped_policy.variables()

  def _action(self, time_step, policy_state, seed):
    q, mask = self._wrapped_policy.q_values(time_step)
    act = self._select(q, mask, -1.0, self._wrapped_policy._action_dtype)
    return policy_step.PolicyStep(act, policy_state, ())
