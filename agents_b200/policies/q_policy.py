"""QPolicy / GreedyPolicy / EpsilonGreedyPolicy / RandomTFPolicy over libb200rl.

Reference: policies/q_policy.py:150-194 (Q-net logits, masked with dtype.min),
policies/greedy_policy.py:70-89 (mode of Categorical == first argmax),
policies/epsilon_greedy_policy.py:120-145 (where(u >= eps, greedy, random)),
policies/random_tf_policy.py:137.  Action selection is ONE launch (b200rl_epsilon_greedy).
"""
import torch

from agents_b200 import _lib
from agents_b200.policies import tf_policy
from agents_b200.trajectories import policy_step
from agents_b200.utils import nest

_POLICY_SEED_TAG = 0x9E3779B97F4A7C15


class QPolicy(tf_policy.TFPolicy):
  """Holds the Q network; `q_values(time_step)` evaluates it (q_policy.py:150-194)."""

  def __init__(self, time_step_spec, action_spec, q_network,
               observation_and_action_constraint_splitter=None, emit_log_probability=False,
               name=None):
    super().__init__(time_step_spec, action_spec, name=name)
    flat = nest.flatten(action_spec)
    if len(flat) > 1:
      raise ValueError('Only scalar actions are supported now.')
    spec = flat[0]
    self._num_actions = int(spec.maximum - spec.minimum + 1)
    self._action_dtype = spec.dtype
    self._q_network = q_network
    self._splitter = observation_and_action_constraint_splitter

  @property
  def observation_and_action_constraint_splitter(self):
    return self._splitter

  @property
  def num_actions(self):
    return self._num_actions

  def variables(self):
    return self._q_network.variables

  def q_values(self, time_step):
    obs, mask = time_step.observation, None
    if self._splitter is not None:
      obs, mask = self._splitter(obs)
    q, _ = self._q_network(obs, step_type=time_step.step_type)
    return q, mask

  def _action(self, time_step, policy_state, seed):
    return GreedyPolicy(self)._action(time_step, policy_state, seed)


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


class EpsilonGreedyPolicy(_Selecting):
  """where(u >= epsilon, greedy, uniform random) (epsilon_greedy_policy.py:120-145)."""

  def __init__(self, policy, epsilon, seed=0, name=None):
    super().__init__(policy.time_step_spec, policy.action_spec, seed=seed, name=name)
    self._wrapped_policy = policy
    self._epsilon = epsilon

  @property
  def wrapped_policy(self):
    return self._wrapped_policy

  def variables(self):
    return self._wrapped_policy.variables()

  def _get_epsilon(self):
    return self._epsilon() if callable(self._epsilon) else self._epsilon

  def _action(self, time_step, policy_state, seed):
    q, mask = self._wrapped_policy.q_values(time_step)
    act = self._select(q, mask, self._get_epsilon(), self._wrapped_policy._action_dtype)
    return policy_step.PolicyStep(act, policy_state, ())


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
