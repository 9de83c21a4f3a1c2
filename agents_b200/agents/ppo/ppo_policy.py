"""PPOPolicy (tf_agents/agents/ppo/ppo_policy.py:40-): ActorPolicy that also records the
distribution parameters (and optionally the value prediction) in `policy_info`
(ppo_policy.py:156-159), so that PPOAgent.train can recompute the behaviour log-probs.

collect=True  -> actions are sampled from Normal(loc, scale) and clipped to the spec
                 (ActorPolicy clip=True); info = {'dist_params': {'loc', 'scale'}[, 'value_prediction']}.
collect=False -> the greedy policy (mode of the distribution = loc), empty info.
"""
import torch

from agents_b200 import _lib
from agents_b200.policies import tf_policy
from agents_b200.specs import tensor_spec
from agents_b200.trajectories import policy_step
from agents_b200.utils import nest

_PPO_SEED_TAG = 0x50504F5F504F4C49


class PPOPolicy(tf_policy.TFPolicy):

  def __init__(self, time_step_spec, action_spec, actor_net, value_net,
               observation_normalizer=None, clip=True, collect=True,
               compute_value_and_advantage_in_train=True, seed=0):
    spec = nest.flatten(action_spec)[0]
    a = spec.shape
    info_spec = ()
    if collect:
      info_spec = {'dist_params': {'loc': tensor_spec.TensorSpec(a, torch.float32, 'loc'),
                                   'scale': tensor_spec.TensorSpec(a, torch.float32, 'scale')}}
      if not compute_value_and_advantage_in_train:
        info_spec['value_prediction'] = tensor_spec.TensorSpec((), torch.float32,
                                                               'value_prediction')
    super().__init__(time_step_spec, action_spec, info_spec=info_spec)
    self._actor_net, self._value_net = actor_net, value_net
    self._observation_normalizer = observation_normalizer
    self._clip, self._collect = clip, collect
    self._value_in_info = collect and not compute_value_and_advantage_in_train
    self._seed = (int(seed) ^ _PPO_SEED_TAG) & 0xFFFFFFFFFFFFFFFF
    self._rng = None

  def variables(self):
    return self._actor_net.variables + self._value_net.variables

  def _normalized(self, observation):
    if self._observation_normalizer is None:
      return observation
    return self._observation_normalizer.normalize(observation)

  def apply_value_network(self, observations, step_types=None, value_state=None, training=False):
    """Value predictions for `[N, ...]` observations (ppo_policy.py apply_value_network)."""
    v, _ = self._value_net(self._normalized(observations))
    return v, ()

  def distribution_params(self, time_step):
    return self._actor_net.distribution_params(self._normalized(time_step.observation))

  def _action(self, time_step, policy_state, seed):
    obs = self._normalized(time_step.observation)
    loc, scale = self._actor_net.distribution_params(obs)
    if not self._collect:
      return policy_step.PolicyStep(loc, policy_state, ())
    n, a = loc.shape
    if self._rng is None or self._rng.device != loc.device:
      self._rng = torch.zeros(2, dtype=torch.int64, device=loc.device)
    act = torch.empty_like(loc)
    _lib.call('b200rl_normal_sample', _lib.ptr(loc), _lib.ptr(scale), a, n, a,
              _lib.ptr(self._actor_net.action_min) if self._clip else None,
              _lib.ptr(self._actor_net.action_max) if self._clip else None, self._seed,
              _lib.ptr(self._rng), _lib.ptr(act), _lib.stream())
    info = {'dist_params': {'loc': loc, 'scale': scale}}
    if self._value_in_info:
      info['value_prediction'], _ = self._value_net(obs)
    return policy_step.PolicyStep(act, policy_state, info)
