"""TFAgent base (tf_agents/agents/tf_agent.py:41-561): public train/loss/initialize wrappers,
LossInfo (:37), spec properties and experience validation (agents/data_converter.py:175-231).
"""
import collections

import torch

from agents_b200.trajectories import trajectory
from agents_b200.utils import common
from agents_b200.utils import nest

LossInfo = collections.namedtuple('LossInfo', ('loss', 'extra'))


def validate_trajectory(value, trajectory_spec, sequence_length, num_outer_dims=2):
  """data_converter._validate_trajectory (agents/data_converter.py:175-231)."""
  flat_v = nest.flatten(value)
  flat_s = nest.flatten(trajectory_spec)
  ok = len(flat_v) == len(flat_s)
  if ok:
    outer = None
    for v, s in zip(flat_v, flat_s):
      if v.dim() != num_outer_dims + len(s.shape) or tuple(v.shape[num_outer_dims:]) != tuple(s.shape):
        ok = False
        break
      o = tuple(v.shape[:num_outer_dims])
      if outer is None:
        outer = o
      elif o != outer:
        ok = False
        break
  if not ok:
    shape_str = 'two outer dimensions' if num_outer_dims == 2 else 'one outer dimension'
    prefix = '[B, T]' if num_outer_dims == 2 else '[B]'
    raise ValueError(
        'All of the Tensors in `value` must have {shape_str}. Specifically, '
        'tensors must have shape `{prefix} + spec.shape`.\n'
        'Full shapes of value tensors:\n  {v}.\n'
        'Expected shapes (excluding the {shape_str}):\n  {s}.'.format(
            shape_str=shape_str, prefix=prefix,
            v=[tuple(x.shape) for x in flat_v], s=[tuple(x.shape) for x in flat_s]))
  if sequence_length is not None and num_outer_dims == 2:
    for v in flat_v:
      if v.shape[1] != sequence_length:
        raise ValueError(
            'The agent was configured to expect a `sequence_length` '
            "of '{seq_len}'. Value is expected to be shaped `[B, T] + "
            'spec.shape` but at least one of the Tensors in `value` has a '
            "time axis dim value '{t_dim}' vs the expected '{seq_len}'.".format(
                seq_len=sequence_length, t_dim=v.shape[1]))


class TFAgent(object):
  """Abstract base class for agents running on libb200rl."""

  def __init__(self, time_step_spec, action_spec, policy, collect_policy, train_sequence_length,
               num_outer_dims=2, training_data_spec=None, debug_summaries=False,
               summarize_grads_and_vars=False, train_step_counter=None, device='cuda'):
    self._time_step_spec = time_step_spec
    self._action_spec = action_spec
    self._policy = policy
    self._collect_policy = collect_policy
    self._train_sequence_length = train_sequence_length
    self._num_outer_dims = num_outer_dims
    self._debug_summaries = debug_summaries
    self._summarize_grads_and_vars = summarize_grads_and_vars
    self._device = torch.device(device)
    if train_step_counter is None:
      train_step_counter = torch.zeros((), dtype=torch.int64, device=self._device)
    self._train_step_counter = train_step_counter
    self._train_step_host = int(train_step_counter.item()) if train_step_counter.numel() else 0
    self._training_data_spec = training_data_spec
    self._initialized = False

  # ---- specs / properties (tf_agent.py:463-561) ---------------------------------------------
  @property
  def time_step_spec(self):
    return self._time_step_spec

  @property
  def action_spec(self):
    return self._action_spec

  @property
  def policy(self):
    return self._policy

  @property
  def collect_policy(self):
    return self._collect_policy

  @property
  def collect_data_spec(self):
    return self._collect_policy.trajectory_spec

  @property
  def training_data_spec(self):
    return self._training_data_spec or self.collect_data_spec

  @property
  def train_sequence_length(self):
    return self._train_sequence_length

  @property
  def train_step_counter(self):
    return self._train_step_counter

  @property
  def debug_summaries(self):
    return self._debug_summaries

  def initialize(self):
    """Initialises the agent (copies weights into target networks) (tf_agent.py:263)."""
    self._initialize()
    self._initialized = True

  def preprocess_sequence(self, experience):
    return self._preprocess_sequence(experience)

  def _preprocess_sequence(self, experience):
    return experience

  def train(self, experience, weights=None, **kwargs):
    """Trains the agent on `[B, T, ...]` experience; returns LossInfo (tf_agent.py:309-358)."""
    loss_info = self._train(experience=experience, weights=weights, **kwargs)
    if not isinstance(loss_info, LossInfo):
      raise TypeError('loss_info is not a subclass of LossInfo: {}'.format(loss_info))
    return loss_info

  def loss(self, experience, weights=None, training=False, **kwargs):
    """Loss without a train step (tf_agent.py:360-415)."""
    loss_info = self._loss(experience, weights=weights, training=training, **kwargs)
    if not isinstance(loss_info, LossInfo):
      raise TypeError('loss_info is not a subclass of LossInfo: {}'.format(loss_info))
    return loss_info

  def _bump_train_step(self, n=1):
    from agents_b200 import _lib
    _lib.call('b200rl_counter_add', _lib.ptr(self._train_step_counter), n, _lib.stream())
    self._train_step_host += n
    common.record_host_effect(lambda: self._note_train_steps(n))

  def _note_train_steps(self, n):
    self._train_step_host += n

  def _initialize(self):
    raise NotImplementedError

  def _train(self, experience, weights):
    raise NotImplementedError

  def _loss(self, experience, weights, training=False):
    raise NotImplementedError
