"""TimeStep record and constructors on torch tensors.

Mirror of tf_agents/trajectories/time_step.py: TimeStep :54, StepType :113-121,
restart :135, transition :209, termination :285, truncation :349, time_step_spec :415.
"""
import collections

import numpy as np
import torch

from agents_b200.specs import tensor_spec
from agents_b200.utils import nest


class StepType(object):
  """FIRST/MID/LAST = int32 0/1/2 (time_step.py:113-121)."""
  FIRST = 0
  MID = 1
  LAST = 2


class TimeStep(collections.namedtuple('TimeStep',
                                      ['step_type', 'reward', 'discount', 'observation'])):
  __slots__ = ()

  def is_first(self):
    return self.step_type == StepType.FIRST

  def is_mid(self):
    return self.step_type == StepType.MID

  def is_last(self):
    return self.step_type == StepType.LAST


def _first_leaf(x):
  return nest.flatten(x)[0]


def _outer_shape_and_device(reward, outer_dims):
  first = _first_leaf(reward)
  if outer_dims is not None:
    shape = list(outer_dims)
  elif first.dim() == 0:
    shape = []
  else:
    shape = [first.shape[0]]
  return shape, first.device


def restart(observation, batch_size=None, reward_spec=None):
  """TimeStep with step_type FIRST, reward 0, discount 1 (time_step.py:135-195)."""
  dev = _first_leaf(observation).device
  shape = [] if batch_size is None else [int(batch_size)]
  step_type = torch.full(shape, StepType.FIRST, dtype=torch.int32, device=dev)
  if reward_spec is None:
    reward = torch.zeros(shape, dtype=torch.float32, device=dev)
  else:
    reward = nest.map_structure(
        lambda r: torch.zeros(shape + list(r.shape), dtype=r.dtype, device=dev), reward_spec)
  discount = torch.ones(shape, dtype=torch.float32, device=dev)
  return TimeStep(step_type, reward, discount, observation)


def _as_discount(discount, shape, dev):
  d = torch.as_tensor(discount, dtype=torch.float32, device=dev)
  if d.dim() == 0:
    d = d.expand(shape).contiguous() if shape else d
  return d


def transition(observation, reward, discount=1.0, outer_dims=None):
  """TimeStep with step_type MID (time_step.py:209-282)."""
  shape, dev = _outer_shape_and_device(reward, outer_dims)
  step_type = torch.full(shape, StepType.MID, dtype=torch.int32, device=dev)
  return TimeStep(step_type, reward, _as_discount(discount, shape, dev), observation)


def termination(observation, reward, outer_dims=None):
  """TimeStep with step_type LAST and discount 0 (time_step.py:285-346)."""
  shape, dev = _outer_shape_and_device(reward, outer_dims)
  step_type = torch.full(shape, StepType.LAST, dtype=torch.int32, device=dev)
  return TimeStep(step_type, reward, torch.zeros(shape, dtype=torch.float32, device=dev),
                  observation)


def truncation(observation, reward, discount=1.0, outer_dims=None):
  """TimeStep with step_type LAST keeping the discount (time_step.py:349-412)."""
  shape, dev = _outer_shape_and_device(reward, outer_dims)
  step_type = torch.full(shape, StepType.LAST, dtype=torch.int32, device=dev)
  return TimeStep(step_type, reward, _as_discount(discount, shape, dev), observation)


def time_step_spec(observation_spec=None, reward_spec=None):
  """TimeStep of specs (time_step.py:415-466)."""
  if observation_spec is None:
    return TimeStep(step_type=(), reward=(), discount=(), observation=())
  return TimeStep(
      step_type=tensor_spec.TensorSpec([], torch.int32, name='step_type'),
      reward=reward_spec or tensor_spec.TensorSpec([], torch.float32, name='reward'),
      discount=tensor_spec.BoundedTensorSpec([], torch.float32, minimum=0.0, maximum=1.0,
                                             name='discount'),
      observation=observation_spec)
