"""Trajectory / Transition records and conversions on torch tensors.

Mirror of tf_agents/trajectories/trajectory.py: Trajectory :36-125, Transition :128-196,
first/mid/last/boundary :262-465, from_transition :614-647, to_transition :650-713,
to_n_step_transition :716-850 (its reward/discount reduction runs in
b200rl_nstep_reduce, csrc/scans.cu).
"""
import collections

import torch

from agents_b200 import _lib
from agents_b200.trajectories import policy_step
from agents_b200.trajectories import time_step as ts
from agents_b200.utils import nest


class Trajectory(collections.namedtuple('Trajectory', [
    'step_type', 'observation', 'action', 'policy_info', 'next_step_type', 'reward',
    'discount'])):
  """Row t holds (step_type_t, obs_t, action_t, policy_info_t, next_step_type_{t+1},
  reward_{t+1}, discount_{t+1}) (trajectory.py:50-72)."""
  __slots__ = ()

  def is_first(self):
    return self.step_type == ts.StepType.FIRST

  def is_mid(self):
    return (self.step_type == ts.StepType.MID) & (self.next_step_type == ts.StepType.MID)

  def is_last(self):
    return self.next_step_type == ts.StepType.LAST

  def is_boundary(self):
    return self.step_type == ts.StepType.LAST

  def replace(self, **kwargs):
    return self._replace(**kwargs)


class Transition(collections.namedtuple('Transition',
                                        ['time_step', 'action_step', 'next_time_step'])):
  __slots__ = ()

  def replace(self, **kwargs):
    return self._replace(**kwargs)


def _create(observation, action, policy_info, reward, discount, step_type, next_step_type):
  """Shared body of first/mid/last/boundary (trajectory.py:198-259): fills step types with
  the shape of `discount`."""
  discount = torch.as_tensor(discount, dtype=torch.float32)
  st = torch.full(discount.shape, step_type, dtype=torch.int32, device=discount.device)
  nst = torch.full(discount.shape, next_step_type, dtype=torch.int32, device=discount.device)
  return Trajectory(st, observation, action, policy_info, nst, reward, discount)


def first(observation, action, policy_info, reward, discount):
  return _create(observation, action, policy_info, reward, discount, ts.StepType.FIRST,
                 ts.StepType.MID)


def mid(observation, action, policy_info, reward, discount):
  return _create(observation, action, policy_info, reward, discount, ts.StepType.MID,
                 ts.StepType.MID)


def last(observation, action, policy_info, reward, discount):
  return _create(observation, action, policy_info, reward, discount, ts.StepType.MID,
                 ts.StepType.LAST)


def single_step(observation, action, policy_info, reward, discount):
  return _create(observation, action, policy_info, reward, discount, ts.StepType.FIRST,
                 ts.StepType.LAST)


def boundary(observation, action, policy_info, reward, discount):
  return _create(observation, action, policy_info, reward, discount, ts.StepType.LAST,
                 ts.StepType.FIRST)


def from_transition(time_step, action_step, next_time_step):
  """Pure re-packaging, no copies (trajectory.py:614-647)."""
  return Trajectory(
      step_type=time_step.step_type,
      observation=time_step.observation,
      action=action_step.action,
      policy_info=action_step.info,
      next_step_type=next_time_step.step_type,
      reward=next_time_step.reward,
      discount=next_time_step.discount)


def _validate_rank(t, min_rank, max_rank=None):
  rank = t.dim()
  if rank < min_rank or (max_rank is not None and rank > max_rank):
    raise ValueError('Expected variable to have rank in [{}, {}], but saw rank {}. '
                     'Shape: {}'.format(min_rank, max_rank, rank, tuple(t.shape)))


def to_transition(trajectory, next_trajectory=None):
  """(time_steps, policy_steps, next_time_steps) (trajectory.py:650-713)."""
  _validate_rank(trajectory.discount, 1, 2)
  if next_trajectory is not None:
    _validate_rank(next_trajectory.discount, 1, 2)
  if next_trajectory is None:
    next_trajectory = nest.map_structure(lambda t: t[:, 1:], trajectory)
    trajectory = nest.map_structure(lambda t: t[:, :-1], trajectory)
  policy_steps = policy_step.PolicyStep(action=trajectory.action, state=(),
                                        info=trajectory.policy_info)
  time_steps = ts.TimeStep(
      trajectory.step_type,
      reward=nest.map_structure(torch.zeros_like, trajectory.reward),
      discount=torch.zeros_like(trajectory.discount),
      observation=trajectory.observation)
  next_time_steps = ts.TimeStep(
      step_type=trajectory.next_step_type,
      reward=trajectory.reward,
      discount=trajectory.discount,
      observation=next_trajectory.observation)
  return Transition(time_steps, policy_steps, next_time_steps)


def to_n_step_transition(trajectory, gamma):
  """N-step transition from `[B, N+1, ...]` frames (trajectory.py:716-850)."""
  _validate_rank(trajectory.discount, 2, 2)
  time_dim = trajectory.discount.shape[1]
  if time_dim in (0, 1):
    raise ValueError('Trajectory frame count must be at least 2, but saw {}.  Shape of '
                     'trajectory.discount: {}'.format(time_dim, tuple(trajectory.discount.shape)))
  first_frame = nest.map_structure(lambda t: t[:, 0], trajectory)
  final_frame = nest.map_structure(lambda t: t[:, -1], trajectory)
  b = trajectory.discount.shape[0]
  reward = trajectory.reward.contiguous()
  discount = trajectory.discount.contiguous()
  out_r = torch.empty(b, dtype=torch.float32, device=reward.device)
  out_d = torch.empty(b, dtype=torch.float32, device=reward.device)
  _lib.call('b200rl_nstep_reduce', _lib.ptr(reward), _lib.ptr(discount), float(gamma),
            _lib.ptr(out_r), _lib.ptr(out_d), b, time_dim, _lib.stream())
  policy_steps = policy_step.PolicyStep(action=first_frame.action, state=(),
                                        info=first_frame.policy_info)
  nan = float('nan')
  time_steps = ts.TimeStep(
      first_frame.step_type,
      reward=nest.map_structure(lambda r: torch.full_like(r, nan), first_frame.reward),
      discount=torch.full_like(first_frame.discount, nan),
      observation=first_frame.observation)
  next_time_steps = ts.TimeStep(step_type=final_frame.step_type, reward=out_r, discount=out_d,
                                observation=final_frame.observation)
  return Transition(time_steps, policy_steps, next_time_steps)
