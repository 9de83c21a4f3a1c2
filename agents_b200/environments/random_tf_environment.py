"""HBM-resident synthetic environments (kernel family i).

RandomTFEnvironment restates tf_agents/environments/random_tf_environment.py:30-129 with two
documented differences: episodes end per environment (the reference ends the whole batch
together, :117-127) and observations are U{0..255} (uint8 specs) or U[-1,1) (float specs)
instead of the reference's spec sampler (float specs there draw from +-dtype.max/8).
CartPoleTFEnvironment is a vectorised CartPole-v1 (gym classic_control dynamics; the reference
reaches it through suite_gym + TFPyEnvironment, agents/dqn/examples/v2/train_eval.py:151).
Each `step` is ONE launch of libb200rl (csrc/env.cu); outputs are written in place into the
environment's own tensors, so a TimeStep returned earlier is only valid until the next step.
"""
import numpy as np
import torch

from agents_b200 import _lib
from agents_b200.environments import tf_environment
from agents_b200.specs import tensor_spec
from agents_b200.trajectories import time_step as ts
from agents_b200.utils import nest

_ENV_SEED_TAG = 0xA5A5A5A55A5A5A5A


class RandomTFEnvironment(tf_environment.TFEnvironment):
  """Randomly generated observations and rewards, per-env episode ends with probability
  `episode_end_probability` (reference default 0.1)."""

  def __init__(self, time_step_spec, action_spec, batch_size=1, episode_end_probability=0.1,
               seed=0, device='cuda', stable_outputs=False):
    super().__init__(time_step_spec, action_spec, batch_size)
    obs_spec = nest.flatten(time_step_spec.observation)
    if len(obs_spec) != 1:
      raise ValueError('RandomTFEnvironment supports a single observation tensor.')
    self._obs_spec = obs_spec[0]
    if self._obs_spec.dtype not in (torch.uint8, torch.float32):
      raise ValueError('observation dtype must be uint8 or float32.')
    self._p = float(episode_end_probability)
    self._seed = (int(seed) ^ _ENV_SEED_TAG) & 0xFFFFFFFFFFFFFFFF
    self._device = torch.device(device)
    self._stable = stable_outputs
    b = batch_size
    self._step_type = torch.full((b,), ts.StepType.LAST, dtype=torch.int32, device=self._device)
    # Output buffers ping-pong: the TimeStep returned by step t stays intact while step t+1 is
    # produced, which is what trajectory.from_transition(time_step, ..., next_time_step) needs.
    self._out = [dict(step_type=torch.zeros(b, dtype=torch.int32, device=self._device),
                      obs=torch.zeros((b,) + self._obs_spec.shape, dtype=self._obs_spec.dtype,
                                      device=self._device),
                      reward=torch.zeros(b, dtype=torch.float32, device=self._device),
                      discount=torch.ones(b, dtype=torch.float32, device=self._device))
                 for _ in range(2)]
    self._cur = 0
    self._rng = torch.zeros(2, dtype=torch.int64, device=self._device)
    self._started = False

  def _launch(self):
    self._cur ^= 1
    o = self._out[self._cur]
    elems = int(np.prod(self._obs_spec.shape)) if self._obs_spec.shape else 1
    _lib.call('b200rl_env_random_step', _lib.ptr(self._step_type), _lib.ptr(o['step_type']),
              _lib.ptr(o['obs']), elems, int(self._obs_spec.dtype == torch.uint8),
              _lib.ptr(o['reward']), _lib.ptr(o['discount']), self._batch_size, self._p,
              self._seed, _lib.ptr(self._rng), _lib.stream())

  def _time_step(self):
    o = self._out[self._cur]
    if self._stable:
      return ts.TimeStep(o['step_type'].clone(), o['reward'].clone(), o['discount'].clone(),
                         o['obs'].clone())
    return ts.TimeStep(o['step_type'], o['reward'], o['discount'], o['obs'])

  def _current_time_step(self):
    if not self._started:
      return self._reset()
    return self._time_step()

  def _reset(self):
    self._step_type.fill_(ts.StepType.LAST)   # every env takes the auto-reset branch
    self._launch()
    self._started = True
    return self._time_step()

  def _step(self, action):
    if not self._started:
      return self._reset()
    self._launch()
    return self._time_step()


class CartPoleTFEnvironment(tf_environment.TFEnvironment):
  """Vectorised CartPole-v1: obs f32[4], actions {0,1}, reward 1 per step, 500-step limit."""

  def __init__(self, batch_size=1, max_episode_steps=500, seed=0, device='cuda',
               action_dtype=torch.int64):
    obs_spec = tensor_spec.BoundedTensorSpec((4,), torch.float32, -3.4e38, 3.4e38, 'observation')
    act_spec = tensor_spec.BoundedTensorSpec((), action_dtype, 0, 1, 'action')
    super().__init__(ts.time_step_spec(obs_spec), act_spec, batch_size)
    self._device = torch.device(device)
    self._seed = (int(seed) ^ _ENV_SEED_TAG) & 0xFFFFFFFFFFFFFFFF
    self._max_steps = int(max_episode_steps)
    b = batch_size
    self._state = torch.zeros(b, 4, dtype=torch.float32, device=self._device)
    self._steps = torch.zeros(b, dtype=torch.int32, device=self._device)
    self._step_type = torch.full((b,), ts.StepType.LAST, dtype=torch.int32, device=self._device)
    self._obs = torch.zeros(b, 4, dtype=torch.float32, device=self._device)
    self._reward = torch.zeros(b, dtype=torch.float32, device=self._device)
    self._discount = torch.ones(b, dtype=torch.float32, device=self._device)
    self._rng = torch.zeros(2, dtype=torch.int64, device=self._device)
    self._zero_action = torch.zeros(b, dtype=torch.int32, device=self._device)
    self._started = False

  def _launch(self, action):
    _lib.call('b200rl_env_cartpole_step', _lib.ptr(self._state), _lib.ptr(self._steps),
              _lib.ptr(self._step_type), _lib.ptr(action), _lib.ptr(self._obs),
              _lib.ptr(self._reward), _lib.ptr(self._discount), self._batch_size,
              self._max_steps, self._seed, _lib.ptr(self._rng), _lib.stream())

  def _time_step(self):
    return ts.TimeStep(self._step_type.clone(), self._reward.clone(), self._discount.clone(),
                       self._obs.clone())

  def _current_time_step(self):
    if not self._started:
      return self._reset()
    return self._time_step()

  def _reset(self):
    self._step_type.fill_(ts.StepType.LAST)
    self._launch(self._zero_action)
    self._started = True
    return self._time_step()

  def _step(self, action):
    if not self._started:
      return self._reset()
    self._launch(action.to(torch.int32).contiguous())
    return self._time_step()
