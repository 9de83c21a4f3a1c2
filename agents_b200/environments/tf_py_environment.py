"""TFPyEnvironment (tf_agents/environments/tf_py_environment.py:68-365): exposes a host
(numpy) environment as a device-resident TFEnvironment.

Reference behaviour kept: non-batched envs are wrapped in a BatchedPyEnvironment of size 1
(:139-143); `current_time_step` resets lazily (:225-229); `step` checks that every action leaf
has `batch_size` as its major dimension when `check_dims` (:300-313); attribute access falls
through to the wrapped env (:181-195); `isolation=True` runs every env interaction in one
dedicated thread (:110-137).

What replaces `tf.numpy_function`: actions are read back with one small D2H copy, the env steps
on the host, and the batched TimeStep is written into PINNED staging buffers and uploaded with
asynchronous copies into one of two device buffer sets (the driver still holds the previous
TimeStep while this one is produced).  An event per staging set guards its reuse, so the host may
run ahead of the GPU by one step.
"""
from multiprocessing import pool as mp_pool

import numpy as np
import torch

from agents_b200.environments import batched_py_environment
from agents_b200.environments import py_environment
from agents_b200.environments import tf_environment
from agents_b200.utils import nest

_NP = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32,
       torch.int64: np.int64, torch.uint8: np.uint8, torch.int8: np.int8, torch.bool: np.bool_,
       torch.int16: np.int16, torch.float16: np.float16}


class TFPyEnvironment(tf_environment.TFEnvironment):
  """Exposes a Python environment as an in-graph (device tensor) environment."""

  def __init__(self, environment, check_dims=False, isolation=False, device='cuda'):
    if not isinstance(environment, py_environment.PyEnvironment):
      raise TypeError('Environment should implement py_environment.PyEnvironment')
    if not environment.batched:
      environment = batched_py_environment.BatchedPyEnvironment([environment],
                                                                multithreading=not isolation)
    self._env = environment
    self._check_dims = check_dims
    self._pool = None
    if isolation:
      self._pool = isolation if hasattr(isolation, 'apply') else mp_pool.ThreadPool(1)
    batch_size = self._env.batch_size if self._env.batch_size else 1
    super(TFPyEnvironment, self).__init__(self._env.time_step_spec(), self._env.action_spec(),
                                          batch_size)
    self._device = torch.device(device)
    self._flat_specs = nest.flatten(self.time_step_spec())
    self._time_step = None
    self._slot = 0
    self._pinned, self._dev, self._done = [], [], []
    for _ in range(2):
      self._pinned.append([torch.empty((batch_size,) + tuple(s.shape), dtype=s.dtype).pin_memory()
                           for s in self._flat_specs])
      self._dev.append([torch.empty((batch_size,) + tuple(s.shape), dtype=s.dtype,
                                    device=self._device) for s in self._flat_specs])
      self._done.append(torch.cuda.Event())
    self._used = [False, False]

  def __getattr__(self, name):
    if name.startswith('_'):
      raise AttributeError(name)
    return getattr(self._env, name)

  @property
  def pyenv(self):
    return self._env

  def close(self):
    self._env.close()
    if self._pool is not None:
      self._pool.close()
      self._pool.join()
      self._pool = None

  def _execute(self, fn, *args):
    if self._pool is None:
      return fn(*args)
    return self._pool.apply(fn, args=args)

  # ---- host -> device staging ---------------------------------------------------------------------
  def _upload(self, time_step_np):
    slot = self._slot
    self._slot ^= 1
    if self._used[slot]:
      self._done[slot].synchronize()          # the previous upload from this pinned set is done
    flat = nest.flatten(time_step_np)
    with torch.cuda.device(self._device):
      for pin, dev, leaf, spec in zip(self._pinned[slot], self._dev[slot], flat, self._flat_specs):
        np.copyto(pin.numpy(), np.asarray(leaf, dtype=_NP[spec.dtype]).reshape(pin.shape))
        dev.copy_(pin, non_blocking=True)
      self._done[slot].record()
    self._used[slot] = True
    return nest.pack_sequence_as(self.time_step_spec(), list(self._dev[slot]))

  def _current_time_step(self):
    if self._time_step is None:
      self._time_step_np = self._execute(self._env.reset)
      self._time_step = self._upload(self._time_step_np)
    return self._time_step

  def _reset(self):
    self._time_step_np = self._execute(self._env.reset)
    self._time_step = self._upload(self._time_step_np)
    return self._time_step

  def _step(self, actions):
    flat_actions = nest.flatten(actions)
    if self._check_dims:
      for action in flat_actions:
        if action.dim() == 0 or action.shape[0] != self.batch_size:
          raise ValueError(
              'Expected actions whose major dimension is batch_size (%d), but saw action with '
              'shape %s:\n   %s' % (self.batch_size, tuple(action.shape), action))
    host = [a.detach().cpu().numpy() for a in flat_actions]      # D2H (synchronises the stream)
    packed = nest.pack_sequence_as(self.action_spec(), host)
    self._time_step_np = self._execute(self._env.step, packed)
    self._time_step = self._upload(self._time_step_np)
    return self._time_step
