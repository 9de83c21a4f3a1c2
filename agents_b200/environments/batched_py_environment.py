"""BatchedPyEnvironment (tf_agents/environments/batched_py_environment.py:38-200): N non-batched
host environments behind one batched interface; optional thread pool, outputs stacked along a
new leading axis (single env: expanded, :150-159)."""
from multiprocessing import pool as mp_pool

import numpy as np

from agents_b200.environments import py_environment
from agents_b200.trajectories import time_step as ts
from agents_b200.utils import nest


def _stack(time_steps):
  flats = [nest.flatten(t) for t in time_steps]
  stacked = [np.stack([np.asarray(f[i]) for f in flats]) for i in range(len(flats[0]))]
  return nest.pack_sequence_as(time_steps[0], stacked)


def _unstack(actions, n):
  flat = nest.flatten(actions)
  return [nest.pack_sequence_as(actions, [np.asarray(a)[i] for a in flat]) for i in range(n)]


class BatchedPyEnvironment(py_environment.PyEnvironment):
  """Batch together multiple py environments and act as a single batch."""

  def __init__(self, envs, multithreading=True):
    if not isinstance(envs, (list, tuple)):
      raise ValueError('envs must be a list or tuple.  Got: %s' % envs)
    if not envs:
      raise ValueError('envs must be non-empty.')
    batched_envs = [(i, env) for i, env in enumerate(envs) if env.batched]
    if batched_envs:
      raise ValueError('Some of the envs are already batched: %s' % batched_envs)
    self._parallel_execution = multithreading
    self._envs = list(envs)
    self._num_envs = len(envs)
    self._action_spec = self._envs[0].action_spec()
    self._observation_spec = self._envs[0].observation_spec()
    self._time_step_spec = self._envs[0].time_step_spec()
    if any(env.action_spec() != self._action_spec for env in self._envs):
      raise ValueError('All environments must have the same action spec.  Saw: %s' %
                       [env.action_spec() for env in self._envs])
    if any(env.time_step_spec() != self._time_step_spec for env in self._envs):
      raise ValueError('All environments must have the same time_step_spec.  Saw: %s' %
                       [env.time_step_spec() for env in self._envs])
    self._pool = mp_pool.ThreadPool(self._num_envs) if multithreading else None
    super(BatchedPyEnvironment, self).__init__()

  def _execute(self, fn, iterable):
    if self._parallel_execution:
      return self._pool.map(fn, iterable)
    return [fn(x) for x in iterable]

  @property
  def batched(self):
    return True

  @property
  def batch_size(self):
    return self._num_envs

  @property
  def envs(self):
    return self._envs

  def observation_spec(self):
    return self._observation_spec

  def action_spec(self):
    return self._action_spec

  def time_step_spec(self):
    return self._time_step_spec

  def _reset(self):
    return _stack(self._execute(lambda env: env.reset(), self._envs))

  def _step(self, actions):
    unstacked = _unstack(actions, self._num_envs)
    if len(unstacked) != self._num_envs:
      raise ValueError('Primary dimension of action items does not match batch size: %d vs. %d' %
                       (len(unstacked), self._num_envs))
    return _stack(self._execute(lambda ea: ea[0].step(ea[1]), list(zip(self._envs, unstacked))))

  def close(self):
    self._execute(lambda env: env.close(), self._envs)
    if self._pool is not None:
      self._pool.close()
      self._pool.join()
      self._pool = None
