"""ParallelPyEnvironment (tf_agents/environments/parallel_py_environment.py:47-420): each host
environment lives in its own process; `step` sends all actions first and then collects the
results, so simulators (MuJoCo, ALE) run concurrently while the GPU trains.

Reference behaviour kept: constructors (callables) build the envs inside the workers (:78-87);
all envs must share action / time-step specs (:97-101); `blocking=True` steps the envs one after
another (:166-170); exceptions raised inside a worker are re-raised in the parent with the
worker's traceback (:361-377); `close()` joins the workers.  Workers are forked
(`multiprocessing.get_context('fork')`), so constructors may be closures; they only touch numpy
and must not use CUDA.
"""
import multiprocessing
import sys
import traceback

import numpy as np

from agents_b200.environments import batched_py_environment
from agents_b200.environments import py_environment

_READY, _RESULT, _EXCEPTION, _CALL, _CLOSE = 1, 2, 3, 4, 5


def _worker(ctor, conn):
  try:
    env = ctor()
    specs = (env.action_spec(), env.observation_spec(), env.time_step_spec())
    conn.send((_READY, specs))
    while True:
      try:
        if not conn.poll(0.1):
          continue
        message, payload = conn.recv()
      except (EOFError, KeyboardInterrupt):
        break
      if message == _CALL:
        name, args = payload
        conn.send((_RESULT, getattr(env, name)(*args)))
      elif message == _CLOSE:
        env.close()
        break
  except Exception:  # pylint: disable=broad-except
    conn.send((_EXCEPTION, ''.join(traceback.format_exception(*sys.exc_info()))))
  finally:
    conn.close()


class ProcessPyEnvironment(object):
  """One environment in a forked worker process, driven through a pipe (:231-420)."""

  def __init__(self, env_constructor):
    self._ctor = env_constructor
    self._conn = None
    self._process = None
    self._specs = None

  def start(self, wait_to_start=True):
    ctx = multiprocessing.get_context('fork')
    self._conn, child = ctx.Pipe()
    self._process = ctx.Process(target=_worker, args=(self._ctor, child), daemon=True)
    self._process.start()
    child.close()
    if wait_to_start:
      self.wait_start()

  def wait_start(self):
    self._specs = self._receive()

  def _receive(self):
    message, payload = self._conn.recv()
    if message == _EXCEPTION:
      raise RuntimeError('Exception in environment process:\n' + payload)
    return payload

  def action_spec(self):
    return self._specs[0]

  def observation_spec(self):
    return self._specs[1]

  def time_step_spec(self):
    return self._specs[2]

  def call(self, name, *args):
    """Sends the call; returns a promise (callable) for its result."""
    self._conn.send((_CALL, (name, args)))
    return self._receive

  def close(self):
    if self._process is None:
      return
    try:
      self._conn.send((_CLOSE, None))
      self._conn.close()
    except (IOError, OSError):
      pass
    self._process.join(5)
    if self._process.is_alive():
      self._process.terminate()
    self._process = None


class ParallelPyEnvironment(py_environment.PyEnvironment):
  """Batch together environments and simulate them in external processes."""

  def __init__(self, env_constructors, start_serially=True, blocking=False, flatten=False):
    super(ParallelPyEnvironment, self).__init__()
    if any(not callable(ctor) for ctor in env_constructors):
      raise TypeError('Found non-callable `env_constructors` in `ParallelPyEnvironment` __init__ '
                      'call. Did you accidentally pass in environment instances instead of '
                      'constructors? Got: {}'.format(env_constructors))
    self._envs = [ProcessPyEnvironment(ctor) for ctor in env_constructors]
    self._num_envs = len(env_constructors)
    self._blocking = blocking
    self._start_serially = start_serially
    self.start()
    self._action_spec = self._envs[0].action_spec()
    self._observation_spec = self._envs[0].observation_spec()
    self._time_step_spec = self._envs[0].time_step_spec()
    self._parallel_execution = True
    if any(env.action_spec() != self._action_spec for env in self._envs):
      raise ValueError('All environments must have the same action spec.')
    if any(env.time_step_spec() != self._time_step_spec for env in self._envs):
      raise ValueError('All environments must have the same time_step_spec.')

  def start(self):
    for env in self._envs:
      env.start(wait_to_start=self._start_serially)
    if not self._start_serially:
      for env in self._envs:
        env.wait_start()

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

  def _gather(self, name, per_env_args):
    if self._blocking:
      return [env.call(name, *args)() for env, args in zip(self._envs, per_env_args)]
    promises = [env.call(name, *args) for env, args in zip(self._envs, per_env_args)]
    return [promise() for promise in promises]

  def _reset(self):
    return batched_py_environment._stack(self._gather('reset', [()] * self._num_envs))

  def _step(self, actions):
    unstacked = batched_py_environment._unstack(actions, self._num_envs)
    return batched_py_environment._stack(self._gather('step', [(a,) for a in unstacked]))

  def close(self):
    for env in self._envs:
      env.close()
