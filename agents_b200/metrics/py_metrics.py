"""Host-side (numpy) step metrics for collection / evaluation loops — the subset of
tf_agents/metrics/py_metrics.py that `train.Actor` wires into `PyDriver`
(`actor.collect_metrics` / `actor.eval_metrics`, train/actor.py:267-330).

Semantics kept: streaming metrics average the last `buffer_size` finished episodes (:94-151);
the episode return is zeroed on FIRST steps, accumulates `trajectory.reward` and is pushed on
`is_last()` (:178-193); episode length counts non-boundary steps (:220-232); EnvironmentSteps
counts non-boundary steps (:249-255); NumberOfEpisodes counts `is_last()` (:272-278).
They are observers: `metric(trajectory)` with numpy Trajectories, batched or not.
"""
import abc

import numpy as np

from agents_b200.utils import nest


class NumpyDeque(object):
  """The newest `maxlen` values of a stream, for windowed means.

  Storage is a flat numpy array used as a circular window: `_count` values are valid and the
  newest sits just before `_next`.  An unbounded deque (`maxlen=np.inf`) never wraps; its array
  doubles when full.  (Role of py_metrics.NumpyDeque in the reference; the metrics below only
  need add / extend / clear / len / mean / last.)
  """

  def __init__(self, maxlen, dtype):
    self._bounded = not np.isinf(maxlen)
    self._capacity = int(maxlen) if self._bounded else 16
    if self._capacity < 1:
      raise ValueError('maxlen must be >= 1.')
    self._store = np.zeros(self._capacity, dtype=dtype)
    self._count = 0          # valid entries
    self._next = 0           # where the next value goes

  def clear(self):
    self._count = 0
    self._next = 0

  def _grow(self):
    bigger = np.zeros(2 * self._capacity, dtype=self._store.dtype)
    bigger[:self._capacity] = self._store
    self._store, self._capacity = bigger, 2 * self._capacity

  def add(self, value):
    if not self._bounded and self._next == self._capacity:
      self._grow()
    self._store[self._next] = value
    self._next += 1
    if self._bounded and self._next == self._capacity:
      self._next = 0
    self._count = min(self._count + 1, self._capacity) if self._bounded else self._count + 1

  def extend(self, values):
    for v in np.asarray(values).reshape(-1):
      self.add(v)

  @property
  def last(self):
    if self._count == 0:
      return None
    return self._store[(self._next - 1) % self._capacity]

  def __len__(self):
    return self._count

  def mean(self, dtype=None):
    # a full bounded window uses the whole array; otherwise the valid part is the prefix
    window = self._store if self._count == self._capacity else self._store[:self._count]
    return np.mean(window, dtype=dtype)


def _batched(trajectory):
  if np.ndim(trajectory.step_type) == 0:
    return nest.map_structure(lambda a: np.asarray(a)[None], trajectory)
  return nest.map_structure(np.asarray, trajectory)


class PyMetric(abc.ABC):

  def __init__(self, name, prefix='Metrics'):
    self.name = name
    self._prefix = prefix

  def __call__(self, *args, **kwargs):
    return self.call(*args, **kwargs)

  @abc.abstractmethod
  def call(self, *args, **kwargs):
    pass

  @abc.abstractmethod
  def reset(self):
    pass

  @abc.abstractmethod
  def result(self):
    pass

  def log(self):
    return '{0} = {1}'.format(self.name, self.result())


class StreamingMetric(PyMetric):
  """A per-episode quantity averaged over the most recent `buffer_size` finished episodes.

  Subclasses keep one accumulator per environment (`_reset(batch_size)` allocates them,
  `_batched_call(trajectory)` advances them and calls `add_to_buffer` when episodes end); the
  accumulators are sized lazily from the first trajectory unless `batch_size` is given."""

  def __init__(self, name='StreamingMetric', buffer_size=10, batch_size=None):
    super(StreamingMetric, self).__init__(name)
    self._window = NumpyDeque(maxlen=buffer_size, dtype=np.float64)
    self._batch_size = batch_size
    self.reset()

  def reset(self):
    self._window.clear()
    if self._batch_size:
      self._reset(self._batch_size)

  @abc.abstractmethod
  def _reset(self, batch_size):
    pass

  def add_to_buffer(self, values):
    self._window.extend(values)

  @property
  def data(self):
    return self._window

  def result(self):
    if len(self._window) == 0:
      return np.array(0.0, dtype=np.float32)
    return self._window.mean(dtype=np.float32)

  @abc.abstractmethod
  def _batched_call(self, trajectory):
    pass

  def call(self, trajectory):
    if not self._batch_size:
      if np.ndim(trajectory.step_type) == 0:
        self._batch_size = 1
      else:
        assert np.ndim(trajectory.step_type) == 1
        self._batch_size = np.shape(trajectory.step_type)[0]
      self.reset()
    self._batched_call(_batched(trajectory))


class AverageReturnMetric(StreamingMetric):

  def __init__(self, name='AverageReturn', buffer_size=10, batch_size=None):
    self._episode_return = np.float64(0)
    super(AverageReturnMetric, self).__init__(name, buffer_size=buffer_size, batch_size=batch_size)

  def _reset(self, batch_size):
    self._episode_return = np.zeros(shape=(batch_size,), dtype=np.float64)

  def _batched_call(self, trajectory):
    episode_return = self._episode_return
    episode_return[np.where(trajectory.is_first())] = 0
    episode_return += trajectory.reward
    self.add_to_buffer(episode_return[np.where(trajectory.is_last())])


class AverageEpisodeLengthMetric(StreamingMetric):

  def __init__(self, name='AverageEpisodeLength', buffer_size=10, batch_size=None):
    self._episode_steps = np.float64(0)
    super(AverageEpisodeLengthMetric, self).__init__(name, buffer_size=buffer_size,
                                                     batch_size=batch_size)

  def _reset(self, batch_size):
    self._episode_steps = np.zeros(shape=(batch_size,), dtype=np.float64)

  def _batched_call(self, trajectory):
    episode_steps = self._episode_steps
    episode_steps[np.where(~trajectory.is_boundary())] += 1
    self.add_to_buffer(episode_steps[np.where(trajectory.is_last())])
    episode_steps[np.where(trajectory.is_last())] = 0


class EnvironmentSteps(PyMetric):
  """Counts the number of (non-boundary) steps taken in the environment."""

  def __init__(self, name='EnvironmentSteps'):
    super(EnvironmentSteps, self).__init__(name)
    self.reset()

  def reset(self, environment_steps=0):
    self._environment_steps = np.int64(environment_steps)

  def result(self):
    return self._environment_steps

  def call(self, trajectory):
    trajectory = _batched(trajectory)
    self._environment_steps += np.sum((~trajectory.is_boundary()).astype(np.int64))


class NumberOfEpisodes(PyMetric):
  """Counts the number of episodes finished in the environment."""

  def __init__(self, name='NumberOfEpisodes'):
    super(NumberOfEpisodes, self).__init__(name)
    self.reset()

  def reset(self):
    self._number_episodes = np.int64(0)

  def result(self):
    return self._number_episodes

  def call(self, trajectory):
    trajectory = _batched(trajectory)
    self._number_episodes += np.sum(trajectory.is_last().astype(np.int64))


class CounterMetric(PyMetric):
  """Counts how often it was called (e.g. the train / eval iteration number)."""

  def __init__(self, name='Counter'):
    super(CounterMetric, self).__init__(name)
    self.reset()

  def reset(self):
    self._count = np.int64(0)

  def call(self):
    self._count += 1

  def result(self):
    return self._count
