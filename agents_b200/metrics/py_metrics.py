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
  """Ring of the last `maxlen` values (py_metrics.py:34-91); `maxlen=np.inf` grows unbounded."""

  def __init__(self, maxlen, dtype):
    self._start_index = np.int64(0)
    self._len = np.int64(0)
    self._maxlen = np.array(maxlen)
    initial_len = 10 if np.isinf(self._maxlen) else int(self._maxlen)
    self._buffer = np.zeros(shape=(initial_len,), dtype=dtype)

  def clear(self):
    self._start_index = np.int64(0)
    self._len = np.int64(0)

  def add(self, value):
    insert_idx = int((self._start_index + self._len) % self._maxlen)
    if np.isinf(self._maxlen) and insert_idx >= self._buffer.shape[0]:
      self._buffer.resize((self._buffer.shape[0] * 2,), refcheck=False)
    self._buffer[insert_idx] = value
    if self._len < self._maxlen:
      self._len += 1
    else:
      self._start_index = np.mod(self._start_index + 1, self._maxlen)

  def extend(self, values):
    for value in values:
      self.add(value)

  @property
  def last(self):
    if self._len == 0:
      return None
    return self._buffer[int((self._start_index + self._len - 1) % self._maxlen)]

  def __len__(self):
    return int(self._len)

  def mean(self, dtype=None):
    if self._len == self._buffer.shape[0]:
      return np.mean(self._buffer, dtype=dtype)
    assert self._start_index == 0
    return np.mean(self._buffer[:self._len], dtype=dtype)


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
  """Average of the metric over the last (up to) `buffer_size` episodes."""

  def __init__(self, name='StreamingMetric', buffer_size=10, batch_size=None):
    super(StreamingMetric, self).__init__(name)
    self._buffer = NumpyDeque(maxlen=buffer_size, dtype=np.float64)
    self._batch_size = batch_size
    self.reset()

  def reset(self):
    self._buffer.clear()
    if self._batch_size:
      self._reset(self._batch_size)

  @abc.abstractmethod
  def _reset(self, batch_size):
    pass

  def add_to_buffer(self, values):
    self._buffer.extend(values)

  @property
  def data(self):
    return self._buffer

  def result(self):
    if len(self._buffer):
      return self._buffer.mean(dtype=np.float32)
    return np.array(0.0, dtype=np.float32)

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
