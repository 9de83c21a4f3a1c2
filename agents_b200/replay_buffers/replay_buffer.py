"""ReplayBuffer abstract base (tf_agents/replay_buffers/replay_buffer.py:31-315)."""
import abc

from agents_b200.utils import nest


class ReplayBuffer(abc.ABC):
  """Abstract base class for TF-Agents-style replay buffers."""

  def __init__(self, data_spec, capacity, stateful_dataset=False):
    self._data_spec = data_spec
    self._capacity = capacity
    self._stateful_dataset = stateful_dataset

  @property
  def data_spec(self):
    return self._data_spec

  @property
  def capacity(self):
    return self._capacity

  @property
  def stateful_dataset(self):
    return self._stateful_dataset

  def num_frames(self):
    return self._num_frames()

  def add_batch(self, items):
    return self._add_batch(items)

  def get_next(self, sample_batch_size=None, num_steps=None, time_stacked=True, **kwargs):
    return self._get_next(sample_batch_size, num_steps, time_stacked, **kwargs)

  def as_dataset(self, sample_batch_size=None, num_steps=None, num_parallel_calls=None,
                 sequence_preprocess_fn=None, single_deterministic_pass=False):
    # replay_buffer.py:211-222: a spec with python lists cannot be gathered.
    def has_list(s):
      if isinstance(s, list):
        return True
      if isinstance(s, dict):
        return any(has_list(v) for v in s.values())
      if isinstance(s, tuple):
        return any(has_list(v) for v in s)
      return False

    if has_list(self._data_spec):
      raise ValueError(
          'Cannot perform gather; data spec contains lists and this conflicts '
          'with gathering operator.  Convert any lists to tuples.  '
          'For example, if your spec looks like [a, b, c], '
          'change it to (a, b, c).  Spec structure is:\n  {}'.format(
              nest.map_structure(lambda spec: spec.dtype, self._data_spec)))
    if single_deterministic_pass:
      return self._single_deterministic_pass_dataset(
          sample_batch_size=sample_batch_size, num_steps=num_steps,
          sequence_preprocess_fn=sequence_preprocess_fn,
          num_parallel_calls=num_parallel_calls)
    return self._as_dataset(
        sample_batch_size=sample_batch_size, num_steps=num_steps,
        sequence_preprocess_fn=sequence_preprocess_fn, num_parallel_calls=num_parallel_calls)

  def gather_all(self):
    return self._gather_all()

  def clear(self):
    return self._clear()

  @abc.abstractmethod
  def _num_frames(self):
    raise NotImplementedError

  @abc.abstractmethod
  def _add_batch(self, items):
    raise NotImplementedError

  @abc.abstractmethod
  def _get_next(self, sample_batch_size, num_steps, time_stacked):
    raise NotImplementedError

  @abc.abstractmethod
  def _as_dataset(self, sample_batch_size, num_steps, sequence_preprocess_fn,
                  num_parallel_calls):
    raise NotImplementedError

  @abc.abstractmethod
  def _single_deterministic_pass_dataset(self, sample_batch_size, num_steps,
                                         sequence_preprocess_fn, num_parallel_calls):
    raise NotImplementedError

  @abc.abstractmethod
  def _gather_all(self):
    raise NotImplementedError

  @abc.abstractmethod
  def _clear(self):
    raise NotImplementedError
