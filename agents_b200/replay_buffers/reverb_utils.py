"""Observers that hand host trajectories to a Reverb-model table server.

Same classes, constructor arguments and item-creation rules as
tf_agents/replay_buffers/reverb_utils.py (`ReverbAddEpisodeObserver` :34-267,
`ReverbAddTrajectoryObserver` :270-506, `ReverbTrajectorySequenceObserver` :509-540).  They are
written against the writer protocol only (`py_client.trajectory_writer(num_keep_alive_refs,
validate_items)` -> `append / history / create_item / end_episode / flush / close`), so they run
unchanged over `reverb_local.Client` (steps land in HBM, see reverb_local.py) or over a real
`reverb.Client`.

Rules kept from the reference (checked by tests/test_reverb_host.py against the call counts its
reverb_utils_test.py:300-452 expects):

* one `append` per observed (unbatched) trajectory; observers are for `PyDriver`;
* trajectory observer: an item over the last `sequence_length` steps as soon as that many are
  cached, then one every `stride_length` calls; a boundary step (`step_type == LAST`) ends the
  writer's episode: the cache is dropped, or — `pad_end_of_episodes` — padded with all-zero
  boundary steps (up to a full window, or `sequence_length - 1` of them when
  `tile_end_of_episodes`), emitting items under the same stride rule;
* sequence observer: same, but boundaries do not cut the sequence;
* episode observer: one item per episode (all its steps incl. the boundary step); an episode
  longer than `max_sequence_length` raises, or is skipped entirely with
  `bypass_partial_episodes`.
"""
import logging

import numpy as np

from agents_b200.trajectories import time_step as ts
from agents_b200.trajectories import trajectory as trajectory_lib
from agents_b200.utils import nest


class _WriterObserver(object):
  """Writer lifetime and table plumbing shared by the three observers."""

  def __init__(self, py_client, table_name, keep_alive, priority):
    self._table_names = [table_name] if isinstance(table_name, str) else table_name
    self._priority = priority
    self._py_client = py_client
    self._keep_alive = keep_alive
    self._writer = None
    self.open()

  @property
  def py_client(self):
    return self._py_client

  def get_table_signature(self):
    # every table of one observer shares the signature of the first
    return self._py_client.server_info()[self._table_names[0]].signature

  def _get_writer(self):
    if self._writer is None:
      raise ValueError('Could not obtain writer from py_client.')
    return self._writer

  def _emit(self, window):
    """One item per table over `history[window]` of every column."""
    w = self._get_writer()
    item = nest.map_structure(lambda column: column[window], w.history)
    for name in self._table_names:
      w.create_item(table=name, trajectory=item, priority=self._priority)

  def flush(self):
    """Pushes pending items to the server (needed before sampling right after collecting)."""
    self._get_writer().flush()

  def open(self):
    """Opens the writer; a no-op if it is already open."""
    if self._writer is None:
      self._writer = self._py_client.trajectory_writer(
          num_keep_alive_refs=self._keep_alive + 1, validate_items=False)
      self._on_open()

  def _on_open(self):
    pass

  def close(self):
    """Ends the episode and closes the writer; the observer must be re-opened before reuse."""
    if self._writer is not None:
      self._writer.end_episode()
      self._writer.close()
      self._writer = None
      self._on_close()

  def _on_close(self):
    pass

  def _next_episode(self):
    if self._writer is None:
      self.open()
    else:
      self._writer.end_episode()


class ReverbAddEpisodeObserver(_WriterObserver):
  """Caches the steps of an episode and writes them as ONE item when it ends."""

  def __init__(self, py_client, table_name, max_sequence_length, priority=1,
               bypass_partial_episodes=False):
    if max_sequence_length <= 0:
      raise ValueError('`max_sequence_length` must be an integer greater equal one.')
    self._max_sequence_length = max_sequence_length
    self._bypass_partial_episodes = bypass_partial_episodes
    self._cached_steps = 0
    self._overflow_episode = False
    self._writer_has_data = False
    super().__init__(py_client, table_name, max_sequence_length, priority)

  def update_priority(self, priority):
    self._priority = priority

  def _on_close(self):
    self._writer_has_data = False

  def __call__(self, trajectory):
    """Caches one unbatched step; at a boundary step the episode becomes an item."""
    if self._cached_steps >= self._max_sequence_length and not self._overflow_episode:
      self._overflow_episode = True
      msg = ('The number of trajectories within the same episode exceeds `max_sequence_length`. '
             'Consider increasing the `max_sequence_length`')
      if not self._bypass_partial_episodes:
        raise ValueError(msg + ' or set `bypass_partial_episodes` to true to bypass the episodes '
                         'with length more than `max_sequence_length`.')
      logging.error('%s. This episode is bypassed and will NOT be written into the replay buffer.',
                    msg)
    boundary = bool(np.all(trajectory.is_boundary()))
    if self._overflow_episode:
      if boundary:                       # the over-long episode is over: forget it
        self.reset(write_cached_steps=False)
      return
    self._get_writer().append(trajectory)
    self._writer_has_data = True
    self._cached_steps += 1
    if boundary:
      self.reset(write_cached_steps=True)

  def _write_cached_steps(self):
    if not self._writer_has_data:
      logging.info('Skipped writing to Reverb because the writer is empty.')
      return
    self._emit(slice(None))
    self._writer_has_data = False

  def reset(self, write_cached_steps=True):
    """Forgets the cached steps, writing them out first unless told otherwise."""
    if write_cached_steps:
      self._write_cached_steps()
    self._cached_steps = 0
    self._overflow_episode = False
    self._next_episode()


class ReverbAddTrajectoryObserver(_WriterObserver):
  """Writes fixed-length windows of consecutive steps; episodes cut the windows."""

  _cut_at_boundaries = True

  def __init__(self, py_client, table_name, sequence_length, stride_length=1, priority=1,
               pad_end_of_episodes=False, tile_end_of_episodes=False):
    if tile_end_of_episodes and not pad_end_of_episodes:
      raise ValueError('Must set `pad_end_of_episodes=True` when using `tile_end_of_episodes`')
    self._sequence_length = sequence_length
    self._stride_length = stride_length
    self._pad_end_of_episodes = pad_end_of_episodes
    self._tile_end_of_episodes = tile_end_of_episodes
    self._cached_steps = 0
    self._last_trajectory = None
    super().__init__(py_client, table_name, sequence_length, priority)

  def _on_open(self):
    self._cached_steps = 0

  def __call__(self, trajectory):
    """Appends one unbatched step and writes the window(s) that became complete."""
    self._last_trajectory = trajectory
    self._get_writer().append(trajectory)
    self._cached_steps += 1
    self._write_cached_steps()
    if self._cut_at_boundaries and bool(np.all(trajectory.is_boundary())):
      self.reset(write_cached_steps=self._pad_end_of_episodes)

  def _sequence_lengths_reached(self):
    extra = self._cached_steps - self._sequence_length
    return extra >= 0 and extra % self._stride_length == 0

  def _write_cached_steps(self):
    """Emits the window ending at the newest step if the stride rule says so (cache untouched)."""
    if self._sequence_lengths_reached():
      self._emit(slice(-self._sequence_length, None))

  def _get_padding_step(self, example_trajectory):
    """All-zero boundary step (LAST -> FIRST) shaped like `example_trajectory`."""
    zeros = lambda x: nest.map_structure(lambda a: np.zeros_like(np.asarray(a)), x)
    e = example_trajectory
    like = np.asarray(e.discount)
    return trajectory_lib.Trajectory(
        step_type=np.full(like.shape, ts.StepType.LAST, np.asarray(e.step_type).dtype),
        observation=zeros(e.observation), action=zeros(e.action), policy_info=zeros(e.policy_info),
        next_step_type=np.full(like.shape, ts.StepType.FIRST, np.asarray(e.next_step_type).dtype),
        reward=zeros(e.reward), discount=zeros(e.discount))

  def reset(self, write_cached_steps=True):
    """Clears the cache (after the last collect call, or internally at episode ends).

    With `write_cached_steps` the tail of the cache is written first: padded when padding is on,
    as it is when a full window is pending; otherwise too few steps remain and this raises."""
    if write_cached_steps and self._last_trajectory is not None:
      if self._pad_end_of_episodes:
        pad = self._get_padding_step(self._last_trajectory)
        n_pad = (self._sequence_length - 1 if self._tile_end_of_episodes
                 else self._sequence_length - self._cached_steps)
        for _ in range(n_pad):
          self._get_writer().append(pad)
          self._cached_steps += 1
          self._write_cached_steps()
      elif self._sequence_lengths_reached():
        self._write_cached_steps()
      else:
        raise ValueError(
            'write_cached_steps is True, but not enough steps remain in the cache to write an '
            'item with sequence_length={}, consider enabling pad_end_of_episodes.'.format(
                self._sequence_length))
    self._cached_steps = 0
    self._last_trajectory = None
    self._next_episode()


class ReverbTrajectorySequenceObserver(ReverbAddTrajectoryObserver):
  """`ReverbAddTrajectoryObserver` whose windows run across episode boundaries, so that a
  boundary may sit anywhere inside a sampled sequence."""

  _cut_at_boundaries = False
