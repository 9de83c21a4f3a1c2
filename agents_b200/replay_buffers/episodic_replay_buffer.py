"""EpisodicReplayBuffer on HBM (tf_agents/replay_buffers/episodic_replay_buffer.py:52-1580).

Same contract as the reference: episodes are identified by ever-increasing int64 ids handed back
by `add_batch(items, episode_ids)` / `add_sequence(items, episode_id)`; an id of -1 (from
`create_episode_ids`) or an item for which `begin_episode_fn` is true starts a new episode in slot
`id % capacity`, evicting whatever lived there; items addressed to an evicted (stale) id are
dropped; `end_episode_fn` marks an episode completed; `get_next` returns one whole episode drawn
uniformly, `as_dataset(num_steps=n)` draws episodes with probability proportional to their length
and returns a random n-step slice; `gather_all` concatenates all episodes to `[1, sum(T_i), ...]`.

What is different from the reference (and why):
  * an episode slot owns a fixed window of `max_episode_length` rows of one `[capacity *
    max_episode_length (+1 trash row), *leaf]` tensor per leaf instead of an unbounded TensorList:
    static shapes, no allocation on the hot path, rows land with the same bulk-copy kernel as the
    uniform buffer; steps beyond the window are dropped and counted in `overflowed()`;
  * ids / lengths / completed flags are device-resident and updated by ONE launch per add
    (`b200rl_ep_assign`); stream order replaces the tf.CriticalSection;
  * reads whose output shape depends on episode lengths (get_next, gather_all, extract) read the
    lengths back to the host first, like TF eager does;
  * random draws come from this library's Philox stream (the reference's `tf.random` stream is
    unpinned); `Episodes.tensor_lists` of `extract` / `extend_episodes` are `[n, max_len, ...]`
    padded tensors.
"""
import collections

import numpy as np
import torch

from agents_b200 import _lib
from agents_b200.replay_buffers import replay_buffer as replay_buffer_base
from agents_b200.replay_buffers import table
from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod
from agents_b200.utils import nest

_INVALID_EPISODE_ID = -1

Episodes = collections.namedtuple('Episodes', ['length', 'completed', 'tensor_lists'])
BufferInfo = collections.namedtuple('BufferInfo', ['ids'])


def _valid_range_ids(last_id, capacity):
  """[min_id, max_id) of live episode ids (episodic_replay_buffer.py:1563-1580)."""
  if last_id < capacity:
    return 0, max(last_id + 1, 0)
  return last_id + 1 - capacity, last_id + 1


class EpisodicReplayBuffer(replay_buffer_base.ReplayBuffer):
  """An episodic ReplayBuffer with uniform sampling."""

  def __init__(self, data_spec, capacity=1000, completed_only=False, buffer_size=8,
               name_prefix='EpisodicReplayBuffer', device='cuda', seed=None,
               begin_episode_fn=None, end_episode_fn=None, dataset_drop_remainder=False,
               dataset_window_shift=None, max_episode_length=1000):
    super(EpisodicReplayBuffer, self).__init__(data_spec, capacity)
    self._device = torch.device(device)
    if self._device.type != 'cuda':
      raise ValueError('EpisodicReplayBuffer stores its episodes in HBM; device must be a CUDA '
                       f'device (got {device!r}). There is no CPU fallback.')
    self._completed_only = completed_only
    self._buffer_size = buffer_size
    self._name_prefix = name_prefix
    self._seed = int(seed or 0) & 0xFFFFFFFFFFFFFFFF
    self._begin_episode_fn = begin_episode_fn or (lambda traj: traj.is_first())
    self._end_episode_fn = end_episode_fn or (lambda traj: traj.is_last())
    self._dataset_drop_remainder = dataset_drop_remainder
    self._dataset_window_shift = dataset_window_shift
    self._max_len = int(max_episode_length)
    if self._max_len < 1:
      raise ValueError('max_episode_length must be >= 1.')
    dev, cap = self._device, int(capacity)
    with torch.cuda.device(dev):
      # one extra row: the trash row written by items whose episode id is stale
      self._data_table = table.Table(data_spec, cap * self._max_len + 1, device=dev)
      self._loc_to_id = torch.full((cap,), -1, dtype=torch.int64, device=dev)
      self._episode_lengths = torch.zeros(cap, dtype=torch.int64, device=dev)
      self._episode_completed = torch.zeros(cap, dtype=torch.uint8, device=dev)
      self._last_episode = torch.full((1,), -1, dtype=torch.int64, device=dev)
      self._num_writes = torch.zeros(1, dtype=torch.int64, device=dev)
      self._overflow = torch.zeros(1, dtype=torch.int32, device=dev)
      self._rng = torch.zeros(2, dtype=torch.int64, device=dev)
    self._flat_specs = nest.flatten(data_spec)
    self._host_rng = np.random.RandomState(self._seed & 0x7FFFFFFF)

  # ---- properties -------------------------------------------------------------------------------
  @property
  def device(self):
    return self._device

  @property
  def name_prefix(self):
    return self._name_prefix

  @property
  def num_writes(self):
    return self._num_writes

  @property
  def max_episode_length(self):
    return self._max_len

  def variables(self):
    return self._data_table.variables() + [self._loc_to_id, self._episode_lengths,
                                           self._episode_completed, self._last_episode]

  def overflowed(self):
    """True when some step was dropped because its episode outgrew `max_episode_length`."""
    return bool(self._overflow.item())

  def _num_frames(self):
    return self._episode_lengths.sum()

  # ---- ids --------------------------------------------------------------------------------------
  def create_episode_ids(self, num_episodes=None):
    """Initial (invalid, -1) episode id(s) to thread through add_batch / add_sequence (:266-330)."""
    if isinstance(num_episodes, torch.Tensor):
      if num_episodes.dim() != 0:
        raise ValueError('num_episodes must be a scalar, but saw shape: {}'.format(
            tuple(num_episodes.shape)))
      num_episodes = int(num_episodes.item())
    shape = ()
    if num_episodes is not None and num_episodes > 0:
      if num_episodes > self._capacity:
        raise ValueError('Buffer cannot create episode_ids when num_episodes {} > capacity '
                         '{}.'.format(num_episodes, self._capacity))
      shape = (num_episodes,)
    return torch.full(shape, _INVALID_EPISODE_ID, dtype=torch.int64, device=self._device)

  def _flags(self, value, n):
    """bool / tensor -> contiguous uint8 [n] device tensor (None for a python False)."""
    if value is None or (isinstance(value, bool) and not value):
      return None
    t = torch.as_tensor(value, device=self._device)
    return t.to(torch.uint8).expand(n).contiguous() if t.dim() == 0 else t.to(torch.uint8).reshape(n).contiguous()

  def _assign(self, ids, begin, end, mask=None, steps=None, steps_all=0, bump=False,
              set_completed=False, want_rows=False):
    n = ids.numel()
    rows = torch.empty(n, dtype=torch.int64, device=self._device) if want_rows else None
    keep = (begin, end, mask, steps)          # alive until the launch is enqueued
    _lib.call('b200rl_ep_assign', _lib.ptr(ids), _lib.ptr(begin), _lib.ptr(end), _lib.ptr(mask),
              _lib.ptr(steps), int(steps_all), n, int(self._capacity), self._max_len,
              _lib.ptr(self._last_episode), _lib.ptr(self._loc_to_id),
              _lib.ptr(self._episode_lengths), _lib.ptr(self._episode_completed),
              _lib.ptr(self._num_writes), int(bump), int(set_completed), _lib.ptr(rows),
              _lib.ptr(self._overflow), _lib.stream())
    del keep
    return rows

  def _get_episode_id(self, episode_id, begin_episode=False, end_episode=False):
    """Scalar form of `_get_batch_episode_ids` (:1047-1107)."""
    ids = torch.as_tensor(episode_id, dtype=torch.int64, device=self._device).reshape(1).clone()
    with torch.cuda.device(self._device):
      self._assign(ids, self._flags(begin_episode, 1), self._flags(end_episode, 1))
    return ids.reshape(())

  def _get_batch_episode_ids(self, batch_episode_ids, begin_episode=False, end_episode=False,
                             mask=None):
    """New consecutive ids for entries that are invalid or begin an episode (:1109-1187)."""
    ids = torch.as_tensor(batch_episode_ids, dtype=torch.int64, device=self._device)
    if ids.dim() != 1:
      raise ValueError('batch_episode_ids must be a vector with 1 dimension')
    ids = ids.clone()
    n = ids.numel()
    with torch.cuda.device(self._device):
      self._assign(ids, self._flags(begin_episode, n), self._flags(end_episode, n),
                   self._flags(mask, n) if mask is not None else None)
    return ids

  def _maybe_end_episode(self, episode_id, end_episode=False):
    ids = torch.as_tensor(episode_id, dtype=torch.int64, device=self._device).reshape(1).clone()
    with torch.cuda.device(self._device):
      self._assign(ids, None, self._flags(end_episode, 1),
                   mask=torch.zeros(1, dtype=torch.uint8, device=self._device))
    loc = ids % self._capacity
    return self._episode_completed[loc].reshape(()) > 0

  def _maybe_end_batch_episodes(self, batch_episode_ids, end_episode=False):
    ids = torch.as_tensor(batch_episode_ids, dtype=torch.int64, device=self._device).clone()
    n = ids.numel()
    with torch.cuda.device(self._device):
      self._assign(ids, None, self._flags(end_episode, n),
                   mask=torch.zeros(n, dtype=torch.uint8, device=self._device))
    return self._episode_completed[ids % self._capacity] > 0

  def _get_episode_id_location(self, episode_id):
    return torch.as_tensor(episode_id, dtype=torch.int64, device=self._device) % self._capacity

  def _get_last_episode_id(self):
    return int(self._last_episode.item())

  def get_valid_ids_mask(self, episode_ids):
    ids = torch.as_tensor(episode_ids, dtype=torch.int64, device=self._device)
    return (ids >= 0) & (self._loc_to_id[ids % self._capacity] == ids)

  def _completed_episodes(self):
    return self._loc_to_id[self._episode_completed == 1]

  # ---- add --------------------------------------------------------------------------------------
  def _prepare(self, items, outer):
    nest.assert_same_structure(items, self._data_spec)
    out = []
    for v, s in zip(nest.flatten(items), self._flat_specs):
      v = torch.as_tensor(v, device=self._device)
      if v.dtype != s.dtype:
        v = v.to(s.dtype)
      if tuple(v.shape) != tuple(outer) + tuple(s.shape):
        raise ValueError('Tensor shape {} vs. expected {} for spec {}.'.format(
            tuple(v.shape), tuple(outer) + tuple(s.shape), s))
      out.append(v.contiguous())
    return out

  def add_batch(self, items, episode_ids):
    """Adds one step to each of `episode_ids` `[num_episodes]`; returns the updated ids (:402-463)."""
    ids = torch.as_tensor(episode_ids, dtype=torch.int64, device=self._device)
    if ids.dim() != 1:
      raise ValueError('episode_ids must be a vector.')
    ids = ids.clone()
    n = ids.numel()
    flat = self._prepare(items, (n,))
    begin = self._flags(self._begin_episode_fn(items), n)
    end = self._flags(self._end_episode_fn(items), n)
    with torch.cuda.device(self._device):
      rows = self._assign(ids, begin, end, steps_all=1, bump=True, want_rows=True)
      self._data_table.write(rows, nest.pack_sequence_as(self._data_spec, flat))
    return ids

  def add_sequence(self, items, episode_id):
    """Appends a `[T, ...]` sequence to one episode; returns the updated id (:332-400)."""
    eid = torch.as_tensor(episode_id, dtype=torch.int64, device=self._device)
    if eid.dim() != 0:
      raise ValueError('episode_id must be a scalar.')
    ids = eid.reshape(1).clone()
    T = int(nest.flatten(items)[0].shape[0])
    flat = self._prepare(items, (T,))
    begin = torch.as_tensor(self._begin_episode_fn(items), device=self._device).to(torch.uint8).reshape(-1)[:1]
    end = torch.as_tensor(self._end_episode_fn(items), device=self._device).to(torch.uint8).reshape(-1)
    end = end.max().reshape(1) if end.numel() else end
    with torch.cuda.device(self._device):
      rows = self._assign(ids, begin.contiguous(), end.contiguous(), steps_all=T, want_rows=True)
      trash = int(self._capacity) * self._max_len
      rows = torch.where(rows[0] == trash, torch.full((T,), trash, dtype=torch.int64, device=self._device),
                         rows[0] + torch.arange(T, dtype=torch.int64, device=self._device))
      self._data_table.write(rows, nest.pack_sequence_as(self._data_spec, flat))
    return ids.reshape(())

  def _add_batch(self, items):
    raise NotImplementedError('add_batch(items) is not implemented in EpisodicReplayBuffer. '
                              'Use add_batch(items, episode_ids) instead')

  # ---- read -------------------------------------------------------------------------------------
  def _episode_rows(self, loc, start, length):
    return loc * self._max_len + start + torch.arange(length, dtype=torch.int64, device=self._device)

  def _get_episode(self, episode_id):
    """All steps of `episode_id`, each leaf `[length, ...]` (:934-969)."""
    eid = int(torch.as_tensor(episode_id).item())
    loc = eid % self._capacity
    at = int(self._loc_to_id[loc].item())
    if eid < 0 or at != eid:
      raise rb_mod.InvalidArgumentError(
          'Episode id {} is not valid.  It points to location {} but the episode at that location '
          'is currently id {}'.format(eid, loc, at))
    n = int(self._episode_lengths[loc].item())
    return self._data_table.read(self._episode_rows(loc, 0, n))

  def _sample_episode_ids(self, shape, weigh_by_episode_length=False, seed=None):
    """Episode ids, uniform over the live range or proportional to length (:1192-1226)."""
    last = self._get_last_episode_id()
    if last < 0:
      raise rb_mod.InvalidArgumentError(
          'EpisodicReplayBuffer is empty. Make sure to add items before sampling the buffer.')
    count = int(np.prod(shape)) if len(shape) else 1
    if weigh_by_episode_length:
      num = min(last + 1, self._capacity)
      lengths = self._episode_lengths[:num].cpu().numpy().astype(np.float64)
      total = lengths.sum()
      p = lengths / total if total > 0 else np.full(num, 1.0 / num)
      locs = self._host_rng.choice(num, size=count, p=p)
      ids = self._loc_to_id.cpu().numpy()[locs]
    else:
      lo, hi = _valid_range_ids(last, self._capacity)
      ids = self._host_rng.randint(lo, hi, size=count)
    return torch.as_tensor(ids.reshape(shape), dtype=torch.int64, device=self._device)

  def _get_next(self, sample_batch_size=None, num_steps=None, time_stacked=None):
    """One whole episode drawn uniformly (:485-507)."""
    eid = int(self._sample_episode_ids(()).item())
    loc = eid % self._capacity
    n = int(self._episode_lengths[loc].item())
    data = self._data_table.read(self._episode_rows(loc, 0, n))
    return data, BufferInfo(ids=torch.full((), eid, dtype=torch.int64, device=self._device))

  def _as_dataset(self, sample_batch_size=None, num_steps=None, sequence_preprocess_fn=None,
                  num_parallel_calls=None):
    """Infinite generator of episodes (num_steps=None) or random num_steps slices (:509-691)."""
    if sequence_preprocess_fn is not None:
      raise NotImplementedError('sequence_preprocess_fn is not supported.')
    if sample_batch_size and num_steps is None:
      raise ValueError('`num_steps` must be set if `sample_batch_size` is set in '
                       'EpisodicReplayBuffer as_dataset.')

    def one():
      while True:
        eid = int(self._sample_episode_ids((), weigh_by_episode_length=num_steps is not None).item())
        loc = eid % self._capacity
        if self._completed_only and not int(self._episode_completed[loc].item()):
          continue
        n = int(self._episode_lengths[loc].item())
        if num_steps is None:
          return self._data_table.read(self._episode_rows(loc, 0, n)), eid
        if n < num_steps:
          continue
        start = int(self._host_rng.randint(0, n - num_steps + 1))
        return self._data_table.read(self._episode_rows(loc, start, num_steps)), eid

    def gen():
      while True:
        if not sample_batch_size:
          data, eid = one()
          yield data, BufferInfo(ids=torch.full((), eid, dtype=torch.int64, device=self._device))
        else:
          got = [one() for _ in range(sample_batch_size)]
          data = nest.map_structure(lambda *xs: torch.stack(xs), *[g[0] for g in got])
          yield data, BufferInfo(ids=torch.as_tensor([g[1] for g in got], dtype=torch.int64,
                                                     device=self._device))
    return gen()

  def _single_deterministic_pass_dataset(self, sample_batch_size=None, num_steps=None,
                                         sequence_preprocess_fn=None, num_parallel_calls=None):
    """Episodes in id order; with num_steps, windows over the concatenated steps (:693-810)."""
    if sequence_preprocess_fn is not None:
      raise NotImplementedError('sequence_preprocess_fn is not supported.')
    if sample_batch_size is not None and num_steps is None:
      raise ValueError('When requesting a batched dataset from EpisodicReplayBuffer, num_steps '
                       'must be provided (but saw num_steps=None).')
    drop, shift = self._dataset_drop_remainder, self._dataset_window_shift
    lo, hi = _valid_range_ids(self._get_last_episode_id(), self._capacity)
    lengths = self._episode_lengths.cpu().numpy()

    def rows_of(ids):
      out = []
      for eid in ids:
        loc = eid % self._capacity
        out.append(loc * self._max_len + np.arange(int(lengths[loc]), dtype=np.int64))
      return np.concatenate(out) if out else np.zeros(0, np.int64)

    def windows(rows, keep_partial):
      step = num_steps if shift is None else shift
      out = []
      for s in range(0, len(rows), step):
        w = rows[s:s + num_steps]
        if len(w) == num_steps or (keep_partial and len(w) > 0):
          out.append(w)
      return out

    read = lambda r: self._data_table.read(torch.as_tensor(r, dtype=torch.int64, device=self._device))
    ids = list(range(lo, hi))
    if sample_batch_size is None:
      if num_steps is None:
        return [read(rows_of([e])) for e in ids]
      return [read(w) for w in windows(rows_of(ids), keep_partial=not drop)]
    # shard the episodes round-robin, window every shard, interleave, batch (:771-808)
    shards = [windows(rows_of(ids[i::sample_batch_size]), keep_partial=False)
              for i in range(sample_batch_size)]
    order = []
    for j in range(max((len(s) for s in shards), default=0)):
      for s in shards:
        if j < len(s):
          order.append(s[j])
    out = []
    for i in range(0, len(order), sample_batch_size):
      batch = order[i:i + sample_batch_size]
      if len(batch) == sample_batch_size or not drop:
        out.append(read(np.stack(batch)))
    return out

  def gather_all(self):
    """All steps of all (completed, if `completed_only`) episodes as `[1, sum(T_i), ...]`."""
    items, _ = self._gather_all()
    return items

  def _gather_all(self):
    lo, hi = _valid_range_ids(self._get_last_episode_id(), self._capacity)
    lengths = self._episode_lengths.cpu().numpy()
    completed = self._episode_completed.cpu().numpy()
    rows, ids = [], []
    for eid in range(lo, hi):
      loc = eid % self._capacity
      if self._completed_only and not completed[loc]:
        continue
      n = int(lengths[loc])
      rows.append(loc * self._max_len + np.arange(n, dtype=np.int64))
      ids.append(np.full(n, eid, dtype=np.int64))
    if not rows or sum(len(r) for r in rows) == 0:
      empty = nest.map_structure(
          lambda s: torch.zeros((0,) + tuple(s.shape), dtype=s.dtype, device=self._device),
          self._data_spec)
      return empty, torch.zeros((), dtype=torch.int64, device=self._device)
    rows = torch.as_tensor(np.concatenate(rows)[None, :], device=self._device)
    return self._data_table.read(rows), torch.as_tensor(np.concatenate(ids)[None, :],
                                                        device=self._device)

  # ---- clear / extract / extend -----------------------------------------------------------------
  def _clear(self, clear_all_variables=False):
    """Drops the stored steps; with clear_all_variables also forgets the ids in flight (:890-917)."""
    self._episode_lengths.zero_()
    self._num_writes.zero_()
    self._overflow.zero_()
    if clear_all_variables:
      self._episode_completed.zero_()
      self._loc_to_id.fill_(_INVALID_EPISODE_ID)
      self._last_episode.fill_(_INVALID_EPISODE_ID)

  def clear(self, clear_all_variables=False):
    return self._clear(clear_all_variables)

  def extract(self, locations, clear_data=False):
    """Episodes(length, completed, padded tensors `[n, max_len, ...]`) at `locations` (:1292-1334)."""
    loc = torch.as_tensor(locations, dtype=torch.int64, device=self._device)
    if loc.dim() != 1:
      raise ValueError('locations must be a vector.')
    rows = loc[:, None] * self._max_len + torch.arange(self._max_len, dtype=torch.int64,
                                                       device=self._device)[None, :]
    out = Episodes(length=self._episode_lengths[loc].clone(),
                   completed=self._episode_completed[loc].clone(),
                   tensor_lists=self._data_table.read(rows))
    if clear_data:
      self._episode_lengths[loc] = 0
      self._episode_completed[loc] = 0
    return out

  def extend_episodes(self, episode_ids, episode_ids_indices, episodes):
    """Appends `episodes.length[i]` steps of `episodes.tensor_lists[i]` to episode
    `episode_ids[episode_ids_indices[i]]`; returns the updated `episode_ids` (:1336-1414)."""
    ids = torch.as_tensor(episode_ids, dtype=torch.int64, device=self._device)
    idx = torch.as_tensor(episode_ids_indices, dtype=torch.int64, device=self._device)
    if ids.dim() != 1 or idx.dim() != 1:
      raise ValueError('episode_ids and episode_ids_indices must be vectors.')
    ids = ids.clone()
    m, n = ids.numel(), idx.numel()
    begin = torch.as_tensor(self._begin_episode_fn(episodes), device=self._device)
    begin = begin.to(torch.uint8).expand(n) if begin.dim() == 0 else begin.to(torch.uint8).reshape(n)
    exp_begin = torch.zeros(m, dtype=torch.uint8, device=self._device)
    exp_begin[idx] = begin
    exp_mask = torch.zeros(m, dtype=torch.uint8, device=self._device)
    exp_mask[idx] = 1
    with torch.cuda.device(self._device):
      self._assign(ids, exp_begin, None, mask=exp_mask)                 # renew ids (:1393-1399)
      sub = ids[idx].clone()
      steps = torch.as_tensor(episodes.length, dtype=torch.int64, device=self._device).reshape(n).contiguous()
      comp = torch.as_tensor(episodes.completed, device=self._device).to(torch.uint8).reshape(n).contiguous()
      rows = self._assign(sub, None, comp, mask=torch.zeros(n, dtype=torch.uint8, device=self._device),
                          steps=steps, set_completed=True, want_rows=True)
      # copy the valid prefix of every padded episode tensor
      steps_h, rows_h = steps.cpu().numpy(), rows.cpu().numpy()
      trash = int(self._capacity) * self._max_len
      flat = [torch.as_tensor(t, device=self._device) for t in nest.flatten(episodes.tensor_lists)]
      for i in range(n):
        k = int(steps_h[i])
        if k == 0 or rows_h[i] == trash:
          continue
        dst = int(rows_h[i]) + torch.arange(k, dtype=torch.int64, device=self._device)
        self._data_table.write(dst, nest.pack_sequence_as(self._data_spec, [t[i, :k] for t in flat]))
    return ids


class StatefulEpisodicReplayBuffer(replay_buffer_base.ReplayBuffer):
  """Keeps the episode ids between calls (episodic_replay_buffer.py:1416-1561), so that
  `add_batch(items)` can be used as a driver observer."""

  def __init__(self, replay_buffer, num_episodes=None):
    super(StatefulEpisodicReplayBuffer, self).__init__(replay_buffer.data_spec,
                                                       replay_buffer.capacity)
    if not isinstance(replay_buffer, EpisodicReplayBuffer):
      raise TypeError('Expected an EpisodicReplayBuffer, saw {}'.format(replay_buffer))
    self._replay_buffer = replay_buffer
    self._episode_ids_var = replay_buffer.create_episode_ids(num_episodes)

  @property
  def episode_ids(self):
    return self._episode_ids_var

  def add_batch(self, items):
    new = self._replay_buffer.add_batch(items=items, episode_ids=self._episode_ids_var)
    self._episode_ids_var.copy_(new)
    return new

  def add_sequence(self, items):
    new = self._replay_buffer.add_sequence(items=items, episode_id=self._episode_ids_var)
    self._episode_ids_var.copy_(new)
    return new

  def extend_episodes(self, episode_ids_indices, episodes):
    new = self._replay_buffer.extend_episodes(episode_ids=self._episode_ids_var,
                                              episode_ids_indices=episode_ids_indices,
                                              episodes=episodes)
    self._episode_ids_var.copy_(new)
    return new

  def _num_frames(self):
    return self._replay_buffer.num_frames()

  def _add_batch(self, items):
    return self.add_batch(items)

  def _get_next(self, sample_batch_size=None, num_steps=None, time_stacked=None):
    return self._replay_buffer.get_next(sample_batch_size, num_steps, time_stacked)

  def _as_dataset(self, sample_batch_size=None, num_steps=None, sequence_preprocess_fn=None,
                  num_parallel_calls=None):
    return self._replay_buffer.as_dataset(sample_batch_size, num_steps,
                                          sequence_preprocess_fn=sequence_preprocess_fn,
                                          num_parallel_calls=num_parallel_calls)

  def _single_deterministic_pass_dataset(self, sample_batch_size=None, num_steps=None,
                                         sequence_preprocess_fn=None, num_parallel_calls=None):
    return self._replay_buffer.as_dataset(sample_batch_size, num_steps,
                                          sequence_preprocess_fn=sequence_preprocess_fn,
                                          num_parallel_calls=num_parallel_calls,
                                          single_deterministic_pass=True)

  def _gather_all(self):
    return self._replay_buffer.gather_all()

  def gather_all(self):
    return self._replay_buffer.gather_all()

  def _clear(self):
    return self._replay_buffer.clear()
