"""ReverbReplayBuffer over the in-process HBM table server (reverb_local.py).

Class surface of tf_agents/replay_buffers/reverb_replay_buffer.py:37-458: data goes IN through
the observers of reverb_utils.py (`add_batch` / `get_next` / `gather_all` raise
NotImplementedError with the reference's messages, :169-207,415-430) and comes OUT through
`as_dataset`:

* `as_dataset(sample_batch_size, num_steps)` is an endless iterator of `(data, SampleInfo)`.
  Items are drawn by the table's sampler on the host; the `[B, T]` matrix of step rows they
  reference is gathered from HBM with ONE `b200rl_rb_read_rows` launch.  For whole items of one
  length from a Uniform / Prioritized table the batch is drawn as a vector (no python work per
  sample: `Table.sample_rows`): 0.06 ms instead of 3.8 ms of host time for a batch of 256.  If `num_steps` differs
  from the item length every item is truncated to a multiple of `num_steps`, cut into
  `[rows, num_steps]` sub-sequences (`truncate_reshape_rows_by_num_steps`, :578-613) and those pass
  through a shuffle buffer of `(sequence_length // num_steps) * batch` entries (100 x when the
  length varies), the rule of :283-294 — here applied to row indices, before any byte moves.
* `as_dataset(single_deterministic_pass=True)` needs a deterministic sampler AND remover
  (:376-381); it ends when the table can give nothing more.
* `sequence_preprocess_fn` sees each whole item (`[S, ...]` device tensors) before it is cut
  (:272-276); that path gathers item by item.

Where dm-reverb would block on the rate limiter: with `rate_limiter_timeout_ms >= 0` the
iterator ends (the reference's dataset ends on the timeout), with the default -1 it raises
`reverb_local.RateLimited` instead of hanging the only process there is.
`SampleInfo` fields are host numpy arrays of shape `[B]` (scalars without a batch).
"""
import numpy as np

from agents_b200.replay_buffers import replay_buffer
from agents_b200.replay_buffers import reverb_local
from agents_b200.utils import nest


def truncate_reshape_rows_by_num_steps(sample, num_steps):
  """`[S, ...]` leaves -> `[S // num_steps, num_steps, ...]` (tail dropped) (:578-613)."""
  first = nest.flatten(sample)[0]
  rows = first.shape[0] // num_steps
  return nest.map_structure(
      lambda t: t[:rows * num_steps].reshape((rows, num_steps) + tuple(t.shape[1:])), sample)


class _ShuffleBuffer(object):
  """tf.data `shuffle(buffer_size)`: fill, then swap a random entry out per request."""

  def __init__(self, upstream, size, rng):
    self._up, self._size, self._rng = upstream, max(1, int(size)), rng
    self._buf = []
    self._dry = False

  def __iter__(self):
    return self

  def __next__(self):
    while not self._dry and len(self._buf) < self._size:
      try:
        self._buf.append(next(self._up))
      except StopIteration:
        self._dry = True
    if not self._buf:
      raise StopIteration
    i = int(self._rng.integers(len(self._buf)))
    out = self._buf[i]
    last = self._buf.pop()
    if i < len(self._buf):
      self._buf[i] = last
    return out


class _Dataset(object):
  """Iterator with the two tf.data conveniences the reference's tests use."""

  def __init__(self, gen):
    self._gen = gen

  def __iter__(self):
    return self

  def __next__(self):
    return next(self._gen)

  def take(self, n):
    def limited():
      for _ in range(n):
        try:
          yield next(self._gen)
        except StopIteration:
          return
    return _Dataset(limited())


class ReverbReplayBuffer(replay_buffer.ReplayBuffer):
  """Reverb-model table exposed as a TF-Agents replay buffer."""

  def __init__(self, data_spec, table_name, sequence_length, server_address=None,
               local_server=None, dataset_buffer_size=None, max_cycle_length=32,
               num_workers_per_iterator=-1, max_samples_per_stream=-1,
               rate_limiter_timeout_ms=-1):
    if (server_address is None) == (local_server is None):
      raise ValueError('Exactly one of the server_address or local_server must be provided.')
    self._table_name = table_name
    self._sequence_length = sequence_length
    self._local_server = local_server
    self._dataset_buffer_size = dataset_buffer_size
    self._max_cycle_length = max_cycle_length
    self._rate_limiter_timeout_ms = rate_limiter_timeout_ms
    del num_workers_per_iterator, max_samples_per_stream     # transport tuning of dm-reverb
    self._py_client = reverb_local.Client(local_server if local_server is not None
                                          else server_address)
    self._server_address = 'localhost:{}'.format(self._py_client.server.port)
    self._table = self._py_client.server._table(table_name)
    if self._table.signature is None:
      self._table.signature = data_spec        # types python scalars appended later
    self._table_info = self.get_table_info()
    self._deterministic_table = (self._table_info.sampler_options.is_deterministic and
                                 self._table_info.remover_options.is_deterministic)
    super().__init__(data_spec=data_spec, capacity=self._table_info.max_size,
                     stateful_dataset=True)

  @property
  def py_client(self):
    return self._py_client

  @property
  def local_server(self):
    return self._local_server

  @property
  def tf_client(self):
    return self._py_client

  def get_table_info(self):
    return self._py_client.server_info()[self._table_name]

  def _num_frames(self):
    return self.get_table_info().current_size

  def add_batch(self, items):
    raise NotImplementedError(
        'ReverbReplayBuffer does not support `add_batch`. See `reverb_utils.ReverbObserver` for '
        'more information on how to add data to the buffer.')

  _add_batch = add_batch

  def get_next(self, sample_batch_size=None, num_steps=None, time_stacked=True):
    raise NotImplementedError('ReverbReplayBuffer does not support `get_next`.')

  _get_next = get_next

  def gather_all(self):
    raise NotImplementedError('ReverbReplayBuffer does not support `gather_all`.')

  _gather_all = gather_all

  def _clear(self):
    self._py_client.reset(self._table_name)

  def update_priorities(self, keys, priorities):
    keys = np.asarray(_to_host(keys)).reshape(-1)
    priorities = np.asarray(_to_host(priorities), np.float64).reshape(-1)
    self._py_client.mutate_priorities(
        self._table_name, updates={int(k): float(p) for k, p in zip(keys, priorities)})

  def _verify_num_steps(self, num_steps):
    if num_steps and self._sequence_length:
      if num_steps > self._sequence_length:
        raise ValueError(
            'Can not guarantee sequential data for num_steps as sequence length of the data is '
            'smaller.  This is not supported.  num_steps > sequence_length ({} vs. {})'.format(
                num_steps, self._sequence_length))
      if self._sequence_length % num_steps != 0:
        raise ValueError(
            'Can not guarantee sequential data since sequence_length is not a multiple of '
            'num_steps ({} vs. {})'.format(num_steps, self._sequence_length))

  # -- sampling pipeline ------------------------------------------------------------------------
  def _draw(self):
    """Endless stream of (item, info, pinned) single draws; ends / raises on the rate limiter."""
    while True:
      try:
        picked = self._table.sample(1)
      except reverb_local.RateLimited:
        if self._rate_limiter_timeout_ms >= 0:
          return
        raise
      yield picked[0]

  def _windows(self, num_steps, preprocess):
    """Stream of (rows | device nest, info, release): one entry per delivered sequence."""
    split = bool(num_steps) and num_steps != self._sequence_length
    for item, info in self._draw():
      done = [False]

      def release(item=item, done=done):
        if not done[0]:
          done[0] = True
          item.store.release(item.rows)
      if preprocess is None:
        rows = item.rows
        if not split:
          yield (item.store, rows), info, release
          continue
        n = rows.shape[0] // num_steps
        if n == 0:
          release()
        pieces = rows[:n * num_steps].reshape(n, num_steps)
        left = [n]

        def release_piece(item=item, left=left):
          left[0] -= 1
          if left[0] == 0:
            item.store.release(item.rows)
        for piece in pieces:
          yield (item.store, piece), info, release_piece
      else:
        data = nest.pack_sequence_as(self._data_spec, item.store.read(item.rows))
        release()
        data = preprocess(data)
        if not split:
          yield data, info, None
          continue
        cut = truncate_reshape_rows_by_num_steps(data, num_steps)
        flat = nest.flatten(cut)
        for r in range(flat[0].shape[0]):
          yield nest.pack_sequence_as(cut, [t[r] for t in flat]), info, None

  def _batches(self, stream, sample_batch_size):
    import torch
    want = sample_batch_size or 1
    while True:
      got = []
      for entry in stream:
        got.append(entry)
        if len(got) == want:
          break
      if len(got) < want:            # upstream ended (deterministic pass / rate-limiter timeout)
        for _, _, release in got:
          if release:
            release()
        return
      payloads, infos, releases = zip(*got)
      if isinstance(payloads[0], tuple) and isinstance(payloads[0][1], np.ndarray):
        store = payloads[0][0]
        lengths = {p[1].shape[0] for p in payloads}
        if len(lengths) != 1:
          raise ValueError('Cannot batch sequences of different lengths {}; use '
                           'sample_batch_size=None or 1 for variable-length episodes.'.format(
                               sorted(lengths)))
        rows = np.stack([p[1] for p in payloads])                 # [B, T]
        flat = store.read(rows if sample_batch_size else rows[0])  # ONE gather launch
        data = nest.pack_sequence_as(self._data_spec, flat)
      else:
        if sample_batch_size:
          stack = lambda *ts: np.stack(ts) if isinstance(ts[0], np.ndarray) else torch.stack(ts)
          data = nest.map_structure(stack, *payloads)
        else:
          data = payloads[0]
      for release in releases:
        if release:
          release()
      if sample_batch_size:
        info = reverb_local.SampleInfo(*[np.asarray(col) for col in zip(*infos)])
      else:
        info = infos[0]
      yield data, info

  def _fast_batches(self, sample_batch_size):
    """Whole batches without python work per sample (`Table.sample_rows`): the positions of a
    batch are drawn as one vector, their `[B, T]` step rows come from the table's dense row
    matrix and go to the store as ONE gather.  Falls back to the item-by-item stream for good as
    soon as the table leaves that regime (items of another length, ...)."""
    while True:
      try:
        drawn = self._table.sample_rows(sample_batch_size)
      except reverb_local.RateLimited:
        if self._rate_limiter_timeout_ms >= 0:
          return
        raise
      if drawn is None:
        for batch in self._batches(self._windows(None, None), sample_batch_size):
          yield batch
        return
      pool, rows, info = drawn
      yield nest.pack_sequence_as(self._data_spec, pool.read(rows)), info

  def _as_dataset(self, sample_batch_size=None, num_steps=None, sequence_preprocess_fn=None,
                  num_parallel_calls=None):
    self._verify_num_steps(num_steps)
    if num_parallel_calls and sample_batch_size and num_parallel_calls > sample_batch_size:
      raise ValueError('num_parallel_calls cannot be bigger than sample_batch_size '
                       '{} > {}'.format(num_parallel_calls, sample_batch_size))
    whole_items = not num_steps or num_steps == self._sequence_length
    if sample_batch_size and whole_items and sequence_preprocess_fn is None:
      return _Dataset(self._fast_batches(sample_batch_size))
    stream = self._windows(num_steps, sequence_preprocess_fn)
    if num_steps and num_steps != self._sequence_length:
      total = sample_batch_size or 1
      per_interleave = total // min(total, self._max_cycle_length)
      size = (self._sequence_length // num_steps) if self._sequence_length else 100
      stream = _ShuffleBuffer(stream, size * per_interleave, self._py_client.server.rng)
    return _Dataset(self._batches(stream, sample_batch_size))

  def _single_deterministic_pass_dataset(self, sample_batch_size=None, num_steps=None,
                                         sequence_preprocess_fn=None, num_parallel_calls=None):
    del num_parallel_calls
    if not self._deterministic_table:
      raise ValueError(
          'Unable to perform a single deterministic pass over the dataset, since either the '
          'sampler or the remover is not deterministic (FIFO or Heap).  Table info:\n{}'.format(
              self._table_info))
    self._verify_num_steps(num_steps)
    stream = self._windows_until_empty(num_steps, sequence_preprocess_fn)
    return _Dataset(self._batches(stream, sample_batch_size))

  def _windows_until_empty(self, num_steps, preprocess):
    """The deterministic pass ends where dm-reverb's sampler would wait for more data."""
    saved = self._rate_limiter_timeout_ms
    self._rate_limiter_timeout_ms = max(saved, 0)
    try:
      for entry in self._windows(num_steps, preprocess):
        yield entry
    finally:
      self._rate_limiter_timeout_ms = saved


def _to_host(x):
  return x.detach().cpu().numpy() if hasattr(x, 'detach') else x
