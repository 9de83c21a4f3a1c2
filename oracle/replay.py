"""numpy restatement of TFUniformReplayBuffer (TEST INFRASTRUCTURE; see oracle/__init__.py).

Follows replay_buffers/tf_uniform_replay_buffer.py:132-154 (storage), :177-209 (num_frames,
add_batch), :211-310 (get_next), :533-557 (gather_all), :559-579 (clear), :582-635 (helpers)
and replay_buffers/table.py:58-137, line for line, on numpy arrays.
`replay_buffers/py_uniform_replay_buffer.py:100-167` was used as a second reading of the
single-segment ring semantics.
"""
import numpy as np

from oracle import philox


def valid_range_ids(last_id, max_length, num_steps=None):
  # tf_uniform_replay_buffer.py:610-635
  if num_steps is None:
    num_steps = 1
  min_id_not_full = 0
  max_id_not_full = max(last_id + 1 - num_steps + 1, 0)
  min_id_full = last_id + 1 - max_length
  max_id_full = last_id + 1 - num_steps + 1
  if last_id < max_length:
    return min_id_not_full, max_id_not_full
  return min_id_full, max_id_full


class UniformReplayOracle(object):
  """leaf_shapes/dtypes: flat lists describing one item (the flattened data_spec)."""

  def __init__(self, leaf_shapes, leaf_dtypes, batch_size, max_length, seed=0):
    self.batch_size = int(batch_size)
    self.max_length = int(max_length)
    self.capacity = self.batch_size * self.max_length
    # table.py:58-72: zeros([capacity] + shape)
    self.storage = [np.zeros((self.capacity,) + tuple(s), dtype=d)
                    for s, d in zip(leaf_shapes, leaf_dtypes)]
    self.id_table = np.zeros((self.capacity,), dtype=np.int64)   # :152
    self.last_id = -1                                            # :153
    self.batch_offsets = np.arange(self.batch_size, dtype=np.int64) * self.max_length  # :141
    self.seed = seed
    self.rng_call = 0

  def num_frames(self):
    # :177-180
    return min((self.last_id + 1) * self.batch_size, self.capacity)

  def add_batch(self, items):
    # :203-208: id = ++last_id; rows = batch_offsets + id % L; scatter id and every leaf
    self.last_id += 1
    id_ = self.last_id
    rows = self.batch_offsets + id_ % self.max_length
    self.id_table[rows] = id_
    for st, v in zip(self.storage, items):
      st[rows] = np.asarray(v, dtype=st.dtype)

  def draw(self, B, T):
    """The two int64 uniform draws (:265-272) with this project's Philox stream."""
    lo, hi = valid_range_ids(self.last_id, self.max_length, T)
    x, y, z, w = philox.philox(np.arange(B, dtype=np.uint64), self.rng_call, self.seed)
    self.rng_call += 1
    ids = philox.uniform_i64(x, y, lo, hi)
    offs = philox.uniform_i64(z, w, 0, self.batch_size)
    return ids, offs

  def get_next(self, sample_batch_size, num_steps, ids=None, batch_offsets=None):
    """Time-stacked batched form: returns (leaves [B,T,...], ids [B,T], rows [B,T], prob [B])."""
    B, T = int(sample_batch_size), int(num_steps)
    lo, hi = valid_range_ids(self.last_id, self.max_length, T)      # :242-244
    if not hi > lo:                                                  # :246-253
      raise ValueError('TFUniformReplayBuffer is empty. Make sure to add items '
                       'before sampling the buffer.')
    num_ids = hi - lo
    prob = np.float32(0.0) if num_ids == 0 else (
        np.float32(1.0) / np.float32(num_ids * self.batch_size))      # :255-264
    if ids is None:
      ids, batch_offsets = self.draw(B, T)
    ids = np.asarray(ids, dtype=np.int64)
    offs = np.asarray(batch_offsets, dtype=np.int64) * self.max_length  # :273
    step_range = np.arange(T, dtype=np.int64)[None, :]                  # :281-284
    rows = np.mod(step_range + ids[:, None], self.max_length) + offs[:, None]  # :289-292
    data = [st[rows] for st in self.storage]                            # table.py:104-110
    data_ids = self.id_table[rows]                                      # :294
    return data, data_ids, rows, np.full((B,), prob, dtype=np.float32)

  def gather_all(self):
    # :533-557
    lo, hi = valid_range_ids(self.last_id, self.max_length)
    ids = np.arange(lo, hi, dtype=np.int64)
    rows = np.mod(np.stack([ids] * self.batch_size), self.max_length)
    rows = rows + self.batch_offsets[:, None]
    return [st[rows] for st in self.storage]

  def clear(self, clear_all_variables=False):
    # :559-579
    self.last_id = -1
    if clear_all_variables:
      for st in self.storage:
        st[...] = 0
      self.id_table[...] = 0

  def deterministic_row_ids(self, sample_batch_size=None, num_steps=None, window_shift=None,
                            drop_remainder=False):
    """Index sequences of _single_deterministic_pass_dataset (:432-513) (ids before the
    `% capacity` of get_data :518-524)."""
    L = self.max_length
    lo, hi = valid_range_ids(self.last_id, L, None)
    frames = np.arange(lo, hi, dtype=np.int64)
    shift = num_steps if window_shift is None else window_shift

    def window_batches(seq, batch_drop):
      out, start = [], 0
      while start < len(seq):                  # Dataset.window(num_steps, shift)
        w = seq[start:start + num_steps]
        if len(w) == num_steps or not batch_drop:   # .batch(num_steps, drop_remainder)
          out.append(w)
        start += shift
      return out

    res = []
    if sample_batch_size is None:
      for b in range(self.batch_size):         # Dataset.range(batch_size).flat_map(row_ids)
        ids = list(b * L + frames)
        if num_steps is None:
          res.extend(ids)
        else:
          res.extend(np.stack(w) for w in window_batches(ids, drop_remainder))
    else:
      segs = list(range(self.batch_size))
      groups = [np.asarray(segs[i:i + sample_batch_size], dtype=np.int64)
                for i in range(0, self.batch_size, sample_batch_size)]
      if drop_remainder:
        groups = [g for g in groups if len(g) == sample_batch_size]
      for g in groups:                         # batched_row_ids
        rows = [f + g * L for f in frames]
        if num_steps is None:
          res.extend(rows)
        else:                                  # group_windows_drop_remainder + transpose
          res.extend(np.stack(w).T for w in window_batches(rows, True))
    return res
