"""A batched replay buffer of nests of Tensors sampled uniformly, resident in HBM.

Drop-in for `tf_agents.replay_buffers.tf_uniform_replay_buffer.TFUniformReplayBuffer`
(reference file replay_buffers/tf_uniform_replay_buffer.py:47-635): same constructor
arguments, methods and return structures; tensors are torch CUDA tensors.

Storage is `B == batch_size` segments of `L == max_length` rows (reference :64-94); the ring
write, the two uniform int64 draws, the `[B,T]` row-id arithmetic and the per-leaf gathers of
`_add_batch` (:182-209) and `_get_next` (:211-310) each run as ONE kernel of libb200rl
(csrc/replay.cu).  `last_id` and the RNG call counter live on the device; the host keeps a
mirror of `last_id` only to raise the reference's "buffer is empty" error without a sync.

Extensions beyond the reference signature (all keyword-only, all optional):
  * `seed` (constructor): Philox key of the sampler (the reference's draws are unseeded).
  * `get_next(..., ids=, batch_offsets=)`: externally supplied draws ("oracle mode") used by
    the parity tests: row ids are then bit-exact functions of the inputs.
"""
import collections
import ctypes

import numpy as np
import torch

from agents_b200 import _lib
from agents_b200.replay_buffers import replay_buffer
from agents_b200.replay_buffers import table
from agents_b200.specs import tensor_spec
from agents_b200.utils import common
from agents_b200.utils import nest

BufferInfo = collections.namedtuple('BufferInfo', ['ids', 'probabilities'])


class InvalidArgumentError(ValueError):
  """Stands in for tf.errors.InvalidArgumentError (raised on an empty buffer, :246-253)."""


def _valid_range_ids(last_id, max_length, num_steps=None):
  """[min_id, max_id) of sampleable ids (reference :610-635), host integers."""
  if num_steps is None:
    num_steps = 1
  if last_id < max_length:
    return 0, max(last_id + 1 - num_steps + 1, 0)
  return last_id + 1 - max_length, last_id + 1 - num_steps + 1


class _SampleDataset(object):
  """Infinite iterable of `get_next` results (reference `_as_dataset`, :329-367)."""

  def __init__(self, buffer, sample_batch_size, num_steps):
    self._buffer = buffer
    self._sample_batch_size = sample_batch_size
    self._num_steps = num_steps

  def __iter__(self):
    return self

  def __next__(self):
    return self._buffer.get_next(self._sample_batch_size, self._num_steps, time_stacked=True)

  def prefetch(self, _):  # tf.data API compatibility: device-side sampling needs no prefetch
    return self

  def take(self, n):
    return (next(self) for _ in range(n))


class _ListDataset(object):
  """Finite re-iterable dataset produced lazily by a factory (deterministic pass)."""

  def __init__(self, factory):
    self._factory = factory

  def __iter__(self):
    return iter(self._factory())

  def prefetch(self, _):
    return self


class TFUniformReplayBuffer(replay_buffer.ReplayBuffer):
  """A TFUniformReplayBuffer with batched adds and uniform sampling."""

  def __init__(self,
               data_spec,
               batch_size,
               max_length=1000,
               scope='TFUniformReplayBuffer',
               device='cuda',
               table_fn=table.Table,
               dataset_drop_remainder=False,
               dataset_window_shift=None,
               stateful_dataset=False,
               seed=0):
    self._batch_size = int(batch_size)
    self._max_length = int(max_length)
    capacity = self._batch_size * self._max_length
    super(TFUniformReplayBuffer, self).__init__(data_spec, capacity, stateful_dataset)
    self._id_spec = tensor_spec.TensorSpec([], torch.int64, name='id')
    self._scope = scope
    self._device = torch.device(device)
    if self._device.type != 'cuda':
      raise ValueError('TFUniformReplayBuffer stores its tables in HBM; device must be a CUDA '
                       f'device (got {device!r}). There is no CPU fallback.')
    self._table_fn = table_fn
    self._dataset_drop_remainder = dataset_drop_remainder
    self._dataset_window_shift = dataset_window_shift
    self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    with torch.cuda.device(self._device):
      self._data_table = table_fn(self._data_spec, capacity, device=self._device)
      self._id_table = table_fn(self._id_spec, capacity, device=self._device)
      self._last_id = torch.full((), -1, dtype=torch.int64, device=self._device)
      # [0] RNG call index (uint64 bit pattern), [1] status, [2] ticket scratch
      self._ctrl = torch.zeros(4, dtype=torch.int64, device=self._device)
    self._last_id_host = -1
    self._flat_specs = nest.flatten(self._data_spec)
    self._fused = isinstance(self._data_table, table.Table) and isinstance(
        self._id_table, table.Table)
    if self._fused:
      self._ring = table.make_ring(
          self._data_table.variables(), self._flat_specs, self._batch_size, self._max_length,
          id_table=self._id_table.variables()[0], last_id=self._last_id,
          ticket=self._ctrl[2:3])

  def variables(self):
    return self._data_table.variables() + self._id_table.variables() + [self._last_id]

  # ---- checkpointing (the reference checkpoints `variables()`: tf_uniform_replay_buffer.py:156-161)
  def state_dict(self):
    """Host copy of the ring: per-leaf storage, id table, last_id and the sampling counters."""
    return {
        'data': [v.detach().cpu() for v in self._data_table.variables()],
        'ids': self._id_table.variables()[0].detach().cpu(),
        'last_id': int(self._last_id.item()),
        'ctrl': self._ctrl.detach().cpu(), 'seed': self._seed,
        'batch_size': self._batch_size, 'max_length': self._max_length,
    }

  def load_state_dict(self, state):
    if (state['batch_size'], state['max_length']) != (self._batch_size, self._max_length):
      raise ValueError('Checkpointed replay buffer has batch_size={}, max_length={}; this one has '
                       'batch_size={}, max_length={}.'.format(
                           state['batch_size'], state['max_length'], self._batch_size,
                           self._max_length))
    dst = self._data_table.variables()
    if len(dst) != len(state['data']):
      raise ValueError('Checkpointed replay buffer has a different data_spec.')
    for d, s in zip(dst, state['data']):
      d.copy_(s.to(d.device))
    self._id_table.variables()[0].copy_(state['ids'].to(self._device))
    self._ctrl.copy_(state['ctrl'].to(self._device))
    self._seed = state.get('seed', self._seed)   # resume the same Philox stream
    self._last_id.fill_(int(state['last_id']))
    self._last_id_host = int(state['last_id'])

  @property
  def device(self):
    return self._device

  @property
  def table_fn(self):
    return self._table_fn

  @property
  def scope(self):
    return self._scope

  @property
  def batch_size(self):
    return self._batch_size

  @property
  def max_length(self):
    return self._max_length

  # ---- counters -------------------------------------------------------------------------
  def _get_last_id(self):
    return self._last_id_host

  def sync_last_id_from_device(self):
    """Re-reads last_id from HBM (after replaying a CUDA graph that contained add_batch)."""
    self._last_id_host = int(self._last_id.item())
    return self._last_id_host

  def _note_adds(self, n):
    """Tells the host mirror that `n` add_batch launches ran outside this object's view."""
    self._last_id_host += int(n)

  def _note_clear(self):
    self._last_id_host = -1

  def _num_frames(self):
    total = (self._get_last_id() + 1) * self._batch_size
    return torch.tensor(min(total, self._capacity), dtype=torch.int64)

  # ---- add ------------------------------------------------------------------------------
  def _prepare_items(self, items, outer):
    nest.assert_same_structure(items, self._data_spec)
    flat = nest.flatten(items)
    out = []
    for v, s in zip(flat, self._flat_specs):
      v = torch.as_tensor(v, device=self._device)
      if v.dtype != s.dtype:
        v = v.to(s.dtype)
      if tuple(v.shape) != tuple(outer) + s.shape:
        raise ValueError(
            'Received a mix of batched and unbatched Tensors, or Tensors are not compatible '
            'with Specs.  Tensor shape {} vs. expected {} for spec {}.'.format(
                tuple(v.shape), tuple(outer) + s.shape, s))
      out.append(v.contiguous())
    return out

  def _add_batch(self, items):
    """Adds a batch of items shaped [batch_size, ...] (reference :182-209)."""
    flat = self._prepare_items(items, (self._batch_size,))
    with torch.cuda.device(self._device):
      if self._fused:
        ptrs = _lib.ptr_array(flat)
        _lib.call('b200rl_rb_add_batch', ctypes.byref(self._ring), ptrs, _lib.stream())
      else:
        id_ = self._last_id_host + 1
        rows = (torch.arange(self._batch_size, dtype=torch.int64, device=self._device) *
                self._max_length + id_ % self._max_length)
        self._id_table.write(rows, torch.full((self._batch_size,), id_, dtype=torch.int64,
                                              device=self._device))
        self._data_table.write(rows, nest.pack_sequence_as(self._data_spec, flat))
        self._last_id.fill_(id_)
    self._last_id_host += 1
    # inside a captured step (common.function) every replay advances the device counter: keep the
    # host mirror in step (gather_all / num_frames / the empty check read it)
    common.record_host_effect(lambda: self._note_adds(1))

  # ---- sample ---------------------------------------------------------------------------
  def _get_next(self, sample_batch_size=None, num_steps=None, time_stacked=True, ids=None,
                batch_offsets=None, out=None):
    """Samples items uniformly (reference :211-310).

    Returns `(data, BufferInfo(ids, probabilities))`; data leaves are `[B, T, ...]`,
    `[B, ...]`, `[T, ...]` or `[...]` exactly as in the reference.

    `out`: a `(data, BufferInfo)` pair returned by an earlier call with the same
    `sample_batch_size` / `num_steps` (both given): the sample is written into those tensors --
    the double buffer of a prefetching input pipeline (`dataset.prefetch` in the reference's
    examples, agents/dqn/examples/v2/train_eval.py:226-232) without an allocation per step.
    """
    T = 1 if num_steps is None else int(num_steps)
    B = 1 if sample_batch_size is None else int(sample_batch_size)
    if T > self._max_length:
      raise ValueError('num_steps ({}) is bigger than max_length ({}).'.format(
          T, self._max_length))
    min_val, max_val = _valid_range_ids(self._get_last_id(), self._max_length, T)
    if max_val <= min_val:
      raise InvalidArgumentError(
          'TFUniformReplayBuffer is empty. Make sure to add items before sampling the buffer.')
    if (ids is None) != (batch_offsets is None):
      raise ValueError('ids and batch_offsets must be given together.')
    dev = self._device
    with torch.cuda.device(dev):
      if ids is not None:
        ids = torch.as_tensor(ids, dtype=torch.int64, device=dev).reshape(B).contiguous()
        batch_offsets = torch.as_tensor(batch_offsets, dtype=torch.int64,
                                        device=dev).reshape(B).contiguous()
      if out is not None:
        if sample_batch_size is None or num_steps is None or not time_stacked or not self._fused:
          raise ValueError('out= needs sample_batch_size, num_steps and the fused sampler.')
        outs = nest.flatten(out[0])
        out_ids, probs = out[1].ids, out[1].probabilities
        for t, s in zip(outs, self._flat_specs):
          if tuple(t.shape) != (B, T) + tuple(s.shape) or t.dtype != s.dtype or not t.is_contiguous():
            raise ValueError('out= does not match this sample shape.')
        if tuple(out_ids.shape) != (B, T) or tuple(probs.shape) != (B,):
          raise ValueError('out= does not match this sample shape.')
      else:
        outs = [torch.empty((B, T) + s.shape, dtype=s.dtype, device=dev)
                for s in self._flat_specs]
        out_ids = torch.empty((B, T), dtype=torch.int64, device=dev)
        probs = torch.empty((B,), dtype=torch.float32, device=dev)
      if self._fused:
        out_ptrs = _lib.ptr_array(outs)
        _lib.call('b200rl_rb_sample', ctypes.byref(self._ring), B, T, _lib.ptr(ids),
                  _lib.ptr(batch_offsets), self._seed, _lib.ptr(self._ctrl[0:1]), out_ptrs,
                  _lib.ptr(out_ids), None, _lib.ptr(probs), _lib.ptr(self._ctrl[1:2]),
                  _lib.stream())
      else:
        outs, out_ids, probs = self._get_next_generic(B, T, ids, batch_offsets, min_val, max_val)
    data = nest.pack_sequence_as(self._data_spec, outs)
    squeeze_b = sample_batch_size is None
    squeeze_t = num_steps is None

    def fix(t):
      if squeeze_t:
        t = t[:, 0]
      if squeeze_b:
        t = t[0]
      return t

    if num_steps is not None and not time_stacked:
      # tuple of per-step items (reference :295-306)
      steps = []
      step_ids = []
      for t in range(T):
        item = nest.map_structure(lambda x: (x[0, t] if squeeze_b else x[:, t]), data)
        steps.append(item)
        step_ids.append(out_ids[0, t] if squeeze_b else out_ids[:, t])
      data = tuple(steps)
      data_ids = tuple(step_ids)
    else:
      data = nest.map_structure(fix, data)
      data_ids = fix(out_ids)
    probabilities = probs[0] if squeeze_b else probs
    return data, BufferInfo(ids=data_ids, probabilities=probabilities)

  def _get_next_generic(self, B, T, ids, offs, min_val, max_val):
    """table_fn injection point (reference :57): draws on device, reads via table.read."""
    dev = self._device
    if ids is None:
      ring = table.make_ring([], [], self._batch_size, self._max_length, last_id=self._last_id,
                             ticket=self._ctrl[2:3])
      ids = torch.empty(B, dtype=torch.int64, device=dev)
      offs = torch.empty(B, dtype=torch.int64, device=dev)
      _lib.call('b200rl_rb_draw', ctypes.byref(ring), B, T, self._seed,
                _lib.ptr(self._ctrl[0:1]), _lib.ptr(ids), _lib.ptr(offs), _lib.stream())
    step = torch.arange(T, dtype=torch.int64, device=dev)[None, :]
    rows = (step + ids[:, None]) % self._max_length + offs[:, None] * self._max_length
    data = self._data_table.read(rows)
    out_ids = self._id_table.read(rows)
    prob = np.float32(1.0) / np.float32((max_val - min_val) * self._batch_size)
    probs = torch.full((B,), float(prob), dtype=torch.float32, device=dev)
    return nest.flatten(data), out_ids, probs

  def _as_dataset(self, sample_batch_size=None, num_steps=None, sequence_preprocess_fn=None,
                  num_parallel_calls=None):
    """Dataset of uniformly sampled items (reference :329-367)."""
    if sequence_preprocess_fn is not None:
      raise NotImplementedError('sequence_preprocess_fn is not supported.')
    return _SampleDataset(self, sample_batch_size, num_steps)

  # ---- deterministic pass -----------------------------------------------------------------
  def _deterministic_row_ids(self, sample_batch_size, num_steps):
    """Host restatement of get_row_ids (reference :432-513): list of int64 index arrays."""
    L, Benv = self._max_length, self._batch_size
    lo, hi = _valid_range_ids(self._get_last_id(), L, None)
    if not lo < hi:
      raise InvalidArgumentError(
          'TFUniformReplayBuffer is empty. Make sure to add items before asking the buffer '
          'for data.')
    frames = np.arange(lo, hi, dtype=np.int64)
    shift = self._dataset_window_shift
    drop = self._dataset_drop_remainder

    def windows(seq, keep_partial):
      # tf.data window(num_steps, shift).flat_map(batch(num_steps, drop_remainder))
      s = num_steps if shift is None else shift
      out = []
      start = 0
      while start < len(seq):
        w = seq[start:start + num_steps]
        if len(w) == num_steps or keep_partial:
          out.append(w)
        start += s
      return out

    result = []
    if sample_batch_size is None:
      for b in range(Benv):
        ids = b * L + frames
        if num_steps is None:
          result.extend(list(ids))
        else:
          result.extend(np.stack(w) for w in windows(list(ids), keep_partial=not drop))
    else:
      segs = np.arange(Benv, dtype=np.int64)
      groups = [segs[i:i + sample_batch_size] for i in range(0, Benv, sample_batch_size)]
      if drop:
        groups = [g for g in groups if len(g) == sample_batch_size]
      for g in groups:
        rows = [frames[j] + g * L for j in range(len(frames))]  # one [len(g)] array per frame
        if num_steps is None:
          result.extend(rows)
        else:
          result.extend(np.stack(w).T for w in windows(rows, keep_partial=False))
    return result

  def _single_deterministic_pass_dataset(self, sample_batch_size=None, num_steps=None,
                                         sequence_preprocess_fn=None, num_parallel_calls=None):
    """Dataset that returns entries in fixed order (reference :369-531)."""
    if sequence_preprocess_fn is not None:
      raise NotImplementedError('sequence_preprocess_fn is not supported.')
    if (self._dataset_drop_remainder and sample_batch_size is not None and
        sample_batch_size > self._batch_size):
      raise ValueError(
          'sample_batch_size ({}) > self.batch_size ({}) and '
          'dataset_drop_remainder is True.  In '
          'this case, ALL data will be dropped by the deterministic dataset.'.format(
              sample_batch_size, self._batch_size))
    if (self._dataset_drop_remainder and num_steps is not None and
        num_steps > self._max_length):
      raise ValueError(
          'num_steps_size ({}) > self.max_length ({}) and '
          'dataset_drop_remainder is True.  In '
          'this case, ALL data will be dropped by the deterministic dataset.'.format(
              num_steps, self._max_length))

    def factory():
      for id_ in self._deterministic_row_ids(sample_batch_size, num_steps):
        id_t = torch.as_tensor(np.asarray(id_), dtype=torch.int64, device=self._device)
        data = self._data_table.read(id_t % self._capacity)  # reference :518-524
        yield data, BufferInfo(ids=id_t, probabilities=())

    return _ListDataset(factory)

  # ---- gather_all / clear -------------------------------------------------------------------
  def _gather_all(self):
    """All items, shape [batch_size, n, ...] in age order (reference :533-557)."""
    lo, hi = _valid_range_ids(self._get_last_id(), self._max_length)
    n = hi - lo
    dev = self._device
    with torch.cuda.device(dev):
      if self._fused:
        outs = [torch.empty((self._batch_size, n) + s.shape, dtype=s.dtype, device=dev)
                for s in self._flat_specs]
        if n > 0:
          _lib.call('b200rl_rb_gather_all', ctypes.byref(self._ring), n, _lib.ptr_array(outs),
                    _lib.stream())
        return nest.pack_sequence_as(self._data_spec, outs)
      ids = torch.arange(lo, hi, dtype=torch.int64, device=dev)
      rows = (ids % self._max_length)[None, :] + (
          torch.arange(self._batch_size, dtype=torch.int64, device=dev) *
          self._max_length)[:, None]
      return self._data_table.read(rows)

  def _clear(self, clear_all_variables=False):
    """Resets the buffer; table contents are only unlinked unless asked (reference :559-579)."""
    with torch.cuda.device(self._device):
      if self._fused:
        _lib.call('b200rl_rb_clear', ctypes.byref(self._ring), int(bool(clear_all_variables)),
                  _lib.stream())
      else:
        self._last_id.fill_(-1)
        if clear_all_variables:
          for v in self._data_table.variables() + self._id_table.variables():
            v.zero_()
    self._last_id_host = -1
    common.record_host_effect(self._note_clear)

  def clear(self, clear_all_variables=False):
    return self._clear(clear_all_variables)
