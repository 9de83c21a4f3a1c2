"""Device-resident table of nested tensors (Table of the reference, replay_buffers/table.py).

`Table(tensor_spec, capacity)` keeps one `[capacity, *leaf.shape]` CUDA tensor per leaf
(table.py:58-72); `read`/`write` (table.py:86-137) go through b200rl_rb_read_rows /
b200rl_rb_write_rows: ONE launch for all leaves instead of one sparse_read/scatter_update
op per leaf.  Not thread-safe, like the reference (table.py:21).
"""
import ctypes

import numpy as np
import torch

from agents_b200 import _lib
from agents_b200.utils import nest


def make_ring(storages, specs, batch_size, max_length, id_table=None, last_id=None,
              ticket=None):
  """Builds the b200rl_ring_t descriptor over per-leaf storage tensors."""
  if len(storages) > _lib.MAX_LEAVES:
    raise ValueError(f'At most {_lib.MAX_LEAVES} leaves are supported, got {len(storages)}.')
  ring = _lib.Ring()
  ring.num_leaves = len(storages)
  ring.batch_size = int(batch_size)
  ring.max_length = int(max_length)
  ring.id_table = _lib.ptr(id_table)
  ring.last_id = _lib.ptr(last_id)
  ring.ticket = _lib.ptr(ticket)
  for i, (st, sp) in enumerate(zip(storages, specs)):
    ring.leaves[i].storage = _lib.ptr(st)
    ring.leaves[i].row_bytes = sp.row_bytes
  return ring


class Table(object):
  """A table that can store Tensors or nested Tensors."""

  def __init__(self, tensor_spec, capacity, scope='Table', device='cuda'):
    self._tensor_spec = tensor_spec
    self._capacity = int(capacity)
    self._device = torch.device(device)
    self._scope = scope
    flat_specs = nest.flatten(tensor_spec)
    if any(s.row_bytes == 0 for s in flat_specs):
      raise ValueError('Table leaves must have at least one element.')
    names, seen = [], {}
    for s in flat_specs:  # unique slot names (table.py:50-56)
      base = s.name or 'slot'
      k = seen.get(base, 0)
      seen[base] = k + 1
      names.append(base if k == 0 else f'{base}_{k}')
    self._slots = nest.pack_sequence_as(tensor_spec, names)
    self._flat_specs = flat_specs
    self._flat_storage = [
        torch.zeros((self._capacity,) + s.shape, dtype=s.dtype, device=self._device)
        for s in flat_specs
    ]
    self._storage = nest.pack_sequence_as(tensor_spec, self._flat_storage)
    self._slot2idx = {n: i for i, n in enumerate(names)}

  @property
  def slots(self):
    return self._slots

  @property
  def capacity(self):
    return self._capacity

  def variables(self):
    return list(self._flat_storage)

  def _select(self, slots):
    slots = slots or self._slots
    idx = [self._slot2idx[s] for s in nest.flatten(slots)]
    return slots, idx

  def _rows(self, rows):
    rows = torch.as_tensor(rows, dtype=torch.int64, device=self._device)
    return rows.shape, rows.reshape(-1).contiguous()

  def read(self, rows, slots=None):
    """Returns values for the given rows (table.py:86-110)."""
    slots, idx = self._select(slots)
    shape, flat_rows = self._rows(rows)
    n = flat_rows.numel()
    specs = [self._flat_specs[i] for i in idx]
    outs = [torch.empty((n,) + s.shape, dtype=s.dtype, device=self._device) for s in specs]
    if n:
      ring = make_ring([self._flat_storage[i] for i in idx], specs, 1, self._capacity)
      out_ptrs = _lib.ptr_array(outs)
      _lib.call('b200rl_rb_read_rows', ctypes.byref(ring), _lib.ptr(flat_rows), n, out_ptrs,
                None, _lib.stream())
    outs = [o.reshape(tuple(shape) + s.shape) for o, s in zip(outs, specs)]
    return nest.pack_sequence_as(slots, outs)

  def write(self, rows, values, slots=None):
    """Writes values at the given rows (table.py:112-137)."""
    slots, idx = self._select(slots)
    shape, flat_rows = self._rows(rows)
    n = flat_rows.numel()
    specs = [self._flat_specs[i] for i in idx]
    flat_values = nest.flatten(values)
    if len(flat_values) != len(idx):
      raise ValueError('values do not match the slots being written.')
    items = []
    for v, s in zip(flat_values, specs):
      v = torch.as_tensor(v, dtype=s.dtype, device=self._device)
      if tuple(v.shape) == s.shape and n >= 1 and len(shape) >= 1:
        v = v.expand((n,) + s.shape)  # scatter_update broadcasts a single value
      items.append(v.reshape((n,) + s.shape).contiguous())
    if n:
      ring = make_ring([self._flat_storage[i] for i in idx], specs, 1, self._capacity)
      item_ptrs = _lib.ptr_array(items)
      _lib.call('b200rl_rb_write_rows', ctypes.byref(ring), _lib.ptr(flat_rows), n, item_ptrs,
                _lib.stream())
