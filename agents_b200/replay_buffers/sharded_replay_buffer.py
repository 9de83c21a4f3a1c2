"""Segment-sharded TFUniformReplayBuffer for data-parallel runs.

The reference's buffer is already `batch_size` independent segments
(replay_buffers/tf_uniform_replay_buffer.py:64-94,141-143) and a sampled window never crosses a
segment (:291), so the ring shards by segment with no traffic on `add_batch`: rank g owns
segments [g * B_env / G, (g + 1) * B_env / G) and the environments that feed them.

Two sampling modes:

* `get_next(B_local, T)` -- the fast path used by bench.py: every rank draws B_local windows over
  ITS segments (own Philox key); no exchange.  Each row keeps the marginal probability
  1 / ((max - min) * B_env); the per-rank counts are fixed instead of multinomial.
* `get_next_global(B, T)` -- parity mode: every rank draws the SAME global `(ids, offsets)[B]`
  (same key and call counter as a single-GPU buffer with all B_env segments), gathers the windows
  whose segment it owns and ONE all_to_all per leaf delivers to rank g the windows at batch
  positions [g * B / G, (g + 1) * B / G).  Concatenating the ranks' results reproduces the
  single-GPU batch bit for bit (tests/dist_parity_main.py).

The routing plan (who sends which batch positions to whom, and the permutation that puts received
rows back into batch order) is a pure function of the draws: `routing_plan` below.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from agents_b200 import _lib
from agents_b200.replay_buffers import table
from agents_b200.replay_buffers import tf_uniform_replay_buffer as rb_mod
from agents_b200.utils import nest


def routing_plan(offsets, batch, world, rank, global_segments):
  """For the draws `offsets[batch]` (global segment index per window) returns, for `rank`:

    send_pos     batch positions this rank owns, ordered by destination rank, then position
    send_counts  [world] how many of them go to each destination
    recv_counts  [world] how many windows of this rank's batch slice come from each owner
    place        permutation: received row i (rows arrive grouped by owner rank, each group in
                 increasing batch position) belongs at slice-local position place[i]
  """
  offsets = np.asarray(offsets, dtype=np.int64)
  if batch % world or global_segments % world:
    raise ValueError('batch and segment count must divide over the replicas.')
  seg_per, b_per = global_segments // world, batch // world
  owner = offsets // seg_per
  dest = np.arange(batch, dtype=np.int64) // b_per
  mine = np.nonzero(owner == rank)[0]
  order = np.lexsort((mine, dest[mine]))
  send_pos = mine[order]
  send_counts = np.bincount(dest[send_pos], minlength=world).astype(np.int64)
  lo = rank * b_per
  slice_pos = np.arange(lo, lo + b_per, dtype=np.int64)
  slice_owner = owner[slice_pos]
  arrive = np.lexsort((slice_pos, slice_owner))            # grouped by owner, then position
  recv_counts = np.bincount(slice_owner, minlength=world).astype(np.int64)
  place = slice_pos[arrive] - lo
  return send_pos, send_counts, recv_counts, place


class ShardedUniformReplayBuffer(object):

  def __init__(self, data_spec, batch_size, max_length=1000, strategy=None, device='cuda', seed=0):
    from agents_b200.train.utils import strategy_utils
    self._strategy = strategy or strategy_utils.get_strategy()
    self._world, self._rank = self._strategy.num_replicas_in_sync, self._strategy.rank
    self._global_segments = int(batch_size)
    lo, hi = self._strategy.shard_range(self._global_segments)
    self._lo, self._local_segments = lo, hi - lo
    self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    # the local ring; its own key is only used by the fast local sampling path
    self._rb = rb_mod.TFUniformReplayBuffer(data_spec, batch_size=self._local_segments,
                                            max_length=max_length, device=device,
                                            seed=(self._seed + 0x9E3779B97F4A7C15 * (self._rank + 1)))
    self._global_calls = torch.zeros(2, dtype=torch.int64, device=self._rb.device)

  @property
  def local(self):
    return self._rb

  @property
  def data_spec(self):
    return self._rb.data_spec

  @property
  def segment_range(self):
    return self._lo, self._lo + self._local_segments

  def add_batch(self, items):
    """`items`: this rank's `[B_env / G, ...]` slice of a driver step."""
    return self._rb.add_batch(items)

  def get_next(self, sample_batch_size=None, num_steps=None):
    return self._rb.get_next(sample_batch_size=sample_batch_size, num_steps=num_steps)

  def draw_global(self, batch, num_steps):
    """The `(ids, offsets)` a single buffer with all segments would draw for this call."""
    rb = self._rb
    ring = table.make_ring([], [], self._global_segments, rb.max_length, last_id=rb._last_id,
                           ticket=rb._ctrl[2:3])
    ids = torch.empty(batch, dtype=torch.int64, device=rb.device)
    offs = torch.empty(batch, dtype=torch.int64, device=rb.device)
    _lib.call('b200rl_rb_draw', ctypes.byref(ring), batch, num_steps, self._seed,
              _lib.ptr(self._global_calls[0:1]), _lib.ptr(ids), _lib.ptr(offs), _lib.stream())
    return ids, offs

  def get_next_global(self, sample_batch_size, num_steps):
    """Rank's `[B / G, T, ...]` slice of the globally drawn batch + BufferInfo (parity mode)."""
    rb, B, T = self._rb, int(sample_batch_size), int(num_steps)
    dev = rb.device
    ids, offs = self.draw_global(B, T)
    offs_h = offs.cpu().numpy()                                # the plan is built on the host
    send_pos, send_counts, recv_counts, place = routing_plan(
        offs_h, B, self._world, self._rank, self._global_segments)
    n_send = int(send_pos.size)
    pos_dev = torch.as_tensor(send_pos, device=dev)
    flat_specs = rb._flat_specs
    b_per = B // self._world
    if n_send:
      data, info = rb.get_next(sample_batch_size=n_send, num_steps=T, ids=ids[pos_dev],
                               batch_offsets=offs[pos_dev] - self._lo)
      send = nest.flatten(data) + [info.ids]
    else:
      send = [torch.empty((0, T) + s.shape, dtype=s.dtype, device=dev) for s in flat_specs]
      send.append(torch.empty((0, T), dtype=torch.int64, device=dev))
    place_dev = torch.as_tensor(place, device=dev)
    out = []
    for leaf in send:
      recv = torch.empty((b_per,) + tuple(leaf.shape[1:]), dtype=leaf.dtype, device=dev)
      if self._world > 1:
        dist.all_to_all_single(recv, leaf.contiguous(), output_split_sizes=recv_counts.tolist(),
                               input_split_sizes=send_counts.tolist())
      else:
        recv.copy_(leaf)
      ordered = torch.empty_like(recv)
      ordered[place_dev] = recv                                  # back into batch order
      out.append(ordered)
    lo_id, hi_id = rb_mod._valid_range_ids(rb._get_last_id(), rb.max_length, T)
    prob = np.float32(1.0) / np.float32((hi_id - lo_id) * self._global_segments)
    probs = torch.full((b_per,), float(prob), dtype=torch.float32, device=dev)
    data = nest.pack_sequence_as(rb.data_spec, out[:-1])
    return data, rb_mod.BufferInfo(ids=out[-1], probabilities=probs)
