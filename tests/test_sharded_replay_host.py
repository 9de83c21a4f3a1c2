"""Routing plan of the parity-mode sharded sampler (replay_buffers/sharded_replay_buffer.py):
emulates all ranks of an all_to_all in one process and checks that the ranks' slices concatenate
to exactly the batch the global draws describe."""
import numpy as np
import pytest

from agents_b200.replay_buffers import sharded_replay_buffer as srb


@pytest.mark.parametrize('world,segments,batch', [(2, 8, 16), (4, 8, 32), (8, 256, 256), (1, 4, 8)])
def test_routing_plan_reconstructs_the_global_batch(world, segments, batch):
  rng = np.random.RandomState(world * 100 + batch)
  offsets = rng.randint(0, segments, size=batch)
  ids = rng.randint(0, 1000, size=batch)
  rows = ids * 10000 + offsets                         # what "gathering window b" returns
  plans = [srb.routing_plan(offsets, batch, world, r, segments) for r in range(world)]
  b_per, seg_per = batch // world, segments // world
  # every window is sent exactly once, by its owner
  sent = np.concatenate([p[0] for p in plans])
  assert sorted(sent.tolist()) == list(range(batch))
  for r, p in enumerate(plans):
    assert (offsets[p[0]] // seg_per == r).all() and p[1].sum() == p[0].size
  # emulate all_to_all_single: rank d receives, from each source s in rank order, the part of s's
  # send buffer addressed to d
  for d in range(world):
    chunks = []
    for s in range(world):
      send_pos, send_counts, _, _ = plans[s]
      start = int(send_counts[:d].sum())
      chunks.append(rows[send_pos[start:start + int(send_counts[d])]])
      assert int(send_counts[d]) == int(plans[d][2][s])          # recv_counts mirror send_counts
    recv = np.concatenate(chunks) if chunks else np.zeros(0, np.int64)
    ordered = np.empty(b_per, dtype=np.int64)
    ordered[plans[d][3]] = recv
    np.testing.assert_array_equal(ordered, rows[d * b_per:(d + 1) * b_per])


def test_routing_plan_rejects_uneven_shards():
  with pytest.raises(ValueError):
    srb.routing_plan(np.zeros(6, np.int64), 6, 4, 0, 8)
