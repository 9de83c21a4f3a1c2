"""Model-based property test of the Reverb-model server's row bookkeeping (reverb_local.py).

Random interleavings of append / create_item / end_episode / sample / reset on a small table are
replayed against a trivial model (python lists of appended values); after every operation

  * what an item reads back from the step store equals the values that were appended,
  * the number of live rows equals the number of distinct steps referenced by items or kept alive
    by the writer window,
  * rows on the free list are referenced by nothing, and one write launch never carries a row twice
    (asserted by the NumpyStepStore double).
"""
import numpy as np
import pytest

hypothesis = pytest.importorskip('hypothesis')
from hypothesis import given, settings, strategies as st  # noqa: E402

from agents_b200.replay_buffers import reverb_local as reverb  # noqa: E402
import py_env_mocks  # noqa: E402

OPS = st.lists(st.tuples(st.sampled_from(['append', 'item', 'end', 'sample', 'batch', 'reset', 'flush']),
                         st.integers(min_value=1, max_value=4)), min_size=1, max_size=60)


@settings(max_examples=100, deadline=None)
@given(ops=OPS, max_size=st.integers(1, 5), keep=st.integers(1, 4), times=st.sampled_from([0, 1, 2]),
       stage=st.integers(1, 5), sampler=st.sampled_from(['uniform', 'prioritized', 'fifo']))
def test_row_pool_matches_model(ops, max_size, keep, times, stage, sampler):
  sampler = {'uniform': reverb.selectors.Uniform(), 'prioritized': reverb.selectors.Prioritized(0.5),
             'fifo': reverb.selectors.Fifo()}[sampler]
  table = reverb.Table('t', sampler=sampler, remover=reverb.selectors.Fifo(),
                       max_size=max_size, max_times_sampled=times,
                       rate_limiter=reverb.rate_limiters.MinSize(1))
  stores = []

  def factory(specs, cap):
    stores.append(py_env_mocks.NumpyStepStore(specs, cap, stage=stage))
    return stores[-1]
  srv = reverb.Server([table], store_factory=factory, initial_step_capacity=2, seed=1)
  writer = srv.localhost_client().trajectory_writer(num_keep_alive_refs=keep)
  episode = []            # model: values of the current episode
  items = {}              # model: key -> list of values
  counter = [0]

  def check():
    # contents
    for it in table._dense:
      got = it.store.read(it.rows)[0]
      np.testing.assert_array_equal(got, items[it.key])
    assert set(items) == {it.key for it in table._dense}
    # dense arrays follow the swap-removes: keys, and the [position, T] row matrix while it exists
    for i, it in enumerate(table._dense):
      assert it.pos == i and table._keys[i] == it.key
      if table._rows2d is not None:
        np.testing.assert_array_equal(table._rows2d[i], it.rows)
    if table._rows2d is None and table._dense:
      assert len({it.rows.shape[0] for it in table._dense}) > 1 or not table._rows_ok
    if table._tree is not None:          # sum tree == p^0.5 at the live positions, 0 beyond them
      n = table.current_size
      for i, it in enumerate(table._dense):
        assert table._tree.get(i) == pytest.approx(it.priority ** 0.5)
      assert table._tree.get(n) == 0.0
      assert table._tree.total == pytest.approx(sum(it.priority ** 0.5 for it in table._dense))
    # live rows = distinct steps referenced by items + the writer's keep-alive window
    (pool,) = srv._pools.values() if srv._pools else (None,)
    if pool is None:
      return
    ref = set()
    for it in table._dense:
      ref.update(int(r) for r in it.rows)
    ref.update(writer._episode_rows[writer._released_upto:])
    assert srv.live_rows() == len(ref)
    assert not (set(pool._free) | set(pool._deferred)) & ref
    assert len(set(pool._free)) == len(pool._free)

  for op, k in ops:
    if op == 'append':
      counter[0] += 1
      writer.append(np.int64(counter[0]))
      episode.append(counter[0])
    elif op == 'item' and episode:
      k = min(k, len(episode), keep)
      before = {it.key for it in table._dense}
      writer.create_item('t', trajectory=writer.history[-k:], priority=(counter[0] * 7 + k) % 4)
      new = [it for it in table._dense if it.key not in before]
      for it in new:
        items[it.key] = episode[-k:]
      for key in before - {it.key for it in table._dense}:      # FIFO eviction
        del items[key]
      if not new:                                             # max_size reached by older items only
        pass
    elif op == 'end':
      writer.end_episode()
      episode = []
    elif op == 'sample' and table.can_sample(1):
      try:
        picked = table.sample(min(k, 2))
      except reverb.RateLimited:          # the first draw used up the last item (max_times_sampled)
        picked = []
      for it, info in picked:
        np.testing.assert_array_equal(it.store.read(it.rows)[0], items[it.key])
        assert info.times_sampled >= 1
      reverb.Table.release_samples(picked)
      alive = {it.key for it in table._dense}
      for key in list(items):
        if key not in alive:                                   # reached max_times_sampled
          del items[key]
    elif op == 'batch' and table.can_sample(1):
      drawn = table.sample_rows(3)
      if drawn is not None:                # fixed-length items, no max_times_sampled
        pool, rows, info = drawn
        by_key = {it.key: it for it in table._dense}
        got = pool.read(rows)[0]
        for b in range(3):
          np.testing.assert_array_equal(rows[b], by_key[int(info.key[b])].rows)
          np.testing.assert_array_equal(got[b], items[int(info.key[b])])
          assert info.times_sampled[b] >= 1 and 0 < info.probability[b] <= 1
      else:
        assert (times > 0 or isinstance(sampler, reverb.selectors.Fifo) or
                len({it.rows.shape[0] for it in table._dense}) > 1 or not table._rows_ok)
    elif op == 'reset':
      table.reset()
      items.clear()
    elif op == 'flush':
      writer.flush()
    check()
  writer.close()
  table.reset()
  assert srv.live_rows() == 0
