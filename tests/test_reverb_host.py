"""Reverb-model observers / table server / ReverbReplayBuffer: host logic, no GPU.

The expected numbers are the ones the reference's own tests hold
(replay_buffers/reverb_utils_test.py:44-627, reverb_replay_buffer_test.py:36-470), replayed
through this package's PyDriver.  The observers run against a recording fake of the writer
protocol (the reference uses mock.MagicMock the same way); the table server runs with the
`NumpyStepStore` test double of tests/py_env_mocks.py in place of the HBM store, so what is checked
here is the item / row bookkeeping (tests/test_reverb_gpu.py repeats the data paths on the GPU).
"""
from unittest import mock

import numpy as np
import pytest
import torch

from agents_b200.drivers import py_driver
from agents_b200.policies import random_py_policy
from agents_b200.replay_buffers import reverb_local as reverb
from agents_b200.replay_buffers import reverb_replay_buffer
from agents_b200.replay_buffers import reverb_utils
from agents_b200.specs import tensor_spec
from agents_b200.trajectories import time_step as ts
from agents_b200.trajectories import trajectory
import py_env_mocks


def _policy(env, seed=0):
  return random_py_policy.RandomPyPolicy(env.time_step_spec(), env.action_spec(), seed=seed)


def _mock_client():
  client, writer = mock.MagicMock(), mock.MagicMock()
  client.trajectory_writer = writer
  writer.return_value = writer
  return client, writer


def _collect(env, observer, max_steps):
  driver = py_driver.PyDriver(env, _policy(env), observers=[observer], max_steps=max_steps)
  driver.run(env.reset())


def _traj(**kw):
  return lambda c: reverb_utils.ReverbAddTrajectoryObserver(c, **kw)


def _episode(**kw):
  return lambda c: reverb_utils.ReverbAddEpisodeObserver(c, **kw)


def _sequence(**kw):
  return lambda c: reverb_utils.ReverbTrajectorySequenceObserver(c, **kw)


_PAD = dict(table_name='test_table', sequence_length=4, pad_end_of_episodes=True,
            tile_end_of_episodes=True)

# (observer factory, episode length, expected items, writers opened, max_steps, appends):
# reverb_utils_test.py:300-398
_WRITES = [
    ('trajectory', _traj(table_name='test_table', sequence_length=2), 3, 3, 1, 4, 5),
    ('episode', _episode(table_name='test_table', max_sequence_length=8, priority=3), 3, 2, 1, 8, 10),
    ('trajectory_stride2', _traj(table_name='test_table', sequence_length=2, stride_length=2),
     3, 2, 1, 4, 5),
    ('pad_stride1', _traj(stride_length=1, **_PAD), 5, 12, 1, 11, 19),
    ('pad_stride2', _traj(stride_length=2, **_PAD), 5, 6, 1, 11, 19),
    ('pad_stride3', _traj(stride_length=3, **_PAD), 5, 4, 1, 11, 19),
    ('pad_stride4', _traj(stride_length=4, **_PAD), 5, 4, 1, 11, 19),
    ('sequence', _sequence(table_name='test_table', sequence_length=2, stride_length=2),
     3, 2, 1, 4, 5),
]


@pytest.mark.parametrize('name,make,ep_len,items,writers,max_steps,appends', _WRITES,
                         ids=[w[0] for w in _WRITES])
def test_observer_writes(name, make, ep_len, items, writers, max_steps, appends):
  client, writer = _mock_client()
  _collect(py_env_mocks.CountingEnv(ep_len), make(client), max_steps)
  assert writer.call_count == writers
  assert writer.append.call_count == appends
  assert writer.create_item.call_count == items


# reverb_utils_test.py:421-478: (factory, write on reset, appends, items, + appends, + items)
_RESETS = [
    ('drop', _traj(table_name='test_table', sequence_length=4, stride_length=4), False, 13, 2, 0, 0),
    ('pad_tile', _traj(stride_length=4, **_PAD), True, 19, 4, 3, 1),
    ('pad_no_tile', _traj(table_name='test_table', sequence_length=4, stride_length=4,
                          pad_end_of_episodes=True, tile_end_of_episodes=False), True, 13, 2, 3, 1),
]


@pytest.mark.parametrize('name,make,write,appends,items,more_appends,more_items', _RESETS,
                         ids=[r[0] for r in _RESETS])
def test_observer_resets(name, make, write, appends, items, more_appends, more_items):
  client, writer = _mock_client()
  observer = make(client)
  _collect(py_env_mocks.CountingEnv(5), observer, 11)
  assert writer.append.call_count == appends
  assert writer.create_item.call_count == items
  observer.reset(write_cached_steps=write)
  assert writer.append.call_count == appends + more_appends
  assert writer.create_item.call_count == items + more_items


def test_reset_with_too_few_cached_steps_raises():
  client, _ = _mock_client()
  observer = reverb_utils.ReverbAddTrajectoryObserver(client, 'test_table', sequence_length=4)
  _collect(py_env_mocks.CountingEnv(10), observer, 2)
  with pytest.raises(ValueError, match='not enough steps remain'):
    observer.reset(write_cached_steps=True)


def test_tile_without_pad_is_rejected():
  client, _ = _mock_client()
  with pytest.raises(ValueError, match='pad_end_of_episodes=True'):
    reverb_utils.ReverbAddTrajectoryObserver(client, 't', 2, tile_end_of_episodes=True)


def test_observer_writes_multi_tables():
  # reverb_utils_test.py:480-502: one item per table per window
  client, writer = _mock_client()
  observer = reverb_utils.ReverbTrajectorySequenceObserver(
      client, table_name=['test_table1', 'test_table2'], sequence_length=3, stride_length=3)
  _collect(py_env_mocks.CountingEnv(3), observer, 6)
  assert writer.create_item.call_count == 2 * (6 // 3)
  tables = [c.kwargs['table'] for c in writer.create_item.call_args_list]
  assert tables == ['test_table1', 'test_table2'] * 2


def test_episodic_observer_overflow_episode_bypass():
  # reverb_utils_test.py:523-548: 3-step episodes fit max_sequence_length=4, 4-step ones never do
  client, writer = _mock_client()
  observer = reverb_utils.ReverbAddEpisodeObserver(client, 'test_table', 4, priority=1,
                                                   bypass_partial_episodes=True)
  _collect(py_env_mocks.CountingEnv(3), observer, 6)
  _collect(py_env_mocks.CountingEnv(4), observer, 6)
  assert writer.create_item.call_count == 1


def test_episodic_observer_overflow_episode_raises():
  client, _ = _mock_client()
  observer = reverb_utils.ReverbAddEpisodeObserver(client, 'test_table', 2, priority=1)
  with pytest.raises(ValueError, match='exceeds `max_sequence_length`'):
    _collect(py_env_mocks.CountingEnv(3), observer, 4)


def test_episodic_observer_validation_and_priority():
  client, _ = _mock_client()
  with pytest.raises(ValueError):
    reverb_utils.ReverbAddEpisodeObserver(client, 'test_table', -1, priority=3)
  observer = reverb_utils.ReverbAddEpisodeObserver(client, 'test_table', 1, priority=3)
  assert observer._priority == 3
  observer.update_priority(4)
  assert observer._priority == 4


def test_close_then_open_gets_a_new_writer():
  client, writer = _mock_client()
  observer = reverb_utils.ReverbAddTrajectoryObserver(client, 'test_table', 2)
  observer.close()
  assert writer.end_episode.call_count == 1 and writer.close.call_count == 1
  with pytest.raises(ValueError, match='Could not obtain writer'):
    observer(None)
  observer.open()
  assert writer.call_count == 2


# ---- the table server with a host store double -------------------------------------------------
def _server(tables, capacity=16, stage=8, **kw):
  stores = []

  def factory(specs, cap):
    stores.append(py_env_mocks.NumpyStepStore(specs, cap, stage=stage))
    return stores[-1]
  srv = reverb.Server(tables, store_factory=factory, initial_step_capacity=capacity, **kw)
  srv.test_stores = stores
  return srv


def _uniform_table(name='uniform_table', max_size=100, min_size=1, **kw):
  return reverb.Table(name, sampler=reverb.selectors.Uniform(), remover=reverb.selectors.Fifo(),
                      max_size=max_size, rate_limiter=reverb.rate_limiters.MinSize(min_size), **kw)


def test_trajectory_observer_on_the_server():
  # reverb_utils_test.py:504-521
  table = _uniform_table()
  client = _server([table]).localhost_client()
  observer = reverb_utils.ReverbAddTrajectoryObserver(client, table.name, sequence_length=2)
  _collect(py_env_mocks.CountingEnv(6), observer, 5)
  assert observer._cached_steps == 5
  assert table.info.current_size == 4


def test_episodic_observer_on_the_server():
  # reverb_utils_test.py:583-603: 3 full episodes and one step
  table = _uniform_table()
  srv = _server([table])
  observer = reverb_utils.ReverbAddEpisodeObserver(srv.localhost_client(), table.name, 8, priority=3)
  _collect(py_env_mocks.CountingEnv(3), observer, 10)
  assert observer._cached_steps == 1
  assert table.info.current_size == 3
  # every episode item holds its 3 steps + the boundary step; the open episode holds one row
  assert srv.live_rows() == 3 * 4 + 1
  observer.close()
  assert srv.live_rows() == 3 * 4


def _scalar_replay(table, server, **kw):
  return reverb_replay_buffer.ReverbReplayBuffer(
      tensor_spec.TensorSpec((), torch.int64), table.name, local_server=server,
      sequence_length=kw.pop('sequence_length', 1), dataset_buffer_size=1, **kw)


def _write_scalars(replay, table_name, n, priority=lambda i: 1):
  with replay.py_client.trajectory_writer(num_keep_alive_refs=1) as writer:
    for i in range(n):
      writer.append(i)
      writer.create_item(table_name, trajectory=writer.history[-1:], priority=priority(i))


def test_queue_table():
  # reverb_utils_test.py:46-71
  table = reverb.Table.queue('test_queue_table', 3)
  replay = _scalar_replay(table, _server([table]))
  _write_scalars(replay, table.name, 3)
  it = iter(replay.as_dataset(sample_batch_size=1, num_steps=None, num_parallel_calls=1))
  for i in range(3):
    sample, info = next(it)
    assert sample.shape == (1, 1) and sample[0, 0] == i and sample.dtype == np.int64
    assert info.times_sampled[0] == 1
  assert table.current_size == 0
  with pytest.raises(reverb.RateLimited):
    next(it)


def test_queue_table_full_rejects_insert():
  table = reverb.Table.queue('q', 2)
  replay = _scalar_replay(table, _server([table]))
  with pytest.raises(reverb.RateLimited, match='is full'):
    _write_scalars(replay, table.name, 3)


def test_uniform_table():
  # reverb_utils_test.py:73-106
  table = _uniform_table(max_size=1000, min_size=3)
  replay = _scalar_replay(table, _server([table]))
  _write_scalars(replay, table.name, 3)
  it = iter(replay.as_dataset(sample_batch_size=1, num_steps=None, num_parallel_calls=1))
  counts = [0] * 3
  for _ in range(1000):
    sample, info = next(it)
    counts[int(sample[0, 0])] += 1
    assert info.probability[0] == pytest.approx(1 / 3) and info.table_size[0] == 3
  assert min(counts) > 200


def test_uniform_table_rate_limited_until_min_size():
  table = _uniform_table(min_size=3)
  replay = _scalar_replay(table, _server([table]))
  _write_scalars(replay, table.name, 2)
  assert not table.can_sample(1)
  with pytest.raises(reverb.RateLimited, match='needs 3'):
    next(iter(replay.as_dataset(sample_batch_size=1)))
  timed = _scalar_replay(table, replay.py_client.server, rate_limiter_timeout_ms=100)
  with pytest.raises(StopIteration):
    next(iter(timed.as_dataset(sample_batch_size=1)))


@pytest.mark.parametrize('sampler', [reverb.selectors.Uniform(), reverb.selectors.Prioritized(1.0)])
def test_table_max_times_sampled(sampler):
  # reverb_utils_test.py:108-153,201-239: 10 batches of 3 exhaust 3 items x 10 samples exactly
  table = reverb.Table('t', sampler=sampler, remover=reverb.selectors.Fifo(), max_size=3,
                       max_times_sampled=10, rate_limiter=reverb.rate_limiters.MinSize(1))
  replay = _scalar_replay(table, _server([table]))
  _write_scalars(replay, table.name, 3, priority=lambda i: i if isinstance(
      sampler, reverb.selectors.Prioritized) else 1)
  assert table.can_sample(3)
  it = iter(replay.as_dataset(sample_batch_size=3, num_parallel_calls=3))
  counts = [0] * 3
  for _ in range(10):
    sample, _ = next(it)
    for v in sample[:, 0]:
      counts[int(v)] += 1
  assert not table.can_sample(3)
  assert counts == [10, 10, 10]
  assert replay.py_client.server.live_rows() == 0


def test_prioritized_table():
  # reverb_utils_test.py:155-199
  table = reverb.Table('p', sampler=reverb.selectors.Prioritized(1.0),
                       remover=reverb.selectors.Fifo(), max_size=3,
                       rate_limiter=reverb.rate_limiters.MinSize(1))
  replay = _scalar_replay(table, _server([table]))
  _write_scalars(replay, table.name, 3, priority=lambda i: i)
  it = iter(replay.as_dataset(sample_batch_size=1, num_steps=None, num_parallel_calls=1))
  counts, keys = [0] * 3, {}
  for _ in range(1000):
    sample, info = next(it)
    counts[int(sample[0, 0])] += 1
    keys[int(sample[0, 0])] = int(info.key[0])
  assert counts[0] == 0 and counts[1] > 250 and counts[2] > 600
  # update_priorities moves all the mass to item 1
  replay.update_priorities(np.array([keys[1], keys[2]]), np.array([5.0, 0.0]))
  assert all(int(next(it)[0][0, 0]) == 1 for _ in range(50))


def test_fifo_remover_evicts_oldest_and_recycles_rows():
  table = _uniform_table(max_size=4)
  srv = _server([table], capacity=8)
  replay = _scalar_replay(table, srv)
  _write_scalars(replay, table.name, 40)
  assert table.current_size == 4
  it = iter(replay.as_dataset(sample_batch_size=4))
  seen = set()
  for _ in range(50):
    seen.update(int(v) for v in next(it)[0][:, 0])
  assert seen == {36, 37, 38, 39}
  assert srv.live_rows() == 4
  assert srv.test_stores[0].capacity == 8            # 40 steps went through 8 rows


def test_step_store_grows_and_keeps_data():
  table = _uniform_table(max_size=1000)
  srv = _server([table], capacity=4, stage=3)
  replay = _scalar_replay(table, srv, sequence_length=None)
  with replay.py_client.trajectory_writer(num_keep_alive_refs=50) as writer:
    for i in range(37):
      writer.append(100 + i)
    writer.create_item(table.name, trajectory=writer.history[:], priority=1)
  assert srv.test_stores[0].capacity >= 37
  sample, _ = next(iter(replay.as_dataset()))
  np.testing.assert_array_equal(sample, 100 + np.arange(37))


def test_writer_keep_alive_window():
  table = _uniform_table()
  client = _server([table]).localhost_client()
  with client.trajectory_writer(num_keep_alive_refs=2) as writer:
    for i in range(4):
      writer.append(i)
    writer.create_item(table.name, trajectory=writer.history[-2:], priority=1)
    with pytest.raises(ValueError, match='kept alive'):
      writer.create_item(table.name, trajectory=writer.history[-3:], priority=1)
    with pytest.raises(ValueError, match='at least one step'):
      writer.create_item(table.name, trajectory=writer.history[2:2], priority=1)
  with pytest.raises(RuntimeError, match='after close'):
    writer.append(5)
  with pytest.raises(ValueError, match='Unknown table'):
    with client.trajectory_writer(1) as w2:
      w2.append(0)
      w2.create_item('nope', trajectory=w2.history[-1:], priority=1)


def test_client_by_address_and_reset():
  table = _uniform_table()
  srv = _server([table])
  client = reverb.Client('localhost:{}'.format(srv.port))
  assert client.server is srv
  with client.trajectory_writer(1) as writer:
    writer.append(np.float32(1.5))
    writer.create_item(table.name, trajectory=writer.history[-1:], priority=1)
  assert client.server_info()[table.name].current_size == 1
  client.reset(table.name)
  assert client.server_info()[table.name].current_size == 0 and srv.live_rows() == 0
  srv.stop()
  with pytest.raises(NotImplementedError, match='No network transport'):
    reverb.Client('localhost:{}'.format(srv.port))
  with pytest.raises(NotImplementedError):
    reverb.Client('10.0.0.1:8000')


# ---- ReverbReplayBuffer (reverb_replay_buffer_test.py) -----------------------------------------
def _data_spec(env):
  tss = env.time_step_spec()
  return trajectory.Trajectory(tss.step_type, tss.observation, env.action_spec(), (),
                               tss.step_type, tss.reward, tss.discount)


class _Fixture(object):

  def __init__(self, steps_per_episode=3):
    self.env = py_env_mocks.EpisodeCountingEnv(steps_per_episode)
    self.spec = _data_spec(self.env)
    self.table = _uniform_table('test_table')
    self.server = _server([self.table], capacity=64)
    self.client = reverb.Client('localhost:{}'.format(self.server.port))

  def insert(self, num_steps, sequence_length=2, env=None):
    env = env or self.env
    obs = reverb_utils.ReverbAddTrajectoryObserver(self.client, self.table.name,
                                                   sequence_length=sequence_length)
    _collect(env, obs, num_steps)
    obs.close()

  def replay(self, sequence_length, **kw):
    return reverb_replay_buffer.ReverbReplayBuffer(self.spec, self.table.name,
                                                   local_server=self.server,
                                                   sequence_length=sequence_length, **kw)


@pytest.mark.parametrize('sequence_length', [None, 2, 4])
def test_dataset_samples_sequential(sequence_length):
  f = _Fixture()
  f.insert(20, sequence_length=sequence_length or 4)
  n = 0
  for sample, _ in f.replay(sequence_length).as_dataset(num_steps=2).take(100):
    episode, step = sample.observation
    assert episode.shape == (2,) and episode[0] == episode[1] and step[0] + 1 == step[1]
    n += 1
  assert n == 100


def test_dataset_with_variable_sequence_length_truncates():
  # reverb_replay_buffer_test.py:126-178
  table = reverb.Table('test_table', sampler=reverb.selectors.Fifo(), remover=reverb.selectors.Fifo(),
                       max_times_sampled=1, max_size=100, rate_limiter=reverb.rate_limiters.MinSize(1))
  server = _server([table])
  client = server.localhost_client()
  for values in ([1, 2, 3], [10, 20, 30, 40, 50]):
    with client.trajectory_writer(10) as writer:
      for v in values:
        writer.append(v)
      writer.create_item('test_table', trajectory=writer.history[-len(values):], priority=5)
  replay = reverb_replay_buffer.ReverbReplayBuffer(
      tensor_spec.TensorSpec((), torch.int64), 'test_table', local_server=server,
      sequence_length=None, rate_limiter_timeout_ms=100)
  it = iter(replay.as_dataset(single_deterministic_pass=True, num_steps=2))
  for want in ([1, 2], [10, 20], [30, 40]):
    data, _ = next(it)
    np.testing.assert_array_equal(data, want)
  with pytest.raises(StopIteration):
    next(it)
  assert server.live_rows() == 0


def test_dataset_with_preprocess():
  f = _Fixture()
  f.insert(10, sequence_length=4)
  replay = f.replay(4)
  for sample, _ in replay.as_dataset(num_steps=2).take(5):
    episode, step = sample.observation
    assert episode[0] == episode[1] and step[0] + 1 == step[1]
    assert step[0] % 2 == 0 and step[1] % 2 == 1

  def preprocess(traj):
    episode, step = traj.observation
    return traj.replace(observation=(episode, step + 1))
  ds = replay.as_dataset(num_steps=2, sample_batch_size=1, sequence_preprocess_fn=preprocess)
  for sample, _ in ds.take(5):
    episode, step = sample.observation
    assert episode[0, 0] == episode[0, 1] and step[0, 0] + 1 == step[0, 1]
    assert step[0, 0] % 2 == 1 and step[0, 1] % 2 == 0


def test_single_episode_dataset():
  f = _Fixture()
  f.insert(3, sequence_length=3)
  for sample, _ in f.replay(None).as_dataset().take(5):
    episode, step = sample.observation
    assert episode.shape == (3,) and step.shape == (3,)
    np.testing.assert_array_equal(episode - episode[:1], [0, 0, 0])
    np.testing.assert_array_equal(step - step[:1], [0, 1, 2])


def test_variable_length_episodes_dataset():
  f = _Fixture()
  for n in range(1, 10):
    f.insert(n, sequence_length=n, env=py_env_mocks.EpisodeCountingEnv(n))
  for sample, _ in f.replay(None).as_dataset(sample_batch_size=1).take(5):
    episode, step = sample.observation
    n = episode.shape[1]
    assert 1 <= n <= 9
    np.testing.assert_array_equal(episode, [[0] * n])
    np.testing.assert_array_equal(step, [list(range(n))])
  with pytest.raises(ValueError, match='different lengths'):
    for _ in f.replay(None).as_dataset(sample_batch_size=8).take(20):
      pass


@pytest.mark.parametrize('sequence_length', [1, 2, 5])
def test_batched_episodes_dataset(sequence_length):
  f = _Fixture()
  f.insert(3 * sequence_length, sequence_length=sequence_length,
           env=py_env_mocks.EpisodeCountingEnv(sequence_length))
  store = f.server.test_stores[0]
  for sample, info in f.replay(None).as_dataset(3).take(5):
    reads = store.reads
    episode, step = sample.observation
    assert episode.shape == (3, sequence_length) and info.key.shape == (3,)
    for n in range(sequence_length):
      np.testing.assert_array_equal(episode[:, 0], episode[:, n])
      np.testing.assert_array_equal(step[:, 0] + n, step[:, n])
  assert store.reads == reads               # ... and each batch was ONE gather
  assert reads <= 5


@pytest.mark.parametrize('num_steps', [1, 2, 5, 10, None])
def test_sequential_ordering(num_steps):
  f = _Fixture()
  f.insert(50, sequence_length=10, env=py_env_mocks.EpisodeCountingEnv(10))
  ds = f.replay(10).as_dataset(5, num_steps=num_steps)
  t = num_steps or 10
  for sample, _ in ds.take(10):
    episode, step = sample.observation
    assert episode.shape == (5, t)
    for n in range(t):
      np.testing.assert_array_equal(episode[:, 0], episode[:, n])
      np.testing.assert_array_equal(step[:, 0] + n, step[:, n])


def test_sample_single_episode():
  f = _Fixture()
  f.insert(100, sequence_length=100, env=py_env_mocks.EpisodeCountingEnv(100))
  n = 0
  for sample, _ in f.replay(100).as_dataset(10, num_steps=5).take(10):
    episode, step = sample.observation
    assert not episode.any()
    for k in range(5):
      np.testing.assert_array_equal(step[:, 0] + k, step[:, k])
    n += 1
  assert n == 10


def test_capacity_size_and_clear():
  f = _Fixture()
  replay = f.replay(None)
  assert replay.capacity == 100 and replay.num_frames() == 0
  f.insert(20)
  assert replay.num_frames() == 19          # reverb_replay_buffer_test.py:410-427
  assert f.replay(20).num_frames() == 19
  replay.clear()
  assert replay.num_frames() == 0 and f.server.live_rows() == 0


def test_argument_errors():
  f = _Fixture()
  with pytest.raises(ValueError, match=r'num_steps > sequence_length'):
    f.replay(2).as_dataset(num_steps=4)
  with pytest.raises(ValueError, match=r'not a multiple of num_steps'):
    f.replay(4).as_dataset(num_steps=3)
  with pytest.raises(ValueError, match=r'either the sampler or the remover is not deterministic'):
    f.replay(None).as_dataset(single_deterministic_pass=True)
  with pytest.raises(ValueError, match='Exactly one of'):
    reverb_replay_buffer.ReverbReplayBuffer(f.spec, 'test_table', 2)
  with pytest.raises(ValueError, match='num_parallel_calls cannot be bigger'):
    f.replay(2).as_dataset(sample_batch_size=2, num_parallel_calls=3)
  replay = f.replay(2)
  for call in (lambda: replay.add_batch(None), replay.get_next, replay.gather_all):
    with pytest.raises(NotImplementedError):
      call()


def test_deterministic_dataset_from_heap_sampler_remover():
  table = reverb.Table('test_table', sampler=reverb.selectors.MaxHeap(),
                       remover=reverb.selectors.MinHeap(), max_size=100, max_times_sampled=0,
                       rate_limiter=reverb.rate_limiters.MinSize(1))
  server = _server([table])
  f = _Fixture()
  replay = reverb_replay_buffer.ReverbReplayBuffer(f.spec, 'test_table', local_server=server,
                                                   sequence_length=None)
  replay.as_dataset(single_deterministic_pass=True)


def test_min_heap_remover_and_max_heap_sampler():
  table = reverb.Table('h', sampler=reverb.selectors.MaxHeap(), remover=reverb.selectors.MinHeap(),
                       max_size=3, rate_limiter=reverb.rate_limiters.MinSize(1))
  replay = _scalar_replay(table, _server([table]))
  _write_scalars(replay, table.name, 6, priority=lambda i: [5, 1, 4, 2, 6, 3][i])
  # inserts evict the lowest priority each time: {5,1,4} -> +2 drops 1 -> +6 drops 2 -> +3 drops 3
  assert sorted(int(it.priority) for it in table._dense) == [4, 5, 6]
  sample, info = next(iter(replay.as_dataset(sample_batch_size=1)))
  assert int(sample[0, 0]) == 4 and info.priority[0] == 6


def test_dead_staged_rows_wait_for_the_flush():
  """A step that dies while still staged (written and dropped inside one staging window) must not
  hand its row to the next append of the same window (NumpyStepStore.commit asserts this)."""
  table = _uniform_table(max_size=1)
  srv = _server([table], capacity=4, stage=8)
  client = srv.localhost_client()
  with client.trajectory_writer(num_keep_alive_refs=1) as writer:
    for i in range(20):                     # keep-alive 1 + max_size 1: rows die almost at once
      writer.append(i)
      writer.create_item(table.name, trajectory=writer.history[-1:], priority=1)
  replay = _scalar_replay(table, srv)
  sample, _ = next(iter(replay.as_dataset(sample_batch_size=1)))
  assert int(sample[0, 0]) == 19
  assert srv.test_stores[0].commits >= 1


def test_mutate_priorities_deletes_and_client_sample():
  table = reverb.Table('m', sampler=reverb.selectors.Prioritized(0.5), remover=reverb.selectors.Lifo(),
                       max_size=4, rate_limiter=reverb.rate_limiters.MinSize(1))
  srv = _server([table])
  client = srv.localhost_client()
  replay = _scalar_replay(table, srv)
  _write_scalars(replay, table.name, 6, priority=lambda i: 1 + i)
  # Lifo remover: every insert beyond max_size evicts the newest item, i.e. the one just written
  assert sorted(int(it.priority) for it in table._dense) == [1, 2, 3, 4]
  keys = {int(it.priority): it.key for it in table._dense}
  client.mutate_priorities(table.name, updates={keys[1]: 0.0, keys[2]: 0.0, 12345: 9.0},
                           deletes=[keys[3], 999])
  assert table.current_size == 3 and srv.live_rows() == 3
  samples = list(client.sample(table.name, num_samples=20))
  assert len(samples) == 20
  assert all(int(s.data[0][0]) == 3 for s in samples)          # value 3 carries priority 4
  assert all(s.info.probability == pytest.approx(1.0) for s in samples)
  assert srv.live_rows() == 3                                    # reader pins were released
  client.update_priorities(table.name, [keys[1]], [7.0])
  seen = {int(s.data[0][0]) for s in client.sample(table.name, num_samples=200)}
  assert seen == {0, 3}


def test_prioritized_all_zero_priorities_sample_uniformly():
  table = reverb.Table('z', sampler=reverb.selectors.Prioritized(1.0), remover=reverb.selectors.Fifo(),
                       max_size=10, rate_limiter=reverb.rate_limiters.MinSize(1))
  replay = _scalar_replay(table, _server([table]))
  _write_scalars(replay, table.name, 4, priority=lambda i: 0)
  it = iter(replay.as_dataset(sample_batch_size=1))
  assert {int(next(it)[0][0, 0]) for _ in range(200)} == {0, 1, 2, 3}


def test_uniform_remover_keeps_the_size_bound():
  table = reverb.Table('u', sampler=reverb.selectors.Uniform(), remover=reverb.selectors.Uniform(),
                       max_size=5, rate_limiter=reverb.rate_limiters.MinSize(1))
  srv = _server([table])
  replay = _scalar_replay(table, srv)
  _write_scalars(replay, table.name, 60)
  assert table.current_size == 5 and srv.live_rows() == 5
  assert len({it.key for it in table._dense}) == 5


def test_two_step_structures_get_two_stores():
  """Writers with differently shaped steps share the server but not a step store."""
  t1, t2 = _uniform_table('a'), _uniform_table('b')
  srv = _server([t1, t2])
  client = srv.localhost_client()
  with client.trajectory_writer(2) as w1, client.trajectory_writer(2) as w2:
    for i in range(3):
      w1.append({'x': np.float32(i), 'y': np.arange(3, dtype=np.int32) + i})
      w1.create_item('a', trajectory=nest_slice(w1.history, -1), priority=1)
      w2.append(np.full((2, 2), i, np.uint8))
      w2.create_item('b', trajectory=w2.history[-1:], priority=1)
    with pytest.raises(ValueError, match='shape'):
      w2.append(np.zeros((3,), np.uint8))
  assert len(srv.test_stores) == 2
  got = next(client.sample('a'))
  assert sorted(s.name for s in srv._pools[next(iter(srv._pools))].store_specs) == ['leaf0', 'leaf1']
  x, y = got.data                      # dict leaves come back in sorted-key order
  assert x.dtype == np.float32 and y.shape == (1, 3) and int(y[0, 0]) == int(x[0])


def nest_slice(history, n):
  from agents_b200.utils import nest
  return nest.map_structure(lambda c: c[n:], history)


def test_prioritized_sum_tree_tracks_inserts_evictions_and_updates():
  """The sampler's sum tree holds p^exponent at every live dense position and 0 elsewhere through
  growth past its initial 64 leaves, FIFO evictions (swap-remove), priority updates and deletes;
  draws follow p^exponent / sum."""
  alpha = 0.7
  table = reverb.Table('p', sampler=reverb.selectors.Prioritized(alpha), remover=reverb.selectors.Fifo(),
                       max_size=150, rate_limiter=reverb.rate_limiters.MinSize(1))
  srv = _server([table], capacity=256)
  replay = _scalar_replay(table, srv)
  rng = np.random.default_rng(0)
  prios = rng.integers(0, 6, size=400).astype(float)

  def check_tree():
    n = table.current_size
    want = np.array([it.priority ** alpha if it.priority > 0 else 0.0 for it in table._dense])
    got = np.array([table._tree.get(i) for i in range(n)])
    np.testing.assert_allclose(got, want, rtol=1e-12)
    assert all(table._tree.get(i) == 0.0 for i in range(n, n + 20))
    assert table._tree.total == pytest.approx(want.sum(), rel=1e-9)

  _write_scalars(replay, table.name, 400, priority=lambda i: prios[i])
  assert table.current_size == 150
  check_tree()
  keys = [it.key for it in table._dense]
  replay.update_priorities(np.array(keys[:40]), rng.integers(0, 9, size=40).astype(float))
  replay.py_client.mutate_priorities(table.name, deletes=keys[40:70])
  assert table.current_size == 120
  check_tree()
  # empirical distribution of 30 000 draws vs p^alpha / sum (5 sigma per item)
  w = np.array([it.priority ** alpha if it.priority > 0 else 0.0 for it in table._dense])
  p = w / w.sum()
  pos = {it.key: i for i, it in enumerate(table._dense)}
  counts = np.zeros(len(p))
  n_draws = 30000
  drawn = table.sample(n_draws)
  for item, info in drawn:
    counts[pos[item.key]] += 1
    assert info.probability == pytest.approx(p[pos[item.key]], rel=1e-9)
  reverb.Table.release_samples(drawn)
  assert srv.live_rows() == table.current_size
  sigma = np.sqrt(n_draws * p * (1 - p)) + 1e-9
  assert np.all(np.abs(counts - n_draws * p) <= 5 * sigma + 1)
  assert counts[p == 0].sum() == 0


@pytest.mark.parametrize('sampler', [reverb.selectors.Uniform(), reverb.selectors.Prioritized(0.8)])
def test_vectorised_batches_follow_the_sampling_law(sampler):
  """`Table.sample_rows` (whole batches drawn as vectors) against the selector's law, its
  SampleInfo arrays and its counters; the table leaves the fast path for good when an item of
  another length arrives and takes it again after a reset."""
  table = reverb.Table('v', sampler=sampler, remover=reverb.selectors.Fifo(), max_size=64,
                       rate_limiter=reverb.rate_limiters.MinSize(1))
  srv = _server([table], capacity=512)
  client = srv.localhost_client()
  rng = np.random.default_rng(5)
  prios = rng.integers(0, 5, size=200).astype(float) + (0 if isinstance(
      sampler, reverb.selectors.Prioritized) else 1)
  with client.trajectory_writer(3) as w:
    for i in range(200):
      w.append(np.int64(i))
      if i:
        w.create_item('v', trajectory=w.history[-2:], priority=prios[i])
  n = table.current_size
  assert n == 64
  if isinstance(sampler, reverb.selectors.Prioritized):
    wgt = np.array([it.priority ** 0.8 if it.priority > 0 else 0.0 for it in table._dense])
  else:
    wgt = np.ones(n)
  p = wgt / wgt.sum()
  pos_of = {it.key: i for i, it in enumerate(table._dense)}
  counts, draws = np.zeros(n), 0
  for _ in range(100):
    pool, rows, info = table.sample_rows(256)
    assert rows.shape == (256, 2) and info.key.shape == (256,) and info.table_size[0] == n
    pos = np.array([pos_of[int(k)] for k in info.key])
    np.testing.assert_array_equal(rows, np.stack([table._dense[i].rows for i in pos]))
    np.testing.assert_allclose(info.probability, p[pos], rtol=1e-9)
    np.testing.assert_array_equal(info.priority, [table._dense[i].priority for i in pos])
    data = pool.read(rows)[0]
    np.testing.assert_array_equal(data[:, 0] + 1, data[:, 1])       # consecutive steps
    counts += np.bincount(pos, minlength=n)
    draws += 256
  sigma = np.sqrt(draws * p * (1 - p))
  assert np.all(np.abs(counts - draws * p) <= 5 * sigma + 1) and counts[p == 0].sum() == 0
  np.testing.assert_array_equal(table._times[:n], counts)
  assert table.info.num_unique_samples == int((counts > 0).sum())
  assert srv.live_rows() == 65 + 0          # 64 overlapping windows of 2 -> 65 steps; no reader pins
  with client.trajectory_writer(4) as w:   # an item of another length: generic path from now on
    for i in range(3):
      w.append(np.int64(1000 + i))
    w.create_item('v', trajectory=w.history[-3:], priority=1)
  assert table.sample_rows(4) is None
  replay = _scalar_replay(table, srv, sequence_length=None)
  it = iter(replay.as_dataset(sample_batch_size=1))    # falls back to item-by-item batches
  lengths = {next(it)[0].shape[1] for _ in range(300)}
  assert lengths == {2, 3} or lengths == {2}
  table.reset()
  with client.trajectory_writer(3) as w:
    for i in range(5):
      w.append(np.int64(i))
      w.create_item('v', trajectory=w.history[-1:], priority=1)
  assert table.sample_rows(4)[1].shape == (4, 1)
