"""Reverb-model table server with its steps in HBM (reverb_local.HbmStepStore): the data paths of
tests/test_reverb_host.py on the GPU — staged appends -> one b200rl_rb_write_rows launch per flush,
sampled `[B, T]` row matrices -> one b200rl_rb_read_rows launch — checked byte for byte against
a host mirror of everything that was appended (Atari-shape rows take the TMA bulk-copy kernels)."""
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

pytestmark = pytest.mark.gpu


def _uniform(name, max_size, **kw):
  return reverb.Table(name, sampler=reverb.selectors.Uniform(), remover=reverb.selectors.Fifo(),
                      max_size=max_size, rate_limiter=reverb.rate_limiters.MinSize(1), **kw)


def test_queue_table_through_hbm(cuda):  # reverb_utils_test.py:46-71
  table = reverb.Table.queue('q', 3)
  server = reverb.Server([table], device=cuda)
  replay = reverb_replay_buffer.ReverbReplayBuffer(
      tensor_spec.TensorSpec((), torch.int64), 'q', local_server=server, sequence_length=1)
  with replay.py_client.trajectory_writer(num_keep_alive_refs=1) as writer:
    for i in range(3):
      writer.append(i)
      writer.create_item('q', trajectory=writer.history[-1:], priority=1)
  it = iter(replay.as_dataset(sample_batch_size=1))
  for i in range(3):
    sample, _ = next(it)
    assert sample.is_cuda and sample.dtype == torch.int64 and sample.cpu().tolist() == [[i]]
  with pytest.raises(reverb.RateLimited):
    next(it)
  assert server.live_rows() == 0


def test_observer_to_dataset_with_driver(cuda):  # reverb_replay_buffer_test.py:99-124,311-353
  env = py_env_mocks.EpisodeCountingEnv(10)
  tss = env.time_step_spec()
  spec = trajectory.Trajectory(tss.step_type, tss.observation, env.action_spec(), (),
                               tss.step_type, tss.reward, tss.discount)
  table = _uniform('test_table', 100)
  server = reverb.Server([table], device=cuda, initial_step_capacity=16, stage_steps=7)
  observer = reverb_utils.ReverbAddTrajectoryObserver(server.localhost_client(), 'test_table',
                                                      sequence_length=10)
  policy = random_py_policy.RandomPyPolicy(tss, env.action_spec(), seed=1)
  py_driver.PyDriver(env, policy, observers=[observer], max_steps=50).run(env.reset())
  observer.close()
  replay = reverb_replay_buffer.ReverbReplayBuffer(spec, 'test_table', local_server=server,
                                                   sequence_length=10)
  for num_steps in (None, 10, 5, 2, 1):
    t = num_steps or 10
    for sample, info in replay.as_dataset(5, num_steps=num_steps).take(6):
      episode, step = (x.cpu().numpy() for x in sample.observation)
      assert episode.shape == (5, t) and sample.reward.shape == (5, t) and info.key.shape == (5,)
      for n in range(t):
        np.testing.assert_array_equal(episode[:, 0], episode[:, n])
        np.testing.assert_array_equal(step[:, 0] + n, step[:, n])

  def preprocess(traj):                      # sees whole [S, ...] device sequences
    episode, step = traj.observation
    assert step.shape == (10,) and step.is_cuda
    return traj.replace(observation=(episode, step + 1000))
  ds = replay.as_dataset(num_steps=2, sample_batch_size=3, sequence_preprocess_fn=preprocess)
  for sample, _ in ds.take(4):
    step = sample.observation[1].cpu().numpy()
    assert step.shape == (3, 2) and (step >= 1000).all()
    np.testing.assert_array_equal(step[:, 0] + 1, step[:, 1])


def _atari_step(i):
  """A numpy Trajectory whose every leaf encodes the step index i."""
  rng = np.random.default_rng(i)
  obs = rng.integers(0, 256, size=(84, 84, 4), dtype=np.uint8)
  return trajectory.Trajectory(
      step_type=np.int32(ts.StepType.MID), observation=obs, action=np.int32(i % 6), policy_info=(),
      next_step_type=np.int32(ts.StepType.MID), reward=np.float32(i) * 0.5, discount=np.float32(1.0))


def test_atari_rows_round_trip_with_growth_and_recycling(cuda):
  """300 Atari-shape steps (28 224-byte observation rows: the TMA bulk-copy write / gather kernels)
  through windows of 4 with stride 1 into a 40-item FIFO table: the store starts at 32 rows and
  must grow, evicted windows must hand their rows back, and every sampled window must equal the
  host copy of the steps it was built from."""
  spec = trajectory.Trajectory(
      tensor_spec.TensorSpec((), torch.int32, 'step_type'),
      tensor_spec.TensorSpec((84, 84, 4), torch.uint8, 'observation'),
      tensor_spec.TensorSpec((), torch.int32, 'action'), (),
      tensor_spec.TensorSpec((), torch.int32, 'next_step_type'),
      tensor_spec.TensorSpec((), torch.float32, 'reward'),
      tensor_spec.TensorSpec((), torch.float32, 'discount'))
  table = _uniform('atari', 40)
  server = reverb.Server([table], device=cuda, initial_step_capacity=32, stage_steps=16, seed=3)
  observer = reverb_utils.ReverbTrajectorySequenceObserver(server.localhost_client(), 'atari',
                                                           sequence_length=4, stride_length=1)
  mirror = [_atari_step(i) for i in range(300)]
  replay = reverb_replay_buffer.ReverbReplayBuffer(spec, 'atari', local_server=server,
                                                   sequence_length=4)

  def check(n_batches, newest):
    seen = set()
    for sample, _ in replay.as_dataset(sample_batch_size=8).take(n_batches):
      obs = sample.observation.cpu().numpy()
      act = sample.action.cpu().numpy()
      rew = sample.reward.cpu().numpy()
      assert obs.shape == (8, 4, 84, 84, 4) and sample.observation.is_cuda
      for b in range(8):
        first = int(round(float(rew[b, 0]) / 0.5))
        seen.add(first)
        for t in range(4):
          want = mirror[first + t]
          np.testing.assert_array_equal(obs[b, t], want.observation)
          assert act[b, t] == want.action and rew[b, t] == want.reward
    # only windows that start within the last 40 can still be in the table
    assert seen and min(seen) >= newest - 3 - 39 and max(seen) <= newest - 3
    return seen

  for i in range(150):
    observer(mirror[i])
  check(6, newest=149)
  for i in range(150, 300):
    observer(mirror[i])
  seen = check(12, newest=299)
  assert len(seen) > 20
  assert table.current_size == 40
  # 40 windows of 4 with stride 1 pin 43 distinct steps; the writer keeps its last 5 alive
  assert server.live_rows() == 43
  observer.close()
  assert server.live_rows() == 43
  replay.clear()
  assert server.live_rows() == 0
  (pool,) = server._pools.values()
  assert 64 <= pool.store.capacity <= 128       # grew from 32, far below the 300 steps written


def test_episode_observer_variable_lengths(cuda):  # reverb_replay_buffer_test.py:262-309
  table = _uniform('episodes', 100)
  server = reverb.Server([table], device=cuda)
  client = server.localhost_client()
  observer = reverb_utils.ReverbAddEpisodeObserver(client, 'episodes', max_sequence_length=16)
  for n in range(1, 8):
    env = py_env_mocks.EpisodeCountingEnv(n)
    policy = random_py_policy.RandomPyPolicy(env.time_step_spec(), env.action_spec(), seed=n)
    py_driver.PyDriver(env, policy, observers=[observer], max_episodes=1).run(env.reset())
  assert table.current_size == 7
  env = py_env_mocks.EpisodeCountingEnv(3)
  tss = env.time_step_spec()
  spec = trajectory.Trajectory(tss.step_type, tss.observation, env.action_spec(), (),
                               tss.step_type, tss.reward, tss.discount)
  replay = reverb_replay_buffer.ReverbReplayBuffer(spec, 'episodes', local_server=server,
                                                   sequence_length=None)
  lengths = set()
  for sample, _ in replay.as_dataset(sample_batch_size=1).take(40):
    step = sample.observation[1].cpu().numpy()
    n = step.shape[1]                      # the episode's n steps + its boundary step
    lengths.add(n)
    assert sample.step_type.cpu().numpy()[0, -1] == ts.StepType.LAST
    np.testing.assert_array_equal(step[0, :n - 1], np.arange(n - 1))
  assert lengths <= set(range(2, 9)) and len(lengths) >= 4
